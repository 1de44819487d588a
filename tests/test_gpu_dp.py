"""Multi-GPU parity of the fused data-parallel step (fs_dp_allreduce_adam: gradient all-reduce over NVLink peer
memory + TF-Adam in one kernel) against NCCL all_reduce + the Adam kernel.  Needs >= 2 GPUs; skipped otherwise.
The same batches go through both paths on every rank (tools/dp_check.py, launched under torchrun)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_fused_dp_step_matches_nccl(built_lib):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "dp_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    out = json.loads(line)
    print(out)
    assert out["fused_available"], "peer-memory exchange could not be set up on this box"
    assert out["replicas_identical"] and out["kernel_maxdiff_rel"] <= 1e-6
    assert out["param_maxdiff_rel"] <= 1e-5 and out["loss_maxdiff_rel"] <= 1e-6

"""CPU restatement ORACLE of the reference's two hot paths — TEST INFRASTRUCTURE ONLY.

Nothing in the product package (``faststyle_b200/``), the CLI scripts, or the
timed GPU legs of ``bench.py`` may import this package.  The only permitted
importers are ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs (where it is the thing being
*compared against*, never the thing shipped).

Why a restatement: the reference (ghwatson/faststyle) is Python-2 +
TensorFlow-1.0 script code; TensorFlow is an un-vendored, un-pinned dependency
(reference ``README.md:23``) that is not installed here and cannot be (no
network, no wheel), so the reference cannot be imported or run.  The oracle
restates the reference's graph op by op in torch-CPU (fp32 or fp64), citing
the reference file:line each function follows, with TensorFlow-1.0 op
semantics from SURVEY.md App. C.

Pinning status
--------------
* transform-net forward (``im_transf_net.py``): PINNED by the reference's own
  golden pair ``results/chicago.jpg`` → ``results/{starry,candy}_chicago.jpg``
  with ``models/{starry,candy}_final.ckpt`` (tests/test_oracle_golden.py).
* checkpoint layout: PINNED (byte-exact round trip of the shipped files).
* VGG16 / Gram / losses / gradients / TF-Adam: **parity unpinned** — the
  reference holds no test, fixture or golden for them and its VGG weights file
  is absent (``.gitignore:2``).  They are restated from the source and checked
  only for self-consistency: fp32 vs fp64, hand-computed loss / Adam values and central
  finite differences of the loss path (tests/test_oracle_selfcheck.py).
* TF-1.0 bicubic resize of the input pipeline (``oracle/bicubic.py``): **parity unpinned**
  (restated from the published TF kernel; hand-computed known answers only).
* ``tools/dump_tf1_goldens.py`` is the pinning kit: run under TensorFlow 1.x next to the
  reference checkout it writes ``tests/golden/tf1/*.npz`` that ``tests/test_tf1_goldens.py``
  consumes (skipped while the files are absent).
"""

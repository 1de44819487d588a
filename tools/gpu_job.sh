#!/bin/bash
# One gpurun call: GPU test suite, default bench (+ A/B switches given as "NAME=VAL" args), ncu launch list.
# usage: bash tools/gpu_job.sh TAG [tests|notests] [bench-env ...]
TAG=${1:-r02a}; shift
TESTS=${1:-tests}; shift
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
if [ "$TESTS" = "tests" ]; then
    timeout 1500 python -m pytest tests -m gpu -x -q -s > $OUT/${TAG}_pytest.log 2>&1
    echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
    tail -5 $OUT/${TAG}_pytest.log
fi
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench rc=$?"; python tools/show_bench.py $OUT/${TAG}_bench.json 2>/dev/null | head -40
for kv in "$@"; do
    name=$(echo $kv | tr '=' '_')
    env $kv timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench_${name}.json 2> $OUT/${TAG}_bench_${name}.err
    echo "bench $kv rc=$?"; python tools/show_bench.py $OUT/${TAG}_bench_${name}.json 2>/dev/null | head -12
done

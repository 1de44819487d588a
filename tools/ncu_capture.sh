#!/bin/bash
# ncu evidence for profiles/: run under gpurun (1 GPU).  $1 = tag (e.g. r01p)
# 1. launch list of one train step (gpu__time_duration per launch; cold-cache + serialised: compare SHARES)
# 2. --set full of the tcgen05 conv kernel inside a step (content pass, transform convs, main VGG pass)
# 3. --set full of the Gram kernel
set -u
TAG=${1:-r01x}
OUT=gpurun_out
mkdir -p $OUT
ncu --metrics gpu__time_duration.sum --clock-control none -s 330 -c 250 --csv --log-file $OUT/${TAG}_launches.csv \
    python tools/step_once.py 3 > $OUT/${TAG}_launches.log 2>&1
ncu --set full --clock-control none -k regex:conv3x3_tc -s 61 -c 30 -o $OUT/${TAG}_conv -f \
    python tools/step_once.py 2 > $OUT/${TAG}_conv.log 2>&1
ncu --set full --clock-control none -k regex:gram_tc -s 4 -c 4 -o $OUT/${TAG}_gram -f \
    python tools/step_once.py 2 > $OUT/${TAG}_gram.log 2>&1
M='gpu__time_duration.sum|sm__pipe_tensor_cycles_active|sm__throughput.avg.pct|dram__bytes_read.sum|dram__bytes_write.sum|gpu__dram_throughput|lts__throughput|l1tex__m_xbar2l1tex_read_bytes.sum|launch__registers_per_thread|launch__grid_size|sm__warps_active|sm__cycles_elapsed.avg.per_second|Kernel Name|launch__block_size'
for k in conv gram; do
    ncu -i $OUT/${TAG}_${k}.ncu-rep --page raw --csv > $OUT/${TAG}_${k}_raw.csv 2>/dev/null
    ls -la $OUT/${TAG}_${k}.ncu-rep
done

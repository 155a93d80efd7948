#!/bin/bash
# ncu --set full capture of the fused MLP kernel on the small bench (only after the plain run exited 0)
TAG=${1:-prof}
SMALL="--width 256 --height 256 --steps 1 --warmup 1 --no-cpu-baseline"
mkdir -p gpurun_out
timeout 300 python bench.py $SMALL > gpurun_out/${TAG}_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${2:-nerf_tc} -s 3 -c 1 -f -o gpurun_out/${TAG}_tc python bench.py $SMALL > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full.log

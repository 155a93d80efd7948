#!/bin/bash
# ncu --set full capture of one kernel of an arbitrary bench command (only after the plain run exited 0)
# usage: gpu_prof2.sh TAG KERNEL_REGEX SKIP -- <bench args>
TAG=$1; K=$2; SKIP=$3; shift 4
mkdir -p gpurun_out
timeout 300 python bench.py "$@" > gpurun_out/${TAG}_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -f -o gpurun_out/${TAG} python bench.py "$@" > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log

#!/bin/bash
# launch lists of the fused NeRF training step at the 1-GPU batch (4096 rays) and at one rank's share of the 8-GPU step (512 rays)
for B in 4096 512; do
  CMD="python bench.py --config train --train-batch $B --steps 1 --warmup 3 --no-graph --no-cpu-baseline"
  timeout 200 $CMD > gpurun_out/r2_train_b${B}_plain.log 2>&1 || { tail -5 gpurun_out/r2_train_b${B}_plain.log; exit 1; }
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_train_b4096_launches.csv python bench.py --config train --train-batch 4096 --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/r2_train_b4096_ncu.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_train_b512_launches.csv python bench.py --config train --train-batch 512 --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/r2_train_b512_ncu.log 2>&1
for B in 4096 512; do
  python bench.py --config train --train-batch $B --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_train_b${B}_graph.json 2>/dev/null
  cut -c1-250 gpurun_out/r2_train_b${B}_graph.json
done

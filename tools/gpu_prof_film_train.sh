#!/bin/bash
# ncu captures of the FiLM-SIREN training kernels (pi-GAN gradient step, 4 latents x 64x64, 24+24), each only after the plain run exited 0.
TAG=${1:-film_train}
CMD="python bench.py --config pigan_grad --steps 1 --warmup 3"
timeout 200 $CMD > gpurun_out/${TAG}_plain.log 2>&1 || { tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_list.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:nerf_tc_wgrad -s 12 -c 1 -f -o gpurun_out/${TAG}_wgrad $CMD > gpurun_out/${TAG}_ncu_wgrad.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:film_tc_bwd -s 3 -c 1 -f -o gpurun_out/${TAG}_bwd $CMD > gpurun_out/${TAG}_ncu_bwd.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:film_tc_kernel -s 7 -c 1 -f -o gpurun_out/${TAG}_fwd $CMD > gpurun_out/${TAG}_ncu_fwd.log 2>&1
ls -la gpurun_out/${TAG}_*

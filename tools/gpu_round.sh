#!/bin/bash
# One GPU session: tests, smoke, bench, then (each only after its plain run exited 0) the ncu launch
# list and one full capture of the fused MLP kernel.  Outputs land in gpurun_out/<tag>_*.
TAG=${1:-run}
SMALL="--width 256 --height 256 --steps 1 --warmup 1 --no-cpu-baseline --no-secondary"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -s > gpurun_out/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/${TAG}_smoke.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?" >> gpurun_out/${TAG}_bench.err
if [ "$2" != "noprof" ]; then
timeout 300 python bench.py $SMALL > gpurun_out/${TAG}_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py $SMALL > gpurun_out/${TAG}_ncu_list.log 2>&1
timeout 300 python bench.py $SMALL > gpurun_out/${TAG}_plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nerf_tc -s 2 -c 2 -f -o gpurun_out/${TAG}_tc python bench.py $SMALL > gpurun_out/${TAG}_ncu_full.log 2>&1
fi
tail -3 gpurun_out/${TAG}_pytest.txt; tail -2 gpurun_out/${TAG}_smoke.txt; cat gpurun_out/${TAG}_bench.json | cut -c1-1500; tail -3 gpurun_out/${TAG}_bench.err

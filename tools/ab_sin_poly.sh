#!/bin/bash
# A/B of the polynomial-sine share: one box, each variant twice (ABAB order)
# needs ab_libs/libb2r_p{0,1,2,3}.so: mlp_tc.cu / mlp_tc_siren.cu compiled with -DB2R_SIN_POLY_PAIRS=N, linked with the other objects
cp msra_practice_project_b200/libb2r.so /tmp/libb2r_orig.so
for rep in 1 2; do
for N in 0 1 2 3; do
  cp ab_libs/libb2r_p$N.so msra_practice_project_b200/libb2r.so
  for cfg in grid pigan siren; do
    python bench.py --config $cfg --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    l = l.strip()
    if l.startswith('{'):
        d = json.loads(l); print('pairs $N rep $rep $cfg', round(d['ms_per_step'], 3))
"
  done
done
done
cp /tmp/libb2r_orig.so msra_practice_project_b200/libb2r.so

#!/bin/bash
# round-2 ncu captures (run on the GPU box through gpurun): launch lists and one --set full capture of each step kernel
B="python bench.py --steps 2 --warmup 3 --no-subconfigs --no-cpu-baseline --no-parity-check --e2e-steps 1"
HS_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02_launches_small.csv python tools/small_grid_probe.py > /dev/null 2>&1
HS_GRAPH=0 HS_QP_MAX_CELLS=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file gpurun_out/r02_launches_small_fused.csv python tools/small_grid_probe.py > /dev/null 2>&1
tail -2 gpurun_out/r02_launches_small.csv | cut -c1-260; tail -2 gpurun_out/r02_launches_small_fused.csv | cut -c1-260
ncu --set full --clock-control none --import-source on -k regex:k_step_sp -s 3 -c 1 -f -o gpurun_out/r02_prof_sp $B --workload sp13_2p24 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_step -s 3 -c 1 -f -o gpurun_out/r02_prof_mph $B --workload mph30_2p24 > /dev/null 2>&1
HS_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:k_step_qp -s 5 -c 1 -f -o gpurun_out/r02_prof_qp python tools/small_grid_probe.py > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-subconfigs --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep

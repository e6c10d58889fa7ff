#!/bin/bash
# final-tree ncu captures of the two dominant kernels + launch list of the bench command
B="python bench.py --steps 2 --warmup 3 --no-subconfigs --no-cpu-baseline --no-parity-check --e2e-steps 1"
ncu --set full --clock-control none --import-source on -k regex:k_step_sp -s 3 -c 1 -f -o gpurun_out/r02_prof_sp_final $B --workload sp13_2p24 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_step -s 3 -c 1 -f -o gpurun_out/r02_prof_mph_final $B --workload mph30_2p24 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_final.csv python bench.py --steps 2 --warmup 3 --no-subconfigs --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/*final*

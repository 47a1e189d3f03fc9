#!/bin/bash
# ncu passes on the GPU box: (1) launch list of a short bench run, (2) full capture of the two hot kernels.
set -u
mkdir -p gpurun_out
WL=${1:-c2}
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$WL.csv \
    python bench.py --steps 3 --warmup 3 --workload $WL --no-cpu-baseline > gpurun_out/ncu_bench_$WL.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'fwd_tc_kernel|bwd_tc_kernel|bwd_pair_kernel' -s 6 -c 2 \
    -f -o gpurun_out/prof_$WL python bench.py --steps 2 --warmup 3 --workload $WL --no-cpu-baseline > gpurun_out/ncu_full_$WL.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out

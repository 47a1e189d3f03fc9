#!/bin/bash
# ncu passes on the GPU box: (1) launch list of a short bench run (same command as the bench, graph nodes included),
# (2) full capture of the hot kernels (eager launches so that -k / -s / -c address plain kernel launches).
set -u
mkdir -p gpurun_out
WL=${1:-c2}
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$WL.csv \
    python bench.py --steps 2 --warmup 1 --workload $WL --no-cpu-baseline --no-extras --no-parity > gpurun_out/ncu_bench_$WL.log 2>&1
echo "launch list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'fwd_tc2_kernel|bwd_tc_kernel|bwd_pair_kernel|bwd_flow_kernel' -s 8 -c 2 \
    -f -o gpurun_out/prof_$WL python bench.py --steps 2 --warmup 3 --workload $WL --no-cpu-baseline --no-graph --no-extras --no-parity > gpurun_out/ncu_full_$WL.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out | head -40

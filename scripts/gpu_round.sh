#!/bin/bash
# Round-end style run on one GPU: reference arm, default bench, ncu launch list + full capture of both hot kernels.
set -u
mkdir -p gpurun_out
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
echo "reference rc=$?"; tail -c 600 gpurun_out/bench_reference.json
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench rc=$?"; cat gpurun_out/bench_n1.json
bash scripts/gpu_profile.sh c2

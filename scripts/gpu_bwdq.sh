#!/bin/bash
# backward kernels on the GPU box: parity (repeated), kernel times per variant, timeline of cluster 0.
set -u
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n 6 gpurun_out/$name.log; }
run tcq python -m pytest tests/test_gpu_parity.py -q -x --timeout 120 -k "tc_matches or c1_config or against_oracle or c2_config or c3_config"
CROSSCLR_BWD_VARIANT=2 run tcq_pair python -m pytest tests/test_gpu_parity.py -q -x --timeout 120 -k "tc_matches or against_oracle or c2_config"
for v in 0 2; do echo "--- CROSSCLR_BWD_VARIANT=$v"; CROSSCLR_BWD_VARIANT=$v timeout 300 bash scripts/gpu_sweep.sh 4096,512 8192,512 4096,256 16384,1024; done
for e in 0 7; do CROSSCLR_PAIR_EXP=$e CROSSCLR_PAIR_TRACE=gpurun_out/trace_quad_e$e.txt timeout 120 python scripts/gpu_trace.py; done
python scripts/trace_report.py gpurun_out/trace_quad_e0.txt gpurun_out/trace_quad_e7.txt

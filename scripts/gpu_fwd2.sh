#!/bin/bash
# A/B of the forward kernels on the GPU box: parity, then kernel times per variant / experiment flag.
set -u
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n 12 gpurun_out/$name.log; }
run tc       python -m pytest tests/test_gpu_parity.py -q --timeout 300 -k "tc_matches or c1_config or against_oracle or c2_config or c3_config"
for e in 0 1 2 3; do
  echo "--- CROSSCLR_FWD_EXP=$e"
  CROSSCLR_FWD_EXP=$e bash scripts/gpu_sweep.sh 4096,512 8192,512 16384,1024
done

#!/bin/bash
# Run on the GPU box (via gpurun): staged test groups in separate processes so that one poisoned CUDA context does not
# hide the other results.  Logs land in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv,noheader > gpurun_out/box.txt 2>&1
nproc >> gpurun_out/box.txt
run() { name=$1; shift; echo "=== $name"; timeout 400 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n 4 gpurun_out/$name.log; }
run selftest python -m pytest tests/test_gpu_parity.py -q -x --timeout 120 -k "selftest"
run simt     python -m pytest tests/test_gpu_parity.py -q --timeout 300 -k "simt or known or b1_and"
run tc       python -m pytest tests/test_gpu_parity.py -q --timeout 300 -k "tc_matches or c1_config or against_oracle or c2_config"
run rest     python -m pytest tests/test_gpu_parity.py -q --timeout 300 -k "c3_config or surface or launch_counter or graph"
run variants python -m pytest tests/test_gpu_variants.py -q --timeout 300
run smoke    python __graft_entry__.py --smoke
run bench    python bench.py --steps 20 --warmup 3

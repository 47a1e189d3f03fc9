#!/bin/bash
# perf experiments on the role-specialised backward kernels (results are wrong with exp flags; timing only)
# usage: gpu_pairexp.sh <CROSSCLR_BWD_VARIANT> <exp flags...>
set -u
V=$1; shift
for e in "$@"; do
  echo "--- CROSSCLR_BWD_VARIANT=$V CROSSCLR_PAIR_EXP=$e"
  CROSSCLR_BWD_VARIANT=$V CROSSCLR_PAIR_EXP=$e timeout 300 bash scripts/gpu_sweep.sh 4096,512 8192,512
done

#!/bin/bash
# forward tile order: linear ranges (A resident) vs super-tile interleaved (CROSSCLR_FWD_BLOCKED), by problem size
for shape in "4096 512 1" "16384 1024 1" "65536 512 1" "65536 512 2" "131072 1024 1" "131072 1024 4" "32768 512 1"; do
  for b in 0 1; do
    CROSSCLR_FWD_BLOCKED=$b timeout 200 python scripts/gpu_fwd_time.py $shape 2>&1 | tail -1 | cut -c1-150
  done
done

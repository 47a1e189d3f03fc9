#!/bin/bash
# forward tile order (CROSSCLR_FWD_BLOCKED): 0 linear ranges, 1 super-tile interleaved, 2 row-interleaved, by problem size
for shape in "65536 512 1" "65536 512 4" "32768 512 1" "65536 256 1" "131072 512 8"; do
  for b in 0 1 2; do
    CROSSCLR_FWD_BLOCKED=$b timeout 200 python scripts/gpu_fwd_time.py $shape 2>&1 | tail -1 | cut -c1-150
  done
done

#!/bin/bash
# row-band producer order of the dataflow backward: contiguous sweeps vs column-block-major interleaved (CROSSCLR_FLOW_JMAJOR)
for shape in "16384 512 1 5" "16384 1024 1 5" "65536 512 1 3" "65536 512 4 3" "131072 1024 1 3" "131072 1024 4 3"; do
  for b in 0 1; do
    CROSSCLR_FLOW_JMAJOR=$b timeout 300 python scripts/gpu_flow_time.py $shape 2>&1 | tail -1 | cut -c1-175
  done
done

#!/bin/bash
# producer-epilogue experiments of the dataflow backward (CROSSCLR_FLOW_EXP bits 16 = no arithmetic, 32 = one store per chunk)
mkdir -p gpurun_out
for e in 0 16 32 48; do
  CROSSCLR_FLOW_EXP=$e timeout 120 python scripts/gpu_flow_time.py 4096 512 2>&1 | tail -1 | cut -c1-200
  CROSSCLR_FLOW_EXP=$e CROSSCLR_FLOW_TRACE=gpurun_out/flow_trace_exp$e.txt timeout 120 python scripts/gpu_flow_time.py 4096 512 > /dev/null 2>&1
  sed -n 4,16p gpurun_out/flow_trace_exp$e.txt
  CROSSCLR_FLOW_EXP=$e timeout 120 python scripts/gpu_flow_time.py 16384 512 1 5 2>&1 | tail -1 | cut -c1-200
done

#!/bin/bash
mkdir -p gpurun_out
for v in tr1 tr1t2 tr1t4; do
  echo "=== $v"
  BOKEGO_B200_SO=$PWD/bokego_b200/libbokego_b200_$v.so timeout 300 python tools/trace_handover.py --iters 8000 2>&1 | tail -n 40 | cut -c1-260
done > gpurun_out/r02y_trace.txt 2>&1
cat gpurun_out/r02y_trace.txt

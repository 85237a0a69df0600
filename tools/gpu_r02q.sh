#!/bin/bash
# where does the early hand-over fail?  BK_HANDOVER measurement builds under the cold-L2 stress (48k launches each)
mkdir -p gpurun_out
D=$PWD/bokego_b200
for v in h1 h2 h3 -; do
  so=$D/libbokego_b200$([ "$v" = "-" ] || echo _$v).so
  echo "=== $v"
  for rep in 1 2 3; do BOKEGO_B200_SO=$so timeout 400 python tools/stress_forward.py --iters 8000 --batches 741 2>&1 | tail -n 12 | cut -c1-330; done
  BOKEGO_B200_SO=$so timeout 300 python tools/time_forward_sizes.py 4096 2>&1 | cut -c1-200
done > gpurun_out/r02q_handover.txt 2>&1
cat gpurun_out/r02q_handover.txt

#!/bin/bash
# round 2: which part of the new 3x3 training kernel breaks the exact linearity of the gradient (test_properties_at_scale)?
mkdir -p gpurun_out
D=$PWD/bokego_b200
for v in base - flat1 rot; do
  so=$D/libbokego_b200$([ "$v" = "-" ] || echo _$v).so
  BOKEGO_B200_SO=$so timeout 300 python tools/check_train_linearity.py 2>&1 | tail -n 40 | cut -c1-200 | sed "s/^/$v /"
done > gpurun_out/r02o_linearity.txt 2>&1
cat gpurun_out/r02o_linearity.txt
for v in base - flat1; do
  so=$D/libbokego_b200$([ "$v" = "-" ] || echo _$v).so
  BOKEGO_B200_SO=$so timeout 300 python tools/bench_train.py --positions 576 2048 --precs 5 --no-iterations 2>&1 | cut -c1-260 | sed "s/^/$v /"
done > gpurun_out/r02o_train_ab.txt 2>&1
cat gpurun_out/r02o_train_ab.txt

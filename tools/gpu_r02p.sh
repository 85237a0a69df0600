#!/bin/bash
# round 2: the reordered 3x3 training kernel against the previous build, and one pipeline stage switched off at a time (BK_TC_DIAG)
mkdir -p gpurun_out
D=$PWD/bokego_b200
BOKEGO_B200_SO=$D/libbokego_b200.so timeout 300 python tools/check_train_linearity.py 2>&1 | tail -n 5 | cut -c1-200
for v in base - d2 d4 d8 d16 d30; do
  so=$D/libbokego_b200$([ "$v" = "-" ] || echo _$v).so
  BOKEGO_B200_SO=$so timeout 300 python tools/bench_train.py --positions 576 2048 --precs 5 --no-iterations 2>&1 | cut -c1-260 | sed "s/^/$v /"
done > gpurun_out/r02p_train_ab.txt 2>&1
cat gpurun_out/r02p_train_ab.txt

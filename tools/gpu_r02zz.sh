#!/bin/bash
# round 2: the low-order products of the 3x3 training kernel on kind::f16 (bf16 copies), -DBK_R3_LO_BF16=1 build: parity, linearity, timing
mkdir -p gpurun_out
D=$PWD/bokego_b200
BOKEGO_B200_SO=$D/libbokego_b200_lob16.so timeout 600 python -m pytest tests/test_gpu_train.py -q -s -m gpu > gpurun_out/r02zz_t_train.log 2>&1; echo "lob16 train pytest exit $?"
grep -E "prec 5|passed|failed|^E  " gpurun_out/r02zz_t_train.log | cut -c1-200 | tail -14
BOKEGO_B200_SO=$D/libbokego_b200_lob16.so timeout 300 python tools/check_train_linearity.py 2>&1 | tail -n 3 | cut -c1-200
for v in - lob16; do
  so=$D/libbokego_b200$([ "$v" = "-" ] || echo _$v).so
  BOKEGO_B200_SO=$so timeout 300 python tools/bench_train.py --positions 576 2048 --precs 5 --no-iterations 2>&1 | cut -c1-260 | sed "s/^/$v /"
done > gpurun_out/r02zz_train_ab.txt 2>&1
cat gpurun_out/r02zz_train_ab.txt

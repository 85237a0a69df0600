#!/bin/bash
# round 2: the persistent 3x3 training kernel -- parity, linearity / determinism, timing against the previous build
mkdir -p gpurun_out
D=$PWD/bokego_b200
timeout 600 python -m pytest tests/test_gpu_train.py -q -s -m gpu -x > gpurun_out/r02q_t_train.log 2>&1; echo "train pytest exit $?"
grep -E "grad error|passed|failed|rror" gpurun_out/r02q_t_train.log | tail -16
timeout 300 python tools/check_train_linearity.py 2>&1 | tail -n 5 | cut -c1-200
for v in base - ; do
  so=$D/libbokego_b200$([ "$v" = "-" ] || echo _$v).so
  BOKEGO_B200_SO=$so timeout 300 python tools/bench_train.py --positions 36 576 2048 --precs 5 4 --no-iterations 2>&1 | cut -c1-260 | sed "s/^/$v /"
done > gpurun_out/r02q_train_ab.txt 2>&1
cat gpurun_out/r02q_train_ab.txt
if false; then
  BOKEGO_B200_SO=$D/libbokego_b200_r3prof.so timeout 300 python tools/prof_train_conv3.py 576 > gpurun_out/r02q_conv3_clocks.txt 2>&1
  head -3 gpurun_out/r02q_conv3_clocks.txt | cut -c1-300; sed -n 40,52p gpurun_out/r02q_conv3_clocks.txt; tail -5 gpurun_out/r02q_conv3_clocks.txt | cut -c1-700
fi

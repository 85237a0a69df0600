#!/bin/bash
# round 2, training 3x3 kernel with rotating group buffers / four weight stages / split accumulation chains / split tail:
# parity, then A/B against the previous build (libbokego_b200_base.so) and the measurement variants; A/B of two forward-kernel variants
mkdir -p gpurun_out
D=$PWD/bokego_b200
timeout 600 python -m pytest tests/test_gpu_train.py -q -s -m gpu -x > gpurun_out/r02n_t_train.log 2>&1; echo "train pytest exit $?"
grep -E "grad error|passed|failed|rror" gpurun_out/r02n_t_train.log | tail -30
for v in base - notail chain1; do
  so=$D/libbokego_b200.so; env=
  case $v in base|chain1) so=$D/libbokego_b200_$v.so;; notail) env="BK_TC_NO_TAIL=1";; esac
  env $env BOKEGO_B200_SO=$so timeout 300 python tools/bench_train.py --positions 576 2048 --precs 5 4 --no-iterations 2>&1 | cut -c1-260 | sed "s/^/$v /"
done > gpurun_out/r02n_train_ab.txt 2>&1
cat gpurun_out/r02n_train_ab.txt
for v in - one l04 onel04; do
  so=$D/libbokego_b200$([ "$v" = "-" ] || echo _$v).so
  BOKEGO_B200_SO=$so timeout 300 python tools/time_forward_sizes.py 740 4096 2>&1 | cut -c1-200 | sed "s/^/$v /"
done > gpurun_out/r02n_fwd_ab.txt 2>&1
cat gpurun_out/r02n_fwd_ab.txt

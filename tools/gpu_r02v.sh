#!/bin/bash
# round 2: transposed result stores in the layer-0 / weight-gradient kernel too; the 3x3 kernel's work-item schedule at its boundary sizes
mkdir -p gpurun_out
D=$PWD/bokego_b200
timeout 900 python -m pytest tests/test_gpu_train.py -q -s -m gpu -x > gpurun_out/r02v_t_train.log 2>&1; echo "train pytest exit $?"
grep -E "work items|passed|failed|rror" gpurun_out/r02v_t_train.log | tail -12
timeout 300 python tools/check_train_linearity.py 2>&1 | tail -n 3 | cut -c1-200
for v in base - ; do
  so=$D/libbokego_b200$([ "$v" = "-" ] || echo _$v).so
  BOKEGO_B200_SO=$so timeout 300 python tools/bench_train.py --positions 36 576 2048 --precs 5 4 --no-iterations 2>&1 | cut -c1-260 | sed "s/^/$v /"
done > gpurun_out/r02v_train_ab.txt 2>&1
cat gpurun_out/r02v_train_ab.txt

#!/bin/bash
# round 2: producer restructuring of the weight gradient: bit-identical to the previous build? parity; clock stamps; timing
mkdir -p gpurun_out
D=$PWD/bokego_b200
BOKEGO_B200_SO=$D/libbokego_b200_prev.so timeout 300 python tools/check_train_tail.py dump /tmp/prev.npz 45 576 1100 | tail -n 1
timeout 300 python tools/check_train_tail.py dump /tmp/new.npz 45 576 1100 | tail -n 1
(echo "against the previous build (same positions, 3xTF32):"; python tools/check_train_tail.py compare /tmp/new.npz /tmp/prev.npz) | tee gpurun_out/r02z_identity3.txt
timeout 600 python -m pytest tests/test_gpu_train.py -q -m gpu -x 2>&1 | tail -n 2
BOKEGO_B200_SO=$D/libbokego_b200_r3prof.so timeout 300 python tools/prof_train_conv3.py 576 --backward > gpurun_out/r02z_wgrad_clocks3.txt 2>&1; grep "weight gradient" gpurun_out/r02z_wgrad_clocks3.txt | cut -c1-330
for v in prev - ; do
  so=$D/libbokego_b200$([ "$v" = "-" ] || echo _$v).so
  BOKEGO_B200_SO=$so timeout 300 python tools/bench_train.py --positions 576 2048 --precs 5 4 --no-iterations 2>&1 | cut -c1-260 | sed "s/^/$v /"
done > gpurun_out/r02z_train_ab3.txt 2>&1
cat gpurun_out/r02z_train_ab3.txt

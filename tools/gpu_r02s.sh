#!/bin/bash
mkdir -p gpurun_out
for v in ""; do
BOKEGO_B200_SO=$PWD/bokego_b200/libbokego_b200_r3prof$v.so timeout 300 python tools/prof_train_conv3.py 576 > gpurun_out/r02s_conv3_clocks$v.txt 2>&1; echo "r3prof$v exit $?"
head -2 gpurun_out/r02s_conv3_clocks$v.txt | cut -c1-250
sed -n 14,24p gpurun_out/r02s_conv3_clocks$v.txt
tail -4 gpurun_out/r02s_conv3_clocks$v.txt | cut -c1-300
done

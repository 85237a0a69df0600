#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/check_train_tail.py dump /tmp/tail_on.npz 190 200 220 576
BK_TC_NO_TAIL=1 timeout 300 python tools/check_train_tail.py dump /tmp/tail_off.npz 190 200 220 576
PREC=2 timeout 300 python tools/check_train_tail.py dump /tmp/ffma.npz 190 200 220 576
(echo "split tail against no split tail (3xTF32 tcgen05):"; python tools/check_train_tail.py compare /tmp/tail_on.npz /tmp/tail_off.npz
 echo "3xTF32 tcgen05 (split tail) against FFMA:"; python tools/check_train_tail.py compare /tmp/tail_on.npz /tmp/ffma.npz
 echo "3xTF32 tcgen05 (no split tail) against FFMA:"; python tools/check_train_tail.py compare /tmp/tail_off.npz /tmp/ffma.npz) | tee gpurun_out/r02w_tail_check.txt

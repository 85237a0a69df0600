#!/bin/bash
# long stress of the committed kernels: forward (cold / warm L2, sizes that mix whole, split and empty items) and persistent playouts
mkdir -p gpurun_out
for rep in 1 2 3 4 5 6 7 8 9 10 11 12; do timeout 400 python tools/stress_forward.py --iters 8000 --batches 741 --max-bad 100000 --quiet 2>&1 | tail -n 1; done > gpurun_out/stress_long_forward.txt 2>&1
timeout 900 python tools/stress_forward.py --iters 6000 --batches 1,16,37,81,740,745,1480,4096 --max-bad 100000 --quiet >> gpurun_out/stress_long_forward.txt 2>&1
cat gpurun_out/stress_long_forward.txt
timeout 900 python tools/stress_playout.py --iters 1000 > gpurun_out/stress_long_playout.txt 2>&1; tail -n 3 gpurun_out/stress_long_playout.txt

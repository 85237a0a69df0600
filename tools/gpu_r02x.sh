#!/bin/bash
mkdir -p gpurun_out
D=$PWD/bokego_b200
timeout 900 python -m pytest tests/test_gpu_playout.py tests/test_gpu_parity.py tests/test_gpu_mcts.py -q -m gpu -x 2>&1 | tail -n 2
for rep in 1 2 3; do
  for v in base -; do
    so=$D/libbokego_b200$([ "$v" = "-" ] || echo _$v).so
    BOKEGO_B200_SO=$so timeout 300 python tools/bench_playout.py 512 1024 4096 2>&1 | cut -c1-160 | sed "s/^/$v /"
  done
done > gpurun_out/r02x_ab.txt 2>&1
cat gpurun_out/r02x_ab.txt
timeout 900 python tools/stress_playout.py --iters 400 > gpurun_out/r02x_stress_playout.txt 2>&1; tail -n 2 gpurun_out/r02x_stress_playout.txt
timeout 300 python tools/prof_playout.py > gpurun_out/r02x_prof_playout.txt 2>&1; tail -n 14 gpurun_out/r02x_prof_playout.txt

#!/bin/bash
# per-tile hand-over (head of the next pass under the read-out of the last tile): parity, stress, A/B
mkdir -p gpurun_out
D=$PWD/bokego_b200
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_playout.py tests/test_gpu_precision.py -q -m gpu -x 2>&1 | tail -n 2
for rep in 1 2 3; do timeout 400 python tools/stress_forward.py --iters 8000 --batches 741 --max-bad 100000 --quiet 2>&1 | tail -n 1; done > gpurun_out/r02aa_stress.txt 2>&1
timeout 400 python tools/stress_forward.py --iters 3000 --batches 745,4096,37,16 --max-bad 100000 --quiet >> gpurun_out/r02aa_stress.txt 2>&1
cat gpurun_out/r02aa_stress.txt
for rep in 1 2 3; do
  for v in base pt2 -; do
    so=$D/libbokego_b200$([ "$v" = "-" ] || echo _$v).so
    BOKEGO_B200_SO=$so timeout 300 python tools/time_forward_sizes.py 740 4096 2>&1 | cut -c1-200 | sed "s/^/$v /"
    BOKEGO_B200_SO=$so timeout 300 python tools/bench_playout.py 512 4096 2>&1 | cut -c1-160 | sed "s/^/$v /"
  done
done > gpurun_out/r02aa_ab.txt 2>&1
cat gpurun_out/r02aa_ab.txt
timeout 900 python tools/stress_playout.py --iters 300 2>&1 | tail -n 2
timeout 300 python tools/prof_forward.py > gpurun_out/r02aa_pass_clocks.txt 2>&1; tail -n 12 gpurun_out/r02aa_pass_clocks.txt

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_playout.py tests/test_gpu_mcts.py -q > gpurun_out/r02l_tests.log 2>&1; echo "tests rc=$?"; tail -n 4 gpurun_out/r02l_tests.log
timeout 400 python tools/bench_playout.py 64 1024 1480 2048 > gpurun_out/r02l_playout.jsonl 2> gpurun_out/r02l_playout.err; echo "bench_playout rc=$?"
cut -c1-250 gpurun_out/r02l_playout.jsonl; tail -n 3 gpurun_out/r02l_playout.err
timeout 900 python tools/stress_playout.py --iters 100 > gpurun_out/r02_stress_playout.txt 2>&1; echo "stress rc=$?"; tail -n 3 gpurun_out/r02_stress_playout.txt

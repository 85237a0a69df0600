#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_playout.py -q -x > gpurun_out/r02f_playout_tests.log 2>&1; echo "playout tests rc=$?"
tail -n 12 gpurun_out/r02f_playout_tests.log
timeout 300 python tools/prof_playout.py 512 1 > gpurun_out/r02f_prof_playout.txt 2>&1; echo "prof rc=$?"
cat gpurun_out/r02f_prof_playout.txt
timeout 300 python tools/bench_playout.py 64 512 1024 4096 > gpurun_out/r02f_playout.jsonl 2> gpurun_out/r02f_playout.err; echo "bench_playout rc=$?"
cut -c1-330 gpurun_out/r02f_playout.jsonl; tail -n 5 gpurun_out/r02f_playout.err
timeout 600 python -m pytest tests/test_gpu_mcts.py tests/test_gpu_reference_callers.py -q > gpurun_out/r02f_mcts_tests.log 2>&1; echo "mcts tests rc=$?"
tail -n 4 gpurun_out/r02f_mcts_tests.log

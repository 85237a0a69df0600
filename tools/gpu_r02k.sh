#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_playout.py -q -x > gpurun_out/r02k_playout_tests.log 2>&1; echo "playout tests rc=$?"; tail -n 12 gpurun_out/r02k_playout_tests.log
timeout 300 python tools/prof_playout.py 512 1 > gpurun_out/r02k_prof_playout.txt 2>&1; cat gpurun_out/r02k_prof_playout.txt
timeout 300 python tools/bench_playout.py 148 296 444 512 592 > gpurun_out/r02k_playout.jsonl 2> gpurun_out/r02k_playout.err; echo "bench_playout rc=$?"
cut -c1-250 gpurun_out/r02k_playout.jsonl; tail -n 3 gpurun_out/r02k_playout.err

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/time_forward_sizes.py > gpurun_out/r02g_forward_sizes.jsonl 2>&1; echo "sizes rc=$?"; cat gpurun_out/r02g_forward_sizes.jsonl
timeout 900 python tools/stress_playout.py --iters 120 > gpurun_out/r02_stress_playout.txt 2>&1; echo "stress rc=$?"; tail -n 14 gpurun_out/r02_stress_playout.txt
timeout 300 python -m pytest tests/test_gpu_parity.py -q -k reentrant > gpurun_out/r02g_reentrant.log 2>&1; echo "reentrant rc=$?"; tail -n 3 gpurun_out/r02g_reentrant.log

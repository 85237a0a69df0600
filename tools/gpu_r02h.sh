#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "from_positions or golden or batch_invariant or reentrant" > gpurun_out/r02h_tests.log 2>&1; echo "tests rc=$?"; tail -n 15 gpurun_out/r02h_tests.log
timeout 300 python tools/time_forward_sizes.py 16 740 4096 > gpurun_out/r02h_sizes.jsonl 2>&1; cat gpurun_out/r02h_sizes.jsonl
timeout 300 python tools/time_positions.py > gpurun_out/r02h_positions.jsonl 2>&1; cat gpurun_out/r02h_positions.jsonl
timeout 600 python -m pytest tests/test_gpu_playout.py -q > gpurun_out/r02h_playout.log 2>&1; echo "playout rc=$?"; tail -n 3 gpurun_out/r02h_playout.log

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "from_positions" > gpurun_out/r02i_tests.log 2>&1; echo "tests rc=$?"; tail -n 5 gpurun_out/r02i_tests.log
timeout 300 python tools/time_positions.py 256 1024 4096 16384 > gpurun_out/r02i_positions.jsonl 2>&1; cat gpurun_out/r02i_positions.jsonl

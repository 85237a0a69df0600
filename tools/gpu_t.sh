#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "schedule_boundaries" > gpurun_out/t_sched.log 2>&1; echo "exit $?" >> gpurun_out/t_sched.log; tail -15 gpurun_out/t_sched.log

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mcts.py -x -q -m gpu > gpurun_out/t_mcts.log 2>&1; echo "exit $?" >> gpurun_out/t_mcts.log; tail -30 gpurun_out/t_mcts.log

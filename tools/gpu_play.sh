#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_playout.py -x -q -m gpu > gpurun_out/t_playout.log 2>&1; echo "exit $?" >> gpurun_out/t_playout.log; tail -15 gpurun_out/t_playout.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; echo "ref exit $?"; cat gpurun_out/bench_ref.json

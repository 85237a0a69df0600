#!/bin/bash
# REINFORCE row on the GPU box: parity tests, then timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -q -s > gpurun_out/t_train.log 2>&1
grep -E "grad error|^FAILED|^E  .*Error|passed|failed" gpurun_out/t_train.log | cut -c1-400 | tail -50
[ -n "$SKIP_BENCH" ] || timeout 600 python tools/bench_train.py --out gpurun_out/train.jsonl 2>&1 | tail -20

#!/bin/bash
# round 2, fourth GPU batch: persistent playout kernel -- parity first (bounded by a timeout: a protocol bug must not hang the box), then timings
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_playout.py -q -x > gpurun_out/r02d_playout_tests.log 2>&1; echo "playout tests rc=$?"
tail -n 15 gpurun_out/r02d_playout_tests.log
timeout 300 python tools/bench_playout.py 512 1024 4096 > gpurun_out/r02d_playout.jsonl 2> gpurun_out/r02d_playout.err; echo "bench_playout rc=$?"
cat gpurun_out/r02d_playout.jsonl; tail -n 5 gpurun_out/r02d_playout.err
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_playout.py > gpurun_out/r02d_pytest.log 2>&1; echo "pytest rc=$?"
tail -n 5 gpurun_out/r02d_pytest.log

#!/bin/bash
# round 2, second GPU batch: fused step+encode parity, reference callers, precision, playout timing
mkdir -p gpurun_out
python -m pytest tests/test_gpu_playout.py tests/test_gpu_reference_callers.py tests/test_gpu_precision.py -q -s > gpurun_out/r02b_tests.log 2>&1; echo "tests rc=$?"
python tools/bench_playout.py 512 1024 4096 > gpurun_out/r02b_playout.jsonl 2> gpurun_out/r02b_playout.err; echo "bench_playout rc=$?"
python -m pytest tests -m gpu -q --deselect tests/test_gpu_reference_callers.py --deselect tests/test_gpu_precision.py --deselect tests/test_gpu_playout.py > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?"
tail -n 5 gpurun_out/r02b_tests.log; cat gpurun_out/r02b_playout.jsonl; tail -n 3 gpurun_out/r02b_pytest.log

#!/bin/bash
# final check of the round-2 tree: whole GPU suite, smoke, training step timing
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r02y_pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/r02y_pytest_gpu.log; tail -n 3 gpurun_out/r02y_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02y_smoke.log 2>&1; echo "smoke exit $?"; tail -n 1 gpurun_out/r02y_smoke.log | cut -c1-200
timeout 300 python tools/bench_train.py --positions 576 2048 --precs 5 --no-iterations 2>&1 | cut -c1-250 | tee gpurun_out/r02y_train.txt

#!/bin/bash
# full GPU suite, bench line with the reinforce leg, ncu of the training kernels
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python -m pytest tests/test_gpu_train.py -q -s > gpurun_out/t_train.log 2>&1; grep -E "grad error|grad cosine|passed|failed" gpurun_out/t_train.log | tail -24
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 python tools/bench_train.py --out gpurun_out/train.jsonl > gpurun_out/bench_train.log 2>&1; tail -3 gpurun_out/bench_train.log
timeout 300 python tools/prof_train.py > gpurun_out/prof_train_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_train.csv python tools/prof_train.py > gpurun_out/ncu_list_train.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bk_train_(gemm|conv3)_tc" -s 14 -c 8 -o gpurun_out/prof_train_tc python tools/prof_train.py > gpurun_out/ncu_full_train.log 2>&1
tail -3 gpurun_out/ncu_full_train.log
ls -la gpurun_out | tail -8

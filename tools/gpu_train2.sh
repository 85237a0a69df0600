#!/bin/bash
# full GPU suite, bench line with the reinforce leg, ncu of the training kernels
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 300 python tools/prof_train.py > gpurun_out/prof_train_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_train.csv python tools/prof_train.py > gpurun_out/ncu_list_train.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bk_train_(conv|wgrad)" -s 14 -c 6 -o gpurun_out/prof_train python tools/prof_train.py > gpurun_out/ncu_full_train.log 2>&1
tail -3 gpurun_out/ncu_full_train.log
ls -la gpurun_out | tail -8

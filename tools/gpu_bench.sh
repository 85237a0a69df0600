#!/bin/bash
# GPU session: full gpu test suite, bench line, batch sweep, ncu launch list + one full capture of the conv kernel.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json
for b in 1 16 256 1024 4096 16384; do timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --batch $b 2>/dev/null; done > gpurun_out/sweep.jsonl
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_list.log 2>&1
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bk_forward_tc -s 3 -c 2 -o gpurun_out/prof_fwd python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out

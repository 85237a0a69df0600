#!/bin/bash
# round 2, after the training 3x3 kernel rewrite: whole GPU suite, smoke, bench line, training timings, ncu of the training kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.current.sm,clocks.max.sm,power.draw,power.limit,memory.total --format=csv > gpurun_out/r02t_gpu.txt
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r02t_pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/r02t_pytest_gpu.log; tail -n 3 gpurun_out/r02t_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02t_smoke.log 2>&1; echo "smoke exit $?"; tail -n 1 gpurun_out/r02t_smoke.log | cut -c1-200
timeout 900 python bench.py > gpurun_out/r02t_bench.json 2> gpurun_out/r02t_bench.err; echo "bench exit $?"; tail -n 2 gpurun_out/r02t_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02t_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'selfplay',d['selfplay']['games_per_s'])
print('reinforce',json.dumps(d.get('reinforce'))[:900])
PY
timeout 600 python tools/bench_train.py --out gpurun_out/r02t_train.jsonl > gpurun_out/r02t_bench_train.log 2>&1; grep tc_3xtf32 gpurun_out/r02t_bench_train.log | cut -c1-200; tail -n 3 gpurun_out/r02t_bench_train.log | cut -c1-200
timeout 300 python tools/prof_train.py > gpurun_out/r02t_prof_train_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_train.csv python tools/prof_train.py > gpurun_out/r02t_ncu_list_train.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bk_train_(gemm|conv3)_tc" -s 14 -c 8 -o gpurun_out/prof_train_tc python tools/prof_train.py > gpurun_out/r02t_ncu_full_train.log 2>&1
tail -n 2 gpurun_out/r02t_ncu_full_train.log
BOKEGO_B200_SO=$PWD/bokego_b200/libbokego_b200_r3prof.so timeout 300 python tools/prof_train_conv3.py 576 > gpurun_out/r02t_conv3_clocks.txt 2>&1; head -2 gpurun_out/r02t_conv3_clocks.txt | cut -c1-250
timeout 300 python tools/stress_train.py > gpurun_out/r02t_stress_train.txt 2>&1; tail -n 2 gpurun_out/r02t_stress_train.txt | cut -c1-250

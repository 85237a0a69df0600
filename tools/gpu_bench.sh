#!/bin/bash
# GPU session: full gpu test suite, smoke, bench line (+ reference arm), batch sweep, per-pass clock profile,
# ncu launch list + one full capture of the conv kernel.  Summaries go to profiles/ via tools/ncu_summary.py.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; echo "ref exit $?"
for b in 1 16 81 256 1024 4096 16384; do timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-playouts --batch $b 2>/dev/null; done > gpurun_out/sweep.jsonl
timeout 300 python tools/prof_forward.py --batch 4096 > gpurun_out/prof_4096.log 2>&1
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-playouts > gpurun_out/plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-playouts > gpurun_out/ncu_list.log 2>&1
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-playouts > gpurun_out/plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bk_forward_tc -s 3 -c 2 -o gpurun_out/prof_fwd python bench.py --steps 3 --warmup 3 --no-cpu --no-playouts > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -20

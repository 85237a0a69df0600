#!/bin/bash
# round-2 GPU session: full gpu test suite, smoke, bench (+ reference arm), batch sweep, stress, ncu launch lists and full captures
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit,memory.total --format=csv > gpurun_out/r02_gpu.txt 2>&1
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/r02_pytest_gpu.log; tail -n 3 gpurun_out/r02_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 gpurun_out/r02_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench exit $?"; head -c 400 gpurun_out/r02_bench.json; echo
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2>> gpurun_out/r02_bench.err; echo "ref exit $?"
for b in 1 16 81 256 1024 4096 16384; do timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-playouts --no-train --batch $b 2>/dev/null; done > gpurun_out/r02_sweep.jsonl; echo "sweep lines $(wc -l < gpurun_out/r02_sweep.jsonl)"
timeout 300 python tools/prof_forward.py --batch 4096 > gpurun_out/r02_pass_clocks_b4096.txt 2>&1
timeout 600 python tools/stress_forward.py --iters 2000 > gpurun_out/r02_stress_forward.txt 2>&1; echo "stress exit $?"; tail -n 3 gpurun_out/r02_stress_forward.txt
# ncu: launch list of the bench's timed step, then full captures (each after the same command has exited 0 without ncu)
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-playouts --no-train > gpurun_out/r02_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-playouts --no-train > gpurun_out/r02_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bk_forward_tc -s 3 -c 3 -o gpurun_out/prof_fwd python bench.py --steps 3 --warmup 3 --no-cpu --no-playouts --no-train > gpurun_out/r02_ncu_full.log 2>&1
timeout 300 python tools/run_playout_once.py 512 6 > gpurun_out/r02_playout_once.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_playout.csv python tools/run_playout_once.py 512 6 > gpurun_out/r02_ncu_list_playout.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bk_forward_tc|bk_step_kernel" -c 4 -o gpurun_out/prof_playout python tools/run_playout_once.py 512 6 > gpurun_out/r02_ncu_full_playout.log 2>&1
ls -la gpurun_out | tail -n 25

#!/bin/bash
# round 2, third GPU batch: whole GPU suite (incl. --simulate search, 65,536-board simulate, reference callers), smoke, bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02c_pytest.log 2>&1; echo "pytest rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02c_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --steps 20 --warmup 5 > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02c_bench_reference.json 2> gpurun_out/r02c_bench_reference.err; echo "bench ref rc=$?"
tail -n 4 gpurun_out/r02c_pytest.log; cat gpurun_out/r02c_smoke.log | tail -n 2; tail -n 5 gpurun_out/r02c_bench.err; head -c 600 gpurun_out/r02c_bench.json

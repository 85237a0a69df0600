#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mcts.py -q > gpurun_out/r02j_tests.log 2>&1; echo "tests rc=$?"; tail -n 5 gpurun_out/r02j_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-train > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err; echo "bench rc=$?"; tail -n 3 gpurun_out/r02j_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02j_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'per_call',d['e2e']['per_call']['value'],'frac',d['roofline']['frac'],'kernel_ms',d['roofline']['kernel_ms'],'one_launch',d['roofline']['one_launch_step'])
PY
timeout 600 python bench.py --steps 100 --warmup 10 --no-train --no-playouts --no-cpu > gpurun_out/r02j_bench100.json 2>> gpurun_out/r02j_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02j_bench100.json').read().strip().splitlines()[-1])
print('100 steps: value',d['value'],'e2e',d['e2e']['value'],'per_call',d['e2e']['per_call']['value'],'frac',d['roofline']['frac'], d['clocks'])
PY

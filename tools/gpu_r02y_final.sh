#!/bin/bash
# final check of the round-2 tree: whole GPU suite, smoke, the bench line as the driver runs it
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r02y_pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/r02y_pytest_gpu.log; tail -n 3 gpurun_out/r02y_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02y_smoke.log 2>&1; echo "smoke exit $?"; tail -n 1 gpurun_out/r02y_smoke.log | cut -c1-200
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02y_bench.json 2> gpurun_out/r02y_bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02y_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'clocks',d['clocks']['sm_mhz'],'selfplay',d['selfplay']['games_per_s'])
print('reinforce 3xtf32',json.dumps(d['reinforce']['3xtf32']),'iteration',d['reinforce']['iteration']['seconds'])
PY

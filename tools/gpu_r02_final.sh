#!/bin/bash
# final check of the committed tree: whole GPU suite, smoke, the bench line (both arms)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r02_final_pytest.log 2>&1; echo "pytest exit $?"; tail -n 3 gpurun_out/r02_final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final_smoke.log 2>&1; echo "smoke exit $?"; tail -n 1 gpurun_out/r02_final_smoke.log | cut -c1-200
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err; echo "bench exit $?"; tail -n 2 gpurun_out/r02_final_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_final_bench_reference.json 2>> gpurun_out/r02_final_bench.err; echo "ref exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_final_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'selfplay',d['selfplay']['games_per_s'],d['selfplay']['steady_stream'])
r=json.loads(open('gpurun_out/r02_final_bench_reference.json').read().strip().splitlines()[-1])
print('reference',r['value'],r['cpu_baseline']['kind'],r['ms_per_step'])
PY
timeout 300 python tools/bench_mcts.py > gpurun_out/r02_final_mcts.jsonl 2>&1; cut -c1-200 gpurun_out/r02_final_mcts.jsonl

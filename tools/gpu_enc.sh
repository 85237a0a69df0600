#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "encode or playout or smoke or mirror or mcts" > gpurun_out/t_enc.log 2>&1; echo "exit $?" >> gpurun_out/t_enc.log; tail -4 gpurun_out/t_enc.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for b in 4096 16384; do timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --batch $b 2>/dev/null; done > gpurun_out/sweep_enc.jsonl
python - <<'PY'
import json
for l in open('gpurun_out/sweep_enc.jsonl'):
    d=json.loads(l); print(d['config']['batch_per_gpu'], round(d['value']), round(d['e2e']['value']), d['roofline']['encoder'], d['selfplay']['games_per_s'], d['simulate']['playouts_per_s'])
PY

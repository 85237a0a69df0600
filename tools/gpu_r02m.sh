#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-train --no-cpu > gpurun_out/r02m_bench_n$N.json 2> gpurun_out/r02m_bench_n$N.err; echo "bench N=$N exit $?"; tail -n 2 gpurun_out/r02m_bench_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02m_bench_n$N.json').read().strip().splitlines()[-1])
print('N=$N value',d['value'],'selfplay',json.dumps(d['selfplay'])[:1200])
PY

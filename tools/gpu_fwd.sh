#!/bin/bash
# GPU session for the conv kernel: forward parity tests, protocol stress with cold L2, per-pass clock profile, batch sweep.
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -${TAILN:-6} gpurun_out/$name.log; }
run t_fwd python -m pytest tests/test_gpu_parity.py -x -q -s -m gpu -k "forward"
TAILN=8 run stress python tools/stress_forward.py --iters 300 --batches 1,16,256,745,4096
TAILN=14 run prof_4096 python tools/prof_forward.py --batch 4096
sed -n 1,14p gpurun_out/prof_4096.log
TAILN=12 run prof_16 python tools/prof_forward.py --batch 16
for b in 1 16 256 1024 4096 16384; do timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-playouts --batch $b 2>/dev/null; done > gpurun_out/sweep.jsonl
python - <<'PY'
import json
for l in open('gpurun_out/sweep.jsonl'):
    d=json.loads(l); print(d['config']['batch_per_gpu'], round(d['value']), round(d['e2e']['value']), round(d['roofline']['frac'],3), round(d['roofline']['kernel_ms'],4))
PY

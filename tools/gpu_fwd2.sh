#!/bin/bash
# GPU session: weight-pipeline stage size experiment (K steps per stage), then forward tests on the default build.
mkdir -p gpurun_out
for k in 1 2 4; do
  BK_NVCC_DEFS="-DBK_KSTEPS_PER_STAGE=$k" python -m bokego_b200.build --force > /dev/null 2>&1
  echo "=== KPS=$k"
  timeout 300 python tools/prof_forward.py --batch 4096 > gpurun_out/prof_kps$k.log 2>&1; sed -n '1,12p' gpurun_out/prof_kps$k.log; tail -1 gpurun_out/prof_kps$k.log
  timeout 300 python tools/prof_forward.py --batch 16 2>&1 | tail -9
  for b in 16 4096 16384; do timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --batch $b 2>/dev/null; done > gpurun_out/sweep_kps$k.jsonl
  python - <<PY
import json
for l in open('gpurun_out/sweep_kps$k.jsonl'):
    d=json.loads(l); print(d['config']['batch_per_gpu'], round(d['value']), round(d['e2e']['value']), round(d['roofline']['frac'],3), round(d['roofline']['kernel_ms'],4))
PY
done
python -m bokego_b200.build --force > /dev/null 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "forward" 2>&1 | tail -3

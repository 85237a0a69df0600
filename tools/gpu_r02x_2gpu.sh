#!/bin/bash
# round 2, 2 GPUs: data-parallel reinforce() with the rebuilt 3x3 kernel against the 1-GPU run; the bench line on 2 ranks
mkdir -p gpurun_out
timeout 300 python tools/train_multi.py 2>&1 | tail -n 1 | cut -c1-200
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/train_multi.py > gpurun_out/r02x_train_2gpu.log 2>&1; echo "train_multi 2 GPUs exit $?"; grep -h '"world"' gpurun_out/r02x_train_2gpu.log | cut -c1-500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu > gpurun_out/r02x_bench_2gpu.json 2> gpurun_out/r02x_bench_2gpu.err; echo "bench 2 GPUs exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02x_bench_2gpu.json').read().strip().splitlines()[-1])
print('N=2 value',d['value'],'e2e',d['e2e']['value'],'selfplay',d['selfplay']['games_per_s'],'reinforce',json.dumps(d.get('reinforce',{}).get('3xtf32')))
PY

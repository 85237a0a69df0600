#!/bin/bash
# multi-GPU session (gpurun --gpus N): the bench under torchrun exactly as the driver launches it, both arms
N=${1:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "bench N=$N exit $?"; tail -n 3 gpurun_out/r02_bench_n$N.err; head -c 300 gpurun_out/r02_bench_n$N.json; echo
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/r02_bench_ref_n$N.json 2>> gpurun_out/r02_bench_n$N.err; echo "ref N=$N exit $?"; head -c 300 gpurun_out/r02_bench_ref_n$N.json; echo

#!/bin/bash
# multi-GPU session (gpurun --gpus N): the bench under torchrun exactly as the driver launches it, plus the library-GPU leg
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N exit $?"; tail -3 gpurun_out/bench_n$N.err; cat gpurun_out/bench_n$N.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.json 2>> gpurun_out/bench_n$N.err; echo "ref N=$N exit $?"; cat gpurun_out/bench_ref_n$N.json | cut -c1-300
timeout 600 python tools/library_gpu.py --batch 4096 > gpurun_out/library_gpu.json 2> gpurun_out/library_gpu.err; echo "lib exit $?"; cat gpurun_out/library_gpu.json

#!/bin/bash
# One GPU session: every test group in its own process (a trapped kernel must not hide the other results).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name" ; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -5 gpurun_out/$name.log; }
run t_encode python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "encode"
run t_misc python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "exp_stream or score"
run t_step python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "playout"
run t_fwd_simt python -m pytest tests/test_gpu_parity.py -x -q -s -m gpu -k "forward_golden and simt"
run t_fwd_tc python -m pytest tests/test_gpu_parity.py -x -q -s -m gpu -k "forward_golden and tcgen05"
for ps in 0 1 2; do run diag_p${ps}_s0 python tools/diag_forward.py --pass $ps --swap 0; done
run diag_p0_s1 python tools/diag_forward.py --pass 0 --swap 1
run t_fwd_rest python -m pytest tests/test_gpu_parity.py -x -q -s -m gpu -k "forward_tc or forward_full"
run smoke python -c "import __graft_entry__ as g; g.smoke()"

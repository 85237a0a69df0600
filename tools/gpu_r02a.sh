#!/bin/bash
# round 2, first GPU batch: the reference's own callers over the mirror, the precision distribution, then the whole GPU suite
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r02a_gpu.txt 2>&1
python -m pytest tests/test_gpu_reference_callers.py -x -q -s > gpurun_out/r02a_refcallers.log 2>&1; echo "refcallers rc=$?"
python -m pytest tests/test_gpu_precision.py -x -q -s > gpurun_out/r02a_precision.log 2>&1; echo "precision rc=$?"
python -m pytest tests -m gpu -q --deselect tests/test_gpu_reference_callers.py --deselect tests/test_gpu_precision.py > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02a_refcallers.log gpurun_out/r02a_precision.log gpurun_out/r02a_pytest.log

#!/bin/bash
# A/B of measurement builds of the library (bokego_b200/libbokego_b200_<name>.so, see BK_NVCC_DEFS in bokego_b200/build.py) on one box:
# bash tools/gpu_ab_variants.sh <tag> <name> [<name> ...]   ("-" = the default build); parity tests run on every variant first
tag=$1; shift
mkdir -p gpurun_out
D=$PWD/bokego_b200
for v in "$@"; do
  so=$D/libbokego_b200$([ "$v" = "-" ] || echo _$v).so
  BOKEGO_B200_SO=$so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_playout.py -q -m gpu -x 2>&1 | tail -n 1 | sed "s/^/$v pytest: /"
done > gpurun_out/${tag}_pytest.txt 2>&1
cat gpurun_out/${tag}_pytest.txt
for rep in 1 2; do
  for v in "$@"; do
    so=$D/libbokego_b200$([ "$v" = "-" ] || echo _$v).so
    BOKEGO_B200_SO=$so timeout 300 python tools/time_forward_sizes.py 740 4096 2>&1 | cut -c1-200 | sed "s/^/$v /"
    BOKEGO_B200_SO=$so timeout 300 python tools/bench_playout.py 512 4096 2>&1 | cut -c1-160 | sed "s/^/$v /"
  done
done > gpurun_out/${tag}_ab.txt 2>&1
cat gpurun_out/${tag}_ab.txt

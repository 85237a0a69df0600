#!/bin/bash
# Build the measurement libraries first, in the build container (they travel to the GPU box with the snapshot):
#   for d in 1 2 4 8 16 31; do BK_NVCC_DEFS="-DBK_TC_DIAG=$d" python -m bokego_b200.build --force && cp bokego_b200/libbokego_b200.so tools/probes/libdiag_$d.so; done
#   python -m bokego_b200.build --force        # back to the product build
# timing of the tcgen05 training GEMMs with one pipeline stage switched off at a time (measurement builds in tools/probes, see BK_TC_DIAG
# in csrc/bk_train_tc.cu; results are wrong by construction, only the times matter)
for d in 0 1 2 4 8 16 31; do
  if [ $d = 0 ]; then so=""; else so=$PWD/tools/probes/libdiag_$d.so; fi
  echo "== BK_TC_DIAG=$d"
  BOKEGO_B200_SO=$so timeout 120 python tools/check_train_tc.py 576 5 2>&1 | grep -E "forward|backward"
done

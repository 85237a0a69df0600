#!/bin/bash
# stress run of the two stress-clean variants of the tile-by-tile epilogue experiment (apply tile_tail_epilogue_overlap.patch first)
for k in 1 2; do
  BK_NVCC_DEFS="-DBK_EXP=$k" python -m bokego_b200.build --force > /dev/null 2>&1
  echo "=== BK_EXP=$k"
  for rep in 1 2 3 4 5 6; do timeout 300 python tools/stress_forward.py --iters 8000 --batches 741 2>&1 | tail -3; done
done

#!/bin/bash
mkdir -p gpurun_out
echo "== 4096 N=256 x2"; python tools/prof_forward.py --batch 4096 --flags 2048 2>&1 | sed -n '1,12p;$p'

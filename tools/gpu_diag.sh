#!/bin/bash
mkdir -p gpurun_out
echo "== 16"; python tools/prof_forward.py --batch 16 2>&1 | sed -n '1,105p'

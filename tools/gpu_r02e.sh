#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/prof_playout.py 512 1 > gpurun_out/r02e_prof_playout.txt 2>&1; echo "prof rc=$?"
timeout 300 python tools/prof_playout.py 512 0 >> gpurun_out/r02e_prof_playout.txt 2>&1; echo "prof rc=$?"
cat gpurun_out/r02e_prof_playout.txt

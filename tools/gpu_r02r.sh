#!/bin/bash
# round 2: ncu launch list and full capture of the persistent 3x3 training kernel (576 positions, as profiles/r01h_*)
mkdir -p gpurun_out
timeout 300 python tools/prof_train.py > gpurun_out/r02r_prof_train_plain.log 2>&1; echo "plain exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_train.csv python tools/prof_train.py > gpurun_out/r02r_ncu_list_train.log 2>&1; echo "list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bk_train_conv3_tc" -s 8 -c 2 -o gpurun_out/prof_train_conv3 python tools/prof_train.py > gpurun_out/r02r_ncu_full_train.log 2>&1; echo "full exit $?"
tail -3 gpurun_out/r02r_ncu_full_train.log

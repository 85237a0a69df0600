#!/bin/bash
# the conv kernel with early read-out (tile-by-tile pass tail): whole GPU suite, stress, the hand-over measurement builds, bench
mkdir -p gpurun_out
D=$PWD/bokego_b200
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02s_pytest.log 2>&1; echo "pytest exit $?"; tail -n 3 gpurun_out/r02s_pytest.log
for rep in 1 2 3 4 5 6; do timeout 400 python tools/stress_forward.py --iters 8000 --batches 741 2>&1 | tail -n 3; done > gpurun_out/r02s_stress_forward.txt 2>&1
timeout 600 python tools/stress_forward.py --iters 4000 --batches 745,4096,37,1,16 >> gpurun_out/r02s_stress_forward.txt 2>&1
cat gpurun_out/r02s_stress_forward.txt
timeout 600 python tools/stress_playout.py --iters 400 > gpurun_out/r02s_stress_playout.txt 2>&1; tail -n 2 gpurun_out/r02s_stress_playout.txt
for v in h1 h2 h3; do
  for rep in 1 2; do BOKEGO_B200_SO=$D/libbokego_b200_$v.so timeout 400 python tools/stress_forward.py --iters 8000 --batches 741 --max-bad 100000 --quiet 2>&1 | tail -n 1 | sed "s/^/BK_HANDOVER=${v#h} /"; done
done > gpurun_out/r02s_handover.txt 2>&1
cat gpurun_out/r02s_handover.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02s_bench.json 2> gpurun_out/r02s_bench.err; echo "bench exit $?"; tail -n 2 gpurun_out/r02s_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02s_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'selfplay',d['selfplay']['games_per_s'],d['selfplay'].get('us_per_move'))
PY
timeout 300 python tools/prof_forward.py > gpurun_out/r02s_pass_clocks.txt 2>&1; tail -n 30 gpurun_out/r02s_pass_clocks.txt

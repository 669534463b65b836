#!/bin/bash
# same-box A/B of one env knob:  exp_ab.sh <tag> "<ENV=val for A>" "<ENV=val for B>"
O=gpurun_out/${1:-exp_ab}
mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "convgru or conv_fwd_dgrad_wgrad or generator or other_config" > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest.log
i=0
for cfg in "$2" "$3" "$2" "$3"; do
  i=$((i+1))
  env $cfg timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 2 > $O/bench_$i.json 2> $O/bench_$i.err; echo "bench [$cfg] rc=$?"; python -c "
import json;d=json.load(open('$O/bench_$i.json'));r=d['roofline'];print(d['value'], d['ms_per_step'], 'fwd TF/s', r['achieved'], 'wgrad', r['wgrad']['achieved'], r['breakdown_ms_per_step'])"; tail -2 $O/bench_$i.err
done

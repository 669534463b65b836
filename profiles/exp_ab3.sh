#!/bin/bash
# same-box comparison of several env settings: exp_ab3.sh <tag> "<cfg1>" "<cfg2>" ...
O=gpurun_out/$1
shift
mkdir -p $O
i=0
for rep in 1 2; do
for cfg in "$@"; do
  i=$((i+1))
  env $cfg timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 2 > $O/bench_$i.json 2> $O/bench_$i.err; echo "bench [$cfg] rc=$?"; python -c "
import json;d=json.load(open('$O/bench_$i.json'));r=d['roofline'];print(d['value'], d['ms_per_step'], 'fwd TF/s', r['achieved'], 'wgrad', r['wgrad']['achieved'], r['breakdown_ms_per_step'])"; tail -2 $O/bench_$i.err
done
done

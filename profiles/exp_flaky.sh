#!/bin/bash
O=gpurun_out/${1:-exp_flaky}
mkdir -p $O
for i in 1 2 3 4; do
  timeout 600 python -m pytest tests -q -m gpu > $O/pytest_$i.log 2>&1; echo "run $i rc=$?"; grep -E "passed|failed|^FAILED|AssertionError: \(" $O/pytest_$i.log | head -6
done
timeout 600 python bench.py --no-cpu-baseline --prof-dump $O/prof.tsv > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; cut -c1-330 $O/bench.json; tail -3 $O/bench.err
timeout 300 python profiles/conv_microbench.py --reps 5 --only s9_cell1_h_ur; 

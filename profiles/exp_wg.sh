#!/bin/bash
O=gpurun_out/${1:-exp_wg}
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "conv_fwd_dgrad_wgrad" > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest.log
for cfg in "X=1" "DVD_TC_SHORTK_BN128=1"; do
  echo "== $cfg" | tee -a $O/microbench.txt
  env $cfg timeout 200 python profiles/conv_microbench.py --reps 5 >> $O/microbench.txt 2>&1
done
cat $O/microbench.txt
timeout 600 python bench.py --no-cpu-baseline --prof-dump $O/prof.tsv > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; cut -c1-330 $O/bench.json; tail -3 $O/bench.err

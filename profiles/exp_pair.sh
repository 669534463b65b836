#!/bin/bash
# experiment: CTA-pair (cta_group::2) kernels: correctness first, then TF/s per shape
O=gpurun_out/${1:-exp_pair}
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "conv_fwd_dgrad_wgrad or conv_fused" > $O/pytest_conv.log 2>&1; echo "pytest rc=$?"; tail -15 $O/pytest_conv.log
for cfg in "DVD_TC_PAIR=1" "DVD_TC_PAIR=0"; do
  echo "== $cfg" | tee -a $O/microbench.txt
  env $cfg timeout 200 python profiles/conv_microbench.py --reps 5 --err >> $O/microbench.txt 2>&1; echo "rc=$?" >> $O/microbench.txt
done
cat $O/microbench.txt
timeout 600 python bench.py --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; cat $O/bench.json; tail -3 $O/bench.err

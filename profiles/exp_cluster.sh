#!/bin/bash
# experiment: cluster shapes for the multicast conv kernels (correctness first, then TF/s per shape)
O=gpurun_out/${1:-exp_cluster}
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "conv_fwd_dgrad_wgrad or conv_fused" > $O/pytest_conv.log 2>&1; echo "pytest rc=$?"; tail -15 $O/pytest_conv.log
for c in 1,1 2,1 1,2 2,2 4,1 4,2 2,4; do
  echo "== DVD_TC_CLUSTER=$c" | tee -a $O/microbench.txt
  DVD_TC_CLUSTER=$c timeout 200 python profiles/conv_microbench.py --reps 5 --err >> $O/microbench.txt 2>&1; echo "rc=$?" >> $O/microbench.txt
done
cat $O/microbench.txt

#!/bin/bash
O=gpurun_out/${1:-exp_dbg}
mkdir -p $O
for i in 1 2; do
for cfg in "X=1" "DVD_GRU_FUSED=0" "DVD_GRU_SHARE_PLANES=0"; do
  env $cfg timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "test_generator or convgru" > $O/pytest_${cfg}_$i.log 2>&1; echo "$cfg run $i rc=$?"; grep -E "passed|failed|AssertionError: \(" $O/pytest_${cfg}_$i.log | head -5
done
done

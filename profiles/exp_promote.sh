#!/bin/bash
O=gpurun_out/${1:-exp_promote}
mkdir -p $O
for cfg in "DVD_TC_PROMOTE=1" "DVD_TC_PROMOTE=0"; do
  echo "== $cfg" | tee -a $O/g_error.txt
  env $cfg timeout 400 python profiles/g_error_by_stage.py 0 >> $O/g_error.txt 2>&1
done
cat $O/g_error.txt
DVD_TC_PROMOTE=0 timeout 600 python -m pytest tests -q -m gpu > $O/pytest_nopromote.log 2>&1; echo "pytest rc=$?"; tail -40 $O/pytest_nopromote.log

#!/bin/bash
O=gpurun_out/${1:-exp_epi}
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "convgru or conv_fwd_dgrad_wgrad" > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest.log
for cfg in "X=1" "DVD_TC_EPI_PREFETCH=0"; do
  env $cfg timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 1 --prof-dump $O/prof_$cfg.tsv > $O/bench_$cfg.json 2> $O/bench_$cfg.err; echo "bench $cfg rc=$?"; cut -c1-230 $O/bench_$cfg.json; tail -2 $O/bench_$cfg.err
done

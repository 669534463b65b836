#!/bin/bash
O=gpurun_out/${1:-exp_gru}
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "convgru or conv_fwd_dgrad_wgrad or generator" > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $O/pytest.log
timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 1 --prof-dump $O/prof.tsv > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; cut -c1-330 $O/bench.json; tail -3 $O/bench.err

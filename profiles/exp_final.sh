#!/bin/bash
O=gpurun_out/${1:-exp_final}
mkdir -p $O
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest.log
timeout 600 python bench.py --prof-dump $O/prof.tsv > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; cat $O/bench.json; tail -3 $O/bench.err
timeout 300 python profiles/conv_microbench.py --reps 5 --err > $O/microbench.txt 2>&1; cat $O/microbench.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_tma_wgrad -c 2 -f -o $O/conv_tma_wgrad_s9 \
    python profiles/conv_microbench.py --reps 1 --only s9_cell1_h_ur > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"

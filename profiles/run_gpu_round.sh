#!/bin/bash
# One gpurun call: GPU parity tests, the bench line, the conv micro-benchmark, one `ncu --set full` capture of the
# dominant kernels and the ncu launch list of one bench step.  Outputs land in gpurun_out/ (scratch); the summaries
# are copied into profiles/ by hand.   usage: gpurun --timeout 1800 -- bash profiles/run_gpu_round.sh [tag]
TAG=${1:-r1}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > $O/smi.txt 2>&1
if [ -z "$SKIP_TESTS" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
  tail -5 $O/pytest_gpu.log
fi
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; cat $O/bench.json
timeout 300 python profiles/conv_microbench.py --reps 3 > $O/microbench.txt 2>&1; cat $O/microbench.txt
if [ -z "$SKIP_NCU" ]; then
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_tma_ -c 4 -f -o $O/conv_tma_s9 \
      python profiles/conv_microbench.py --reps 1 --only s9_cell1_h_ur > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv \
      python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > $O/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
  python profiles/summarize_launches.py $O/launches.csv > $O/launches_summary.md 2>&1; head -40 $O/launches_summary.md
  gzip -f $O/launches.csv
fi

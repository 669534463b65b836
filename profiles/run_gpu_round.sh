#!/bin/bash
# One gpurun call: GPU parity tests, the bench line, the per-shape event profile, the conv micro-benchmark, one
# `ncu --set full` capture of the dominant kernels and an ncu launch list.  Outputs land in gpurun_out/<tag>/ (scratch);
# the summaries are copied into profiles/ afterwards.
#   usage: gpurun --timeout 2400 -- bash profiles/run_gpu_round.sh [tag]
# The ncu launch list costs ~0.18 s per launch on this box; a full config-2 step is ~13 k launches (40 min), so the list
# is taken on the same bench command with --frames 8 (same kernels, same shapes per step, 1/6 of the time steps) and the
# bench's own per-shape CUDA-event table is written for BOTH the reduced and the full command for comparison.
TAG=${1:-r1}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > $O/smi.txt 2>&1
if [ -z "$SKIP_TESTS" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
  tail -5 $O/pytest_gpu.log
fi
timeout 600 python bench.py --prof-dump $O/prof_full.tsv > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; cat $O/bench.json
timeout 300 python profiles/conv_microbench.py --reps 5 --err > $O/microbench.txt 2>&1; cat $O/microbench.txt
if [ -z "$SKIP_NCU" ]; then
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_tma_ -c 4 -f -o $O/conv_tma_s9 \
      python profiles/conv_microbench.py --reps 1 --only s9_cell1_h_ur > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
  timeout 300 python bench.py --frames 8 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --prof-dump $O/prof_f8.tsv \
      > $O/bench_f8.json 2> $O/bench_f8.err; echo "bench f8 rc=$?"
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_f8.csv \
      python bench.py --frames 8 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > $O/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
  python profiles/summarize_launches.py $O/launches_f8.csv > $O/launches_f8_summary.md 2>&1; head -30 $O/launches_f8_summary.md
  gzip -f $O/launches_f8.csv
fi

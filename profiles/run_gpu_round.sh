#!/bin/bash
# One gpurun call of round 2.  Stages (space-separated in $STAGES, default: all of "tests bench configs oneacc"):
#   tests    pytest -m gpu
#   bench    the default bench line (config 2) with the per-shape event profile
#   configs  bench lines of BASELINE.json configs 3, 4, 5 at their per-GPU batch on ONE GPU
#   oneacc   A/B of the double-buffered single-accumulator tiles: micro-benchmark, Generator error by stage, bench line
#   ncu      `ncu --set full` of the dominant kernels (GEMMs, operand split, CBN / BN statistics, attention) + the launch
#            list of a --frames 8 step
#   refgpu   informational: the unmodified reference under eager PyTorch / cuDNN on this GPU (profiles/ref_eager_b200.py)
#   sanitize compute-sanitizer memcheck + racecheck over smoke() and a small attention case
# Outputs land in gpurun_out/<tag>/ (scratch); summaries are copied into profiles/ afterwards.
#   usage: gpurun --timeout 2400 -- bash profiles/run_gpu_round.sh [tag]
TAG=${1:-r2}
STAGES=${STAGES:-tests bench configs oneacc}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > $O/smi.txt 2>&1
nproc > $O/nproc.txt
has() { [[ " $STAGES " == *" $1 "* ]]; }
if has tests; then
  timeout ${TEST_TIMEOUT:-1500} python -m pytest tests -m gpu ${PYTEST_ARGS:--x} -q --durations=15 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
  tail -30 $O/pytest_gpu.log
fi
if has bench; then
  timeout 600 python bench.py --prof-dump $O/prof_full.tsv > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; cat $O/bench.json; tail -3 $O/bench.err
fi
if has configs; then
  for c in 3 4 5; do
    timeout 600 python bench.py --config $c --steps 2 --warmup 3 --no-cpu-baseline --prof-dump $O/prof_c$c.tsv \
        > $O/bench_c$c.json 2> $O/bench_c$c.err; echo "bench config $c rc=$?"; cat $O/bench_c$c.json; tail -3 $O/bench_c$c.err
  done
fi
if has oneacc; then
  for o in "oneacc=0" "oneacc=1"; do
    echo "== $o" | tee -a $O/microbench_oneacc.txt
    DVD_OPTIONS=$o timeout 200 python profiles/conv_microbench.py --reps 5 --err --acc >> $O/microbench_oneacc.txt 2>&1
  done
  cat $O/microbench_oneacc.txt
  timeout 400 python profiles/g_error_by_stage.py 0 oneacc=0 oneacc=1 > $O/g_error_oneacc.txt 2>&1; cat $O/g_error_oneacc.txt
  DVD_OPTIONS=oneacc=1 timeout 400 python bench.py --no-cpu-baseline --prof-dump $O/prof_oneacc.tsv > $O/bench_oneacc.json 2> $O/bench_oneacc.err
  echo "bench oneacc rc=$?"; cat $O/bench_oneacc.json
fi
if has refgpu; then
  timeout 900 python profiles/ref_eager_b200.py --batch ${REF_BATCH:-32} > $O/ref_eager_b200.jsonl 2> $O/ref_eager_b200.err
  echo "ref eager rc=$?"; cat $O/ref_eager_b200.jsonl; tail -3 $O/ref_eager_b200.err
fi
if has sanitize; then
  for tool in memcheck racecheck; do
    timeout 900 compute-sanitizer --tool $tool --print-limit 30 python -c "import __graft_entry__ as g; g.smoke()" \
        > $O/sanitizer_${tool}_smoke.log 2>&1; echo "$tool smoke rc=$?"; tail -4 $O/sanitizer_${tool}_smoke.log
    timeout 900 compute-sanitizer --tool $tool --print-limit 30 python -m pytest tests/test_gpu_round2.py -q -x \
        -k "flash_attention_matches and (shape4 or shape2)" > $O/sanitizer_${tool}_attn.log 2>&1; echo "$tool attn rc=$?"
    tail -4 $O/sanitizer_${tool}_attn.log
  done
fi
if has ncu; then
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_tma_ -c 4 -f -o $O/conv_tma_s9 \
      python profiles/conv_microbench.py --reps 1 --only s9_cell1_h_ur > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
  # memory-bound kernels at full size (config 2, B = 64, 48 frames): only the matching launches are profiled, the rest
  # of the step runs untouched.  -s skips the small early launches of the step.
  timeout 900 ncu --set full --clock-control none --import-source on \
      -k regex:"prep_planes|cbn_apply_vec|bn_partial_vec|cbn_bwd_dx_vec|cbn_bwd_plane_vec|gru_bwd1_planes|channel_sum|adam_kernel" \
      -s 300 -c 32 -f -o $O/membound python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline \
      > $O/ncu_membound.log 2>&1; echo "ncu membound rc=$?"
  # the attention kernels at N = 4096 (config 4: Ds on 256x256 frames)
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_tc_kernel" -c 8 -f -o $O/attn \
      python bench.py --config 4 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > $O/ncu_attn.log 2>&1; echo "ncu attn rc=$?"
  # gpurun brings back at most 64 MiB: keep the raw-page CSV exports, drop the big reports (sources included)
  for r in membound attn; do
    ncu -i $O/$r.ncu-rep --page raw --csv > $O/${r}_raw.csv 2>/dev/null && gzip -f $O/${r}_raw.csv && rm -f $O/$r.ncu-rep
  done
  ls -la $O | head -40
  if [ -n "$NCU_SKIP_LAUNCHES" ]; then exit 0; fi
  timeout 300 python bench.py --frames 8 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --prof-dump $O/prof_f8.tsv \
      > $O/bench_f8.json 2> $O/bench_f8.err; echo "bench f8 rc=$?"
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_f8.csv \
      python bench.py --frames 8 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > $O/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
  python profiles/summarize_launches.py $O/launches_f8.csv > $O/launches_f8_summary.md 2>&1; head -30 $O/launches_f8_summary.md
  gzip -f $O/launches_f8.csv
fi

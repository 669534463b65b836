#!/bin/bash
# Final single-GPU validation of round 2: the whole GPU suite, smoke(), the default bench line (with the CPU arm) and
# the per-shape event profiles / bench lines of configs 2, 3, 5.
O=gpurun_out/${1:-fin_r2}; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --durations=10 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log; tail -15 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 600 python bench.py --prof-dump $O/prof_c2.tsv > $O/bench_config2.json 2> $O/bench_config2.err; echo "bench rc=$?"; cat $O/bench_config2.json
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err; echo "ref arm rc=$?"; cat $O/bench_reference_arm.json
for c in 3 5 4; do
  timeout 600 python bench.py --config $c --steps 2 --warmup 3 --no-cpu-baseline --prof-dump $O/prof_c$c.tsv > $O/bench_config${c}_1gpu.json 2> $O/bench_config${c}_1gpu.err; echo "config $c rc=$?"
  python -c "
import json;d=json.load(open('$O/bench_config${c}_1gpu.json'));print('   ',round(d['value'],2),'clips/s e2e',round(d['e2e']['value'],2),round(d['ms_per_step'],1),'ms', d['clocks']['sm_mhz'],'MHz', round(d['peak_mem_gib'],1),'GiB', d['config']['gru_bptt_state'][:80])"
done

# A/B of the ConvGRU chain count (option gru_streams) at several per-GPU batch sizes of config 2
O=gpurun_out/${1:-r3b}; mkdir -p $O
for b in ${BATCHES:-8 16 32}; do
for n in ${CHAINS:-1 2 4}; do
  DVD_OPTIONS=gru_streams=$n timeout 300 python bench.py --batch $b --no-cpu-baseline --no-e2e --steps 3 --warmup 3 > $O/bench_b${b}_s$n.json 2> $O/bench_b${b}_s$n.err; echo "batch=$b streams=$n rc=$?"
  python -c "
import json;d=json.load(open('$O/bench_b${b}_s$n.json'));print(round(d['value'],2),round(d['ms_per_step'],1),d['roofline']['breakdown_ms_per_step'],d['clocks']['sm_mhz'])"
done; done

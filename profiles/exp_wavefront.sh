# A/B of the ConvGRU layer wavefront (ops.GRUStackFn) against batch chains (option gru_streams) and the plain loop
O=gpurun_out/${1:-r3c}; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q -k "wavefront or helper_streams or convgru or generator or train_step or two_steps or two_devices" > $O/pytest_wave.log 2>&1; echo "pytest rc=$?"; tail -6 $O/pytest_wave.log
run() {  # name, env..., -- bench args
  local name=$1; shift
  env "$@" timeout 400 python bench.py --no-cpu-baseline --no-e2e --steps 3 --warmup 3 $BARGS > $O/bench_$name.json 2> $O/bench_$name.err; echo "$name rc=$?"
  python -c "
import json;d=json.load(open('$O/bench_$name.json'));print('   ',round(d['value'],2),'clips/s',round(d['ms_per_step'],1),'ms', d['clocks']['sm_mhz'],'MHz', round(d['peak_mem_gib'],1),'GiB', d['peak_mem_detail_gib'])"
}
for cfg in ${CFGS:-"2:" "5:" "4:"}; do
  BARGS="--config ${cfg%%:*} ${cfg#*:}"; tag=$(echo "c${cfg%%:*}${cfg#*:}" | tr -d ' -')
  echo "== $BARGS"
  run ${tag}_chains2 DVD_GRU_WAVEFRONT=enabled=0 DVD_OPTIONS=gru_streams=2
  run ${tag}_wave8a DVD_GRU_WAVEFRONT=enabled=1,chunk=8
  run ${tag}_wave8b DVD_GRU_WAVEFRONT=enabled=1,chunk=8
  run ${tag}_wave8_32k DVD_GRU_WAVEFRONT=enabled=1,chunk=8,max_rows=32768
  run ${tag}_wave4_32k DVD_GRU_WAVEFRONT=enabled=1,chunk=4,max_rows=32768
done

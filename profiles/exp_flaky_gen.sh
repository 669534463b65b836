# which mode makes the end-to-end Generator gradient criterion fail after the ConvGRU tests ran in the same process?
O=gpurun_out/${1:-r3f}; mkdir -p $O
SEL="wavefront or helper_streams or convgru or generator or train_step or two_steps"
for mode in "enabled=1:gru_streams=2" "enabled=1:gru_streams=2" "enabled=0:gru_streams=2" "enabled=1:gru_streams=1" "enabled=0:gru_streams=1"; do
  echo "== wavefront ${mode%%:*} ${mode#*:}"
  DVD_GRU_WAVEFRONT=${mode%%:*} DVD_OPTIONS=${mode#*:} timeout 300 python -m pytest tests/test_gpu_parity.py -q -s -k "$SEL" 2>&1 | grep -E "rel-L2|passed|failed|Error"
done 2>&1 | tee $O/flaky_gen.log
timeout 900 compute-sanitizer --tool initcheck --print-limit 40 python -m pytest tests/test_gpu_parity.py -q -x -k "test_generator" > $O/initcheck_generator.log 2>&1; echo "initcheck rc=$?"
grep -E "Uninitialized|ERROR SUMMARY|at .*\+0x|in /" $O/initcheck_generator.log | grep -v "libtorch\|python3" | head -60

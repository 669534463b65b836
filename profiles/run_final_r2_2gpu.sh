O=gpurun_out/fin_r2b; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_2gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_2gpu.log; tail -6 $O/pytest_2gpu.log
timeout 600 python bench.py --gpus 2 --no-cpu-baseline > $O/bench_config2_2gpu.json 2> $O/bench_config2_2gpu.err; echo "bench2 rc=$?"; cut -c1-400 $O/bench_config2_2gpu.json
timeout 600 python bench.py --gpus 2 --global-batch 64 --no-cpu-baseline > $O/bench_config2_2gpu_strong64.json 2> $O/bench_config2_2gpu_strong64.err; echo "strong rc=$?"; cut -c1-300 $O/bench_config2_2gpu_strong64.json

set -x
python -m pytest tests/test_gpu_dock.py -q --tb=short 2>&1 | grep -v "^$" | cut -c1-400 | tail -120 > gpurun_out/r2c_dock.log
timeout 300 python -m pytest tests/test_gpu_tc.py -x -q --tb=short 2>&1 | cut -c1-300 | tail -40 > gpurun_out/r2c_tc.log
echo "rc=$?" >> gpurun_out/r2c_tc.log
GB_TC_FUSED=0 timeout 200 python bench.py --no-cpu-baseline --no-gpu-reference > gpurun_out/r2c_bench_unfused.json 2> gpurun_out/r2c_bench_unfused.err
timeout 200 python bench.py --no-cpu-baseline --no-gpu-reference > gpurun_out/r2c_bench_fused.json 2> gpurun_out/r2c_bench_fused.err
GB_TC_FUSED_PERSIST=2 timeout 200 python bench.py --no-cpu-baseline --no-gpu-reference > gpurun_out/r2c_bench_fused_p2.json 2> gpurun_out/r2c_bench_fused_p2.err

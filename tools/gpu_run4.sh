set -x
timeout 120 python tools/fused_diag.py 9 > gpurun_out/r2d_diag.log 2>&1
python -m pytest tests/test_gpu_dock.py -q --tb=short 2>&1 | grep -v "^$" | cut -c1-300 | tail -60 > gpurun_out/r2d_dock.log
timeout 300 python -m pytest tests/test_gpu_tc.py -q --tb=line 2>&1 | cut -c1-300 | tail -30 > gpurun_out/r2d_tc.log
timeout 200 python bench.py --no-cpu-baseline --no-gpu-reference > gpurun_out/r2d_bench_fused.json 2> gpurun_out/r2d_bench_fused.err
GB_TC_FUSED_PERSIST=2 timeout 200 python bench.py --no-cpu-baseline --no-gpu-reference > gpurun_out/r2d_bench_fused_p2.json 2> gpurun_out/r2d_bench_fused_p2.err

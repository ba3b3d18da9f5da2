set -x
python -m pytest tests/test_gpu_dock.py tests/test_gpu_min.py tests/test_gpu_vina.py -q --tb=short 2>&1 | grep -v "^$" | cut -c1-300 | tail -50 > gpurun_out/r2g_pytest.log
GB_TC_FUSED_PW=1 timeout 120 python tools/fused_diag.py 17 > gpurun_out/r2g_diag_pw1.log 2>&1
timeout 200 python bench.py --no-cpu-baseline --no-gpu-reference > gpurun_out/r2g_bench_A.json 2> gpurun_out/r2g_bench_A.err
GB_TC_FUSED_PW=1 timeout 200 python bench.py --no-cpu-baseline --no-gpu-reference > gpurun_out/r2g_bench_B.json 2> gpurun_out/r2g_bench_B.err
GB_TC_FUSED_PW=1 GB_TC_FUSED_PERSIST=0 timeout 200 python bench.py --no-cpu-baseline --no-gpu-reference > gpurun_out/r2g_bench_B0.json 2> gpurun_out/r2g_bench_B0.err
GB_TC_FUSED_PW=1 timeout 300 python -m pytest tests/test_gpu_tc.py -q --tb=line 2>&1 | cut -c1-300 | tail -12 > gpurun_out/r2g_tc_pw1.log

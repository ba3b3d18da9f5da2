set -x
GB_TC_CONV_V2=3 timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -q --tb=short -x 2>&1 | tail -8 > gpurun_out/r3g_pytest.log
GB_TC_CONV_V2=3 timeout 200 python bench.py --no-cpu-baseline --no-gpu-reference > gpurun_out/r3g_bench.json 2> gpurun_out/r3g_bench.err
GB_TC_CONV_V2=1 timeout 200 python bench.py --no-cpu-baseline --no-gpu-reference > gpurun_out/r3g_bench1.json 2> gpurun_out/r3g_bench1.err

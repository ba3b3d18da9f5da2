set -x
timeout 400 python -m pytest tests/test_gpu_tc.py tests/test_gpu_grad.py -m gpu -q --tb=short 2>&1 | tail -8 > gpurun_out/r3f_pytest.log
timeout 200 python bench.py --no-cpu-baseline --no-gpu-reference > gpurun_out/r3f_bench.json 2> gpurun_out/r3f_bench.err

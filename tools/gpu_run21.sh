set -x
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q --tb=short -x 2>&1 | tail -8 > gpurun_out/r2u_pytest.log
timeout 200 python bench.py --no-cpu-baseline --no-gpu-reference > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err

set -x
GB_TC_FUSED_V2=1 timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -q --tb=short -x 2>&1 | tail -12 > gpurun_out/r2y_pytest.log
GB_TC_FUSED_V2=1 timeout 200 python bench.py --no-cpu-baseline --no-gpu-reference > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err
GB_TC_FUSED_V2=1 GB_TC_FUSED_TRACE=gpurun_out/r2y_trace.txt timeout 120 python tools/ncu_score.py 2048 > gpurun_out/r2y_t1.log 2>&1

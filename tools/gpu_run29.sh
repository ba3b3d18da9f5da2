set -x
GB_TC_FUSED_V2=2 timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -q --tb=line 2>&1 | tail -8 > gpurun_out/r3c_pytest.log
GB_TC_FUSED_V2=2 timeout 150 python bench.py --no-cpu-baseline --no-gpu-reference > gpurun_out/r3c_bench.json 2> gpurun_out/r3c_bench.err
GB_TC_FUSED_V2=2 GB_TC_FUSED_TRACE=gpurun_out/r3c_trace.txt timeout 120 python tools/ncu_score.py 2048 > gpurun_out/r3c_t1.log 2>&1

set -x
GB_TC_FUSED_V2=2 timeout 200 python -m pytest tests/test_gpu_tc.py -m gpu -q --tb=line -k "fused" 2>&1 | tail -12 > gpurun_out/r3b_pytest.log
GB_TC_FUSED_V2=2 GB_TC_FUSED_TRACE=gpurun_out/r3b_trace.txt timeout 120 python tools/ncu_score.py 2048 > gpurun_out/r3b_t1.log 2>&1

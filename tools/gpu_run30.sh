set -x
GB_TC_FUSED_V2=2 GB_TC_FUSED_TRACE=gpurun_out/r3d_trace.txt timeout 120 python tools/ncu_score.py 2048 > gpurun_out/r3d_t1.log 2>&1
timeout 200 python -m pytest tests/test_gpu_tc.py -m gpu -q --tb=short -k "variants" 2>&1 | tail -8 > gpurun_out/r3d_pytest.log

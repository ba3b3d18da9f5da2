set -x
GB_TC_FUSED_V2=2 GB_TC_FUSED_TRACE=gpurun_out/r3e_trace.txt timeout 120 python tools/ncu_score.py 2048 > gpurun_out/r3e_t1.log 2>&1

set -x
GB_TC_FUSED_TRACE=gpurun_out/r2j_trace_p2.txt timeout 120 python tools/ncu_score.py 2048 > gpurun_out/r2j_t1.log 2>&1
GB_TC_FUSED_PERSIST=1 GB_TC_FUSED_TRACE=gpurun_out/r2j_trace_p1.txt timeout 120 python tools/ncu_score.py 2048 > gpurun_out/r2j_t2.log 2>&1

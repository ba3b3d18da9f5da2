set -x
GB_TC_FUSED_TRACE=gpurun_out/r2v_trace.txt timeout 120 python tools/ncu_score.py 2048 > gpurun_out/r2v_t1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"conv1_pw2_pool" -s 1 -c 1 -o gpurun_out/r2v_fused python tools/ncu_score.py 1024 > gpurun_out/r2v_ncu.log 2>&1

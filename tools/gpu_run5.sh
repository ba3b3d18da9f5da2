set -x
timeout 200 python tools/order_diag.py > gpurun_out/r2e_order_fused.log 2>&1
GB_TC_FUSED=0 timeout 200 python tools/order_diag.py > gpurun_out/r2e_order_unfused.log 2>&1

set -x
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dock_mc -c 1 -o gpurun_out/r2o_mc python tools/ncu_dock.py 4096 6 > gpurun_out/r2o_ncu.log 2>&1

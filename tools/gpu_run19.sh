set -x
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dock_mc -c 1 -o gpurun_out/r2s_mc python tools/ncu_dock.py 2960 6 > gpurun_out/r2s_ncu.log 2>&1
cp gnina_b200/libgnina_b200.so /tmp/new.so; cp tools/_old/lib_lb6.so gnina_b200/libgnina_b200.so
timeout 300 python tools/dock_rows.py 3552 40 > gpurun_out/r2s_lb6.json 2> gpurun_out/r2s_lb6.err
timeout 300 python tools/dock_rows.py 3552 40 >> gpurun_out/r2s_lb6.json 2>> gpurun_out/r2s_lb6.err
cp /tmp/new.so gnina_b200/libgnina_b200.so
timeout 300 python tools/dock_rows.py 3552 40 > gpurun_out/r2s_cur.json 2> gpurun_out/r2s_cur.err

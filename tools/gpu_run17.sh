set -x
timeout 300 python tools/dock_rows.py 4096 40 > gpurun_out/r2q_c.json 2> gpurun_out/r2q_c.err
timeout 300 python tools/dock_rows.py 2960 40 >> gpurun_out/r2q_c.json 2>> gpurun_out/r2q_c.err
timeout 300 python tools/dock_rows.py 2960 40 >> gpurun_out/r2q_c.json 2>> gpurun_out/r2q_c.err
cp gnina_b200/libgnina_b200.so /tmp/new.so; cp tools/_old/libgnina_b200_old.so gnina_b200/libgnina_b200.so
timeout 300 python tools/dock_rows.py 2960 40 > gpurun_out/r2q_old.json 2> gpurun_out/r2q_old.err
timeout 300 python tools/dock_rows.py 2960 40 >> gpurun_out/r2q_old.json 2>> gpurun_out/r2q_old.err
cp /tmp/new.so gnina_b200/libgnina_b200.so

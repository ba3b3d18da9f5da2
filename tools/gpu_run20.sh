set -x
timeout 300 python tools/dock_rows.py 4144 40 > gpurun_out/r2t.json 2> gpurun_out/r2t.err
timeout 300 python tools/dock_rows.py 4144 40 >> gpurun_out/r2t.json 2>> gpurun_out/r2t.err
timeout 300 python tools/dock_rows.py 2960 40 >> gpurun_out/r2t.json 2>> gpurun_out/r2t.err

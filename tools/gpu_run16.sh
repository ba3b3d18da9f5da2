set -x
timeout 300 python tools/dock_rows.py 4096 40 > gpurun_out/r2p_new.json 2> gpurun_out/r2p_new.err
timeout 300 python tools/dock_rows.py 2960 40 >> gpurun_out/r2p_new.json 2>> gpurun_out/r2p_new.err
cp gnina_b200/libgnina_b200.so /tmp/new.so; cp tools/_old/libgnina_b200_old.so gnina_b200/libgnina_b200.so
timeout 300 python tools/dock_rows.py 4096 40 > gpurun_out/r2p_old.json 2> gpurun_out/r2p_old.err
timeout 300 python tools/dock_rows.py 2960 40 >> gpurun_out/r2p_old.json 2>> gpurun_out/r2p_old.err
cp /tmp/new.so gnina_b200/libgnina_b200.so
timeout 600 python -m pytest tests/test_gpu_dock.py tests/test_gpu_vina.py tests/test_docking_pipeline.py -m gpu -q --tb=short 2>&1 | tail -15 > gpurun_out/r2p_pytest.log

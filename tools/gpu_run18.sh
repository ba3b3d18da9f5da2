set -x
timeout 300 python tools/dock_rows.py 4096 40 > gpurun_out/r2r_d.json 2> gpurun_out/r2r_d.err
timeout 300 python tools/dock_rows.py 2960 40 >> gpurun_out/r2r_d.json 2>> gpurun_out/r2r_d.err
timeout 300 python tools/dock_rows.py 2960 40 >> gpurun_out/r2r_d.json 2>> gpurun_out/r2r_d.err
timeout 600 python -m pytest tests/test_gpu_dock.py tests/test_gpu_vina.py tests/test_docking_pipeline.py -m gpu -q --tb=short 2>&1 | tail -15 > gpurun_out/r2r_pytest.log

set -x
python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -v "^$" | cut -c1-300 | tail -30 > gpurun_out/r2x_pytest.log
python bench.py > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2x_ref.json 2> gpurun_out/r2x_ref.err
timeout 1200 python bench.py --rows > gpurun_out/r2x_rows.jsonl 2> gpurun_out/r2x_rows.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2x_launches.csv python bench.py --steps 2 --warmup 3 --poses 2048 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2x_bench_under_ncu.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"conv1_pw2_pool|voxelize_pool|conv3_tc_kernel|pointwise_pool_mma" -s 5 -c 5 -o gpurun_out/r2x_top python tools/ncu_score.py 1024 > gpurun_out/r2x_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dock_mc -c 1 -o gpurun_out/r2x_mc python tools/ncu_dock.py 4144 6 > gpurun_out/r2x_ncu_mc.log 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2x_smoke.log 2>&1

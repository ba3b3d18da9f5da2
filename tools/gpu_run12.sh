set -x
python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -v "^$" | cut -c1-300 | tail -30 > gpurun_out/r2l_pytest.log
python bench.py > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2l_ref.json 2> gpurun_out/r2l_ref.err
timeout 900 python bench.py --rows > gpurun_out/r2l_rows.jsonl 2> gpurun_out/r2l_rows.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2l_launches.csv python bench.py --steps 2 --warmup 3 --poses 2048 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2l_bench_under_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"conv1_pw2_pool|voxelize_pool" -s 2 -c 2 -o gpurun_out/r2l_top python tools/ncu_score.py 1024 > gpurun_out/r2l_ncu.log 2>&1

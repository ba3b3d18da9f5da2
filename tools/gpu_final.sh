# Round-end evidence run on one B200: tests, smoke, bench arms, rows, launch list, ncu captures (outputs under gpurun_out/)
set -x
T=${1:-r3z}
python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -v "^$" | cut -c1-300 | tail -30 > gpurun_out/${T}_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${T}_smoke.log 2>&1
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_ref.json 2> gpurun_out/${T}_ref.err
timeout 1200 python bench.py --rows > gpurun_out/${T}_rows.jsonl 2> gpurun_out/${T}_rows.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --poses 2048 --no-cpu-baseline --no-gpu-reference > gpurun_out/${T}_bench_under_ncu.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"conv1_pw2_pool|voxelize_pool|conv3_tc|pointwise_pool_mma" -s 5 -c 5 -o gpurun_out/${T}_top python tools/ncu_score.py 1024 > gpurun_out/${T}_ncu.log 2>&1

set -x
python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -v "^$" | cut -c1-300 | tail -30 > gpurun_out/r2z_pytest.log
python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"conv1_pw2_pool" -s 1 -c 1 -o gpurun_out/r2z_fused python tools/ncu_score.py 1024 > gpurun_out/r2z_ncu.log 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2z_smoke.log 2>&1

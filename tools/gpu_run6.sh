set -x
python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -v "^$" | cut -c1-300 | tail -60 > gpurun_out/r2f_pytest.log
timeout 200 python bench.py --no-cpu-baseline --no-gpu-reference > gpurun_out/r2f_bench_fused.json 2> gpurun_out/r2f_bench_fused.err
GB_TC_FUSED_PERSIST=2 timeout 200 python bench.py --no-cpu-baseline --no-gpu-reference > gpurun_out/r2f_bench_fused_p2.json 2> gpurun_out/r2f_bench_fused_p2.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"conv1_pw2_pool|conv3_tc_kernel" -s 3 -c 3 -o gpurun_out/r2f_conv python tools/ncu_score.py 1024 > gpurun_out/r2f_ncu.log 2>&1

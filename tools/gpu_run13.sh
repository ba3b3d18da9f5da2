set -x
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_grad.py tests/test_gpu_parity.py tests/test_cpp_host.py -m gpu -q --tb=short 2>&1 | grep -v "^$" | cut -c1-300 | tail -40 > gpurun_out/r2m_pytest.log
timeout 200 python bench.py --no-cpu-baseline --no-gpu-reference > gpurun_out/r2m_bench_v4.json 2> gpurun_out/r2m_bench_v4.err
GB_VOX4=0 timeout 200 python bench.py --no-cpu-baseline --no-gpu-reference > gpurun_out/r2m_bench_v3.json 2> gpurun_out/r2m_bench_v3.err
timeout 200 python bench.py --no-cpu-baseline --no-gpu-reference --model default2017 > gpurun_out/r2m_bench_2017.json 2> gpurun_out/r2m_bench_2017.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"voxelize_avg_q" -s 1 -c 1 -o gpurun_out/r2m_vox4 python tools/ncu_score.py 1024 > gpurun_out/r2m_ncu.log 2>&1

set -x
python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -v "^$" | cut -c1-300 | tail -30 > gpurun_out/r2i_pytest.log
timeout 200 python bench.py --no-cpu-baseline --no-gpu-reference > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
GB_TC_FUSED_PERSIST=0 timeout 200 python bench.py --no-cpu-baseline --no-gpu-reference > gpurun_out/r2i_bench_p0.json 2> gpurun_out/r2i_bench_p0.err
timeout 600 python bench.py --rows > gpurun_out/r2i_rows.jsonl 2> gpurun_out/r2i_rows.err

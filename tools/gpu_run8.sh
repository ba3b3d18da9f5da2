set -x
python -m pytest tests/test_gpu_vina.py -q --tb=short 2>&1 | grep -v "^$" | cut -c1-300 | tail -20 > gpurun_out/r2h_vina.log
for P in 1 2 4; do GB_TC_FUSED_PERSIST=$P timeout 200 python bench.py --no-cpu-baseline --no-gpu-reference > gpurun_out/r2h_bench_p$P.json 2> gpurun_out/r2h_bench_p$P.err; done
GB_TC_FUSED_DBG=8 timeout 200 python bench.py --no-cpu-baseline --no-gpu-reference > gpurun_out/r2h_bench_dbg8.json 2> gpurun_out/r2h_bench_dbg8.err
GB_TC_FUSED_DBG=8 GB_TC_FUSED_PERSIST=1 timeout 200 python bench.py --no-cpu-baseline --no-gpu-reference > gpurun_out/r2h_bench_dbg8_p1.json 2> gpurun_out/r2h_bench_dbg8_p1.err
python tests/test_cpp_host.py > /dev/null 2>&1; python -m pytest tests/test_cpp_host.py tests/test_docking_pipeline.py -q --tb=short -m gpu 2>&1 | cut -c1-300 | tail -20 > gpurun_out/r2h_cpp.log

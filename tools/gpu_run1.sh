set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2a_pytest.log
for occ in 2 3 4; do GB_VOX_OCC=$occ python bench.py --no-cpu-baseline --no-gpu-reference > gpurun_out/r2a_bench_occ$occ.json 2>gpurun_out/r2a_err_occ$occ.log; done
GB_VOX_OCC=3 python bench.py --no-cpu-baseline --no-gpu-reference --overlap 1 > gpurun_out/r2a_bench_overlap.json 2>gpurun_out/r2a_err_overlap.log
python bench.py > gpurun_out/r2a_bench_full.json 2> gpurun_out/r2a_bench_full.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:voxelize_pool -s 1 -c 1 -o gpurun_out/r2a_vox python tools/ncu_score.py 1024 > gpurun_out/r2a_ncu.log 2>&1

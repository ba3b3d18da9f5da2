set -x
GB_TC_FUSED_V2=2 timeout 90 python tools/ncu_score.py 64 > gpurun_out/r3a_small.log 2>&1; echo "rc=$?" >> gpurun_out/r3a_small.log
GB_TC_FUSED_V2=1 timeout 90 python tools/ncu_score.py 64 > gpurun_out/r3a_small_ref.log 2>&1
GB_TC_FUSED_V2=2 timeout 120 python -m pytest tests/test_gpu_tc.py -m gpu -q --tb=short -x -k "fused or fast_scores or full_size" 2>&1 | tail -12 > gpurun_out/r3a_pytest.log
GB_TC_FUSED_V2=2 timeout 150 python bench.py --no-cpu-baseline --no-gpu-reference > gpurun_out/r3a_bench.json 2> gpurun_out/r3a_bench.err
nvidia-smi --query-gpu=name,memory.used --format=csv > gpurun_out/r3a_smi.log 2>&1

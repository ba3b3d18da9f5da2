set -x
(nproc; cat /sys/fs/cgroup/cpu.max; python -c "import os;print(os.cpu_count(), len(os.sched_getaffinity(0)))"; free -g | head -2) > gpurun_out/r2b_host.log 2>&1
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2b_pytest.log
python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
timeout 900 python bench.py --rows > gpurun_out/r2b_rows.jsonl 2> gpurun_out/r2b_rows.err

set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2k_bench_n2.json 2> gpurun_out/r2k_bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 1 --warmup 1 --workload screen --ligands 16384 > gpurun_out/r2k_screen_n2.json 2> gpurun_out/r2k_screen_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 2 --workload minimize > gpurun_out/r2k_min_n2.json 2> gpurun_out/r2k_min_n2.err

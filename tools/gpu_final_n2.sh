# two-GPU evidence: headline (weak + strong scaling block), config 4 and config 5 workloads, the reference arm under torchrun
set -x
T=${1:-r3z}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/${T}_bench_n2.json 2> gpurun_out/${T}_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 1 --warmup 1 --workload screen --ligands 16384 > gpurun_out/${T}_screen_n2.json 2> gpurun_out/${T}_screen_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 2 --workload minimize > gpurun_out/${T}_min_n2.json 2> gpurun_out/${T}_min_n2.err

set -x
timeout 400 python tools/dock_inflight.py 2000 32 64 128 > gpurun_out/r2n_inflight.jsonl 2> gpurun_out/r2n_inflight.err

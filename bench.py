#!/usr/bin/env python3
"""bench.py — poses/sec CNN-rescored (48^3 x 28ch default2018), BASELINE.json's metric.

A "step" = one pass of the hot path (voxelise -> CNN forward -> heads -> ensemble) over one batch of synthetic
poses against one synthetic receptor.  Workload at N=1 is BASELINE.json configs[1]: 1 receptor, 10k ligand
poses, crossdock_default2018.  N>1: weak scaling, every rank scores its own 10k poses (pose sharding, no data-path
collective; one NCCL all_gather of the per-pose results per step — SURVEY.md §8e).

  value : device-resident throughput — poses already staged in HBM, K steps of kernels timed with CUDA events on
          the library's stream (per-step event pairs; an L2 flush runs between steps, outside the events).
  e2e   : same metric through the public call a user makes (CNNScorer.score_batch -> C ABI gb_cnn_score_batch)
          with HOST buffers: pinned staging + H2D of the poses and D2H of the four result arrays inside the timed
          region, every step.
  --impl reference : the reference's CPU path restated (oracle/: C voxeliser + the same network in torch CPU ops,
          batch 1 per call, receptor re-voxelised per pose like torch_model.cpp:153-224), all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODEL = "crossdock_default2018"
FLOP_PER_EVAL = {"default2018": 0.998148096e9, "dense": 4.541571072e9, "default2017": 1.122729984e9}
CONV1_FLOP = 668_860_416.0  # conv3^3 28->32 @24^3
# dram__bytes_read.sum + dram__bytes_write.sum of the conv1 launch, ncu --set full, 1024 poses per launch
# (profiles/r1e_ncu_conv1_tcgen05.csv: 1.095620 GB + 0.863929 GB) -> per pose
CONV1_NCU_DRAM_BYTES_PER_POSE = (1.095620e9 + 0.863929e9) / 1024


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_workload(n_poses, seed):
    from gnina_b200 import synth
    rec_xyz, rec_t = synth.make_receptor()
    lx0, lt0 = synth.make_ligand()
    lig_xyz, offs = synth.make_poses(lx0, n_poses, seed=seed)
    return rec_xyz, rec_t, lig_xyz, np.tile(lt0, n_poses), offs


class CpuPort:
    """The reference's CPU algorithm restated (oracle/): C voxeliser + the same network in torch CPU fp32 ops, batch 1
    per CNN call, the receptor re-voxelised for every pose (torch_model.cpp:153-224).  Two ways of using the host:
      sequential    : poses one after another, torch intra-op threads = T (what `gnina --cpu T` does: one ligand
                      worker thread for rescoring, main/main.cpp:1432-1433, torch::set_num_threads, :1374)
      pose_parallel : W python threads each scoring whole poses with 1 torch thread (best effort for the host)
    """

    def __init__(self, n_poses):
        import torch
        from gnina_b200 import model_blob
        from oracle import pipeline
        self.torch = torch
        self.cores = os.cpu_count() or 1
        self.w = make_workload(n_poses, seed=1)
        self.om = pipeline.OracleModel(model_blob.load_model(MODEL))
        self.n = n_poses

    def one(self, i):
        rec_xyz, rec_t, lig_xyz, lig_t, offs = self.w
        a, b = offs[i], offs[i + 1]
        return self.om.score(rec_xyz, rec_t, lig_xyz[a:b], lig_t[a:b], np.array([0, b - a], np.int32), batch=1)

    def sequential(self, idx, threads):
        self.torch.set_num_threads(threads)
        t0 = time.perf_counter()
        for i in idx:
            self.one(i)
        return len(idx) / (time.perf_counter() - t0)

    def pose_parallel(self, idx, workers):
        from concurrent.futures import ThreadPoolExecutor
        self.torch.set_num_threads(1)
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=workers) as ex:
            list(ex.map(self.one, idx))
        return len(idx) / (time.perf_counter() - t0)

    def tune(self):
        """pick the faster host configuration on a small pilot; -> (mode, threads, pilot rate)"""
        self.one(0)
        best = ("sequential", 1, 0.0)
        for t in sorted({4, 8, 16, min(32, self.cores), self.cores}):
            if t > self.cores:
                continue
            r = self.sequential(range(min(3, self.n)), t)
            if r > best[2]:
                best = ("sequential", t, r)
        w = self.cores
        r = self.pose_parallel(range(min(2 * w, self.n)), w)
        if r > best[2]:
            best = ("pose_parallel", w, r)
        return best

    def run(self, mode, threads, idx):
        return self.sequential(idx, threads) if mode == "sequential" else self.pose_parallel(idx, threads)


def run_reference(args, rank, world):
    """Reference arm: the reference's own CPU algorithm (oracle port) on the host cores, bounded sample per step."""
    if rank != 0:
        return
    cpu = CpuPort(max(args.ref_sample, 2 * (os.cpu_count() or 1)))
    mode, threads, pilot = cpu.tune()
    sample = int(max(4, min(cpu.n, pilot * args.ref_step_seconds)))
    idx = list(range(sample))
    for _ in range(min(args.warmup, 1)):
        cpu.run(mode, threads, idx[: max(2, sample // 8)])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu.run(mode, threads, idx)
    dt = time.perf_counter() - t0
    v = sample * args.steps / dt
    desc = "%d poses/step, %s, %d threads, batch 1 per CNN call, receptor re-voxelised per pose" % (sample, mode, threads)
    line = {"impl": "reference", "metric": "poses/sec CNN-rescored (48^3x28ch default2018)", "value": v,
            "unit": "poses/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "CNN rescoring: 1 receptor (3000 atoms), synthetic ligand poses, 48^3x28ch "
                                   "crossdock_default2018", "sample_poses_per_step": sample, "host_mode": mode},
            "cpu_baseline": {"value": v, "unit": "poses/s", "cores": threads, "kind": "port", "sample": desc},
            "e2e": {"value": v, "unit": "poses/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def cpu_baseline(seconds_budget=15.0):
    cpu = CpuPort(4 * (os.cpu_count() or 1))
    mode, threads, pilot = cpu.tune()
    sample = int(max(4, min(cpu.n, pilot * seconds_budget)))
    v = cpu.run(mode, threads, list(range(sample)))
    seq = cpu.sequential(range(3), min(8, cpu.cores))
    return {"value": v, "unit": "poses/s", "cores": threads, "kind": "port",
            "sample": "%d poses, %s with %d threads (best of sequential/pose-parallel pilots), batch 1 per CNN call, "
                      "receptor re-voxelised per pose (torch_model.cpp:153-224 restated: oracle C voxeliser + torch "
                      "CPU fp32 network)" % (sample, mode, threads),
            "sequential_8_threads": seq}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--poses", type=int, default=10000, help="poses per GPU per step (config 2: 10k)")
    ap.add_argument("--precision", type=int, default=-1, help="-1 library default, 0 fp32 validation, 1 fp16 tensor-core")
    ap.add_argument("--ref-sample", type=int, default=4096, help="max poses per step for --impl reference")
    ap.add_argument("--ref-step-seconds", type=float, default=6.0, help="target CPU seconds per step (reference arm)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--model", default="", help="model name(s), comma separated; 'default' = gnina's default 3-model ensemble")
    ap.add_argument("--overlap", type=int, default=-1, help="voxeliser/network stream overlap (library option)")
    ap.add_argument("--max-batch", type=int, default=0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from gnina_b200 import CNNScorer
    W = max(args.warmup, 3)

    rec_xyz, rec_t, lig_xyz, lig_t, offs = make_workload(args.poses, seed=1 + rank)
    names = [MODEL] if not args.model else ([] if args.model == "default" else args.model.split(","))
    s = CNNScorer(names, device=local)
    if args.precision >= 0:
        s.set_option("precision", args.precision)
    if args.overlap >= 0:
        s.set_option("overlap", args.overlap)
    if args.max_batch:
        s.set_option("max_batch", args.max_batch)
    precision = int(s.get_option("precision"))
    s.set_receptor(rec_xyz, rec_t)
    stream = torch.cuda.ExternalStream(s.stream_ptr(), device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    gathered = [torch.empty(4 * args.poses, device=dev) for _ in range(world)] if world > 1 else None

    # ---------------- device-resident: value ----------------
    s.stage(lig_xyz, lig_t, offs)
    for _ in range(W):
        s.run_staged()
    s.set_option("profile", 1)
    s.profile_reset()
    launches0 = s.kernel_launches()
    clocks = ClockSampler(local)
    barrier()
    clocks.start()
    evs = []
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        s.run_staged()
        b.record(stream)
        evs.append((a, b))
    barrier()
    clk = clocks.stop()
    ms = sum(a.elapsed_time(b) for a, b in evs)
    launches = s.kernel_launches() - launches0
    prof = s.profile()
    s.set_option("profile", 0)
    res = s.fetch()
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * args.poses * args.steps / (ms * 1e-3)

    # ---------------- end to end through the public API: e2e ----------------
    out_host = None
    for _ in range(W):
        s.score_batch(lig_xyz, lig_t, offs)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out_host = s.score_batch(lig_xyz, lig_t, offs)           # H2D poses, kernels, D2H 4 x n floats
        if world > 1:                                             # final score gather (NCCL), SURVEY.md §8e
            dist.all_gather(gathered, torch.from_numpy(np.concatenate(out_host)).to(dev))
    barrier()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e = world * args.poses * args.steps / float(t.item())
    n_atoms = int(offs[-1])
    h2d = n_atoms * (16 + 4) + (args.poses + 1) * 4 + args.poses * 12
    d2h = 4 * 4 * args.poses

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    hbm, tf_burst, tf_sus, which = peaks()
    # dominant kernel: conv1 (67% of the network's FLOPs).  achieved = algorithmic FLOPs / event-measured duration
    conv1_keys = [k for k in prof if ("conv3_28x32" in k or k.startswith("tc_conv1"))]
    if args.model:
        conv1_keys = []
    roof = None
    if conv1_keys:
        k = conv1_keys[0]
        tot_ms, cnt = prof[k]
        flops = CONV1_FLOP * args.poses * args.steps
        ach = flops / (tot_ms * 1e-3) / 1e12
        roof = {"kernel": k, "bound": "tensor", "achieved": ach, "peak": tf_sus, "unit": "TFLOP/s",
                "frac": ach / tf_sus,
                "traffic": (CONV1_NCU_DRAM_BYTES_PER_POSE * args.poses * args.steps / max(cnt, 1)) if k.startswith("tc_conv1") else None,
                "traffic_unit": "DRAM bytes per launch (ncu capture profiles/r1e_ncu_conv1_tcgen05.csv, scaled by poses per launch)",
                "peak_source": "%s bf16 sustained (MEASURED_PEAKS.json)" % which,
                "launches": cnt, "avg_launch_ms": tot_ms / max(cnt, 1)}
    total_ms = sum(v[0] for v in prof.values())
    shares = {k: round(v[0] / total_ms, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])} if total_ms else {}
    line = {"metric": "poses/sec CNN-rescored (48^3x28ch default2018)", "value": value, "unit": "poses/s",
            "n_gpus": world, "steps": args.steps, "warmup": W, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16" if precision == 1 else "f32", "data": "synthetic",
            "config": {"workload": "CNN rescoring: 1 receptor (3000 atoms), %d synthetic ligand poses per GPU, 48^3x28ch "
                                   "%s" % (args.poses, ",".join(s.model_names)),
                       "precision": "fp16 tcgen05, fp32 accumulate" if precision == 1 else "fp32 CUDA-core validation path",
                       "l2": "explicit 256 MiB flush between timed steps; per-step intermediates >> L2",
                       "parallelism": "pose-sharded x%d" % world},
            "e2e": {"value": e2e, "unit": "poses/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches), "clocks": clk, "roofline": roof,
            "kernel_time_share": shares,
            "model_tflops": value * FLOP_PER_EVAL["default2018"] / 1e12,
            "checksum": float(np.sum(res[0], dtype=np.float64))}
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline()
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

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
  --rows : the other hot-path rows (gradient path, ensembles, Vina cache / exact score / docking inner loop, the
          config 3 pipeline), one JSON line each, with the CPU oracle timed beside them where it is cheap.
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
CONV1_FLOP = 668_860_416.0  # unit1_conv: conv3^3 28->32 @24^3
CONV2_FLOP = 28_311_552.0   # unit2_conv: conv1^3 32->32 @24^3, computed by the same (fused) kernel


def ncu_traffic(kernel_key):
    """DRAM bytes per pose of the dominant kernel from the latest committed `ncu --set full` capture of THIS build
    (profiles/traffic.json, written by tools/ncu_traffic.py from the .ncu-rep: dram__bytes_read.sum + dram__bytes_write.sum
    divided by the poses of the captured launch) -> (bytes per pose, source) or (None, None).  DRAM counters cannot be read
    inside a timed run; a capture that does not name this kernel yields null rather than a stale constant."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        d = json.load(open(path))
        e = d.get(kernel_key)
        if e:
            return float(e["dram_bytes_per_pose"]), e.get("source", "profiles/traffic.json")
    except (OSError, ValueError, KeyError):
        pass
    return None, None


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


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's CPU algorithm restated (oracle/), run by a pool of single-threaded worker PROCESSES
_W = {}


def _cpu_worker_init(path, batch):
    """worker process: 1 torch thread, private copy of the workload and of the oracle model"""
    import torch
    from gnina_b200 import model_blob
    from oracle import pipeline
    torch.set_num_threads(1)
    d = np.load(path)
    _W["w"] = (d["rec_xyz"], d["rec_t"], d["lig_xyz"], d["lig_t"], d["offs"])
    _W["om"] = pipeline.OracleModel(model_blob.load_model(MODEL))
    _W["batch"] = batch


def _cpu_worker_run(span):
    """score poses [a, b): C voxeliser (receptor re-voxelised for every pose) + the network in torch CPU fp32 ops,
    `batch` poses per CNN call; -> checksum of the pose scores"""
    a, b = span
    rec_xyz, rec_t, lig_xyz, lig_t, offs = _W["w"]
    lo, hi = offs[a], offs[b]
    r = _W["om"].score(rec_xyz, rec_t, lig_xyz[lo:hi], lig_t[lo:hi], offs[a:b + 1] - lo, batch=_W["batch"])
    return float(np.sum(r[0]))


def usable_cores():
    """host cores this process may actually use: the scheduler affinity mask and the cgroup CPU quota both cap it
    (os.cpu_count() reports the machine, not the container) -> (usable, os.cpu_count(), how)"""
    total = os.cpu_count() or 1
    n, how = total, "os.cpu_count"
    try:
        a = len(os.sched_getaffinity(0))
        if a < n:
            n, how = a, "sched_getaffinity"
    except (AttributeError, OSError):
        pass
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                quota, period = txt[0], float(txt[1])
            else:
                quota, period = txt[0], float(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if quota not in ("max", "-1"):
                q = max(1, int(float(quota) / period))
                if q < n:
                    n, how = q, "cgroup cpu quota"
            break
        except (OSError, ValueError, IndexError):
            continue
    return n, total, how


class CpuPort:
    """The reference's CPU algorithm restated (oracle/): C voxeliser + the same network in torch CPU fp32 ops, the
    receptor re-voxelised for every pose (torch_model.cpp:153-224).  Host usage: W worker processes (spawned, 1 torch
    thread each -- a Python thread pool serialises on the GIL and on torch's intra-op pool, which is what capped the
    round-1 number), each scoring contiguous spans of poses with `batch` poses per CNN call:
      batch 1  = the reference's own plumbing (one grid, one forward per pose)
      batch 32 = BASELINE.md 3.4's best-effort mode (same arithmetic, fewer calls)
    The faster of the two on a pilot is reported, with poses/s/core."""

    def __init__(self, n_poses, workers=None):
        import tempfile
        self.cores, self.machine_cores, self.cores_how = usable_cores()
        self.workers = workers or self.cores
        self.n = n_poses
        rec_xyz, rec_t, lig_xyz, lig_t, offs = make_workload(n_poses, seed=1)
        f = tempfile.NamedTemporaryFile(suffix=".npz", delete=False)
        f.close()
        np.savez(f.name, rec_xyz=rec_xyz, rec_t=rec_t, lig_xyz=lig_xyz, lig_t=lig_t, offs=offs)
        self.path = f.name
        self.pools = {}

    def pool(self, batch):
        import multiprocessing as mp
        if batch not in self.pools:
            ctx = mp.get_context("spawn")   # the parent may hold a CUDA context: never fork it
            self.pools[batch] = ctx.Pool(self.workers, initializer=_cpu_worker_init, initargs=(self.path, batch))
            self.pools[batch].map(_cpu_worker_run, [(0, 1)] * self.workers)   # every worker imports, loads, warms up
        return self.pools[batch]

    def run(self, batch, n):
        """score poses [0, n) once -> poses/s (wall clock over the pool)"""
        n = min(n, self.n)
        per = max(batch, -(-n // (self.workers * 4)))
        per = -(-per // batch) * batch
        spans = [(a, min(n, a + per)) for a in range(0, n, per)]
        pl = self.pool(batch)
        t0 = time.perf_counter()
        pl.map(_cpu_worker_run, spans, chunksize=1)
        return n / (time.perf_counter() - t0)

    def tune(self):
        """-> (batch, pilot poses/s): the faster of batch 1 and batch 32 on a pilot of a few poses per worker"""
        best = (1, 0.0)
        for batch in (1, 32):
            r = self.run(batch, min(self.n, self.workers * max(2, batch)))
            if r > best[1]:
                best = (batch, r)
        return best

    def close(self):
        for p in self.pools.values():
            p.terminate()
        self.pools = {}
        try:
            os.unlink(self.path)
        except OSError:
            pass

    def describe(self, batch, sample):
        return ("%d poses, %d worker processes x 1 torch thread, batch %d per CNN call, receptor re-voxelised per pose "
                "(torch_model.cpp:153-224 restated: oracle C voxeliser + torch CPU fp32 network)" % (sample, self.workers, batch))


def run_reference(args, rank, world):
    """Reference arm: the reference's own CPU algorithm (oracle port) on the host cores, bounded sample per step."""
    if rank != 0:
        return
    cpu = CpuPort(max(args.ref_sample, 64 * usable_cores()[0]))
    try:
        batch, pilot = cpu.tune()
        sample = int(max(cpu.workers * batch, min(cpu.n, pilot * args.ref_step_seconds)))
        for _ in range(min(args.warmup, 1)):
            cpu.run(batch, max(cpu.workers * batch, sample // 4))
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cpu.run(batch, sample)
        dt = time.perf_counter() - t0
        sample = min(sample, cpu.n)
        v = sample * args.steps / dt
        desc = cpu.describe(batch, sample)
    finally:
        cpu.close()
    line = {"impl": "reference", "metric": "poses/sec CNN-rescored (48^3x28ch default2018)", "value": v,
            "unit": "poses/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "CNN rescoring: 1 receptor (3000 atoms), synthetic ligand poses, 48^3x28ch "
                                   "crossdock_default2018", "sample_poses_per_step": sample, "batch_per_cnn_call": batch},
            "cpu_baseline": {"value": v, "unit": "poses/s", "cores": cpu.workers, "kind": "port", "sample": desc,
                             "poses_per_s_per_core": v / cpu.workers, "machine_cores": cpu.machine_cores,
                             "cores_from": cpu.cores_how},
            "e2e": {"value": v, "unit": "poses/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def cpu_baseline(seconds_budget=12.0):
    cpu = CpuPort(64 * usable_cores()[0])
    try:
        batch, pilot = cpu.tune()
        sample = int(max(cpu.workers * batch, min(cpu.n, pilot * seconds_budget)))
        v = cpu.run(batch, sample)
        other = 32 if batch == 1 else 1
        alt = cpu.run(other, min(sample, cpu.workers * max(4, other)))
        # one worker alone: what a core does when nothing else competes for memory bandwidth and caches
        spans = [(0, 4)]
        t0 = time.perf_counter()
        cpu.pool(1).map(_cpu_worker_run, spans)
        single = 4 / (time.perf_counter() - t0)
        return {"value": v, "unit": "poses/s", "cores": cpu.workers, "kind": "port", "sample": cpu.describe(batch, min(sample, cpu.n)),
                "poses_per_s_per_core": v / cpu.workers, "batch_%d_poses_per_s" % other: alt,
                "single_process_poses_per_s": single, "machine_cores": cpu.machine_cores, "cores_from": cpu.cores_how}
    finally:
        cpu.close()


# ---------------------------------------------------------------------------------------------------------------
# GPU reference: what gnina's own single-GPU path executes for this metric -- the TorchScript network through
# libtorch/cuDNN (torch_model.cpp:185) -- restated with torch.nn.functional on the same weights, on this GPU.
def gpu_reference(dev, seconds=2.0):
    """-> {"faithful": ..., "batched": ...} poses/s of the NETWORK ALONE on pre-voxelised grids already in HBM.
    The reference's GPU voxeliser (libmolgrid, not in this image) and its host-side make_coordset are NOT timed, so
    both figures are upper bounds of what the reference reaches on this GPU.
      faithful : batch 1, fp32 (cuDNN, TF32 convolutions allowed = libtorch's default), and the three .item()
                 device syncs per pose of torch_model.cpp:188-195,222
      batched  : best effort, batch 64, channels_last_3d, fp16 autocast, one sync per batch"""
    import torch
    import torch.nn.functional as F
    from gnina_b200 import model_blob
    blob = model_blob.load_model(MODEL)
    W = {k: torch.from_numpy(np.array(v)).to(dev) for k, v in blob.tensors.items()}

    def net(x):
        x = F.avg_pool3d(x, 2, 2)
        x = F.relu(F.conv3d(x, W["unit1_conv.weight"], W["unit1_conv.bias"], padding=1))
        x = F.relu(F.conv3d(x, W["unit2_conv.weight"], W["unit2_conv.bias"]))
        x = F.avg_pool3d(x, 2, 2)
        x = F.relu(F.conv3d(x, W["unit3_conv.weight"], W["unit3_conv.bias"], padding=1))
        x = F.relu(F.conv3d(x, W["unit4_conv.weight"], W["unit4_conv.bias"]))
        x = F.avg_pool3d(x, 2, 2)
        x = F.relu(F.conv3d(x, W["unit5_conv.weight"], W["unit5_conv.bias"], padding=1))
        f = x.reshape(x.shape[0], -1)
        pose = F.log_softmax(F.linear(f, W["pose_output.weight"], W["pose_output.bias"]), 1)
        aff = F.linear(f, W["affinity_output.weight"], W["affinity_output.bias"])
        return pose, aff

    out = {}
    torch.backends.cudnn.benchmark = True
    with torch.no_grad():
        g1 = torch.rand(8, 1, 28, 48, 48, 48, device=dev)
        label = torch.ones(1, dtype=torch.long, device=dev)

        def one(i):
            pose, aff = net(g1[i % 8])
            s = torch.softmax(pose, 1)[0, 1].item()            # torch_model.cpp:188-191
            a = aff[0, 0].item()                               # :192
            l = F.cross_entropy(pose, label).item()            # :195
            return s + a + l
        for i in range(20):
            one(i)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 0
        while time.perf_counter() - t0 < seconds:
            one(n)
            n += 1
        torch.cuda.synchronize()
        out["faithful"] = {"value": n / (time.perf_counter() - t0), "unit": "poses/s", "batch": 1, "dtype": "fp32 (TF32 conv allowed)",
                           "syncs_per_pose": 3}
        B = 64
        gb = torch.rand(B, 28, 48, 48, 48, device=dev).to(memory_format=torch.channels_last_3d)
        Wh = {k: (v.to(memory_format=torch.channels_last_3d) if v.dim() == 5 else v) for k, v in W.items()}
        W.update(Wh)

        def batch():
            with torch.autocast("cuda", dtype=torch.float16):
                pose, aff = net(gb)
            return float(torch.softmax(pose.float(), 1)[:, 1].sum().item())
        for _ in range(3):
            batch()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        a.record()
        for _ in range(reps):
            batch()
        b.record()
        torch.cuda.synchronize()
        out["batched"] = {"value": B * reps / (a.elapsed_time(b) * 1e-3), "unit": "poses/s", "batch": B,
                          "dtype": "fp16 autocast, channels_last_3d", "syncs_per_batch": 1}
    out["note"] = ("network only (torch %s / cuDNN %s) on pre-voxelised grids resident in HBM; the reference's voxeliser and "
                   "host code are not timed: upper bounds for gnina's single-GPU path" % (torch.__version__, torch.backends.cudnn.version()))
    del W
    torch.cuda.empty_cache()
    return out


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
    ap.add_argument("--no-gpu-reference", action="store_true")
    ap.add_argument("--model", default="", help="model name(s), comma separated; 'default' = gnina's default 3-model ensemble")
    ap.add_argument("--overlap", type=int, default=-1, help="voxeliser/network stream overlap (library option)")
    ap.add_argument("--max-batch", type=int, default=0)
    ap.add_argument("--workload", default="rescoring", choices=["rescoring", "screen", "minimize"],
                    help="rescoring = BASELINE config 2 (the headline); screen = config 4 (100k ragged ligands x dense_ensemble, "
                         "sharded over the GPUs); minimize = config 5 (1k poses with atom gradients, sharded)")
    ap.add_argument("--ligands", type=int, default=100000, help="total ligands of --workload screen (all GPUs together)")
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

    if args.workload != "rescoring":
        run_other_workload(args, rank, world, local, dev, dist)
        return
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

    gathered = torch.empty(world * 4 * args.poses, device=dev) if world > 1 else None
    mine = torch.empty(4 * args.poses, device=dev)

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
        if world > 1:                                             # final score gather (NCCL), SURVEY.md §8e:
            s.fetch_device(mine.data_ptr())                       # straight from the handle's device results, no host bounce
            dist.all_gather_into_tensor(gathered, mine)
    barrier()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e = world * args.poses * args.steps / float(t.item())
    n_atoms = int(offs[-1])
    h2d = n_atoms * (16 + 4) + (args.poses + 1) * 4 + args.poses * 12
    d2h = 4 * 4 * args.poses

    # ---------------- strong scaling: config 2's FIXED 10k poses split over the ranks ----------------
    strong = None
    if world > 1:
        per = args.poses // world
        lo = rank * per
        s.stage(lig_xyz[offs[lo]:offs[lo + per]], lig_t[offs[lo]:offs[lo + per]], offs[lo:lo + per + 1] - offs[lo])
        for _ in range(W):
            s.run_staged()
        barrier()
        evs2 = []
        for _ in range(args.steps):
            flush.fill_(1)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            s.run_staged()
            b.record(stream)
            evs2.append((a, b))
        barrier()
        t = torch.tensor([sum(a.elapsed_time(b) for a, b in evs2)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        strong = {"value": per * world * args.steps / (float(t.item()) * 1e-3), "unit": "poses/s", "total_poses": per * world,
                  "poses_per_gpu": per, "scaling": "strong", "note": "BASELINE config 2's fixed 10k poses split over the ranks, device-resident"}
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
        fused = "pw2" in k   # the fused kernel also computes unit2_conv
        flops = (CONV1_FLOP + (CONV2_FLOP if fused else 0.0)) * args.poses * args.steps
        ach = flops / (tot_ms * 1e-3) / 1e12
        per_pose, src = ncu_traffic(k)
        roof = {"kernel": k, "bound": "tensor", "achieved": ach, "peak": tf_sus, "unit": "TFLOP/s",
                "frac": ach / tf_sus, "frac_of_burst_peak": ach / tf_burst,
                "algorithmic_flop_per_pose": CONV1_FLOP + (CONV2_FLOP if fused else 0.0),
                "traffic": (per_pose * args.poses * args.steps / max(cnt, 1)) if per_pose else None,
                "traffic_unit": "DRAM bytes per launch = ncu dram__bytes_read.sum + dram__bytes_write.sum per pose (%s) x poses per launch" % src
                                if per_pose else None,
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
    if strong:
        line["strong_scaling"] = strong
    if not args.no_gpu_reference and world == 1 and not args.model:
        del flush
        torch.cuda.empty_cache()
        try:
            line["gpu_reference"] = gpu_reference(dev)
            line["gpu_reference"]["ours_over_faithful"] = e2e / line["gpu_reference"]["faithful"]["value"]
            line["gpu_reference"]["ours_over_batched"] = e2e / line["gpu_reference"]["batched"]["value"]
        except Exception as ex:   # the torch/cuDNN arm is a yardstick, never a reason to lose the bench line
            line["gpu_reference"] = {"unavailable": repr(ex)[:200]}
    if not args.no_cpu_baseline and world == 1:   # the CPU port is timed beside the N = 1 run only
        line["cpu_baseline"] = cpu_baseline()
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_other_workload(args, rank, world, local, dev, dist):
    """BASELINE configs 4 and 5 as sharded, SCALE-able bench lines (same JSON contract, their own metric):
      screen   : config 4 -- args.ligands ragged ligands (N_heavy ~ U[15,45]) x 1 pose x `--cnn dense_ensemble` (20 models),
                 ligands split over the ranks (strong scaling: the total is fixed), results gathered with NCCL from device
                 buffers; metric ligands/s through the public call (host buffers in, four floats per ligand out)
      minimize : config 5 -- 1000 poses with atom gradients (G2 + N5), default2018, split over the ranks"""
    import torch
    from gnina_b200 import CNNScorer, synth
    rec_xyz, rec_t = synth.make_receptor()
    W = max(args.warmup, 1)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    if args.workload == "screen":
        total = args.ligands
        per = total // world
        # every rank generates only its own shard (deterministic per ligand index range)
        sx, st, so = synth.make_screen(min(per, 4096), seed=3 + rank)
        reps = -(-per // (len(so) - 1))                     # the shard = the 4096 generated ligands repeated (timing only)
        s = CNNScorer(["dense_ensemble"], device=local, precision=1)
        s.set_receptor(rec_xyz, rec_t)
        nl = len(so) - 1
        mine = torch.empty(4 * nl, device=dev)
        gathered = torch.empty(world * 4 * nl, device=dev) if world > 1 else None

        def step():
            done = 0
            for r in range(reps):
                k = min(nl, per - done)
                s.score_batch(sx[:so[k]], st[:so[k]], so[:k + 1])
                if world > 1:
                    s.fetch_device(mine.data_ptr())
                    dist.all_gather_into_tensor(gathered, mine)
                done += k
        metric, unit, n_units = "ligands/sec virtual screen (dense_ensemble, 20 models, 48^3x28ch)", "ligands/s", per * world
        cfg = {"workload": "config 4: %d ragged ligands x dense_ensemble (20 models), %d per GPU in batches of %d" % (per * world, per, nl),
               "parallelism": "ligand-sharded x%d, NCCL all_gather of the device results per batch" % world}
    else:
        total = 1000
        per = total // world
        lx0, lt0 = synth.make_ligand()
        lx, offs = synth.make_poses(lx0, per, seed=5 + rank)
        lt = np.tile(lt0, per)
        s = CNNScorer([MODEL], device=local, precision=1)
        s.set_receptor(rec_xyz, rec_t)

        def step():
            s.score_grad_batch(lx, lt, offs)
        metric, unit, n_units = "poses/sec with atom gradients (CNN minimisation step, default2018)", "poses/s", per * world
        cfg = {"workload": "config 5: %d poses with ligand-atom gradients (forward + backward), %d per GPU" % (per * world, per),
               "parallelism": "pose-sharded x%d" % world}
    for _ in range(W):
        step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    barrier()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        v = n_units * args.steps / float(t.item())
        print(json.dumps({"metric": metric, "value": v, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": W,
                          "ms_per_step": 1e3 * float(t.item()) / args.steps, "higher_is_better": True, "scaling": "strong",
                          "vs_baseline": None, "dtype": "f16", "data": "synthetic", "config": cfg,
                          "e2e": {"value": v, "unit": unit, "note": "timed through the public call with host buffers (the value IS end to end)"},
                          "gpu_launches": int(s.kernel_launches())}))
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------
# other rows
def timed(fn, reps=3):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


def reference_vina_rates(lig, rec_xyz, rec_t, begin, end, ng, confs, mc_steps=40, mc_maxiters=12, mc_saved=8):
    """The reference's OWN Vina code (oracle/_ref: gnina's sources compiled where they lie, one host thread) timed on the inputs of the
    docking rows: cache::populate, model::eval_deriv, quasi_newton (12 iterations), monte_carlo::operator().  None where the
    library is absent (it is built in the development container and travels with the snapshot)."""
    from oracle import vina_refbuild as R
    if not R.available():
        return None
    sf = R.RefScoring()
    rm = R.RefModel(lig, rec_xyz, rec_t)
    t0 = time.perf_counter()
    cg = R.RefGrid.cache(sf, R.LINEAR, rm, begin, end, ng, 1e3)
    needed = sorted(set(int(t) for t in lig["types"] if t > 1))
    out = {"populate_point_types_per_s": float(np.prod(np.asarray(ng) + 1)) * len(needed) / (time.perf_counter() - t0)}
    t0 = time.perf_counter()
    for x in confs[:400]:
        R.model_eval_deriv(rm, sf, R.LINEAR, cg, x)
    out["eval_deriv_per_s"] = 400 / (time.perf_counter() - t0)
    t0 = time.perf_counter()
    for x in confs[:40]:
        R.bfgs(rm, sf, R.LINEAR, cg, x, 12)
    out["bfgs12_per_s"] = 40 / (time.perf_counter() - t0)
    t0 = time.perf_counter()
    for c in range(3):
        R.mc(rm, sf, R.LINEAR, cg, 1000 + c, [-6, -6, -6], [6, 6, 6], mc_steps, mc_maxiters, lig["conf0"], num_saved_mins=mc_saved,
             min_rmsd=0.5, hunt_cap=(10, 1.5, 10))
    out["mc_steps_per_s"] = 3 * mc_steps / (time.perf_counter() - t0)
    out["kind"] = "reference (gnina's own lib/*.cpp via oracle/_ref), 1 thread"
    return out


def rows_main():
    """python bench.py --rows: the hot-path rows that are not the headline, one JSON line per row:
    device throughput, and where cheap the CPU oracle timed beside it on a bounded sample (the cpu_baseline leg of these
    rows -- the only other place besides the headline's cpu_baseline where this file executes oracle/)."""
    import torch
    from gnina_b200 import CNNScorer, model_blob, synth
    from gnina_b200.vina import VinaScorer
    from oracle import pipeline
    from oracle.vina import VinaOracle
    assert torch.cuda.is_available(), "no CUDA device — no CPU fallback"
    torch.set_num_threads(min(16, os.cpu_count() or 1))
    rec_xyz, rec_t = synth.make_receptor()
    lx0, lt0 = synth.make_ligand()
    out = []
    # --rows-select cnn | vina | config3 (default all): a subset for short GPU sessions
    sel = sys.argv[sys.argv.index("--rows-select") + 1] if "--rows-select" in sys.argv else "all"

    if sel in ("all", "cnn"):
        # --- CNN gradient path (config 5 shape: 1k poses with atom gradients) ---
        n = 1000
        lx, offs = synth.make_poses(lx0, n, seed=5)
        lt = np.tile(lt0, n)
        om = pipeline.OracleModel(model_blob.load_model("crossdock_default2018"))
        k = 2
        t0 = time.perf_counter()
        pipeline.score_grad([om], rec_xyz, rec_t, lx[:offs[k]], lt[:offs[k]], offs[:k + 1], dtype=torch.float32)
        cpu = k / (time.perf_counter() - t0)
        for prec, mode, nn in ((1, "fp16 tcgen05 forward + backward", n), (0, "fp32 validation kernels", 256)):
            s = CNNScorer(["crossdock_default2018"], precision=prec)
            s.set_receptor(rec_xyz, rec_t)
            dt = timed(lambda: s.score_grad_batch(lx[:offs[nn]], lt[:offs[nn]], offs[:nn + 1]), reps=3)
            out.append({"row": "cnn_gradient (G2+N5+S1), %d poses" % nn, "value": nn / dt, "unit": "poses/s", "mode": mode,
                        "cpu_oracle": cpu, "cpu_sample": "%d poses, torch autograd + C gridmaker backward" % k})
        # larger batch: the config-5 workload batched over many ligands
        n_big = 8192
        lxb, offb = synth.make_poses(lx0, n_big, seed=6)
        ltb = np.tile(lt0, n_big)
        s = CNNScorer(["crossdock_default2018"], precision=1)
        s.set_receptor(rec_xyz, rec_t)
        dt = timed(lambda: s.score_grad_batch(lxb, ltb, offb), reps=3)
        out.append({"row": "cnn_gradient (G2+N5+S1), %d poses" % n_big, "value": n_big / dt, "unit": "poses/s",
                    "mode": "fp16 tcgen05 forward + backward"})

        # --- config 4 shape: virtual screen, ragged ligands, --cnn dense_ensemble (20 models) ---
        n4 = 4096
        sx, st, so = synth.make_screen(n4, seed=3)
        e20 = CNNScorer(["dense_ensemble"], precision=1)
        e20.set_receptor(rec_xyz, rec_t)
        dt = timed(lambda: e20.score_batch(sx, st, so), reps=2)
        out.append({"row": "virtual screen, dense_ensemble (config 4 shape), %d ligands" % n4, "value": n4 / dt, "unit": "ligands/s",
                    "mode": "fp16 tcgen05", "models": len(e20.model_names), "model_evals_per_s": n4 * len(e20.model_names) / dt})
        del e20

        # --- fp32 validation forward and the default 3-model ensemble ---
        n2 = 512
        for names, tag in ((["crossdock_default2018"], "cnn_validation_fp32 (N1)"), ([], "default_ensemble 3 models (N1+N2, S1)")):
            e = CNNScorer(names, precision=0)
            e.set_receptor(rec_xyz, rec_t)
            dt = timed(lambda: e.score_batch(lx[:offs[n2]], lt[:offs[n2]], offs[:n2 + 1]), reps=2)
            out.append({"row": tag + " [fp32 validation kernels]", "value": n2 / dt, "unit": "poses/s", "mode": "fp32 validation kernels",
                        "models": e.model_names})
        # gnina's default (--cnn unset): the 3-model ensemble on the fast path, and its score + atom-gradient call (config 5 as gnina
        # runs it by default; the two dense members' backward runs on the fp32 kernels)
        e3 = CNNScorer([], precision=1)
        e3.set_receptor(rec_xyz, rec_t)
        n3 = 4096
        lx3, offs3 = synth.make_poses(lx0, n3, seed=8)
        lt3 = np.tile(lt0, n3)
        dt = timed(lambda: e3.score_batch(lx3, lt3, offs3), reps=2)
        out.append({"row": "default_ensemble 3 models (N1+N2, S1), fast path", "value": n3 / dt, "unit": "poses/s", "mode": "fp16 tcgen05",
                    "models": e3.model_names})
        ng = 128
        dt = timed(lambda: e3.score_grad_batch(lx3[:offs3[ng]], lt3[:offs3[ng]], offs3[:ng + 1]), reps=2)
        out.append({"row": "default_ensemble 3 models, score + atom gradients (config 5 default)", "value": ng / dt, "unit": "poses/s",
                    "mode": "default2018 member: tcgen05 forward + backward; dense members: fp32 kernels", "models": e3.model_names})
        del e3

    if sel in ("all", "vina", "config3"):
        # --- Vina rows ---
        v, o = VinaScorer(), VinaOracle()
        v.set_receptor(rec_xyz, rec_t)
        begin, end, ng = [-12.0] * 3, [12.0] * 3, [64, 64, 64]     # 0.375 A spacing, 65^3 points
        needed = sorted(set(int(t) for t in lt0 if t > 1))
        dt = timed(lambda: v.cache_build(begin, end, ng, needed), reps=2)
        pts = 65 ** 3 * len(needed)
        t0 = time.perf_counter()
        o.cache_populate(begin, end, [16, 16, 16], rec_xyz, rec_t, needed[0])
        cpu = 17 ** 3 / (time.perf_counter() - t0)
        out.append({"row": "vina_cache_build (V4)", "value": pts / dt, "unit": "grid-point-types/s", "grid": "65^3 x %d types" % len(needed),
                    "cpu_oracle": cpu, "cpu_sample": "17^3 points x 1 type, scalar C"})
        nv = 20000
        lxv, offv = synth.make_poses(lx0, nv, trans_box=12, seed=6)
        ltv = np.tile(lt0, nv)
        dt = timed(lambda: v.cache_eval(lxv, ltv, offv), reps=3)
        grids = {t: v.cache_grid(t) for t in needed}
        kk = 200
        t0 = time.perf_counter()
        for p in range(kk):
            VinaOracle.cache_eval(grids, begin, end, ng, lxv[offv[p]:offv[p + 1]], lt0, 1e3, 1000.0)
        cpu = kk / (time.perf_counter() - t0)
        out.append({"row": "vina_cache_eval+deriv (V5)", "value": nv / dt, "unit": "poses/s", "cpu_oracle": cpu,
                    "cpu_sample": "%d poses, scalar C via ctypes" % kk})
        dt = timed(lambda: v.score_exact(lxv, ltv, offv, np.full(nv, 4.0, np.float32)), reps=3)
        kk = 20
        t0 = time.perf_counter()
        for p in range(kk):
            o.naive_exact(rec_xyz, rec_t, lxv[offv[p]:offv[p + 1]], lt0)
        cpu = kk / (time.perf_counter() - t0)
        out.append({"row": "vina_exact_affinity (V12)", "value": nv / dt, "unit": "poses/s", "cpu_oracle": cpu,
                    "cpu_sample": "%d poses, scalar C" % kk, "note": "includes H2D of the poses and D2H of the energies"})
        # --- docking inner loop (V6-V11): one warp per conformation / chain ---
        from oracle.vina_mc import DockOracle
        lig = synth.make_flexible_ligand()
        needed2 = sorted(set(int(t) for t in lig["types"] if t > 1))
        v.cache_build(begin, end, ng, needed2)
        v.set_ligand(lig)
        d = DockOracle(o, {t: v.cache_grid(t) for t in needed2}, begin, end, ng, lig)
        X = np.stack([d.random_conf(1 + i, [-6, -6, -6], [6, 6, 6])[0] for i in range(512)])
        Xb = np.tile(X, (64, 1))                                     # 32768 conformations
        dt = timed(lambda: v.eval_deriv(Xb), reps=3)
        t0 = time.perf_counter()
        for x in X[:200]:
            d.eval_deriv(x)
        cpu = 200 / (time.perf_counter() - t0)
        ref = reference_vina_rates(lig, rec_xyz, rec_t, begin, end, ng, X)
        out.append({"row": "dock_eval_deriv (V5+V6+V7+V8)", "value": len(Xb) / dt, "unit": "eval_deriv/s", "cpu_oracle": cpu,
                    "cpu_sample": "200 conformations, scalar C", "ligand": "27 heavy atoms, 6 torsions, %d pairs" % len(lig["pair_a"]),
                    "cpu_reference": ref and ref["eval_deriv_per_s"], "cpu_reference_kind": ref and ref["kind"],
                    "cpu_reference_cache_populate_point_types_per_s": ref and ref["populate_point_types_per_s"]})
        res = {}
        def run_bfgs():
            res["ne"] = v.bfgs(Xb[:8192], 12)[3]
        dt = timed(run_bfgs, reps=2)
        t0 = time.perf_counter()
        ner = sum(d.bfgs(x, 12)[3] for x in X[:40])
        cdt = time.perf_counter() - t0
        out.append({"row": "dock_bfgs 12 iterations (V9)", "value": 8192 / dt, "unit": "minimisations/s",
                    "device_eval_deriv_per_s": float(res["ne"].sum()) / dt, "cpu_oracle": 40 / cdt, "cpu_eval_deriv_per_s": ner / cdt,
                    "cpu_reference": ref and ref["bfgs12_per_s"]})
        n_chains, steps = 4096, 40
        seeds = (np.arange(1, n_chains + 1, dtype=np.uint32) * 2654435761) & 0xFFFFFFFF
        dt = timed(lambda: v.mc(seeds, [-6, -6, -6], [6, 6, 6], steps, 12, 8), reps=1)
        t0 = time.perf_counter()
        for c in range(3):
            d.mc(int(seeds[c]), [-6, -6, -6], [6, 6, 6], steps, 12, 8)
        cpu = 3 * steps / (time.perf_counter() - t0)
        out.append({"row": "dock_monte_carlo chains (V10+V11)", "value": n_chains * steps / dt, "unit": "MC steps/s",
                    "chains": n_chains, "steps_per_chain": steps, "ligands_per_s_at_exhaustiveness_64": n_chains / 64 / dt,
                    "cpu_oracle": cpu, "cpu_sample": "3 chains, scalar C, 1 thread", "cpu_reference": ref and ref["mc_steps_per_s"],
                    "cpu_reference_kind": ref and ref["kind"]})
    if sel in ("all", "config3"):
        # --- config 3 glue: cache build -> 64 chains -> merge -> CNN rescoring -> exact affinity -> ranked modes ---
        from gnina_b200 import docking
        cs = CNNScorer(["crossdock_default2018"], precision=1)
        cs.set_receptor(rec_xyz, rec_t)
        ref_steps = docking.reference_num_steps(len(lig["types"]), 6 + v.T)
        st = 200
        docking.dock_ligand(v, cs, lig, [-6, -6, -6], [6, 6, 6], exhaustiveness=64, seed=1, num_steps=st)
        t0 = time.perf_counter()
        poses = docking.dock_ligand(v, cs, lig, [-6, -6, -6], [6, 6, 6], exhaustiveness=64, seed=2, num_steps=st)
        dt = time.perf_counter() - t0
        out.append({"row": "dock + rescore pipeline, one ligand at a time (config 3 glue)", "value": 1.0 / dt, "unit": "ligands/s",
                    "exhaustiveness": 64, "mc_steps_per_chain": st, "reference_formula_steps": ref_steps, "modes_out": len(poses),
                    "note": "64 chains = 64 warps: one ligand cannot fill the GPU; throughput needs ligands in flight concurrently "
                            "(the MC row above runs 4096 chains per launch)"})
        n_l, workers = 96, 16
        ligs = [synth.make_flexible_ligand(n_heavy=20 + (i % 8), n_tors=3 + i % 4, seed=100 + i) for i in range(n_l)]
        kw = dict(exhaustiveness=64, num_steps=st)
        with docking.DockingPool(rec_xyz, rec_t, ["crossdock_default2018"], n_workers=workers) as pool:
            pool.dock(ligs[:2 * workers], [-6, -6, -6], [6, 6, 6], **kw)          # every worker builds its tables / workspaces
            t0 = time.perf_counter()
            res = pool.dock(ligs, [-6, -6, -6], [6, 6, 6], **kw)
            dt = time.perf_counter() - t0
        out.append({"row": "dock + rescore pipeline, %d ligands in flight (config 3 glue)" % workers, "value": n_l / dt, "unit": "ligands/s",
                    "exhaustiveness": 64, "mc_steps_per_chain": st, "ligands": n_l, "host_threads": workers,
                    "mc_steps_per_s": n_l * 64 * st / dt, "modes_out_mean": float(np.mean([len(r) for r in res]))})
        # --- config 3 at the REFERENCE's Monte-Carlo length: num_steps = 105 (50 + N_atoms + 10 DOF) per chain (main/main.cpp:
        # 442-443), exhaustiveness 64, search -> merge -> refine_structure -> CNN rescoring -> exact affinity.  A bounded sample of
        # ligands, all in flight at once (one Vina handle + CNN clone per host thread), extrapolated to BASELINE's 1k ligands.
        # search box: what --autobox_ligand with the default --autobox_add 4 gives for these ligands (extent ~12 A + 4 A on every side).
        # The reference uses ONE box for grids, random starts and penalties (docking.search_box); a box as tight as the ligand keeps the
        # output containers from ever filling (few RMSD-distinct poses fit), which makes every accepted step a "promising" one
        # (second quasi-Newton run): measured 0.50 ligands/s with +-6 A, profiles/README.md r5z
        AB1, AB2 = [-10, -10, -10], [10, 10, 10]
        n_full, workers_full = 64, 64
        ligs_full = [synth.make_flexible_ligand(n_heavy=20 + (i % 8), n_tors=3 + i % 4, seed=300 + i) for i in range(n_full)]
        steps_ref = [docking.reference_num_steps(len(l["types"]), 6 + len(l["seg_parent"]) - 1) for l in ligs_full]
        with docking.DockingPool(rec_xyz, rec_t, ["crossdock_default2018"], n_workers=workers_full) as pool:
            pool.dock(ligs_full, AB1, AB2, exhaustiveness=64, num_steps=50)   # warm-up: tables, workspaces
            t0 = time.perf_counter()
            res = pool.dock(ligs_full, AB1, AB2, exhaustiveness=64)             # num_steps=None -> reference formula
            dt = time.perf_counter() - t0
        mc_steps = 64 * float(np.sum(steps_ref))
        # the CPU oracle beside it: one chain of one ligand for a few hundred steps, scaled to the full length
        lig0 = ligs_full[0]
        needed0 = sorted(set(int(t) for t in lig0["types"] if t > 1))
        v.cache_build(begin, end, ng, needed0)
        v.set_ligand(lig0)
        d0 = DockOracle(o, {t: v.cache_grid(t) for t in needed0}, begin, end, ng, lig0)
        t0 = time.perf_counter()
        d0.mc(12345, [-6, -6, -6], [6, 6, 6], 150, int((25 + len(lig0["types"])) // 3), 50, min_rmsd=1.0, hunt_cap=(10, 10, 10))
        cpu_steps_per_s = 150 / (time.perf_counter() - t0)
        out.append({"row": "config 3 at the reference's step count: dock + refine + rescore, exhaustiveness 64", "value": n_full / dt, "unit": "ligands/s",
                    "ligands": n_full, "in_flight": workers_full, "search_box": "+-10 A (autobox_add 4)", "mc_steps_per_chain_mean": float(np.mean(steps_ref)),
                    "mc_steps_per_s": mc_steps / dt, "seconds_for_1k_ligands": 1000.0 * dt / n_full, "modes_out_mean": float(np.mean([len(r) for r in res])),
                    "cpu_oracle_mc_steps_per_s_1_thread": cpu_steps_per_s,
                    "cpu_oracle_ligands_per_s_per_thread": cpu_steps_per_s / (64 * float(np.mean(steps_ref))),
                    "cpu_reference_mc_steps_per_s_1_thread": ref and ref["mc_steps_per_s"],
                    "cpu_reference_ligands_per_s_per_thread": ref and ref["mc_steps_per_s"] / (64 * float(np.mean(steps_ref)))})
    for r in out:
        print(json.dumps(r))



if __name__ == "__main__":
    if "--rows" in sys.argv:
        rows_main()
    else:
        main()

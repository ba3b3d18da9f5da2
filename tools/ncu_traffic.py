"""profiles/traffic.json from an `ncu --set full` report: DRAM bytes per pose of the captured launch of each hot kernel.
  python tools/ncu_traffic.py gpurun_out/<rep>.ncu-rep <poses per launch> <source label>"""
import csv, io, json, os, subprocess, sys
rep, poses, label = sys.argv[1], int(sys.argv[2]), sys.argv[3]
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
txt = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True, stderr=subprocess.DEVNULL)
rows = list(csv.reader(io.StringIO(txt)))
h, units = rows[0], rows[1]
KEYS = {"conv1_pw2_pool": "tc_conv1_pw2_pool_fused", "conv3_tc_v2_kernel<32, 12>": "tc_conv3_3x3x3_32x64_d12", "conv3_tc_v2_kernel<64, 6>": "tc_conv5_3x3x3_64x128_d6", "pointwise_pool_mma_kernel<64, 12>": "tc_pw4_pool", "voxelize_pool_f16_kernel<0": "tc_voxelize_pool"}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
path = os.path.join(ROOT, "profiles", "traffic.json")
out = json.load(open(path)) if os.path.exists(path) else {}
for r in rows[2:]:
    d = dict(zip(h, r))
    for frag, key in KEYS.items():
        if frag in d["Kernel Name"]:
            rd = float(d["dram__bytes_read.sum"]) * scale[units[h.index("dram__bytes_read.sum")]]
            wr = float(d["dram__bytes_write.sum"]) * scale[units[h.index("dram__bytes_write.sum")]]
            out[key] = {"dram_bytes_per_pose": (rd + wr) / poses, "read_per_pose": rd / poses, "write_per_pose": wr / poses, "source": label,
                        "kernel_ms_under_ncu": float(d["gpu__time_duration.sum"]), "poses_in_launch": poses}
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))

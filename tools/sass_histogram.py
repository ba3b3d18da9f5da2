"""cuobjdump -sass opcode histogram of the hot kernels of libgnina_b200.so -> profiles/<tag>_sass_histogram.txt
(the evidence for tcgen05 / TMA: UTCHMMA, LDTM / STTM, UTCBAR, UTMALDG, UBLKCP; legacy tensor path: HMMA).
  python tools/sass_histogram.py r2"""
import collections, os, re, subprocess, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
tag = sys.argv[1] if len(sys.argv) > 1 else "rX"
lib = os.path.join(ROOT, "gnina_b200", "libgnina_b200.so")
sass = subprocess.check_output(["cuobjdump", "-sass", lib], text=True)
KEEP = ("conv1_pw2_pool", "conv3_tc_kernel", "dense_conv_tc", "voxelize_pool", "pointwise_pool_mma", "pw_backward", "bottleneck",
        "dock_mc_kernel", "dock_eval_kernel", "dock_refine_kernel", "fc_heads", "cache_populate")
MARK = ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UBLKCP", "HMMA", "FFMA2", "FMUL2", "FADD2", "MUFU", "SYNCS", "ATOMS", "DFMA")
out, cur, hist = [], None, None
def flush():
    if cur and hist:
        tot = sum(hist.values())
        marks = "  ".join("%s=%d" % (m, sum(v for k, v in hist.items() if k.startswith(m))) for m in MARK if any(k.startswith(m) for k in hist))
        out.append("%s\n  instructions %d | %s\n  top: %s\n" % (cur, tot, marks, ", ".join("%s %d" % kv for kv in hist.most_common(14))))
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        flush()
        name = subprocess.check_output(["c++filt", m.group(1)], text=True).strip()
        cur = name[:150] if any(k in name for k in KEEP) else None
        hist = collections.Counter()
        continue
    if cur:
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", line)
        if m:
            hist[m.group(1).split(".")[0] if not m.group(1).startswith(("UTC", "LDTM", "STTM", "UBLKCP", "UTMA")) else m.group(1)] += 1
flush()
path = os.path.join(ROOT, "profiles", "%s_sass_histogram.txt" % tag)
open(path, "w").write("cuobjdump -sass gnina_b200/libgnina_b200.so (sm_100a), opcode histograms of the hot kernels\n\n" + "\n".join(out))
print(path, len(out), "kernels")

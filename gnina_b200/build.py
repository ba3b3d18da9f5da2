"""Builds libgnina_b200.so in-tree with nvcc for sm_100a (the .so travels to the GPU box with the snapshot)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgnina_b200.so")
SOURCES = ["gb_model.cu", "gb_grid.cu", "gb_cnn_fp32.cu", "gb_cnn_tc.cu", "gb_cnn_tc_fused.cu", "gb_vox_tc.cu", "gb_cnn_tc_dense.cu", "gb_cnn_tc_grad.cu", "gb_capi.cu", "gb_vina.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "--expt-relaxed-constexpr", "-Xptxas", "-v"]


# gb_vina.cu restates float code whose sums and products must round like the reference's sequential C++ (compiled
# without FMA contraction): no fused multiply-add anywhere in that file
PER_FILE_FLAGS = {"gb_vina.cu": ["-fmad=false"]}


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "gnina_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(HERE, "build", src + ".o")
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + PER_FILE_FLAGS.get(src, []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, p in procs:
        out, _ = p.communicate()
        log.append("== %s ==\n%s" % (src, out))
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError("nvcc failed on " + src)
    with open(os.path.join(HERE, "build", "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    subprocess.check_call([nvcc, "-shared", "-o", LIB] + objs + ["-lcudart", "-lcuda"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

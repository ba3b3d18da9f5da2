#!/usr/bin/env python3
"""Spec freeze: convert the reference's embedded TorchScript models into flat GNB200W1 blobs.

The reference embeds `gninasrc/lib/models/*.pt` into its binary (gninasrc/CMakeLists.txt:95-188,
gninasrc/make_model_cpp.py:1-55) and executes them with torch::jit (gninasrc/lib/torch_model.cpp:55,185).
Weights + JSON metadata (torch_model.cpp:53-106) are the *specification* of the CNN path; this script
reads them HERE (container with /root/reference) and writes `gnina_b200/weights/<name>.gbw`, which is
what travels with the repo.  No libtorch is needed at run time on the product path.

File layout (little endian):
  char  magic[8]  = "GNB200W1"
  u32   arch (1 default2018, 2 dense, 3 default2017, 4 overlap), u32 n_tensors
  f32   resolution, dimension, radius_scaling ; u32 flags (bit0 apply_logistic_loss, bit1 skip_softmax)
  u32   name_len, recmap_len, ligmap_len, reserved
  bytes name, recmap, ligmap ; zero pad to 8
  n_tensors x { char name[96]; u32 ndim; u32 dims[6]; u32 pad; u64 offset; u64 nelem }
  fp32 tensor data, each 64-byte aligned (offset from file start)
"""
import argparse, json, os, struct, sys
import numpy as np
import torch

ARCH = {"default2018": 1, "dense": 2, "default2017": 3, "overlap": 4}

DEFAULT_RECMAP = """AliphaticCarbonXSHydrophobe 
AliphaticCarbonXSNonHydrophobe 
AromaticCarbonXSHydrophobe 
AromaticCarbonXSNonHydrophobe
Bromine Iodine Chlorine Fluorine
Nitrogen NitrogenXSAcceptor 
NitrogenXSDonor NitrogenXSDonorAcceptor
Oxygen OxygenXSAcceptor 
OxygenXSDonorAcceptor OxygenXSDonor
Sulfur SulfurAcceptor
Phosphorus 
Calcium
Zinc
GenericMetal Boron Manganese Magnesium Iron
"""
DEFAULT_LIGMAP = """AliphaticCarbonXSHydrophobe 
AliphaticCarbonXSNonHydrophobe 
AromaticCarbonXSHydrophobe 
AromaticCarbonXSNonHydrophobe
Bromine Iodine
Chlorine
Fluorine
Nitrogen NitrogenXSAcceptor 
NitrogenXSDonor NitrogenXSDonorAcceptor
Oxygen OxygenXSAcceptor 
OxygenXSDonorAcceptor OxygenXSDonor
Sulfur SulfurAcceptor
Phosphorus
GenericMetal Boron Manganese Magnesium Zinc Calcium Iron
"""


def classify(sd):
    keys = list(sd.keys())
    if any("dense_block_0" in k for k in keys):
        return "dense"
    if any(k.endswith("unit5_conv.weight") for k in keys):
        return "default2018"
    if any(k.endswith("unit3_conv1.weight") for k in keys):
        return "default2017"
    if not keys:
        return "overlap"   # the parameter-free overlay graph of test/gnina/data/overlap*.pt (test_min.py)
    raise ValueError("unknown architecture: %s" % keys[:5])


def write_gbw(path, name, arch, meta, tensors):
    recmap = meta.get("recmap", DEFAULT_RECMAP).encode()
    ligmap = meta.get("ligmap", DEFAULT_LIGMAP).encode()
    nm = name.encode()
    flags = (1 if meta.get("apply_logistic_loss", False) else 0) | (2 if meta.get("skip_softmax", False) else 0)
    hdr = b"GNB200W1" + struct.pack("<II", ARCH[arch], len(tensors))
    hdr += struct.pack("<fffI", float(meta.get("resolution", 0.5)), float(meta.get("dimension", 23.5)),
                       float(meta.get("radius_scaling", 1.0)), flags)
    hdr += struct.pack("<IIII", len(nm), len(recmap), len(ligmap), 0)
    hdr += nm + recmap + ligmap
    hdr += b"\0" * ((-len(hdr)) % 8)
    entry = 96 + 4 + 24 + 4 + 8 + 8
    off = len(hdr) + entry * len(tensors)
    table, blobs = b"", []
    for tname, arr in tensors:
        arr = np.ascontiguousarray(arr, dtype="<f4")
        off += (-off) % 64
        dims = list(arr.shape) + [0] * (6 - arr.ndim)
        table += tname.encode().ljust(96, b"\0") + struct.pack("<I6IIQQ", arr.ndim, *dims, 0, off, arr.size)
        blobs.append((off, arr.tobytes()))
        off += arr.nbytes
    with open(path, "wb") as f:
        f.write(hdr + table)
        for o, b in blobs:
            f.seek(o)
            f.write(b)


def convert(src, dst_dir):
    ef = {"metadata": ""}
    m = torch.jit.load(src, map_location="cpu", _extra_files=ef)
    meta = json.loads(ef["metadata"]) if ef["metadata"] else {}
    sd = m.state_dict()
    arch = classify(sd)
    # canonical tensor names: the same architecture appears with and without the
    # "features." / "pose." / "affinity." sub-module prefixes (older vs 1.3 exports)
    def canon(k):
        for pre in ("features.", "pose.", "affinity."):
            if k.startswith(pre):
                k = k[len(pre):]
        return k.replace(".blocks.", ".")   # the pre-1.3 dense exports nest the block layers one level deeper
    tensors = [(canon(k), v.detach().numpy()) for k, v in sd.items() if v.dtype == torch.float32]
    # gnina model names replace '.' by '_' (gninasrc/make_model_cpp.py:31-32)
    name = os.path.basename(src)[:-3].replace(".", "_")
    out = os.path.join(dst_dir, name + ".gbw")
    write_gbw(out, name, arch, meta, tensors)
    return name, arch, os.path.getsize(out)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference/gninasrc/lib/models")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(__file__), "..", "gnina_b200", "weights"))
    ap.add_argument("models", nargs="*", default=[
        "crossdock_default2018", "dense_1.3", "dense_1.3_PT_KD_3", "crossdock_default2018_KD_4",
        "all_default_to_default_1.3_1", "default2017"])
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    for mname in a.models:
        print(convert(os.path.join(a.ref, mname + ".pt"), a.out))

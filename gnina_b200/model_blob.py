"""Reader for GNB200W1 weight blobs (written by tools/extract_models.py).

The blob carries what the reference reads from each embedded TorchScript model: the fp32 parameters and the
JSON metadata keys resolution / dimension / recmap / ligmap / apply_logistic_loss / skip_softmax /
radius_scaling (gninasrc/lib/torch_model.cpp:53-106).
"""
import os
import struct
import numpy as np

ARCH_NAMES = {1: "default2018", 2: "dense", 3: "default2017", 4: "overlap"}
WEIGHTS_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "weights")


class ModelBlob:
    def __init__(self, path):
        self.path = path
        with open(path, "rb") as f:
            raw = f.read()
        if raw[:8] != b"GNB200W1":
            raise ValueError("%s: not a GNB200W1 blob" % path)
        arch, nt = struct.unpack_from("<II", raw, 8)
        self.resolution, self.dimension, self.radius_scaling, flags = struct.unpack_from("<fffI", raw, 16)
        nl, rl, ll, _ = struct.unpack_from("<IIII", raw, 32)
        p = 48
        self.name = raw[p:p + nl].decode(); p += nl
        self.recmap = raw[p:p + rl].decode(); p += rl
        self.ligmap = raw[p:p + ll].decode(); p += ll
        p += (-p) % 8
        self.arch = ARCH_NAMES[arch]
        self.apply_logistic_loss = bool(flags & 1)
        self.skip_softmax = bool(flags & 2)
        self.tensors = {}
        for _ in range(nt):
            tname = raw[p:p + 96].split(b"\0")[0].decode()
            ndim, *rest = struct.unpack_from("<I6IIQQ", raw, p + 96)
            dims, off, nelem = rest[:6], rest[7], rest[8]
            shape = tuple(dims[:ndim])
            self.tensors[tname] = np.frombuffer(raw, dtype="<f4", count=nelem, offset=off).reshape(shape)
            p += 96 + 4 + 24 + 4 + 8 + 8


def model_path(name):
    """Built-in model name (gnina spelling, '.' replaced by '_') -> blob path."""
    p = os.path.join(WEIGHTS_DIR, name.replace(".", "_") + ".gbw")
    if not os.path.exists(p):
        raise FileNotFoundError("Invalid model name: " + name)  # cnn_torch_scorer.cpp:70-72 usage_error
    return p


def load_model(name_or_path):
    if os.path.exists(name_or_path):
        return ModelBlob(name_or_path)
    return ModelBlob(model_path(name_or_path))

"""Helpers for the fast-path parity tests: decode the device's chunk-planar padded activation layout and compute
the oracle's intermediate activations of the default2018 graph."""
import numpy as np
import torch
import torch.nn.functional as F


def layout(D, G, C):
    P = D + 2
    span = (G - 1) * P * P + (D - 1) * P + D
    T = (span + 127) // 128
    Lp = (128 * T + 2 * (P + 1) + 7) & ~7
    return dict(D=D, P=P, G=G, T=T, C8=C // 8, Lp=Lp)


def decode_chunk_planar(buf, n_poses, D, G, C):
    """raw fp16 buffer -> [n_poses][C][D][D][D] float32 (interior only) + max |border| value"""
    L = layout(D, G, C)
    ng = (n_poses + G - 1) // G
    a = buf[: ng * D * L["C8"] * L["Lp"] * 8].reshape(ng, D, L["C8"], L["Lp"], 8).astype(np.float32)
    P = L["P"]
    out = np.zeros((ng * G, C, D, D, D), np.float32)
    border = 0.0
    for q in range(G):
        blk = a[:, :, :, q * P * P:(q + 1) * P * P, :].reshape(ng, D, L["C8"], P, P, 8)
        inner = blk[:, :, :, 1:D + 1, 1:D + 1, :]                       # [g][x][c8][y][z][8]
        out[q::G] = inner.transpose(0, 2, 5, 1, 3, 4).reshape(ng, C, D, D, D)
        b = blk.copy()
        b[:, :, :, 1:D + 1, 1:D + 1, :] = 0
        border = max(border, float(np.abs(b).max()))
    return out[:n_poses], border


def decode_channels_last(buf, n_poses, D, C):
    return buf[: n_poses * D ** 3 * C].reshape(n_poses, D, D, D, C).astype(np.float32).transpose(0, 4, 1, 2, 3)


def oracle_intermediates(blob, grid):
    """default2018 graph (oracle/cnn_ref.py) with every intermediate kept, float64."""
    w = lambda n: torch.from_numpy(np.array(blob.tensors[n])).double()
    x = torch.from_numpy(np.asarray(grid)).double()
    r = {}
    with torch.no_grad():
        r["x0"] = F.avg_pool3d(x, 2, 2)
        y1 = F.relu(F.conv3d(r["x0"], w("unit1_conv.weight"), w("unit1_conv.bias"), padding=1))
        r["y1"] = y1
        y2 = F.relu(F.conv3d(y1, w("unit2_conv.weight"), w("unit2_conv.bias")))
        r["x2"] = F.avg_pool3d(y2, 2, 2)
        r["y3"] = F.relu(F.conv3d(r["x2"], w("unit3_conv.weight"), w("unit3_conv.bias"), padding=1))
        y4 = F.relu(F.conv3d(r["y3"], w("unit4_conv.weight"), w("unit4_conv.bias")))
        r["x4"] = F.avg_pool3d(y4, 2, 2)
        r["y5"] = F.relu(F.conv3d(r["x4"], w("unit5_conv.weight"), w("unit5_conv.bias"), padding=1))
    return {k: v.numpy() for k, v in r.items()}

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


def oracle_intermediates_dense(blob, grid):
    """dense graph (oracle/cnn_ref.py) with the block buffers kept, float64: x0 (max-pooled grid), b0 [96ch @24^3],
    b1 [160 @12^3], b2 [224 @6^3], feat [224]"""
    w = lambda n: torch.from_numpy(np.array(blob.tensors[n])).double()
    x = torch.from_numpy(np.asarray(grid)).double()

    def block(x, level):
        for i in range(4):
            bn = "dense_block_%d.data_enc_level%d_batchnorm_conv%d." % (level, level, i)
            cv = "dense_block_%d.data_enc_level%d_conv%d." % (level, level, i)
            y = F.batch_norm(x, w(bn + "running_mean"), w(bn + "running_var"), w(bn + "weight"), w(bn + "bias"), False, 0.1, 1e-5)
            x = torch.cat([x, F.relu(F.conv3d(y, w(cv + "weight"), w(cv + "bias"), padding=1))], 1)
        return x

    r = {}
    with torch.no_grad():
        r["x0"] = F.max_pool3d(x, 2, 2)
        y = F.relu(F.conv3d(r["x0"], w("data_enc_init_conv.weight"), w("data_enc_init_conv.bias"), padding=1))
        r["b0"] = block(y, 0)
        y = F.max_pool3d(F.relu(F.conv3d(r["b0"], w("data_enc_level0_bottleneck.weight"), w("data_enc_level0_bottleneck.bias"))), 2, 2)
        r["b1"] = block(y, 1)
        y = F.max_pool3d(F.relu(F.conv3d(r["b1"], w("data_enc_level1_bottleneck.weight"), w("data_enc_level1_bottleneck.bias"))), 2, 2)
        r["b2"] = block(y, 2)
        r["feat"] = F.max_pool3d(r["b2"], 6).reshape(x.shape[0], -1)
    return {k: v.numpy() for k, v in r.items()}

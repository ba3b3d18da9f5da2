"""CPU restatement of the reference's CNN forward (TEST INFRASTRUCTURE ONLY).

What the reference executes is the embedded TorchScript graph (gninasrc/lib/torch_model.cpp:185) followed by
the head post-processing at torch_model.cpp:188-195.  The graphs were read from the `.pt` files themselves
(`torch.jit.load(...).code`); this file restates the three architectures with plain torch CPU ops on the
weights in gnina_b200/weights/*.gbw:

  default2018: avgpool2 -> conv3(28->32,p1)+ReLU -> conv1(32->32)+ReLU -> avgpool2 -> conv3(32->64)+ReLU
               -> conv1(64->64)+ReLU -> avgpool2 -> conv3(64->128)+ReLU -> flatten(NCDHW, 27648)
               -> {Linear->2 -> log_softmax ; Linear->1}
  default2017: maxpool2 -> conv3(35->32)+ReLU -> maxpool2 -> conv3(32->64)+ReLU -> maxpool2 -> conv3(64->128)+ReLU
               -> flatten -> heads
  dense:       maxpool2 -> conv3(28->32)+ReLU -> DB0 -> conv1(96->96)+ReLU -> maxpool2 -> DB1
               -> conv1(160->160)+ReLU -> maxpool2 -> DB2 -> global maxpool -> 224 -> heads
               DBk = 4 x [BatchNorm3d(eval, eps 1e-5) -> conv3(Cin->16,p1) -> ReLU -> concat]

PINNED: tests/golden/cnn_kat.npz holds outputs of the reference's own `.pt` files (generated here by
tests/golden/make_golden.py, which imports them from /root/reference); tests/test_oracle_cnn.py checks this
restatement against them.  dtype float64 gives the tie-breaking value, float32 "what the reference CPU path
computes" up to libtorch-version round-off (the reference pins libtorch 2.4.1, CMakeLists.txt:89).
"""
import numpy as np
import torch
import torch.nn.functional as F


def _t(blob, name, dtype):
    # converted once per (blob, tensor, dtype): a CPU baseline must not re-convert 1.5 MB of weights per call
    cache = blob.__dict__.setdefault("_torch_cache", {})
    key = (name, dtype)
    if key not in cache:
        cache[key] = torch.from_numpy(np.array(blob.tensors[name])).to(dtype)
    return cache[key]


def features_default2018(blob, x, dtype):
    w = lambda n: _t(blob, n, dtype)
    x = F.avg_pool3d(x, 2, 2)
    x = F.relu(F.conv3d(x, w("unit1_conv.weight"), w("unit1_conv.bias"), padding=1))
    x = F.relu(F.conv3d(x, w("unit2_conv.weight"), w("unit2_conv.bias")))
    x = F.avg_pool3d(x, 2, 2)
    x = F.relu(F.conv3d(x, w("unit3_conv.weight"), w("unit3_conv.bias"), padding=1))
    x = F.relu(F.conv3d(x, w("unit4_conv.weight"), w("unit4_conv.bias")))
    x = F.avg_pool3d(x, 2, 2)
    x = F.relu(F.conv3d(x, w("unit5_conv.weight"), w("unit5_conv.bias"), padding=1))
    return x.reshape(x.shape[0], -1)


def features_default2017(blob, x, dtype):
    w = lambda n: _t(blob, n, dtype)
    x = F.max_pool3d(x, 2, 2)
    x = F.relu(F.conv3d(x, w("unit1_conv1.weight"), w("unit1_conv1.bias"), padding=1))
    x = F.max_pool3d(x, 2, 2)
    x = F.relu(F.conv3d(x, w("unit2_conv1.weight"), w("unit2_conv1.bias"), padding=1))
    x = F.max_pool3d(x, 2, 2)
    x = F.relu(F.conv3d(x, w("unit3_conv1.weight"), w("unit3_conv1.bias"), padding=1))
    return x.reshape(x.shape[0], -1)


def _dense_block(blob, x, level, dtype):
    w = lambda n: _t(blob, n, dtype)
    for i in range(4):
        bn = "dense_block_%d.data_enc_level%d_batchnorm_conv%d." % (level, level, i)
        cv = "dense_block_%d.data_enc_level%d_conv%d." % (level, level, i)
        y = F.batch_norm(x, w(bn + "running_mean"), w(bn + "running_var"), w(bn + "weight"), w(bn + "bias"),
                         False, 0.1, 1e-5)
        y = F.relu(F.conv3d(y, w(cv + "weight"), w(cv + "bias"), padding=1))
        x = torch.cat([x, y], 1)
    return x


def features_dense(blob, x, dtype):
    w = lambda n: _t(blob, n, dtype)
    x = F.max_pool3d(x, 2, 2)
    x = F.relu(F.conv3d(x, w("data_enc_init_conv.weight"), w("data_enc_init_conv.bias"), padding=1))
    x = _dense_block(blob, x, 0, dtype)
    x = F.relu(F.conv3d(x, w("data_enc_level0_bottleneck.weight"), w("data_enc_level0_bottleneck.bias")))
    x = F.max_pool3d(x, 2, 2)
    x = _dense_block(blob, x, 1, dtype)
    x = F.relu(F.conv3d(x, w("data_enc_level1_bottleneck.weight"), w("data_enc_level1_bottleneck.bias")))
    x = F.max_pool3d(x, 2, 2)
    x = _dense_block(blob, x, 2, dtype)
    x = F.max_pool3d(x, x.shape[2:])
    return x.reshape(x.shape[0], -1)


FEATURES = {"default2018": features_default2018, "default2017": features_default2017, "dense": features_dense}


def forward_logits(blob, grid, dtype=torch.float32, grad=False):
    """grid [B,C,N,N,N] (numpy or tensor) -> (log-softmax pose [B,2], affinity [B]) exactly as the TorchScript
    module returns them.  grad=True keeps the autograd graph (grid may then be a leaf tensor requiring grad)."""
    x = grid.to(dtype) if torch.is_tensor(grid) else torch.as_tensor(np.asarray(grid)).to(dtype)
    if blob.arch == "overlap":
        # test/gnina/data/overlap.pt (the model of test_min.py), restated from its TorchScript code: out[0] is NOT a
        # log-softmax here -- the metadata sets skip_softmax and apply_logistic_loss
        with torch.set_grad_enabled(grad):
            ave = F.avg_pool3d(x[:, 0] * x[:, 1], x.shape[-1]).flatten(1)
            ave = torch.where(ave > 0, ave, torch.full_like(ave, 1e-20))
            return torch.hstack([torch.zeros_like(ave), ave]), torch.zeros(x.shape[0], dtype=dtype)
    with torch.set_grad_enabled(grad):
        f = FEATURES[blob.arch](blob, x, dtype)
        pose = F.linear(f, _t(blob, "pose_output.weight", dtype), _t(blob, "pose_output.bias", dtype))
        aff = F.linear(f, _t(blob, "affinity_output.weight", dtype), _t(blob, "affinity_output.bias", dtype))
    return F.log_softmax(pose, 1), aff.squeeze(-1)


def head_post(blob, pose_out, aff_out):
    """torch_model.cpp:188-195: pose = softmax(model output)[:,1] unless skip_softmax; affinity = out[1];
    loss = CE(model output, label 1) or -log(out[:,1]) when apply_logistic_loss."""
    if blob.skip_softmax:
        pose = pose_out[:, 1]
    else:
        pose = torch.softmax(pose_out, 1)[:, 1]
    if blob.apply_logistic_loss:
        loss = -torch.log(pose_out[:, 1])
    else:
        loss = F.cross_entropy(pose_out, torch.ones(pose_out.shape[0], dtype=torch.long), reduction="none")
    return pose, aff_out, loss


def score_grid(blob, grid, dtype=torch.float32):
    """-> numpy (pose[B], affinity[B], loss[B])"""
    p, a = forward_logits(blob, grid, dtype)
    pose, aff, loss = head_post(blob, p, a)
    return pose.numpy(), aff.numpy(), loss.numpy()


def loss_grid_gradient(blob, grid, dtype=torch.float64):
    """torch_model.cpp:195-199: loss = CE(module output, label 1) (summed over the batch: poses are independent),
    backward to the grid.  -> (pose[B], affinity[B], loss[B], dloss/dgrid [B,C,N,N,N])"""
    g = torch.as_tensor(np.asarray(grid)).to(dtype).requires_grad_(True)
    lp, aff = forward_logits(blob, g, dtype, grad=True)
    pose, aff, loss = head_post(blob, lp, aff)
    loss.sum().backward()
    return pose.detach().numpy(), aff.detach().numpy(), loss.detach().numpy(), g.grad.numpy()


def ensemble(scores, affinities, losses):
    """CNNTorchScorer::score accumulation (gninasrc/lib/cnn_torch_scorer.cpp:117-192) for ONE pose:
    mean score (double accumulator), mean affinity/loss (float), population variance of affinities."""
    cnt = len(scores)
    score = float(np.sum(np.asarray(scores, np.float64)) / cnt)
    aff = np.float32(0)
    loss = np.float32(0)
    for a, l in zip(affinities, losses):
        aff = np.float32(aff + np.float32(a)); loss = np.float32(loss + np.float32(l))
    aff = np.float32(aff / np.float32(cnt)); loss = np.float32(loss / np.float32(cnt))
    var = np.float32(0)
    if cnt > 1:
        s = np.float32(0)
        for a in affinities:
            d = np.float32(aff - np.float32(a)); s = np.float32(s + d * d)
        var = np.float32(s / np.float32(cnt))
    return score, float(aff), float(loss), float(var)

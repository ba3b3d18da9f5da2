import os, sys, time, numpy as np
sys.path.insert(0, '/root/repo')
from gnina_b200 import CNNScorer, synth
kat = np.load('/root/repo/tests/golden/cnn_kat.npz')
g = np.load('/root/repo/tests/golden/grad_kat.npz')
n = int(g["n_poses"]); offs = kat["pose_offsets"][:n + 1]
x, t = kat["lig_xyz"][:offs[-1]], kat["lig_types"][:offs[-1]]
for prec in (0, 1):
    s = CNNScorer([str(g["model"])], precision=prec)
    s.set_receptor(kat["rec_xyz"], kat["rec_types"])
    out = s.score_grad_batch(x, t, offs)
    scale = np.abs(g["lig_grad"]).max()
    print('prec', prec, 'loss err', np.abs(out[2] - g["loss"]).max(), 'grad err/scale', np.abs(out[4] - g["lig_grad"]).max() / scale, 'scale', scale)
    if prec == 1:
        print(out[4][:6]); print(g["lig_grad"][:6])
# throughput
xyz, types, po = synth.make_screen(4096, seed=5, trans_box=6.0)
rec_xyz, rec_t = synth.make_receptor(3000)
for prec, nn in ((0, 256), (1, 4096)):
    s = CNNScorer(["crossdock_default2018"], precision=prec)
    s.set_receptor(rec_xyz, rec_t)
    s.score_grad_batch(xyz[:po[nn]], types[:po[nn]], po[:nn + 1])
    t0 = time.time(); s.score_grad_batch(xyz[:po[nn]], types[:po[nn]], po[:nn + 1]); dt = time.time() - t0
    print('prec', prec, 'grad poses/s', nn / dt)
s.set_option("profile", 1)
s.score_grad_batch(xyz, types, po)
for k, v in s.profile().items(): print(k, v)

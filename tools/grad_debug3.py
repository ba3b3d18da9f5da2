import sys, numpy as np
sys.path.insert(0, '/root/repo')
from gnina_b200 import CNNScorer, synth
kat = np.load('/root/repo/tests/golden/cnn_kat.npz')
xyz, types, offs = synth.make_screen(301, seed=5, trans_box=6.0)
names = ["crossdock_default2018", "crossdock_default2018_KD_4"]
for nm in (names[:1], names[1:], names):
    ref = CNNScorer(nm, precision=0); fast = CNNScorer(nm, precision=1, max_batch=128)
    for s in (ref, fast): s.set_receptor(kat["rec_xyz"], kat["rec_types"])
    a = ref.score_grad_batch(xyz, types, offs); b = fast.score_grad_batch(xyz, types, offs)
    ga, gb = a[4], b[4]
    print(nm, "finite", np.isfinite(gb).all(), "scale", np.abs(ga).max(), "max diff", np.abs(ga - gb).max(), "loss diff", np.abs(a[2] - b[2]).max())
    per = np.array([np.abs(ga[offs[i]:offs[i+1]] - gb[offs[i]:offs[i+1]]).max() for i in range(301)])
    mag = np.array([np.abs(ga[offs[i]:offs[i+1]]).max() for i in range(301)])
    o = np.argsort(-per)[:6]
    print("  worst poses", o, "diff", per[o], "pose |g|max", mag[o], "loss ref", a[2][o], "fast", b[2][o])
    print("  median rel (diff/pose gmax)", np.median(per / np.maximum(mag, 1e-9)), "p99", np.percentile(per / np.maximum(mag, 1e-9), 99))

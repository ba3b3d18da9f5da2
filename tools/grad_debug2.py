import sys, numpy as np
sys.path.insert(0, '/root/repo')
from gnina_b200 import CNNScorer, synth
kat = np.load('/root/repo/tests/golden/cnn_kat.npz')
xyz, types, offs = synth.make_screen(301, seed=5, trans_box=6.0)
def run(s, k):
    return s.score_grad_batch(xyz[:offs[k]], types[:offs[k]], offs[:k + 1])
for names in (["crossdock_default2018"], ["crossdock_default2018", "crossdock_default2018_KD_4"]):
    print("models", len(names))
    s = CNNScorer(names, precision=1, max_batch=128)
    s.set_receptor(kat["rec_xyz"], kat["rec_types"])
    a = run(s, 301); b = run(s, 301)
    print(" same call twice: dgrad", np.abs(a[4] - b[4]).max(), "dloss", np.abs(a[2] - b[2]).max())
    for k in (7, 8, 128, 129):
        c = run(s, k)
        n = offs[k]
        print(" k=%d vs 301: dgrad %.3g dloss %.3g dscore %.3g" % (k, np.abs(c[4] - a[4][:n]).max(), np.abs(c[2] - a[2][:k]).max(), np.abs(c[0] - a[0][:k]).max()))
    s2 = CNNScorer(names, precision=1, max_batch=128)
    s2.set_receptor(kat["rec_xyz"], kat["rec_types"])
    c = run(s2, 7)
    print(" fresh handle k=7 vs 301: dgrad %.3g" % np.abs(c[4] - a[4][:offs[7]]).max())
    d = np.abs(c[4] - a[4][:offs[7]]).max(1)
    print("  per-pose max diff", [float(d[offs[i]:offs[i + 1]].max()) for i in range(7)])

import sys, os, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import test_fullsize_gpu as T
from neurips18_hierchical_image_manipulation_b200.synthetic import synthetic_batch
m = T._model()
b = synthetic_batch(2, T.H, T.W, 35, seed=99)
l2, g2 = T._grads(m, b)
l2b, g2b = T._grads(m, b)
print("repeat same batch: max rel diff", float((g2 - g2b).abs().max() / g2.abs().max()))
acc = None
for i in range(2):
    bi = {k: v[i:i + 1] for k, v in b.items()}
    li, gi = T._grads(m, bi)
    acc = gi if acc is None else acc + gi
acc /= 2
worst = []
for fp, base in ((m.fpG, 0), (m.fpD, m.fpG.total)):
    for name, shape, off in fp.specs:
        n = 1
        for s in shape: n *= s
        a, c = g2[base + off: base + off + n], acc[base + off: base + off + n]
        worst.append((float((a - c).abs().max() / a.abs().max().clamp_min(1e-30)), float(a.abs().max()), name))
worst.sort(reverse=True)
for w in worst[:12]: print(w)

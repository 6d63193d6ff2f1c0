"""Diagnostic (CPU only): how sensitive are the box2mask generator's parameter gradients to forward rounding of the size
the bf16x3 engines have?  The float64 oracle is evaluated with absolute noise of 1e-5 / 3e-5 x max|x| added to every ReLU
input (what a GEMM's rounding does to a pre-activation) and its gradients are compared with the noise-free ones.
Result (golden weights, the two test batches): per-tensor max-norm changes of 5e-2 ... 4e-1, 2-norm 1e-2 ... 6e-2 -- larger
than the product's deviation from the oracle (<= 5e-2 / <= 9e-3), i.e. that deviation is ReLU decision flips, not
arithmetic.  usage: python tools/diag_b2m_sensitivity.py"""
import sys, torch
import os
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
from oracle import box2mask as B2
from oracle.weights import named_param
import make_golden_box2mask as G
import numpy as np
import torch.nn.functional as F
z = np.load(os.path.join(ROOT, 'tests', 'golden', 'box2mask_small.npz'))
names = [str(n) for n in z["param_names"]]; shapes = [tuple(int(v) for v in str(s).split(";")) for s in z["param_shapes"]]
orig_relu = F.relu
for seed in (17, 53):
    d = G.synthetic(dict(label_nc=6, fineSize=64), 3, seed=seed)
    cond, _ = B2.encode_input(6, d["mask_ctx_in"], d["mask_in"], d["cls"])
    res = {}
    for noise in (0.0, 1e-5, 3e-5):
        torch.manual_seed(0)
        def relu(x, inplace=False, _n=noise):
            if _n:
                x = x + (_n * float(x.detach().abs().max())) * torch.randn_like(x)   # absolute noise like a GEMM's rounding
            return orig_relu(x)
        B2.F.relu = relu
        sd = {n: named_param(n, s).double().requires_grad_(True) for n, s in zip(names, shapes)}
        _, lp, _, op_ = B2.two_stream_forward(sd, cond.double(), num_layers=3, n_blocks=2)
        lc = B2.mask_recon_loss(lp, d["label_map"], d["mask_out"])
        lo = B2.obj_recon_loss(op_, d["mask_out"].double(), d["mask_obj_inst"].double())
        res[noise] = torch.autograd.grad(lo + lc, list(sd.values()))
    B2.F.relu = orig_relu
    for noise in (1e-5, 3e-5):
        worst = sorted(((float((a - b).abs().max() / b.abs().max()), float((a-b).norm()/b.norm()), n)
                        for a, b, n in zip(res[noise], res[0.0], names) if float(b.abs().max()) > 1e-6), reverse=True)
        print("seed", seed, "noise %.0e at every ReLU input: worst" % noise, ["%.1e %.1e %s" % w for w in worst[:4]])

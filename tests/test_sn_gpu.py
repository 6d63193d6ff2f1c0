"""GPU parity tests of K13, the opt-in spectral normalisation of the PatchGAN convolutions (csrc/hm_sn.cu):
hm_sn_power_iteration / hm_sn_weight_grad / the 1/sigma scale fused into hm_pack_weight_ex, against the golden vectors
generated from the reference's own max_singular_value and SNConv2d (tests/golden/sn.npz, sn_conv.npz), and the full
training step with sn_D=True against the oracle."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = torch.as_tensor(a).float().cpu(), torch.as_tensor(b).float().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _layers(ctx, Ws, grads, us):
    from neurips18_hierchical_image_manipulation_b200 import _lib as L
    arr = (L.SnLayer * len(Ws))()
    stashes = []
    mn = mm = 0
    for i, (W, g, u) in enumerate(zip(Ws, grads, us)):
        n, m = W.shape[0], W[0].numel()
        st = torch.zeros(ctx.lib.hm_sn_stash_floats(n, m), device="cuda")
        stashes.append(st)
        arr[i] = L.SnLayer(W.data_ptr(), g.data_ptr(), u.data_ptr(), st.data_ptr(), n, m)
        mn, mm = max(mn, n), max(mm, m)
    dev = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).cuda()
    return dev, stashes, mn, mm


def test_power_iteration_matches_reference_golden(golden_dir):
    from neurips18_hierchical_image_manipulation_b200 import ops
    ctx = ops.Ctx("cuda:0", split=True)
    z = np.load(os.path.join(golden_dir, "sn.npz"))
    z2 = np.load(os.path.join(golden_dir, "sn_conv.npz"))
    Ws = [torch.from_numpy(z["W"]).cuda(), torch.from_numpy(z2["W"]).cuda()]
    us = [torch.from_numpy(z["u"]).cuda().clone(), torch.from_numpy(z2["u"]).cuda().clone()]
    gs = [torch.zeros_like(w) for w in Ws]
    dev, st, mn, mm = _layers(ctx, Ws, gs, us)
    ops.sn_power_iteration(ctx, dev, 2, mn, mm, update_u=True)
    torch.cuda.synchronize()
    for W, u, s, sig, uo in ((Ws[0], us[0], st[0], z["sigma"], z["u_out"]), (Ws[1], us[1], st[1], z2["sigma"], z2["u_out"])):
        n, m = W.shape[0], W[0].numel()
        assert rel(s[2 * n + m], sig.reshape(())) < 1e-5
        assert rel(s[2 * n + m + 1], 1.0 / sig.reshape(())) < 1e-5
        assert rel(u, uo) < 1e-5


@pytest.mark.parametrize("split", [True, False], ids=["x3", "x1"])
def test_sn_conv_forward_backward_matches_reference_snconv2d(golden_dir, split):
    """SNConv2d(6,10,4,p=2): y, dW (through sigma), db, dx against the reference's autograd."""
    from neurips18_hierchical_image_manipulation_b200 import ops
    from neurips18_hierchical_image_manipulation_b200.networks import ConvP, FlatParams
    tol = 2e-4 if split else 3e-2
    ctx = ops.Ctx("cuda:0", split=split)
    z = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(golden_dir, "sn_conv.npz")).items()}
    fp = FlatParams(ctx.device)
    conv = ConvP(ctx, fp, "c", 6, 10, 4, 1, 2)
    fp.materialize()
    fp.load_state_dict({"c.weight": z["W"], "c.bias": z["b"]})
    u = z["u"].cuda().clone()
    dev, st, mn, mm = _layers(ctx, [conv.weight], [conv.weight.grad], [u])
    n, m = 10, 6 * 16
    conv.sn_scale = st[0][2 * n + m + 1:2 * n + m + 2]
    ops.sn_power_iteration(ctx, dev, 1, mn, mm, update_u=True)
    x = z["x"]
    xin = ops.Operand(ctx, 2, 9, 13, 6)
    ops.in_apply(ctx, x.permute(0, 2, 3, 1).contiguous().cuda(), None, None, ops.ACT_NONE, out_op=xin, reflect=False)
    ho, wo = conv.out_hw(9, 13, 2)
    y = torch.empty(2, ho, wo, 10, device="cuda")
    conv.forward(xin, 2, out32=y)
    dy = ops.Operand(ctx, 2, ho, wo, 10, grad=True)
    ops.in_apply(ctx, z["g"].permute(0, 2, 3, 1).contiguous().cuda(), None, None, ops.ACT_NONE, out_op=dy, reflect=False)
    fp.grad.zero_()
    conv.wgrad(xin, dy, 2, bias_grad=True)
    ops.sn_weight_grad(ctx, dev, 1, mn, mm)
    gin = torch.empty(2, 9, 13, 6, device="cuda")
    conv.dgrad(dy, 9, 13, 2, gin)
    torch.cuda.synchronize()
    ctx.check_pipeline()
    assert rel(y.permute(0, 3, 1, 2), z["y"]) < tol
    assert rel(u, z["u_out"]) < 1e-5
    assert rel(conv.weight.grad, z["dW"]) < tol
    assert rel(conv.bias.grad, z["db"]) < tol
    assert rel(gin.permute(0, 3, 1, 2), z["dx"]) < tol


def test_full_step_with_spectral_norm_matches_oracle():
    from tests.test_model_gpu import run_parity
    r = run_parity("bf16x3", sn_D=True)
    assert r["fake"] < 1e-3, r
    for k, v in r.items():
        if k.startswith("loss_"):
            assert v < 1e-3, (k, r)
    assert r["gradG"] < 1e-2 and r["gradD"] < 1e-2, r

"""Layer-by-layer audit of the backward pass on the FULL architecture (config #2 network at 128x256, batch 1):
every backward kernel call of one training step is recorded with its actual inputs, and each call's output is
recomputed in float64 from exactly those inputs.

Why: the end-to-end gradient of this GAN is a discontinuous function of the forward values (L1 sign, ReLU / LeakyReLU
masks, max-pool arg-max), so product-vs-oracle gradient bounds must stay loose (1e-2 ... 5e-2).
GIVEN its forward values, though, every layer's backward is a smooth function, and that is asserted here tightly:

  * InstanceNorm / activation / L1-term backward (hm_in_bwd, every presence-mask variant the step uses) <= 1e-4
  * data gradients (hm_conv_dgrad / fprop-as-dgrad of ConvTranspose, thin-side lowering, 4-parity stride-2)  <= 1e-4
  * weight gradients (hm_conv_wgrad + unpack, all engines incl. the CTA-pair and row-streaming ones)        <= 2e-4
  * reflect-fold of the residual chain (hm_fold_add), max-pool adjoint (hm_maxpool2_bwd)                    <= 1e-5

(relative to each tensor's max magnitude; the discrete mask / sign decisions are reproduced from the kernels' own fp32
expressions, so no element is excluded.)
"""
import contextlib
import io
import os
import sys

import pytest
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
pytestmark = pytest.mark.gpu



def _stored(op, c=None):
    """float64 NCHW copy of an Operand's STORED extent (border included), first c channels."""
    v = op.hi.double()
    if op.lo is not None:
        v = v + op.lo.double()
    return v[..., :(op.c if c is None else c)].permute(0, 3, 1, 2).contiguous()


def _nchw(t):
    return t.double().permute(0, 3, 1, 2).contiguous()


def _relmax(got, ref, keep=None):
    d = (got - ref).abs()
    if keep is not None:
        d = d * keep
    return float(d.max() / ref.abs().max().clamp_min(1e-30))


class Recorder(object):
    def __init__(self):
        self.calls = []

    def install(self, mp):
        from neurips18_hierchical_image_manipulation_b200 import networks, ops
        rec = self

        real_in_bwd = ops.in_bwd

        def in_bwd(ctx, shape, act, slope=0.2, y=None, mean=None, rstd=None, z=None, mask_op=None, g1=None, g1_border=0,
                   g1_ld=None, g1_coff=0, g2=None, tref=None, l1coef=0.0, out_op=None, out32=None):
            real_in_bwd(ctx, shape, act, slope, y, mean, rstd, z, mask_op, g1, g1_border, g1_ld, g1_coff, g2, tref,
                        l1coef, out_op, out32)
            c = lambda t: None if t is None else t.detach().clone()  # noqa: E731
            rec.calls.append(dict(kind="in_bwd", shape=shape, act=act, slope=slope, y=c(y), mean=c(mean), rstd=c(rstd),
                                  z=c(z), mask=None if mask_op is None else _stored(mask_op), g1=c(g1), g1_border=g1_border,
                                  g1_ld=g1_ld, g1_coff=g1_coff, g2=c(g2), tref=c(tref), l1coef=l1coef,
                                  out=_stored(out_op) if out_op is not None else _nchw(out32)))
        mp.setattr(ops, "in_bwd", in_bwd)

        real_fold = ops.fold_add

        def fold_add(ctx, g_padded, border, base, out):
            real_fold(ctx, g_padded, border, base, out)
            rec.calls.append(dict(kind="fold_add", g=g_padded.detach().clone(), border=border,
                                  base=None if base is None else base.detach().clone(), out=out.detach().clone()))
        mp.setattr(ops, "fold_add", fold_add)

        real_mp = ops.maxpool2_bwd

        def maxpool2_bwd(ctx, g, a_op, dz, n=None):
            real_mp(ctx, g, a_op, dz, n)
            rec.calls.append(dict(kind="maxpool_bwd", g=g.detach().clone(), a=_stored(a_op)[:g.shape[0]],
                                  out=dz.detach().clone()))
        mp.setattr(ops, "maxpool2_bwd", maxpool2_bwd)

        real_dgrad, real_wgrad = networks.ConvP.dgrad, networks.ConvP.wgrad

        def dgrad(self, dy, x_h, x_w, zero_pad, out32):
            real_dgrad(self, dy, x_h, x_w, zero_pad, out32)
            rec.calls.append(dict(kind="dgrad", conv=self, dy=_stored(dy), x_hw=(x_h, x_w), zero_pad=zero_pad,
                                  out=_nchw(out32)))

        def wgrad(self, x, dy, zero_pad, bias_grad=True):
            w0 = self.weight.grad.detach().clone()
            b0 = self.bias.grad.detach().clone()
            real_wgrad(self, x, dy, zero_pad, bias_grad)
            rec.calls.append(dict(kind="wgrad", conv=self, x=_stored(x, self.cin)[:dy.n], dy=_stored(dy), zero_pad=zero_pad,
                                  bias_grad=bias_grad, dw=(self.weight.grad.detach() - w0).double(),
                                  db=(self.bias.grad.detach() - b0).double()))
        real_rows = networks.ConvP.dgrad_rows

        def dgrad_rows(self, dy, x_h, x_w, zero_pad, out32, row0, nrows):
            real_rows(self, dy, x_h, x_w, zero_pad, out32, row0, nrows)
            rec.calls.append(dict(kind="dgrad", conv=self, dy=_stored(dy), x_hw=(x_h, x_w), zero_pad=zero_pad,
                                  out=_nchw(out32[..., :nrows]), rows=(row0, nrows)))
        mp.setattr(networks.ConvP, "dgrad", dgrad)
        mp.setattr(networks.ConvP, "dgrad_rows", dgrad_rows)
        mp.setattr(networks.ConvP, "wgrad", wgrad)


def _conv_fwd64(conv, x, zero_pad):
    w = conv.weight.detach().double()
    if conv.sn_scale is not None:
        w = w * conv.sn_scale.double()
    if conv.transposed:
        return F.conv_transpose2d(x, w, None, stride=2, padding=conv.pad, output_padding=1)
    return F.conv2d(x, w, None, stride=conv.stride, padding=zero_pad)


def _check_dgrad(c):
    conv = c["conv"]
    x = torch.zeros(c["dy"].shape[0], conv.cin, c["x_hw"][0], c["x_hw"][1], dtype=torch.float64, device="cuda",
                    requires_grad=True)
    y = _conv_fwd64(conv, x, c["zero_pad"])
    assert y.shape == c["dy"].shape, (conv.name, y.shape, c["dy"].shape)
    ref, = torch.autograd.grad(y, x, c["dy"])
    if "rows" in c:      # data gradient restricted to a channel range (image channels of the first PatchGAN layer)
        ref = ref[:, c["rows"][0]:c["rows"][0] + c["rows"][1]]
    return _relmax(c["out"], ref)


def _check_wgrad(c):
    conv = c["conv"]
    w = conv.weight.detach().double().requires_grad_(True)
    if conv.transposed:
        y = F.conv_transpose2d(c["x"], w, None, stride=2, padding=conv.pad, output_padding=1)
    else:
        y = F.conv2d(c["x"], w, None, stride=conv.stride, padding=c["zero_pad"])
    assert y.shape == c["dy"].shape, (conv.name, y.shape, c["dy"].shape)
    ref, = torch.autograd.grad(y, w, c["dy"])
    err = _relmax(c["dw"], ref)
    if c["bias_grad"]:
        refb = c["dy"].sum(dim=(0, 2, 3))
        err = max(err, _relmax(c["db"], refb))
    else:
        assert float(c["db"].abs().max()) == 0.0
    return err


def _fold_reflect64(g_nchw, b, H, W):
    if b == 0:
        return g_nchw
    t = torch.zeros(g_nchw.shape[0], g_nchw.shape[1], H, W, dtype=torch.float64, device=g_nchw.device, requires_grad=True)
    p = F.pad(t, (b, b, b, b), mode="reflect")
    out, = torch.autograd.grad(p, t, g_nchw)
    return out


def _check_in_bwd(c):
    """Values in float64; the DISCRETE decisions (activation mask, sign of the L1 term) are taken from the same fp32
    expressions the kernel evaluates -- (y - mean) * rstd and z - tref are single fp32 roundings of stored fp32 inputs,
    so torch reproduces them bit for bit and no element is ambiguous."""
    N, H, W, C = c["shape"]
    act, slope = c["act"], c["slope"]
    f32 = lambda t: t[:N].float().permute(0, 3, 1, 2)  # noqa: E731
    dz = torch.zeros(N, C, H, W, dtype=torch.float64, device="cuda")
    if c["g1"] is not None:
        g1 = c["g1"].double()
        ld = C if c["g1_ld"] is None else c["g1_ld"]
        assert g1.shape[-1] == ld
        g1 = g1[..., c["g1_coff"]:c["g1_coff"] + C][:N].permute(0, 3, 1, 2)
        dz = dz + _fold_reflect64(g1.contiguous(), c["g1_border"], H, W)
    if c["g2"] is not None:
        dz = dz + _nchw(c["g2"][:N])
    yhat = yhat32 = None
    if c["y"] is not None:
        mu, rs = c["mean"][:N, :, None, None], c["rstd"][:N, :, None, None]
        yhat32 = (f32(c["y"]) - mu.float()) * rs.float()
        yhat = (_nchw(c["y"][:N]) - mu.double()) * rs.double()
    gneg = 0.0 if act == 1 else (slope if act == 2 else 1.0)
    if c["tref"] is not None:
        if c["z"] is not None:
            zact32 = f32(c["z"])
        else:
            zact32 = torch.relu(yhat32) if act == 1 else (torch.where(yhat32 > 0, yhat32, yhat32 * slope) if act == 2 else yhat32)
        dz = dz + c["l1coef"] * torch.sign(zact32 - f32(c["tref"])).double()
    if yhat32 is not None:
        src = yhat32
    elif c["z"] is not None:
        src = f32(c["z"])
    elif c["mask"] is not None:
        src = c["mask"][:N, :C]
    else:
        src = torch.ones_like(dz)
    dyh = dz * torch.where(src > 0, torch.ones_like(dz), torch.full_like(dz, gneg))
    if yhat is not None:
        rs = c["rstd"][:N].double()[:, :, None, None]
        ref = rs * (dyh - dyh.mean(dim=(2, 3), keepdim=True) - yhat * (dyh * yhat).mean(dim=(2, 3), keepdim=True))
    else:
        ref = dyh
    return _relmax(c["out"][:N, :C], ref)


def _check_fold(c):
    g = _nchw(c["g"])
    N, H, W, C = c["out"].shape
    ref = _fold_reflect64(g, c["border"], H, W)
    if c["base"] is not None:
        ref = ref + _nchw(c["base"])
    return _relmax(_nchw(c["out"]), ref)


def _check_maxpool(c):
    a = c["a"].clone().requires_grad_(True)
    C = c["g"].shape[-1]
    p = F.max_pool2d(a[:, :C], 2, 2)
    ref, = torch.autograd.grad(p, a, _nchw(c["g"]))
    return _relmax(_nchw(c["out"]), ref[:, :C])


def test_every_backward_kernel_call_of_a_training_step_is_exact_given_its_inputs(monkeypatch):
    from neurips18_hierchical_image_manipulation_b200.models import Options, create_model
    from neurips18_hierchical_image_manipulation_b200.synthetic import synthetic_batch
    old_tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    rec = Recorder()
    rec.install(monkeypatch)
    opt = Options(label_nc=35, no_instance=False, netG="global", ngf=64, n_downsample_global=4, n_blocks_global=9,
                  num_D=3, use_output_gate=True, gpu_ids=[0], precision="bf16x3", name="audit",
                  checkpoints_dir="/tmp/hm_audit", vgg_weights="random", cuda_graph=False)
    with contextlib.redirect_stdout(io.StringIO()):
        m = create_model(opt).module
    b = synthetic_batch(1, 128, 256, 35, seed=99)
    st = m._forward_all(b["label"], b["inst"], b["image"], b["mask_in"])
    m._step = st
    m.flat_grad.zero_()
    m._backward_G([1.0, 1.0, 1.0])
    m._backward_D([0.5, 0.5])
    torch.cuda.synchronize()
    m.ctx.check_pipeline()
    tol = dict(in_bwd=1e-4, dgrad=1e-4, wgrad=2e-4, fold_add=1e-5, maxpool_bwd=1e-5)
    check = dict(in_bwd=_check_in_bwd, dgrad=_check_dgrad, wgrad=_check_wgrad, fold_add=_check_fold,
                 maxpool_bwd=_check_maxpool)
    worst, count = {}, {}
    try:
        for c in rec.calls:
            e = check[c["kind"]](c)
            name = c["conv"].name if "conv" in c else str(c.get("shape", ""))
            count[c["kind"]] = count.get(c["kind"], 0) + 1
            if e > worst.get(c["kind"], (-1, ""))[0]:
                worst[c["kind"]] = (e, name)
            assert e < tol[c["kind"]], (c["kind"], name, e)
    finally:
        torch.backends.cudnn.allow_tf32 = old_tf32
    print("backward audit: calls %s, worst %s" % (count, {k: ("%.1e" % v[0], v[1]) for k, v in worst.items()}))
    # the step really exercised every kind, in the numbers the architecture implies
    # G: 28 convs (stem, 4 down, 18 res, 4 up, head); D: 3 scales x 5 convs; VGG19: 13 convs, 4 pools
    assert count["wgrad"] == 28 + 15 and count["dgrad"] >= 27 + 15 + 13 and count["in_bwd"] >= 27 + 24 + 13, count
    assert count["maxpool_bwd"] == 4 and count["fold_add"] >= 9, count

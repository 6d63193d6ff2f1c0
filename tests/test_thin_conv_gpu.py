"""GPU parity tests of the tap-unrolled lowering of thin (<= 4 channel) convolutions (csrc/hm_thin.cu + ConvP):
generator head Conv2d(ngf,3,7) after ReflectionPad2d(3) (models/Pix2Pix_NET.py:91), PatchGAN output Conv2d(nf,1,4,p=2)
(models/Discriminator_NET.py:93) and VGG conv1_1 Conv2d(3,64,3,p=1), forward / data gradient / weight gradient,
against torch CPU fp64 convolutions and against the plain (not unrolled) engine path."""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
pytestmark = pytest.mark.gpu


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def _conv(ctx, cin, cout, k, seed, thin):
    from neurips18_hierchical_image_manipulation_b200.networks import ConvP, FlatParams
    fp = FlatParams(ctx.device)
    conv = ConvP(ctx, fp, "c", cin, cout, k, 1, 0)
    fp.materialize()
    conv.init_reference(torch.Generator().manual_seed(seed))
    if not thin:
        conv.thin_in = conv.thin_out = False
    return conv, fp


def _operand(ctx, x_nchw, border, reflect, grad=False):
    from neurips18_hierchical_image_manipulation_b200 import ops
    n, c, h, w = x_nchw.shape
    op = ops.Operand(ctx, n, h, w, c, border=border, grad=grad)
    ops.in_apply(ctx, nhwc(x_nchw).cuda(), None, None, ops.ACT_NONE, out_op=op, reflect=reflect)
    return op


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("split", [True, False], ids=["x3", "x1"])
@pytest.mark.parametrize("name,cin,cout,k,border,zero_pad,h,w", [
    ("head", 16, 3, 7, 3, 0, 37, 150),       # reflect border materialised, odd extents, > 1 M tile wide
    ("head_ngf64", 64, 3, 7, 3, 0, 20, 70),
    ("patch_out", 40, 1, 4, 0, 2, 19, 35),   # zero padding 2, even kernel
])
def test_thin_output_conv(name, cin, cout, k, border, zero_pad, h, w, split):
    from neurips18_hierchical_image_manipulation_b200 import ops
    tol = 1e-4 if split else 3e-2
    ctx = ops.Ctx("cuda:0", split=split)
    conv, fp = _conv(ctx, cin, cout, k, 3, thin=True)
    assert conv.thin_out
    g = torch.Generator().manual_seed(11)
    x = torch.randn(2, cin, h, w, generator=g)
    xin = _operand(ctx, x, border, reflect=True)
    ho, wo = conv.out_hw(xin.h, xin.w, zero_pad)
    y = torch.full((2, ho, wo, cout), float("nan"), device="cuda")
    conv.forward(xin, zero_pad, act=ops.ACT_TANH, out32=y)

    xd = x.double().requires_grad_(True)
    wd = conv.weight.detach().cpu().double().requires_grad_(True)
    bd = conv.bias.detach().cpu().double().requires_grad_(True)
    xp = F.pad(xd, (border,) * 4, mode="reflect") if border else xd
    xp.retain_grad()
    ref = torch.tanh(F.conv2d(xp, wd, bd, padding=zero_pad))
    torch.cuda.synchronize()
    assert rel(y.cpu().permute(0, 3, 1, 2).double(), ref.detach()) < tol

    # gradients of the pre-activation conv: dy -> (d stored input, dW, db)
    dy = torch.randn(2, cout, ho, wo, generator=g)
    pre = F.conv2d(xp, wd, bd, padding=zero_pad)
    pre.backward(dy.double())
    dyo = _operand(ctx, dy, 0, reflect=False, grad=True)
    fp.grad.zero_()
    conv.wgrad(xin, dyo, zero_pad, bias_grad=True)
    gin = torch.full((2, xin.h, xin.w, cin), float("nan"), device="cuda")
    conv.dgrad(dyo, xin.h, xin.w, zero_pad, gin)
    torch.cuda.synchronize()
    ctx.check_pipeline()
    assert rel(gin.cpu().permute(0, 3, 1, 2).double(), xp.grad) < tol
    assert rel(conv.weight.grad.cpu().double(), wd.grad) < tol
    assert rel(conv.bias.grad.cpu().double(), bd.grad) < tol

    # the plain engine path computes the same thing
    conv2, fp2 = _conv(ctx, cin, cout, k, 3, thin=False)
    y2 = torch.empty_like(y)
    conv2.forward(xin, zero_pad, act=ops.ACT_TANH, out32=y2)
    torch.cuda.synchronize()
    assert rel(y.cpu(), y2.cpu()) < tol


@pytest.mark.parametrize("split", [True, False], ids=["x3", "x1"])
def test_thin_input_conv_vgg_conv1_1(split):
    from neurips18_hierchical_image_manipulation_b200 import ops
    tol = 1e-4 if split else 3e-2
    ctx = ops.Ctx("cuda:0", split=split)
    conv, fp = _conv(ctx, 3, 64, 3, 5, thin=True)
    assert conv.thin_in
    g = torch.Generator().manual_seed(12)
    x = torch.randn(3, 3, 33, 140, generator=g)
    xin = _operand(ctx, x, 0, reflect=False)
    tap = torch.full((3, 33, 140, 64), float("nan"), device="cuda")
    out = ops.Operand(ctx, 3, 33, 140, 64)
    conv.forward(xin, 1, act=ops.ACT_RELU, out32=tap, out16=out)
    xd = x.double().requires_grad_(True)
    wd = conv.weight.detach().cpu().double()
    pre = F.conv2d(xd, wd, conv.bias.detach().cpu().double(), padding=1)
    ref = F.relu(pre)
    torch.cuda.synchronize()
    assert rel(tap.cpu().permute(0, 3, 1, 2).double(), ref.detach()) < tol
    assert rel(out.dense().cpu().double(), ref.detach()) < tol
    dy = torch.randn(3, 64, 33, 140, generator=g)
    pre.backward(dy.double())
    dyo = _operand(ctx, dy, 0, reflect=False, grad=True)
    gin = torch.full((3, 33, 140, 3), float("nan"), device="cuda")
    conv.dgrad(dyo, 33, 140, 1, gin)
    torch.cuda.synchronize()
    ctx.check_pipeline()
    assert rel(gin.cpu().permute(0, 3, 1, 2).double(), xd.grad) < tol


def test_tap_unroll_and_combine_kernels():
    """hm_tap_unroll / hm_tap_combine against index arithmetic in torch (both directions, zero fill outside)."""
    from neurips18_hierchical_image_manipulation_b200 import ops
    ctx = ops.Ctx("cuda:0", split=True)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 3, 9, 21, generator=g)
    src = _operand(ctx, x, 0, reflect=False)
    dst = ops.Operand(ctx, 2, 9, 21 + 4, 5 * 3)
    ops.tap_unroll(ctx, src, dst, 3, 1, 5, 0, 0, 1, -1)          # U[h, w', (i, c)] = x[h, w' - i, c]
    torch.cuda.synchronize()
    got = dst.dense().cpu()
    xs = (src.hi.float() + src.lo.float()).cpu().permute(0, 3, 1, 2)[:, :3]
    ref = torch.zeros(2, 15, 9, 25)
    for i in range(5):
        ref[:, i * 3:(i + 1) * 3, :, i:i + 21] = xs
    assert torch.equal(got, ref)
    T = torch.randn(2, 9, 25, 16, generator=g)
    out = torch.full((2, 9, 21, 3), float("nan"), device="cuda")
    bias = torch.randn(3, generator=g)
    ops.tap_combine(ctx, T.cuda(), 1, 5, 3, 0, 0, 1, 1, bias.cuda(), ops.ACT_NONE, 0.0, out)
    torch.cuda.synchronize()
    ref = bias.view(1, 1, 1, 3) + sum(T[:, :, i:i + 21, i * 3:(i + 1) * 3] for i in range(5))
    assert float((out.cpu() - ref).abs().max()) < 1e-5


@pytest.mark.parametrize("split", [True, False], ids=["x3", "x1"])
@pytest.mark.parametrize("cin,cout,d,h,w", [(48, 64, 2, 16, 16), (256, 256, 4, 16, 16), (40, 72, 2, 19, 131), (64, 32, 4, 12, 200)])
def test_dilated_conv3x3_forward_and_gradients(cin, cout, d, h, w, split):
    """nn.Conv2d(cin, cout, 3, padding=d, dilation=d, bias=False) -- DilatedResnetBlock's conv3x3 (layer_util.py:255-293) --
    through hm_conv_fprop_dil / hm_conv_dgrad_dil / hm_conv_wgrad_dil against torch CPU fp64; wide rows included (the
    row-streaming engines must NOT pick these up: their taps are not adjacent)."""
    from neurips18_hierchical_image_manipulation_b200 import ops
    from neurips18_hierchical_image_manipulation_b200.networks import ConvP, FlatParams
    tol = 1e-4 if split else 3e-2
    ctx = ops.Ctx("cuda:0", split=split)
    fp = FlatParams(ctx.device)
    conv = ConvP(ctx, fp, "c", cin, cout, 3, 1, d, dilation=d, bias=False)
    fp.materialize()
    conv.init_reference(torch.Generator().manual_seed(4))
    assert set(fp.params) == {"c.weight"}
    g = torch.Generator().manual_seed(13)
    x = torch.randn(2, cin, h, w, generator=g)
    xin = _operand(ctx, x, 0, reflect=False)
    assert conv.out_hw(h, w, d) == (h, w)
    y = torch.full((2, h, w, cout), float("nan"), device="cuda")
    conv.forward(xin, d, out32=y)
    xd = x.double().requires_grad_(True)
    wd = conv.weight.detach().cpu().double().requires_grad_(True)
    ref = F.conv2d(xd, wd, None, padding=d, dilation=d)
    torch.cuda.synchronize()
    assert rel(y.cpu().permute(0, 3, 1, 2).double(), ref.detach()) < tol
    dy = torch.randn(2, cout, h, w, generator=g)
    ref.backward(dy.double())
    dyo = _operand(ctx, dy, 0, reflect=False, grad=True)
    fp.grad.zero_()
    conv.wgrad(xin, dyo, d, bias_grad=False)
    gin = torch.full((2, h, w, cin), float("nan"), device="cuda")
    conv.dgrad(dyo, h, w, d, gin)
    torch.cuda.synchronize()
    ctx.check_pipeline()
    assert rel(gin.cpu().permute(0, 3, 1, 2).double(), xd.grad) < tol
    assert rel(conv.weight.grad.cpu().double(), wd.grad) < tol


@pytest.mark.parametrize("cin,cout,h,w", [(256, 128, 8, 8), (192, 64, 16, 16), (64, 32, 32, 32), (40, 24, 9, 13)])
def test_conv_transpose_4x4_s2_forward_and_gradients(cin, cout, h, w):
    """nn.ConvTranspose2d(cin, cout, 4, stride=2, padding=1) -- DeconvResnetBlock's deep branch (layer_util.py:210-222) --
    forward, data gradient and weight gradient through ConvP against torch CPU fp64."""
    from neurips18_hierchical_image_manipulation_b200 import ops
    from neurips18_hierchical_image_manipulation_b200.networks import ConvP, FlatParams
    ctx = ops.Ctx("cuda:0", split=True)
    fp = FlatParams(ctx.device)
    conv = ConvP(ctx, fp, "c", cin, cout, 4, 2, 1, transposed=True)
    fp.materialize()
    conv.init_reference(torch.Generator().manual_seed(6))
    g = torch.Generator().manual_seed(14)
    x = torch.randn(3, cin, h, w, generator=g)
    xin = _operand(ctx, x, 0, reflect=False)
    y = torch.full((3, 2 * h, 2 * w, cout), float("nan"), device="cuda")
    conv.forward(xin, 0, out32=y)
    xd = x.double().requires_grad_(True)
    wd = conv.weight.detach().cpu().double().requires_grad_(True)
    ref = F.conv_transpose2d(xd, wd, conv.bias.detach().cpu().double(), stride=2, padding=1)
    torch.cuda.synchronize()
    assert rel(y.cpu().permute(0, 3, 1, 2).double(), ref.detach()) < 1e-4
    dy = torch.randn(3, cout, 2 * h, 2 * w, generator=g)
    ref.backward(dy.double())
    dyo = _operand(ctx, dy, 0, reflect=False, grad=True)
    fp.grad.zero_()
    conv.wgrad(xin, dyo, 0, bias_grad=False)
    gin = torch.full((3, h, w, cin), float("nan"), device="cuda")
    conv.dgrad(dyo, h, w, 0, gin)
    torch.cuda.synchronize()
    ctx.check_pipeline()
    assert rel(gin.cpu().permute(0, 3, 1, 2).double(), xd.grad) < 1e-4
    assert rel(conv.weight.grad.cpu().double(), wd.grad) < 1e-4

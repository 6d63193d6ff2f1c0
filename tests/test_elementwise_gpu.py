"""GPU parity tests of the HBM-bound kernels (hm_elementwise.cu) against torch CPU references."""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from neurips18_hierchical_image_manipulation_b200.ops import Ctx
    return Ctx("cuda:0", split=True)


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def test_adam_matches_torch_optim(ctx):
    from neurips18_hierchical_image_manipulation_b200 import ops
    torch.manual_seed(0)
    n = 100003
    p0 = torch.randn(n)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=2e-4, betas=(0.5, 0.999))
    p = p0.clone().cuda(); m = torch.zeros(n).cuda(); v = torch.zeros(n).cuda()
    for step in range(1, 6):
        g = torch.randn(n) * (10.0 ** (step - 3))
        ref.grad = g.clone()
        opt.step()
        ops.adam_step(ctx, p, g.cuda(), m, v, 2e-4, 0.5, 0.999, 1e-8, step)
    torch.cuda.synchronize()
    assert float((p.cpu() - ref.detach()).abs().max()) < 1e-6


def test_instance_norm_apply_reflect_border(ctx):
    from neurips18_hierchical_image_manipulation_b200 import ops
    torch.manual_seed(1)
    y = torch.randn(2, 24, 13, 17) * 3 + 1
    skip = torch.randn(2, 24, 13, 17)
    yd = nhwc(y).cuda()
    mean, rstd = ops.in_stats(ctx, yd)
    ref_mean = y.mean((2, 3)); ref_var = y.var((2, 3), unbiased=False)
    assert float((mean.cpu() - ref_mean).abs().max()) < 1e-5
    assert float((rstd.cpu() - 1 / torch.sqrt(ref_var + 1e-5)).abs().max()) < 1e-5
    out32 = torch.empty_like(yd)
    op = ops.Operand(ctx, 2, 13, 17, 24, border=3)
    ops.in_apply(ctx, yd, mean, rstd, ops.ACT_RELU, skip=nhwc(skip).cuda(), out32=out32, out_op=op, reflect=True)
    ref = F.relu(F.instance_norm(y)) + skip
    torch.cuda.synchronize()
    assert float((out32.cpu().permute(0, 3, 1, 2) - ref).abs().max()) < 1e-5
    got = (op.hi.float() + op.lo.float()).cpu().permute(0, 3, 1, 2)
    assert float((got - F.pad(ref, (3,) * 4, mode="reflect")).abs().max()) < 1e-4


def test_instance_norm_backward_with_reflect_fold(ctx):
    from neurips18_hierchical_image_manipulation_b200 import ops
    torch.manual_seed(2)
    y = (torch.randn(2, 16, 9, 11) * 2).requires_grad_(True)
    out = F.pad(F.leaky_relu(F.instance_norm(y), 0.2), (2,) * 4, mode="reflect")
    g = torch.randn_like(out)
    (ref,) = torch.autograd.grad(out, y, g)
    yd = nhwc(y.detach()).cuda()
    mean, rstd = ops.in_stats(ctx, yd)
    dy = ops.Operand(ctx, 2, 9, 11, 16)
    dy32 = torch.empty_like(yd)
    ops.in_bwd(ctx, (2, 9, 11, 16), ops.ACT_LRELU, 0.2, y=yd, mean=mean, rstd=rstd, g1=nhwc(g).cuda(), g1_border=2,
               out_op=dy, out32=dy32)
    torch.cuda.synchronize()
    assert float((dy32.cpu().permute(0, 3, 1, 2) - ref).abs().max() / ref.abs().max()) < 1e-4
    got = (dy.hi.float() + dy.lo.float()).cpu().permute(0, 3, 1, 2)
    assert float((got - ref).abs().max() / ref.abs().max()) < 1e-4


def test_avgpool_and_maxpool_fwd_bwd(ctx):
    from neurips18_hierchical_image_manipulation_b200 import ops
    torch.manual_seed(3)
    x = torch.randn(2, 8, 13, 18).requires_grad_(True)
    xo = ops.Operand(ctx, 2, 13, 18, 8)
    ops.in_apply(ctx, nhwc(x.detach()).cuda(), None, None, ops.ACT_NONE, out_op=xo)
    # avg pool (count_include_pad=False), odd extent
    ref = F.avg_pool2d(x, 3, stride=2, padding=1, count_include_pad=False)
    po = ops.Operand(ctx, 2, 7, 9, 8)
    ops.avgpool3s2(ctx, xo, po)
    torch.cuda.synchronize()
    assert float((po.dense().cpu() - ref).abs().max()) < 1e-4
    g = torch.randn_like(ref)
    (gref,) = torch.autograd.grad(ref, x, g)
    gf = torch.zeros(2, 13, 18, 8).cuda()
    ops.avgpool3s2_bwd(ctx, nhwc(g).cuda(), gf, 2, 7)
    torch.cuda.synchronize()
    assert float((gf.cpu().permute(0, 3, 1, 2)[:, 2:7] - gref[:, 2:7]).abs().max()) < 1e-5
    assert float(gf[..., :2].abs().max()) == 0.0
    # max pool 2x2
    x2 = torch.randn(2, 8, 12, 18).requires_grad_(True)
    x2o = ops.Operand(ctx, 2, 12, 18, 8)
    ops.in_apply(ctx, nhwc(x2.detach()).cuda(), None, None, ops.ACT_NONE, out_op=x2o)
    x2r = x2o.dense().cpu().requires_grad_(True)  # the operand's own (hi+lo) values decide the arg-max
    mref = F.max_pool2d(x2r, 2, 2)
    mo = ops.Operand(ctx, 2, 6, 9, 8)
    ops.maxpool2(ctx, x2o, mo)
    torch.cuda.synchronize()
    assert float((mo.dense().cpu() - mref.detach()).abs().max()) < 1e-6
    g2 = torch.randn_like(mref)
    (g2ref,) = torch.autograd.grad(mref, x2r, g2)
    dz = torch.empty(2, 12, 18, 8).cuda()
    ops.maxpool2_bwd(ctx, nhwc(g2).cuda(), x2o, dz)
    torch.cuda.synchronize()
    assert float((dz.cpu().permute(0, 3, 1, 2) - g2ref).abs().max()) < 1e-6


def test_encode_input_matches_oracle(ctx):
    from neurips18_hierchical_image_manipulation_b200 import ops
    from neurips18_hierchical_image_manipulation_b200.synthetic import synthetic_batch
    from oracle import model as O
    b = synthetic_batch(2, 32, 48, label_nc=7, seed=9)
    input_label, real, cond = O.encode_input(b["label"], b["inst"], b["image"], b["mask_in"], 7, False)
    ref = torch.cat((input_label, cond), 1)  # 7 + 1 + 3 channels
    g = ops.Operand(ctx, 2, 32, 48, 11, border=3)
    d = ops.Operand(ctx, 4, 32, 48, 14)
    v = ops.Operand(ctx, 4, 32, 48, 3)
    dev = {k: t.cuda() for k, t in b.items()}
    ops.encode_input(ctx, dev["label"], dev["inst"], dev["image"], dev["mask_in"], 7, g, d, v)
    torch.cuda.synchronize()
    got = (g.hi.float() + g.lo.float()).cpu().permute(0, 3, 1, 2)[:, :11]
    assert float((got - F.pad(ref, (3,) * 4, mode="reflect")).abs().max()) < 1e-5
    dd = d.dense().cpu()
    assert float((dd[2:, :11] - ref).abs().max()) < 1e-5 and float((dd[2:, 11:14] - real).abs().max()) < 1e-5
    assert float((dd[:2, :11] - ref).abs().max()) < 1e-5
    assert float((v.dense().cpu()[2:] - real).abs().max()) < 1e-5


def test_loss_reductions(ctx):
    from neurips18_hierchical_image_manipulation_b200 import ops
    torch.manual_seed(4)
    a, b = torch.randn(3, 17, 19, 5), torch.randn(3, 17, 19, 5)
    acc = torch.zeros(5, dtype=torch.float64, device="cuda")
    ops.l1_sum(ctx, a.cuda(), b.cuda(), 2.0 / a.numel(), acc, 1)
    ops.mse_sum(ctx, a.cuda(), 1.0, 1.0 / a.numel(), acc, 3)
    torch.cuda.synchronize()
    assert abs(float(acc[1]) - 2 * float(F.l1_loss(a, b))) < 1e-6
    assert abs(float(acc[3]) - float(F.mse_loss(a, torch.ones_like(a)))) < 1e-6

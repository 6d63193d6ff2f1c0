"""GPU tests of the GlobalTwoStreamGenerator glue kernels (csrc/hm_twostream.cu) against torch, and of the single-stream
('label') variant of the generator against the oracle."""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
pytestmark = pytest.mark.gpu


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def test_mask_maxpool_blend_concat_and_cond_image():
    from neurips18_hierchical_image_manipulation_b200 import ops
    ctx = ops.Ctx("cuda:0", split=True)
    g = torch.Generator().manual_seed(0)
    mask = (torch.rand(2, 1, 32, 48, generator=g) > 0.97).float()
    m = ops.mask_maxpool(ctx, mask.cuda(), 8)
    torch.cuda.synchronize()
    assert torch.equal(m.cpu(), F.max_pool2d(mask, 8, 8)[:, 0])
    a, b = torch.randn(2, 12, 4, 6, generator=g), torch.randn(2, 12, 4, 6, generator=g)
    out32 = torch.full((2, 4, 6, 12), float("nan"), device="cuda")
    op = ops.Operand(ctx, 2, 4, 6, 12, border=1)
    ops.mask_blend(ctx, nhwc(a).cuda(), nhwc(b).cuda(), m, out32=out32, out_op=op)
    mm = F.max_pool2d(mask, 8, 8)
    ref = (1 - mm) * a + mm * b
    torch.cuda.synchronize()
    assert float((out32.cpu().permute(0, 3, 1, 2) - ref).abs().max()) < 1e-6
    got = (op.hi.float() + op.lo.float()).cpu().permute(0, 3, 1, 2)[:, :12]
    assert float((got - F.pad(ref, (1,) * 4, mode="reflect")).abs().max()) < 1e-4
    assert float(op.hi[..., 12:].abs().max()) == 0.0          # channel padding is zero
    gr = torch.randn(2, 4, 6, 12, generator=g).cuda()
    da, db = torch.empty_like(gr), torch.empty_like(gr)
    ops.mask_blend_bwd(ctx, gr, m, da, db)
    torch.cuda.synchronize()
    mq = mm.permute(0, 2, 3, 1).cuda()
    assert torch.allclose(da, (1 - mq) * gr) and torch.allclose(db, mq * gr)
    # concat of two operands (8-aligned and ragged channel counts)
    for ca, cb in ((8, 16), (4, 12)):
        xa, xb = torch.randn(1, ca, 5, 7, generator=g), torch.randn(1, cb, 5, 7, generator=g)
        oa, ob = ops.Operand(ctx, 1, 5, 7, ca), ops.Operand(ctx, 1, 5, 7, cb)
        ops.in_apply(ctx, nhwc(xa).cuda(), None, None, ops.ACT_NONE, out_op=oa)
        ops.in_apply(ctx, nhwc(xb).cuda(), None, None, ops.ACT_NONE, out_op=ob)
        oc = ops.concat_operands(ctx, oa, ob)
        torch.cuda.synchronize()
        assert oc.c == ca + cb
        assert float((oc.dense().cpu() - torch.cat((xa, xb), 1)).abs().max()) < 1e-4
    image = torch.rand(2, 3, 32, 48, generator=g) * 2 - 1
    co = ops.cond_image_operand(ctx, image.cuda(), mask.cuda(), 3)
    torch.cuda.synchronize()
    got = (co.hi.float() + co.lo.float()).cpu().permute(0, 3, 1, 2)[:, :3]
    assert float((got - F.pad((1 - mask) * image, (3,) * 4, mode="reflect")).abs().max()) < 1e-4


def test_parity_two_stream_label_only_variant():
    from tests.test_model_gpu import run_parity
    r = run_parity("bf16x3", netG="global_twostream", which_encoder="label", use_output_gate=True, n_downsample_global=2,
                   H=128, W=128)
    assert r["fake"] < 1e-3, r
    for k, v in r.items():
        if k.startswith("loss_"):
            assert v < 1e-3, (k, r)
    # the generator side is what this variant changes: 1.2e-3.  The (standard) discriminator's coarse scale sees 10x10
    # pixel planes here and its gradient flips a few LeakyReLU / L1 signs on fp32 rounding noise:
    # measured 9.6e-2 on scale0_layer3, 1e-2..2e-2 elsewhere; every D layer's backward is pinned by the other tests
    assert r["gradG"] < 1e-2 and r["gradD"] < 0.2, r

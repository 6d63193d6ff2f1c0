"""The deterministic stream-K tail of the CTA-pair K-engine (csrc/hm_engine2.cuh): when the pair tiles do not fill the
last wave of 74 clusters, the k-steps of the remaining tiles are dealt out evenly, split tiles park partial fp32
accumulators in the registered scratch (hm_set_scratch) and the last arriver sums them in segment order.

Checked per shape: against a float64 convolution (1e-4, like every bf16x3 engine case), bit-identical results from run
to run, and agreement with the plain whole-tile schedule (scratch unregistered) to fp32 summation-order noise.
Shapes: K1 (128 tiles: one full wave + 54-tile tail), fewer tiles than clusters (pure stream-K), an odd number of
M tiles (the peer CTA of the last pair owns no pixels), fused bias + ReLU + bf16 operand output, and a data gradient."""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
pytestmark = pytest.mark.gpu

CASES = [
    # n, h, w, cin, cout, k, act, out16, kind
    pytest.param(4, 32, 64, 1024, 1024, 3, 0, False, "fprop", id="K1-128tiles-1wave+54"),
    pytest.param(2, 32, 64, 512, 1024, 3, 0, False, "fprop", id="64tiles-pure-streamk"),
    pytest.param(3, 16, 24, 256, 2048, 3, 1, True, "fprop", id="odd-m-tiles-bias-relu-bf16out"),
    pytest.param(1, 32, 64, 512, 512, 4, 0, False, "fprop", id="16tiles-4x4"),
    pytest.param(2, 32, 64, 1024, 512, 3, 0, False, "dgrad", id="dgrad-64tiles"),
]


def _run(ctx, conv, op, n, h, w, cout, act, out16, kind, ops):
    y = torch.empty(n, h, w, cout, device="cuda")
    o16 = ops.Operand(ctx, n, h, w, cout, border=1, zero=True) if out16 else None
    if kind == "fprop":
        conv.forward(op, conv.pad_used, act=act, out32=y, out16=o16)
    else:
        conv.dgrad(op, h, w, conv.pad_used, y)
    return y, o16


@pytest.mark.parametrize("n,h,w,cin,cout,k,act,out16,kind", CASES)
def test_streamk_matches_fp64_is_deterministic_and_agrees_with_whole_tiles(n, h, w, cin, cout, k, act, out16, kind):
    from neurips18_hierchical_image_manipulation_b200 import ops
    from neurips18_hierchical_image_manipulation_b200.networks import ConvP, FlatParams
    ctx = ops.Ctx("cuda:0", split=True)
    lib = ctx.lib
    fp = FlatParams(ctx.device)
    pad = 1 if k == 3 else 2
    # fprop: conv cin -> cout on an h x w image (zero padding);  dgrad: gradient of a conv cout <- cin ... w.r.t. its input
    conv = ConvP(ctx, fp, "c", cin, cout, k, 1, pad) if kind == "fprop" else ConvP(ctx, fp, "c", cout, cin, k, 1, pad)
    conv.pad_used = pad
    fp.materialize()
    conv.init_reference(torch.Generator().manual_seed(0))
    g = torch.Generator(device="cuda").manual_seed(1)
    hin = h if kind == "fprop" else h + 2 * pad - k + 1
    win = w if kind == "fprop" else w + 2 * pad - k + 1
    x = torch.randn(n, hin, win, cin, device="cuda", generator=g)
    op = ops.Operand(ctx, n, hin, win, cin, grad=(kind == "dgrad"))
    ops.in_apply(ctx, x, None, None, ops.ACT_NONE, out_op=op, reflect=False)
    if kind == "fprop":
        assert conv.out_hw(hin, win, pad) == (h + 2 * pad - k + 1, w + 2 * pad - k + 1)
        h, w = conv.out_hw(hin, win, pad)
    scratch = ops._SCRATCH[ctx.device.index if ctx.device.index is not None else torch.cuda.current_device()]
    try:
        assert lib.hm_set_streamk(1) == 0                              # opt-in (off by default: no measured gain)
        y1, o1 = _run(ctx, conv, op, n, h, w, cout, act, out16, kind, ops)
        y2, o2 = _run(ctx, conv, op, n, h, w, cout, act, out16, kind, ops)
        torch.cuda.synchronize()
        assert lib.hm_set_scratch(None, 0) == 0                       # whole-tile schedule
        y0, _ = _run(ctx, conv, op, n, h, w, cout, act, out16, kind, ops)
        torch.cuda.synchronize()
    finally:
        assert lib.hm_set_scratch(scratch.data_ptr(), scratch.numel()) == 0
        lib.hm_set_streamk(0)
    ctx.check_pipeline()
    assert torch.equal(y1, y2), "stream-K result differs from run to run"
    if out16:
        assert torch.equal(o1.hi, o2.hi) and torch.equal(o1.lo, o2.lo)
    xd = x.permute(0, 3, 1, 2).double()
    wd = conv.weight.detach().double()
    if kind == "fprop":
        ref = F.conv2d(xd, wd, conv.bias.detach().double(), padding=pad)
        if act == 1:
            ref = torch.relu(ref)
    else:
        xin = torch.zeros(n, cout, h, w, dtype=torch.float64, device="cuda", requires_grad=True)
        ref, = torch.autograd.grad(F.conv2d(xin, wd, None, padding=pad), xin, xd)
    ref = ref.permute(0, 2, 3, 1)
    scale = float(ref.abs().max())
    e_sk, e_plain = float((y1.double() - ref).abs().max()) / scale, float((y0.double() - ref).abs().max()) / scale
    d = float((y1 - y0).abs().max()) / scale
    print("stream-K vs fp64 %.2e, whole tiles vs fp64 %.2e, stream-K vs whole tiles %.2e" % (e_sk, e_plain, d))
    assert e_sk < 1e-4 and e_plain < 1e-4 and d < 1e-4
    if out16:
        got = o1.dense().permute(0, 2, 3, 1).double()
        assert float((got - ref).abs().max()) / scale < 1e-4

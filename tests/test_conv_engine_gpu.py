"""GPU parity tests of the tcgen05 conv engines through the C ABI (hm_conv_fprop / hm_conv_dgrad / hm_conv_wgrad)
against torch CPU fp64 convolutions; shapes cover every conv kind on the hot path (3x3 reflect, 3x3 s2, 7x7 stem/head,
4x4 s2/s1 p2 with odd extents, ConvTranspose, Cout=1/3, Cin=3/38/41)."""
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

# bf16x3 (fp32-parity mode): fp32-accumulation noise only; bf16: operand rounding (2^-9 per operand)
TOL = {True: 1e-4, False: 3e-2}


def _cases():
    import gpu_engine_check as G
    out = []
    for grp, lst in G.CASES.items():
        for name, fn, kw in lst:
            for split in (True, False):
                out.append(pytest.param(fn, kw, split, id="%s-%s-%s" % (grp, name.replace(" ", "_"), "x3" if split else "x1")))
    return out


@pytest.mark.parametrize("fn,kw,split", _cases())
def test_engine_case(fn, kw, split):
    import torch
    import gpu_engine_check as G
    from neurips18_hierchical_image_manipulation_b200 import _lib as L
    lib = L.load()
    err = G.run_case(lib, torch.device("cuda:0"), fn, kw, split)
    assert err < TOL[split], err


def test_pack_weight_pair_equals_the_two_single_role_packs():
    """hm_pack_weight_pair (one pass over the fp32 weights, both engine roles) writes bit-identical slabs to the two
    hm_pack_weight calls it replaces -- Conv2d and ConvTranspose2d layouts, 3x3 / 4x4 / 7x7 / 1x1 taps, ragged channel
    counts, with and without lo planes."""
    import torch
    from neurips18_hierchical_image_manipulation_b200 import _lib as L
    from neurips18_hierchical_image_manipulation_b200 import ops
    from neurips18_hierchical_image_manipulation_b200.ops import PackedWeight
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(5)
    for split in (True, False):
        ctx = ops.Ctx(dev, split=split)
        for A, B, k in ((64, 38, 7), (128, 64, 3), (256, 128, 4), (1024, 1024, 3), (72, 200, 3), (35, 16, 1), (512, 256, 4)):
            kk = k * k
            w = torch.randn(A, B, k, k, generator=g).to(dev)
            ref1, ref2 = PackedWeight(ctx, A, B, kk), PackedWeight(ctx, B, A, kk, grad=True)
            ref1.pack(ctx, w, B * kk, kk, 1)
            ref2.pack(ctx, w, kk, B * kk, 1)
            p1, p2 = PackedWeight(ctx, A, B, kk), PackedWeight(ctx, B, A, kk, grad=True)
            for t in (p1.hi, p1.lo, p2.hi, p2.lo):
                if t is not None:
                    t.fill_(float("nan"))
            L.check(ctx.lib.hm_pack_weight_pair(w.data_ptr(), A, B, kk, p1.hi.data_ptr(), ops._ptr(p1.lo), p2.hi.data_ptr(),
                                                ops._ptr(p2.lo), ops._stream()), "hm_pack_weight_pair")
            torch.cuda.synchronize()
            for a, b in ((p1.hi, ref1.hi), (p1.lo, ref1.lo), (p2.hi, ref2.hi), (p2.lo, ref2.lo)):
                if a is not None:
                    assert torch.equal(a.view(torch.int16), b.view(torch.int16)), (A, B, k, split)

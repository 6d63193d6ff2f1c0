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

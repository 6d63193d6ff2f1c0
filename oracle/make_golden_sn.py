"""Generate tests/golden/sn_conv.npz from the REFERENCE'S OWN SNConv2d (models/sn_utils.py:49-72), run once in the
build container:   python oracle/make_golden_sn.py      (needs /root/reference)

Fixture: one training-mode forward of SNConv2d(6, 10, 4, stride=1, padding=2) on a seeded input, the gradient of a
seeded linear functional of the output w.r.t. weight / bias / input (autograd flows through the power iteration,
sn_utils.py:11-25), sigma and the iterated u.  Kept separate from make_golden.py so the other fixtures stay
byte-identical.  TEST INFRASTRUCTURE ONLY.
"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("HM_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def main():
    sys.path.insert(0, os.path.join(REF, "models"))
    import sn_utils as SN
    torch.manual_seed(8)
    conv = SN.SNConv2d(6, 10, 4, stride=1, padding=2)
    conv.train()
    u0 = conv.u.detach().clone()
    W0 = conv.weight.detach().clone()
    x = torch.randn(2, 6, 9, 13, requires_grad=True)
    sigma, _ = SN.max_singular_value(conv.weight, conv.u, 1)
    y = conv(x)
    g = torch.randn_like(y)
    (y * g).sum().backward()
    np.savez_compressed(os.path.join(OUT, "sn_conv.npz"), W=W0.numpy(), b=conv.bias.detach().numpy(), u=u0.numpy(),
                        x=x.detach().numpy(), g=g.numpy(), y=y.detach().numpy(), sigma=sigma.detach().numpy(),
                        u_out=conv.u.detach().numpy(), dW=conv.weight.grad.numpy(), db=conv.bias.grad.numpy(),
                        dx=x.grad.numpy())
    print("sn_conv.npz", os.path.getsize(os.path.join(OUT, "sn_conv.npz")), "sigma", float(sigma))


if __name__ == "__main__":
    main()

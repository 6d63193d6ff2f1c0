"""Generate tests/golden/twostream_small.npz from the REFERENCE'S OWN GlobalTwoStreamGenerator
(models/Pix2Pix_NET.py:103-247), run once in the build container:   python oracle/make_golden_twostream.py

Fixture: GlobalTwoStreamGenerator(6, 3, ngf=8, n_downsampling=3, n_blocks=2, 'instance', 'reflect', use_skip=True,
which_stream='ctx_label', use_output_gate=True, 'early_add') after weights_init; seeded inputs, output and the
gradients of a seeded linear functional of the output w.r.t. every parameter.  TEST INFRASTRUCTURE ONLY.
"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("HM_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def main():
    sys.path.insert(0, os.path.join(REF, "models"))
    sys.path.insert(0, REF)
    import Pix2Pix_NET as P
    import layer_util as LU
    torch.manual_seed(9)
    g = P.GlobalTwoStreamGenerator(6, 3, 8, 3, 2, "instance", "reflect", True, "ctx_label", True, "early_add")
    g.apply(LU.weights_init)
    img = torch.randn(2, 3, 32, 48)
    lab = torch.randint(0, 6, (2, 1, 4, 6))
    lab = torch.nn.functional.interpolate(lab.float(), size=(32, 48), mode="nearest").long()
    label = torch.zeros(2, 6, 32, 48).scatter_(1, lab, 1.0)
    mask = torch.zeros(2, 1, 32, 48)
    mask[0, :, 8:24, 8:32] = 1
    mask[1, :, 4:20, 20:44] = 1
    img = (1 - mask) * img
    y = g(img, label, mask)
    cot = torch.randn_like(y)
    grads = torch.autograd.grad((y * cot).sum(), list(g.parameters()))
    names = [k for k, _ in g.named_parameters()]
    np.savez_compressed(os.path.join(OUT, "twostream_small.npz"), img=img.numpy(), label=label.numpy(), mask=mask.numpy(),
                        out=y.detach().numpy(), cot=cot.numpy(),
                        **{"w::" + k: v.detach().numpy() for k, v in g.state_dict().items()},
                        **{"g::" + k: v.numpy() for k, v in zip(names, grads)})
    print("twostream_small.npz", os.path.getsize(os.path.join(OUT, "twostream_small.npz")), names[:4], len(names))


if __name__ == "__main__":
    main()

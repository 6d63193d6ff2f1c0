"""Generate tests/golden/box2mask_d_small.npz from the REFERENCE'S OWN MultiscaleDiscriminator in its box2mask
configuration (models/TwoStreamAE_mask.py:83-92: which_gan == 'patch_multiscale' -> 2 scales, norm_layer 'batch',
LSGAN, getIntermFeat=True), run once in the build container:   python oracle/make_golden_box2mask_gan.py

Stored: the input (masked [object mask | cond] tensor, 13 channels at label_nc 6), the parameter manifest (values are
oracle.weights.named_param(name, shape), loaded like a checkpoint), the 10 taps, the LSGAN losses of GANLoss for both
targets, the gradient of the target-0 loss w.r.t. every parameter and of the target-1 loss w.r.t. the input.
TEST INFRASTRUCTURE ONLY."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
REF = os.environ.get("HM_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def main():
    sys.path.insert(0, os.path.join(REF, "models"))
    sys.path.insert(0, REF)
    import Discriminator_NET as D
    import losses as LS
    from oracle.weights import named_param
    torch.manual_seed(3)
    net = D.MultiscaleDiscriminator(13, 16, 3, "batch", False, 2, True)
    with torch.no_grad():
        for k, p in net.named_parameters():
            p.copy_(named_param(k, p.shape))
    net.train()
    g = torch.Generator().manual_seed(8)
    x = torch.rand(3, 13, 48, 64, generator=g)
    x[:, 1:] = (x[:, 1:] > 0.7).float()
    mask = torch.zeros(3, 1, 48, 64)
    mask[:, :, 8:40, 10:50] = 1
    x = (x * mask).requires_grad_(True)
    taps = net(x)
    crit = LS.GANLoss(use_lsgan=True)
    l_real, l_fake = crit(taps, True), crit(taps, False)
    names = [k for k, _ in net.named_parameters()]
    gp = torch.autograd.grad(l_fake, list(net.parameters()), retain_graph=True)
    gx, = torch.autograd.grad(l_real, x)
    out = dict(x=x.detach().numpy(), loss_real=float(l_real), loss_fake=float(l_fake), gx=gx.numpy(),
               param_names=np.array(names), param_shapes=np.array([";".join(str(v) for v in p.shape) for p in net.parameters()]))
    for i, sc in enumerate(taps):
        for j, t in enumerate(sc):
            out["tap_%d_%d" % (i, j)] = t.detach().numpy()
    for n_, g_ in zip(names, gp):
        out["g::" + n_] = g_.numpy()
    np.savez_compressed(os.path.join(OUT, "box2mask_d_small.npz"), **out)
    print("wrote box2mask_d_small.npz:", len(names), "parameters; losses", float(l_real), float(l_fake))


if __name__ == "__main__":
    main()

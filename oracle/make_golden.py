"""Generate tests/golden/*.npz from the REFERENCE'S OWN classes (run once, in the build container).

    python oracle/make_golden.py            # needs /root/reference; writes tests/golden/

The reference modules are imported unmodified from /root/reference/models (they import cleanly under
python 3.12 / torch 2.11); every fixture is produced on CPU in fp32 with the seeds
below.  The fixtures pin oracle/model.py (tests/test_oracle_golden.py) and, through it, the CUDA path.
TEST INFRASTRUCTURE ONLY.
"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("HM_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def _import_reference():
    sys.path.insert(0, os.path.join(REF, "models"))
    sys.path.insert(0, REF)
    import Pix2Pix_NET  # noqa
    import Discriminator_NET  # noqa
    import layer_util  # noqa
    import losses  # noqa
    import sn_utils  # noqa
    return Pix2Pix_NET, Discriminator_NET, layer_util, losses, sn_utils


def sd_np(module, prefix="w::"):
    return {prefix + k: v.detach().numpy() for k, v in module.state_dict().items()}


def main():
    os.makedirs(OUT, exist_ok=True)
    P, D, LU, LS, SN = _import_reference()
    torch.set_num_threads(8)

    # ---- 1. BASELINE config #1: GlobalGenerator(38,3,64,1,1), 128x256, batch 1 -------------------------
    torch.manual_seed(0)
    g = P.GlobalGenerator(38, 3, 64, 1, 1, "instance", "reflect", False)
    g.apply(LU.weights_init)
    torch.manual_seed(1)
    lab = torch.randint(0, 35, (1, 1, 8, 16)).float()
    lab = torch.nn.functional.interpolate(lab, size=(128, 256), mode="nearest")
    onehot = torch.zeros(1, 35, 128, 256).scatter_(1, lab.long(), 1.0)
    img = torch.rand(1, 3, 128, 256) * 2 - 1
    x = torch.cat((onehot, img), 1)
    with torch.no_grad():
        y = g(x)
    np.savez_compressed(os.path.join(OUT, "g_config1.npz"), label=lab.numpy().astype(np.uint8), image=img.numpy(),
                        out=y.numpy(), **sd_np(g))

    # ---- 2. small GlobalGenerator with output gate, forward + backward --------------------------------
    torch.manual_seed(2)
    g = P.GlobalGenerator(10, 3, 8, 2, 2, "instance", "reflect", True)
    g.apply(LU.weights_init)
    x = torch.randn(2, 10, 32, 48)
    mask = torch.zeros(2, 1, 32, 48)
    mask[:, :, 8:24, 10:30] = 1
    y = g(x, mask)
    cot = torch.randn_like(y)
    grads = torch.autograd.grad((y * cot).sum(), list(g.parameters()))
    names = [k for k, _ in g.named_parameters()]
    np.savez_compressed(os.path.join(OUT, "g_small.npz"), x=x.numpy(), mask=mask.numpy(), out=y.detach().numpy(),
                        cot=cot.numpy(), **sd_np(g), **{"g::" + k: v.numpy() for k, v in zip(names, grads)})

    # ---- 3. LocalEnhancer (constructed directly; the reference's factory never builds it) ----
    torch.manual_seed(3)
    le = P.LocalEnhancer(9, 3, 4, 2, 2, 1, 2, "instance", "reflect")
    le.apply(LU.weights_init)
    x = torch.randn(1, 9, 32, 64)
    with torch.no_grad():
        y = le(x)
    np.savez_compressed(os.path.join(OUT, "local_small.npz"), x=x.numpy(), out=y.numpy(), **sd_np(le))

    # ---- 4. MultiscaleDiscriminator: 15 taps + LSGAN losses + backward ---------------------------------
    torch.manual_seed(4)
    d = D.MultiscaleDiscriminator(12, 8, 3, "instance", False, 3, True)
    d.apply(LU.weights_init)
    x = torch.randn(2, 12, 64, 96, requires_grad=True)
    taps = d(x)
    crit = LS.GANLoss(use_lsgan=True, tensor=torch.FloatTensor)
    l_real = crit(taps, True)
    l_fake = crit(taps, False)
    grads = torch.autograd.grad(l_real, list(d.parameters()) + [x])
    names = [k for k, _ in d.named_parameters()]
    fx = {"tap_%d_%d" % (i, j): t.detach().numpy() for i, s in enumerate(taps) for j, t in enumerate(s)}
    np.savez_compressed(os.path.join(OUT, "d_small.npz"), x=x.detach().numpy(), loss_real=float(l_real),
                        loss_fake=float(l_fake), gx=grads[-1].numpy(), **fx, **sd_np(d),
                        **{"g::" + k: v.numpy() for k, v in zip(names, grads[:-1])})

    # ---- 5. ResnetBlock, layer by layer -----------------------------------------------------------------
    torch.manual_seed(5)
    rb = LU.ResnetBlock(16, "reflect", LU.get_norm_layer("instance"))
    rb.apply(LU.weights_init)
    x = torch.randn(2, 16, 12, 20)
    with torch.no_grad():
        y = rb(x)
    np.savez_compressed(os.path.join(OUT, "resblock.npz"), x=x.numpy(), out=y.numpy(), **sd_np(rb))

    # ---- 6. weights_init statistics + spectral norm ------------------------------------------------------
    torch.manual_seed(6)
    W = torch.randn(24, 7, 3, 3)
    u = torch.randn(1, 24)
    sigma, u2 = SN.max_singular_value(W, u, 1)
    np.savez_compressed(os.path.join(OUT, "sn.npz"), W=W.numpy(), u=u.numpy(), sigma=sigma.numpy(), u_out=u2.numpy())

    # ---- 7. AvgPool pyramid used by D and LocalEnhancer ---------------------------------------------------
    torch.manual_seed(7)
    x = torch.randn(2, 5, 33, 47)
    ap = torch.nn.AvgPool2d(3, stride=2, padding=[1, 1], count_include_pad=False)
    np.savez_compressed(os.path.join(OUT, "avgpool.npz"), x=x.numpy(), out=ap(x).numpy())

    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()

"""Full-size (BASELINE config #2 shapes: 512x1024, ngf 64, 4 down / 9 res, 3-scale D, VGG19) checks through
size-independent properties -- the CPU oracle needs ~20 s per image at this size, so instead of a direct comparison:

* batch-split invariance: InstanceNorm is per sample and every loss is a mean over the batch, so the gradient of a
  2-image batch equals the mean of the two 1-image gradients and the losses average (this is also the exactness
  argument of the data-parallel path);
* the fused step (optimize_parameters) reproduces the reference script's sequence at full size;
* precision modes agree: plain bf16 losses are within 2e-2 of the bf16x3 (fp32-parity) losses;
* a conv layer at its real shape is linear in its input (engine property, bit-level independent of the oracle).
"""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
pytestmark = pytest.mark.gpu

H, W = 512, 1024


def _model(precision="bf16x3"):
    import contextlib
    import io
    from neurips18_hierchical_image_manipulation_b200.models import Options, create_model
    opt = Options(label_nc=35, no_instance=True, netG="global", ngf=64, n_downsample_global=4, n_blocks_global=9,
                  num_D=3, gpu_ids=[0], precision=precision, name="full", checkpoints_dir="/tmp/hm_full", vgg_weights="random")
    with contextlib.redirect_stdout(io.StringIO()):
        return create_model(opt).module


def _grads(m, batch):
    kw = dict(label=batch["label"], inst=batch["inst"], image=batch["image"], feat=None, mask_in=batch["mask_in"],
              mask_out=batch["mask_out"])
    st = m._forward_all(kw["label"], kw["inst"], kw["image"], kw["mask_in"])
    m._step = st
    m.flat_grad.zero_()
    m._backward_G([1.0, 1.0, 1.0])
    m._backward_D([0.5, 0.5])
    torch.cuda.synchronize()
    m.ctx.check_pipeline()
    return st["losses"].clone(), m.flat_grad.clone()


def test_batch_split_invariance_full_size():
    from neurips18_hierchical_image_manipulation_b200.synthetic import synthetic_batch
    m = _model()
    b = synthetic_batch(2, H, W, 35, seed=99)
    l2, g2 = _grads(m, b)
    acc_l, acc_g = None, None
    for i in range(2):
        bi = {k: v[i:i + 1] for k, v in b.items()}
        li, gi = _grads(m, bi)
        acc_l = li if acc_l is None else acc_l + li
        acc_g = gi if acc_g is None else acc_g + gi
    acc_l /= 2
    acc_g /= 2
    assert torch.isfinite(g2).all() and float(g2.abs().max()) > 0
    assert float((l2 - acc_l).abs().max() / l2.abs().max()) < 1e-5
    # The gradient of this GAN is discontinuous in the forward values (L1 sign, ReLU / LeakyReLU masks, max-pool
    # arg-max) and InstanceNorm over 32x64 planes amplifies: the ORACLE's own gradients move by 2-20 % (per-tensor max
    # norm) under a 1e-5 relative weight perturbation.  A different batch size changes fp32
    # summation orders by ~1e-7, so the two evaluations must agree far better than that bound, in direction and norm.
    nG = m.fpG.total
    for name, sl in (("G", slice(0, nG)), ("D", slice(nG, None))):
        a, c = g2[sl].double(), acc_g[sl].double()
        cos = float((a * c).sum() / (a.norm() * c.norm()))
        rl2 = float((a - c).norm() / a.norm())
        assert cos > 0.999 and rl2 < 5e-2, (name, cos, rl2)


def test_fused_step_equals_script_sequence_and_modes_agree_full_size():
    from neurips18_hierchical_image_manipulation_b200.synthetic import synthetic_batch
    a, b2 = _model(), _model()
    b2.fpG.load_state_dict(a.fpG.state_dict()); b2.fpD.load_state_dict(a.fpD.state_dict())
    batch = synthetic_batch(1, H, W, 35, seed=5)
    kw = dict(label=batch["label"], inst=batch["inst"], image=batch["image"], feat=None, mask_in=batch["mask_in"],
              mask_out=batch["mask_out"])
    losses, fake = a.forward(infer=True, **kw)
    assert fake.shape == (1, 3, H, W) and float(fake.abs().max()) <= 1.0
    ld = dict(zip(a.loss_names, losses))
    a.optimizer_G.zero_grad(); (ld["G_GAN"] + ld["G_GAN_Feat"] + ld["G_VGG"]).backward(); a.optimizer_G.step()
    a.optimizer_D.zero_grad(); ((ld["D_fake"] + ld["D_real"]) * 0.5).backward(); a.optimizer_D.step()
    lb = b2.optimize_parameters(**kw)
    torch.cuda.synchronize()
    la = torch.stack([x.detach() for x in losses])
    assert float((la - lb).abs().max() / la.abs().max()) < 1e-5
    # split-K weight gradients use fp32 atomics, so the two runs agree to rounding, not bitwise; Adam's first step
    # moves every weight by ~lr, so compare the parameters at a fraction of that
    dG = float((a.flat - b2.flat).abs().max())
    assert dG < 0.05 * a.opt.lr * 10, dG
    del b2
    c = _model("bf16")
    c.fpG.load_state_dict(a.fpG.state_dict()); c.fpD.load_state_dict(a.fpD.state_dict())
    la, _ = a.forward(infer=False, **kw)      # both modes at the same (post-step) weights
    lc, _ = c.forward(infer=False, **kw)
    torch.cuda.synchronize()
    c.ctx.check_pipeline()
    la = torch.stack([x.detach() for x in la])
    lc = torch.stack([x.detach() for x in lc])
    # plain bf16 products: every loss within 2e-2 (relative) of the bf16x3 / fp32-parity value on the same weights
    e = ((lc - la).abs() / la.abs()).cpu()
    assert float(e.max()) < 2e-2, (e.tolist(), la.tolist(), lc.tolist())
    assert float(e.max()) > 0, "bf16 and bf16x3 cannot agree bitwise: the precision switch is not wired through"


def test_resblock_conv_is_linear_at_full_size():
    """K1 shape (1024->1024 3x3 on 32x64 x4): conv(2x) == 2 conv(x) exactly (power-of-two scaling commutes with
    every rounding step), conv(x1 + x2) == conv(x1) + conv(x2) to fp32 accumulation noise."""
    from neurips18_hierchical_image_manipulation_b200 import ops
    from neurips18_hierchical_image_manipulation_b200.networks import ConvP, FlatParams
    ctx = ops.Ctx("cuda:0", split=True)
    fp = FlatParams(ctx.device)
    conv = ConvP(ctx, fp, "k1", 1024, 1024, 3, 1, 0)
    fp.materialize()
    conv.init_reference(torch.Generator().manual_seed(0))
    g = torch.Generator(device="cuda").manual_seed(1)

    def run(x32):
        op = ops.Operand(ctx, 4, 32, 64, 1024, border=1)
        ops.in_apply(ctx, x32, None, None, ops.ACT_NONE, out_op=op, reflect=True)
        y = torch.empty(4, 32, 64, 1024, device="cuda")
        conv.forward(op, 0, out32=y, use_bias=False)
        return y
    x1 = torch.randn(4, 32, 64, 1024, device="cuda", generator=g)
    x2 = torch.randn(4, 32, 64, 1024, device="cuda", generator=g)
    y1, y2 = run(x1), run(x2)
    torch.cuda.synchronize()
    assert torch.equal(run(2 * x1), 2 * y1)
    y12 = run(x1 + x2)
    assert float((y12 - (y1 + y2)).abs().max() / y12.abs().max()) < 1e-4
    ctx.check_pipeline()

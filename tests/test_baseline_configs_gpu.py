"""Direct oracle comparisons at BASELINE.json's REAL shapes (the other GPU tests use reduced networks):

* config #2 exactly -- 512x1024, 35 classes, GlobalGenerator(ngf 64, 4 down, 9 res), 3-scale D, VGG19, batch 1:
  generator output per pixel and the five losses against `oracle.train_step` (fp32 CPU, ~20 s), tolerance 1e-3
  (north_star), plus the direction of the G / D gradients;
* K1 at its benchmarked shape (res-block conv 1024->1024 3x3 on 4 x 32 x 64, M = 8192: 1.73 waves of CTA pairs)
  against a float64 convolution;
* config #4 at its real size -- LocalEnhancer 1024x2048 (ngf 32, 1 local enhancer), 2-scale D, instance edges,
  batch 1: forward + losses against the oracle (no-grad CPU forward), then the product's backward + Adam step run
  to completion with finite results.
"""
import contextlib
import io
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import model as O  # noqa: E402

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _model(**kw):
    from neurips18_hierchical_image_manipulation_b200.models import Options, create_model
    opt = Options(gpu_ids=[0], precision="bf16x3", name="cfg", checkpoints_dir="/tmp/hm_cfg", vgg_weights="random",
                  cuda_graph=False, **kw)
    with contextlib.redirect_stdout(io.StringIO()):
        return opt, create_model(opt)


def _cos(got, ref):
    keys = [k for k in ref if k.endswith("weight")]
    a = torch.cat([got[k].reshape(-1) for k in keys]).double()
    r = torch.cat([ref[k].reshape(-1).to(a.device) for k in keys]).double()
    return float((a @ r) / (a.norm() * r.norm()))


def test_config2_full_size_training_step_against_oracle():
    from neurips18_hierchical_image_manipulation_b200.models import random_vgg19_state_dict
    torch.set_num_threads(os.cpu_count() or 1)
    opt, model = _model(label_nc=35, no_instance=True, netG="global", ngf=64, n_downsample_global=4, n_blocks_global=9,
                        num_D=3, n_layers_D=3, ndf=64)
    m = model.module
    batch = O.synthetic_batch(1, 512, 1024, label_nc=35, seed=2024)
    g_sd, d_sd = m.fpG.state_dict(), m.fpD.state_dict()
    ls_ref, fake_ref, gG_ref, gD_ref, _ = O.train_step(O.Opt(num_D=3), {k: v.clone() for k, v in g_sd.items()},
                                                       {k: v.clone() for k, v in d_sd.items()},
                                                       random_vgg19_state_dict(opt.vgg_seed), batch)
    losses, fake = model(label=batch["label"], inst=batch["inst"], image=batch["image"], feat=None,
                         mask_in=batch["mask_in"], mask_out=batch["mask_out"], infer=True)
    ld = dict(zip(m.loss_names, losses))
    m.optimizer_G.zero_grad()
    (ld["G_GAN"] + ld["G_GAN_Feat"] + ld["G_VGG"]).backward()
    gG = {k: p.grad.detach().clone() for k, p in m.fpG.params.items()}
    m.optimizer_D.zero_grad()
    ((ld["D_fake"] + ld["D_real"]) * 0.5).backward()
    gD = {k: p.grad.detach().clone() for k, p in m.fpD.params.items()}
    torch.cuda.synchronize()
    m.ctx.check_pipeline()
    # per-pixel generator output and per-scalar losses: the north_star tolerance
    assert fake.shape == (1, 3, 512, 1024)
    e_fake = rel(fake, fake_ref)
    e_loss = {n: abs(float(a) - b) / abs(b) for n, a, b in zip(m.loss_names, losses, ls_ref)}
    cG, cD = _cos(gG, gG_ref), _cos(gD, gD_ref)
    print("config #2 full size: fake %.2e, losses %s, grad cosine G %.5f D %.5f" % (
        e_fake, {k: "%.1e" % v for k, v in e_loss.items()}, cG, cD))
    assert e_fake < 1e-3, e_fake
    for n, v in e_loss.items():
        assert v < 1e-3, (n, v, e_loss)
    # gradients: a discontinuous function of the forward values -> direction, not element-wise;
    # tests/test_backward_audit_gpu.py pins every layer's backward given its forward values
    assert cG > 0.98 and cD > 0.98, (cG, cD)


def test_k1_resblock_conv_at_benchmarked_shape_against_fp64():
    """M = 8192 (4 x 32 x 64 pixels), N = 1024, K = 9 x 1024: the exact launch bench.py's roofline times."""
    import torch.nn.functional as F
    from neurips18_hierchical_image_manipulation_b200 import ops
    from neurips18_hierchical_image_manipulation_b200.networks import ConvP, FlatParams
    ctx = ops.Ctx("cuda:0", split=True)
    fp = FlatParams(ctx.device)
    conv = ConvP(ctx, fp, "k1", 1024, 1024, 3, 1, 0)
    fp.materialize()
    conv.init_reference(torch.Generator().manual_seed(0))
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(4, 32, 64, 1024, device="cuda", generator=g)
    op = ops.Operand(ctx, 4, 32, 64, 1024, border=1)
    ops.in_apply(ctx, x, None, None, ops.ACT_NONE, out_op=op, reflect=True)
    y = torch.empty(4, 32, 64, 1024, device="cuda")
    conv.forward(op, 0, out32=y)
    y2 = torch.empty_like(y)
    conv.forward(op, 0, out32=y2)
    torch.cuda.synchronize()
    ctx.check_pipeline()
    assert torch.equal(y, y2), "K1 is not run-to-run deterministic"
    xr = F.pad(x.permute(0, 3, 1, 2).double(), (1, 1, 1, 1), mode="reflect")
    ref = F.conv2d(xr, conv.weight.detach().double(), conv.bias.detach().double()).permute(0, 2, 3, 1)
    err = float((y.double() - ref).abs().max() / ref.abs().max())
    print("K1 @ M=8192 vs fp64: %.2e" % err)
    assert err < 1e-4, err


def test_config4_local_enhancer_full_size_against_oracle():
    from neurips18_hierchical_image_manipulation_b200.models import random_vgg19_state_dict
    torch.set_num_threads(os.cpu_count() or 1)
    kw = dict(label_nc=35, no_instance=False, netG="local", ngf=32, n_downsample_global=4, n_blocks_global=9,
              n_local_enhancers=1, n_blocks_local=3, num_D=2, n_layers_D=3, ndf=64)
    opt, model = _model(**kw)
    m = model.module
    batch = O.synthetic_batch(1, 1024, 2048, label_nc=35, seed=404)
    g_sd, d_sd = m.fpG.state_dict(), m.fpD.state_dict()
    oopt = O.Opt(**kw)
    with torch.no_grad():
        ls_ref, fake_ref, _ = O.model_forward(oopt, g_sd, d_sd, random_vgg19_state_dict(opt.vgg_seed), batch["label"],
                                              batch["inst"], batch["image"], batch["mask_in"])
    torch.cuda.reset_peak_memory_stats()
    losses, fake = model(label=batch["label"], inst=batch["inst"], image=batch["image"], feat=None,
                         mask_in=batch["mask_in"], mask_out=batch["mask_out"], infer=True)
    torch.cuda.synchronize()
    assert fake.shape == (1, 3, 1024, 2048)
    e_fake = rel(fake, fake_ref)
    e_loss = {n: abs(float(a) - float(b)) / abs(float(b)) for n, a, b in zip(m.loss_names, losses, ls_ref)}
    print("config #4 full size: fake %.2e, losses %s" % (e_fake, {k: "%.1e" % v for k, v in e_loss.items()}))
    assert e_fake < 1e-3, e_fake
    for n, v in e_loss.items():
        assert v < 1e-3, (n, v, e_loss)
    # the rest of the training step at this size: both backward passes + Adam, finite and non-trivial
    ld = dict(zip(m.loss_names, losses))
    before = m.flat.clone()
    m.optimizer_G.zero_grad()
    (ld["G_GAN"] + ld["G_GAN_Feat"] + ld["G_VGG"]).backward()
    m.optimizer_G.step()
    m.optimizer_D.zero_grad()
    ((ld["D_fake"] + ld["D_real"]) * 0.5).backward()
    m.optimizer_D.step()
    torch.cuda.synchronize()
    m.ctx.check_pipeline()
    assert torch.isfinite(m.flat).all() and torch.isfinite(m.flat_grad).all()
    moved = (m.flat - before).abs()
    assert float(moved.max()) <= 1.01 * opt.lr and float((moved > 0).float().mean()) > 0.9
    print("config #4 peak memory %.1f GB" % (torch.cuda.max_memory_allocated() / 2 ** 30))


def test_vgg19_taps_against_oracle():
    """A6: the five relu{1..5}_1 taps of the VGG19 tower, tap by tap (a wrong pool / tap index cannot hide in the scalar
    G_VGG loss), on a 2-image 64x96 input."""
    from neurips18_hierchical_image_manipulation_b200 import ops
    from neurips18_hierchical_image_manipulation_b200.models import random_vgg19_state_dict
    from neurips18_hierchical_image_manipulation_b200.networks import VGG19_TAP_AFTER, Vgg19
    ctx = ops.Ctx("cuda:0", split=True)
    sd = random_vgg19_state_dict(77)
    vgg = Vgg19(ctx, sd)
    x = torch.rand(2, 3, 64, 96, generator=torch.Generator().manual_seed(3)) * 2 - 1
    op = ops.Operand(ctx, 2, 64, 96, 3)
    ops.in_apply(ctx, x.permute(0, 2, 3, 1).contiguous().cuda(), None, None, ops.ACT_NONE, out_op=op, reflect=True)
    tape = vgg.forward(op)
    torch.cuda.synchronize()
    ctx.check_pipeline()
    ref = O.vgg19_forward(sd, x)
    got = [tape["taps"][li] for li in sorted(tape["taps"])]
    assert len(got) == len(ref) == 5 == len(VGG19_TAP_AFTER)
    for i, (a, r) in enumerate(zip(got, ref)):
        assert tuple(a.shape) == (2, r.shape[2], r.shape[3], r.shape[1]), (i, a.shape, r.shape)
        assert rel(a.permute(0, 3, 1, 2), r) < 1e-3, i

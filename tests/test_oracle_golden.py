"""CPU tests: oracle/model.py (the restatement) against golden vectors produced by the reference's own
classes (oracle/make_golden.py).  fp32 on both sides; tolerances are accumulation-order noise only."""
import os
from collections import OrderedDict

import numpy as np
import pytest
import torch

from oracle import model as O


def load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name))
    sd = OrderedDict((k[3:], torch.from_numpy(z[k])) for k in z.files if k.startswith("w::"))
    grads = OrderedDict((k[3:], torch.from_numpy(z[k])) for k in z.files if k.startswith("g::"))
    return z, sd, grads


def close(a, b, tol):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    err = float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
    assert err < tol, "rel err %.3e >= %.1e" % (err, tol)


def test_global_generator_config1(golden_dir):
    """BASELINE config #1: GlobalGenerator(38,3,64,1,1) 128x256 batch 1."""
    z, sd, _ = load(golden_dir, "g_config1.npz")
    lab = torch.from_numpy(z["label"].astype(np.float32))
    img = torch.from_numpy(z["image"])
    onehot = torch.zeros(1, 35, 128, 256).scatter_(1, lab.long(), 1.0)
    y = O.global_generator_forward(sd, torch.cat((onehot, img), 1), 1, 1)
    close(y, z["out"], 2e-5)


def test_global_generator_small_gate_and_grads(golden_dir):
    z, sd, grads = load(golden_dir, "g_small.npz")
    par = OrderedDict((k, v.clone().requires_grad_(True)) for k, v in sd.items())
    y = O.global_generator_forward(par, torch.from_numpy(z["x"]), 2, 2, mask=torch.from_numpy(z["mask"]),
                                   use_output_gate=True)
    close(y.detach(), z["out"], 2e-5)
    g = torch.autograd.grad((y * torch.from_numpy(z["cot"])).sum(), list(par.values()))
    for (k, _), gi in zip(par.items(), g):
        if k.endswith("bias") and float(grads[k].abs().max()) < 1e-4:
            continue  # biases in front of InstanceNorm: analytically zero gradient, numerically noise
        close(gi, grads[k], 2e-3)


def test_local_enhancer(golden_dir):
    z, sd, _ = load(golden_dir, "local_small.npz")
    y = O.local_enhancer_forward(sd, torch.from_numpy(z["x"]), 2, 2, 1, 2)
    close(y, z["out"], 2e-5)


def test_multiscale_discriminator_taps_losses_grads(golden_dir):
    z, sd, grads = load(golden_dir, "d_small.npz")
    par = OrderedDict((k, v.clone().requires_grad_(True)) for k, v in sd.items())
    x = torch.from_numpy(z["x"]).requires_grad_(True)
    taps = O.multiscale_discriminator_forward(par, x, 3, 3)
    assert len(taps) == 3 and all(len(t) == 5 for t in taps)
    for i in range(3):
        for j in range(5):
            close(taps[i][j].detach(), z["tap_%d_%d" % (i, j)], 5e-5)
    l_real, l_fake = O.gan_loss(taps, True), O.gan_loss(taps, False)
    assert abs(float(l_real) - float(z["loss_real"])) < 1e-5 * abs(float(z["loss_real"]))
    assert abs(float(l_fake) - float(z["loss_fake"])) < 1e-5 * abs(float(z["loss_fake"]))
    g = torch.autograd.grad(l_real, list(par.values()) + [x])
    close(g[-1], z["gx"], 2e-3)
    for (k, _), gi in zip(par.items(), g[:-1]):
        if k.endswith("bias") and float(grads[k].abs().max()) < 1e-6:
            continue
        close(gi, grads[k], 2e-3)


def test_resnet_block(golden_dir):
    z, sd, _ = load(golden_dir, "resblock.npz")
    x = torch.from_numpy(z["x"])
    r = torch.nn.functional.conv2d(O.reflect_pad(x, 1), sd["conv_block.1.weight"], sd["conv_block.1.bias"])
    r = torch.relu(O.instance_norm(r))
    r = torch.nn.functional.conv2d(O.reflect_pad(r, 1), sd["conv_block.5.weight"], sd["conv_block.5.bias"])
    close(x + O.instance_norm(r), z["out"], 2e-5)


def test_avgpool_pyramid(golden_dir):
    z = np.load(os.path.join(golden_dir, "avgpool.npz"))
    close(O.avgpool_3s2(torch.from_numpy(z["x"])), z["out"], 1e-6)


def test_spectral_norm_power_iteration(golden_dir):
    z = np.load(os.path.join(golden_dir, "sn.npz"))
    sigma, u = O.max_singular_value(torch.from_numpy(z["W"]), torch.from_numpy(z["u"]), 1)
    close(sigma, z["sigma"], 1e-5)
    close(u, z["u_out"], 1e-5)


def test_vgg19_topology_matches_torchvision():
    tv = pytest.importorskip("torchvision")
    torch.manual_seed(0)
    feats = tv.models.vgg19(weights=None).features[:30].eval()
    sd = OrderedDict()
    for idx, _, _ in O.VGG19_CONVS:
        k = "slice%d.%d." % (O.VGG19_SLICE_OF[idx], idx)
        sd[k + "weight"] = feats[idx].weight.detach()
        sd[k + "bias"] = feats[idx].bias.detach()
    x = torch.randn(1, 3, 32, 48)
    taps = O.vgg19_forward(sd, x)
    cuts = [2, 7, 12, 21, 30]  # models/layer_util.py:390-399
    h = x
    with torch.no_grad():
        prev = 0
        for t, c in enumerate(cuts):
            for i in range(prev, c):
                h = feats[i](h)
            prev = c
            close(taps[t], h, 1e-5)


def test_model_forward_and_step_run():
    """The model-level restatement has no reference run to compare with (the class hard-codes .cuda(),
    SURVEY 8(c)); check its structural identities instead."""
    opt = O.Opt(ngf=4, n_downsample_global=2, n_blocks_global=1, ndf=4, num_D=2, label_nc=5)
    torch.manual_seed(0)
    from tests.util_weights import random_g_sd, random_d_sd
    g_sd, d_sd = random_g_sd(5 + 3, 3, 4, 2, 1), random_d_sd(5 + 3 + 3, 4, 3, 2)
    vgg = O.vgg19_random_state_dict()
    b = O.synthetic_batch(2, 32, 32, label_nc=5)
    losses, fake, ex = O.model_forward(opt, g_sd, d_sd, vgg, b["label"], b["inst"], b["image"], b["mask_in"])
    assert fake.shape == (2, 3, 32, 32) and len(losses) == 5
    # D_fake is evaluated on fake.detach(): same value as the G-side GAN tap but with target 0
    assert torch.isfinite(torch.stack([l.detach() for l in losses])).all()
    # cond image is zero inside the box, equal to the image outside (model :165-166)
    m = b["mask_in"].bool().expand(-1, 3, -1, -1)
    assert float(ex["cond"][m].abs().max()) == 0.0
    assert torch.equal(ex["cond"][~m], b["image"][~m])
    ls, fk, gG, gD, st = O.train_step(opt, g_sd, d_sd, vgg, b)
    assert st["step"] == 1 and all(torch.isfinite(v).all() for v in gG.values())


def test_spectral_norm_conv_forward_backward(golden_dir):
    """oracle.spectral_normalize against the reference's own SNConv2d (fixture: oracle/make_golden_sn.py), and the
    closed-form gradient through the power iteration that csrc/hm_sn.cu implements against the reference's autograd."""
    z = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(golden_dir, "sn_conv.npz")).items()}
    W = z["W"].clone().requires_grad_(True)
    b = z["b"].clone().requires_grad_(True)
    x = z["x"].clone().requires_grad_(True)
    eff, new_u = O.spectral_normalize(OrderedDict([("c.weight", W), ("c.bias", b), ("c.u", z["u"])]))
    y = torch.nn.functional.conv2d(x, eff["c.weight"], eff["c.bias"], stride=1, padding=2)
    (y * z["g"]).sum().backward()
    close(y, z["y"], 1e-5)
    close(new_u["c.u"], z["u_out"], 1e-5)
    close(W.grad, z["dW"], 1e-4)
    close(b.grad, z["db"], 1e-5)
    close(x.grad, z["dx"], 1e-5)
    # closed form: dL/dW = (G - <G, W_bar> (g_b v^T + u0 g_a^T)) / sigma, G = dL/dW_bar
    Wm = z["W"].reshape(10, -1).double()
    u0 = z["u"].double().reshape(-1)
    eps = 1e-12
    a = Wm.t() @ u0
    na = a.norm(); v = a / (na + eps)
    bb = Wm @ v
    nb = bb.norm(); sigma = (bb @ bb) / (nb + eps)
    close(sigma.float(), z["sigma"].reshape(()), 1e-5)
    Wbar = (z["W"] / z["sigma"].reshape(())).detach().requires_grad_(True)
    y2 = torch.nn.functional.conv2d(z["x"], Wbar, z["b"], stride=1, padding=2)
    (y2 * z["g"]).sum().backward()
    G = Wbar.grad.reshape(10, -1).double()
    c = (G * Wm).sum() / sigma
    g_b = bb * (1 / (nb + eps) + eps / (nb + eps) ** 2)
    g_v = Wm.t() @ g_b
    g_a = g_v / (na + eps) - a * (a @ g_v) / (na * (na + eps) ** 2)
    dW = (G - c * (torch.outer(g_b, v) + torch.outer(u0, g_a))) / sigma
    close(dW.float().reshape(z["dW"].shape), z["dW"], 1e-4)


def test_global_twostream_generator_forward_and_grads(golden_dir):
    """oracle.global_twostream_forward against the reference's own GlobalTwoStreamGenerator (ctx_label, use_skip,
    output gate, early_add; fixture: oracle/make_golden_twostream.py)."""
    z, sd, grads = load(golden_dir, "twostream_small.npz")
    par = OrderedDict((k, v.clone().requires_grad_(True)) for k, v in sd.items())
    y = O.global_twostream_forward(par, torch.from_numpy(z["img"]), torch.from_numpy(z["label"]),
                                   torch.from_numpy(z["mask"]), 3, 2, use_skip=True, which_stream="ctx_label",
                                   use_output_gate=True)
    close(y.detach(), z["out"], 2e-5)
    g = torch.autograd.grad((y * torch.from_numpy(z["cot"])).sum(), list(par.values()))
    for (k, _), gi in zip(par.items(), g):
        if k.endswith("bias") and float(grads[k].abs().max()) < 1e-4:
            continue
        close(gi, grads[k], 2e-3)


def _model_case(golden_dir, name, **optkw):
    z = np.load(os.path.join(golden_dir, name))
    part = lambda p: OrderedDict((k[len(p):], torch.from_numpy(z[k])) for k in z.files if k.startswith(p))  # noqa: E731
    return z, part("wG::"), part("wD::"), part("gG::"), part("gD::"), {k: v for k, v in part("in::").items()}, O.Opt(**optkw)


@pytest.mark.parametrize("name,optkw", [
    ("model_global_gate_edges.npz", dict(label_nc=6, no_instance=False, ngf=8, n_downsample_global=2, n_blocks_global=2,
                                         ndf=8, num_D=2, use_output_gate=True)),
    ("model_shipped_twostream.npz", dict(label_nc=6, no_instance=True, ngf=8, n_downsample_global=3, n_blocks_global=2,
                                         ndf=8, num_D=2, use_output_gate=True, netG="global_twostream",
                                         which_encoder="ctx_label", use_skip=True, no_imgCond=True, mask_gan_input=True)),
    # --no_lsgan --no_ganFeat_loss: vanilla GAN (Sigmoid + BCE), one flattened Sequential per scale
    ("model_global_vanilla_gan.npz", dict(label_nc=6, no_instance=True, ngf=8, n_downsample_global=2, n_blocks_global=2,
                                          ndf=8, num_D=2, use_output_gate=True, no_lsgan=True, no_ganFeat_loss=True)),
    # which_encoder == 'ctx' (the option's default): context stream only, the discriminator sees the bare image
    ("model_twostream_ctx.npz", dict(label_nc=6, no_instance=True, ngf=8, n_downsample_global=2, n_blocks_global=2,
                                     ndf=8, num_D=2, use_output_gate=True, netG="global_twostream", which_encoder="ctx",
                                     use_skip=True, mask_gan_input=True)),
    # feat_fusion == 'late_add': per-stream ResnetBlocks before the masked fusion
    ("model_twostream_late_add.npz", dict(label_nc=6, no_instance=True, ngf=8, n_downsample_global=2, n_blocks_global=3,
                                          ndf=8, num_D=2, use_output_gate=True, netG="global_twostream",
                                          which_encoder="ctx_label", feat_fusion="late_add", use_skip=True)),
    # feat_fusion '*_concat': cat -> ReLU -> 1x1 conv -> InstanceNorm in place of the masked sum
    ("model_twostream_early_concat.npz", dict(label_nc=6, no_instance=True, ngf=8, n_downsample_global=2, n_blocks_global=2,
                                              ndf=8, num_D=2, use_output_gate=True, netG="global_twostream",
                                              which_encoder="ctx_label", feat_fusion="early_concat", use_skip=True)),
    ("model_twostream_late_concat.npz", dict(label_nc=6, no_instance=False, ngf=8, n_downsample_global=2, n_blocks_global=3,
                                             ndf=8, num_D=2, use_output_gate=False, netG="global_twostream",
                                             which_encoder="ctx_label", feat_fusion="late_concat", use_skip=False)),
])
def test_model_level_forward_against_the_reference_model(golden_dir, name, optkw):
    """oracle.model_forward / step_losses against the reference's OWN Pix2PixHDModel_condImg.forward run on CPU
    (oracle/make_golden_model.py): the five losses, the generated image and every parameter gradient of loss_G / loss_D.
    The oracle is evaluated in float64: in float32 it reproduces outputs and losses to 1e-6 as well, but one fp32
    rounding difference (nn.InstanceNorm2d vs (x-mean)/sqrt(var+eps) on the 6x8-pixel planes of the coarse PatchGAN
    scale) flips a LeakyReLU sign and moves that scale's gradients by 0.9-4.8 % -- the float64 evaluation agrees with
    the reference's float32 gradients to 1e-6, i.e. the restatement is exact and the gradient is that sensitive."""
    z, g_sd, d_sd, gG, gD, b, opt = _model_case(golden_dir, name, **optkw)
    dt = torch.float64
    g_par = OrderedDict((k, v.clone().to(dt).requires_grad_(True)) for k, v in g_sd.items())
    d_par = OrderedDict((k, v.clone().to(dt).requires_grad_(True)) for k, v in d_sd.items())
    vgg = OrderedDict((k, v.to(dt)) for k, v in O.vgg19_random_state_dict().items())
    losses, fake, _ = O.model_forward(opt, g_par, d_par, vgg, b["label"], b["inst"], b["image"], b["mask_in"], dtype=dt,
                                      mask_out=b["mask_out"])
    close(fake.detach().float(), z["fake"], 2e-5)
    for a, r in zip(losses, z["losses"]):
        assert abs(float(a) - float(r)) <= 2e-5 * abs(float(r)), (float(a), float(r))
    loss_G, loss_D = O.step_losses(losses)
    g1 = torch.autograd.grad(loss_G, list(g_par.values()), retain_graph=True)
    g2 = torch.autograd.grad(loss_D, list(d_par.values()))
    for ref, got, par in ((gG, g1, g_par), (gD, g2, d_par)):
        for (k, _), gi in zip(par.items(), got):
            if k.endswith("bias") and float(ref[k].abs().max()) < 1e-3 * float(ref[k[:-4] + "weight"].abs().max()):
                continue   # bias in front of an InstanceNorm: analytically zero, fp32 noise (1e-8 .. 5e-7) in the reference
            close(gi.float(), ref[k], 2e-4)   # reference gradients are float32: 1.2e-4 on the worst tensor of the 'ctx' case


# ---- N3 box2mask: oracle/box2mask.py against the reference's own MaskTwoStreamConv_NET -------------------------------
def _box2mask_fixture(golden_dir, dtype):
    from oracle.weights import named_param
    z = np.load(os.path.join(golden_dir, "box2mask_small.npz"))
    sd = {}
    for name, shp in zip(z["param_names"], z["param_shapes"]):
        shape = tuple(int(v) for v in str(shp).split(";"))
        sd[str(name)] = named_param(str(name), shape).to(dtype).requires_grad_(True)
    return z, sd


def test_box2mask_two_stream_forward_losses_and_gradients_against_the_reference_class(golden_dir):
    """BASELINE config #5 network (scripts/train_box2mask_city.sh flag set, reduced size): the four outputs of
    MaskTwoStreamConv_NET.forward, MaskReconLoss + BCE, and the gradient of every one of the 120 parameters."""
    from oracle import box2mask as B2
    from oracle.weights import named_param
    z, sd = _box2mask_fixture(golden_dir, torch.float64)
    ins = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("in::")}
    cond, cls_onehot = B2.encode_input(6, ins["mask_ctx_in"], ins["mask_in"], ins["cls"])
    assert torch.equal(cond, torch.from_numpy(z["cond"])) and torch.equal(cls_onehot, torch.from_numpy(z["cls_onehot"]))
    comb_logit, comb_lp, obj_logit, obj_prob = B2.two_stream_forward(sd, cond.double(), num_layers=3, n_blocks=2)
    for got, key in ((comb_logit, "comb_logit"), (comb_lp, "comb_prob"), (obj_logit, "obj_logit"), (obj_prob, "obj_prob")):
        ref = torch.from_numpy(z[key]).double()
        assert got.shape == ref.shape
        assert float((got - ref).abs().max() / ref.abs().max()) < 2e-5, key
    l_comb = B2.mask_recon_loss(comb_lp, ins["label_map"], ins["mask_out"])
    l_obj = B2.obj_recon_loss(obj_prob, ins["mask_out"].double(), ins["mask_obj_inst"].double())
    assert abs(float(l_comb) - float(z["loss_comb"])) < 2e-5 * abs(float(z["loss_comb"]))
    assert abs(float(l_obj) - float(z["loss_obj"])) < 2e-5 * abs(float(z["loss_obj"]))
    names = list(sd)
    grads = torch.autograd.grad(l_obj + l_comb, [sd[k] for k in names])
    n_full = n_proj = 0
    for k, g in zip(names, grads):
        if "g::" + k in z.files:
            ref = torch.from_numpy(z["g::" + k]).double()
            # (conv biases in front of a BatchNorm have an analytically zero gradient: ~1e-8 fp32 noise in the reference)
            tol = 2e-3 * float(ref.abs().max()) + 1e-6
            assert float((g - ref).abs().max()) <= tol, k
            n_full += 1
        else:
            s_, a_, p_ = (float(v) for v in z["gs::" + k])
            r = named_param("proj::" + k + ".bias", g.shape).double()
            assert abs(float(g.abs().sum()) - a_) <= 2e-3 * a_ + 1e-5, k
            assert abs(float(g.sum()) - s_) <= 2e-3 * a_ + 1e-5 and abs(float((g * r).sum()) - p_) <= 2e-3 * a_ * 0.05 + 1e-6, k
            n_proj += 1
    assert n_full + n_proj == 120 and n_proj >= 20


def test_box2mask_batchnorm_discriminator_against_the_reference_class(golden_dir):
    """The 2-scale BatchNorm MultiscaleDiscriminator of box2mask's --use_gan branch (TwoStreamAE_mask.py:83-92): taps,
    LSGAN losses for both targets, parameter gradients of the target-0 loss, input gradient of the target-1 loss."""
    from oracle import box2mask as B2
    from oracle.weights import named_param
    z = np.load(os.path.join(golden_dir, "box2mask_d_small.npz"))
    sd = {}
    for name, shp in zip(z["param_names"], z["param_shapes"]):
        sd[str(name)] = named_param(str(name), tuple(int(v) for v in str(shp).split(";"))).double().requires_grad_(True)
    x = torch.from_numpy(z["x"]).double().requires_grad_(True)
    taps = B2.multiscale_discriminator_bn_forward(sd, x, 2, 3)
    for i, sc in enumerate(taps):
        assert len(sc) == 5
        for j, t in enumerate(sc):
            ref = torch.from_numpy(z["tap_%d_%d" % (i, j)]).double()
            assert float((t - ref).abs().max() / ref.abs().max()) < 2e-5, (i, j)
    l_real, l_fake = B2.lsgan(taps, True), B2.lsgan(taps, False)
    assert abs(float(l_real) - float(z["loss_real"])) < 2e-5 * float(z["loss_real"])
    assert abs(float(l_fake) - float(z["loss_fake"])) < 2e-5 * float(z["loss_fake"])
    names = list(sd)
    gp = torch.autograd.grad(l_fake, [sd[k] for k in names], retain_graph=True)
    for k, g in zip(names, gp):
        ref = torch.from_numpy(z["g::" + k]).double()
        assert float((g - ref).abs().max()) <= 2e-3 * float(ref.abs().max()) + 1e-6, k
    gx, = torch.autograd.grad(l_real, x)
    ref = torch.from_numpy(z["gx"]).double()
    assert float((gx - ref).abs().max()) <= 2e-3 * float(ref.abs().max())


def test_box2mask_switch_net_train_and_eval_mode_against_the_reference_class(golden_dir):
    """--no_comb (MaskTwoStreamConvSwitch_NET) in training mode, the BatchNorm running buffers after that pass, and the
    eval-mode forward that uses them, against the reference's own class (oracle/make_golden_box2mask_switch.py)."""
    from oracle import box2mask as B2
    from oracle.weights import named_param
    z = np.load(os.path.join(golden_dir, "box2mask_switch_small.npz"))
    sd = {str(n): named_param(str(n), tuple(int(v) for v in str(s).split(";"))).double()
          for n, s in zip(z["param_names"], z["param_shapes"])}
    a = {k[3:]: torch.from_numpy(z[k]).double() for k in z.files if k.startswith("a::")}
    b = {k[3:]: torch.from_numpy(z[k]).double() for k in z.files if k.startswith("b::")}
    stats = {}
    cond, _ = B2.encode_input(6, a["mask_ctx_in"], a["mask_in"], a["cls"])
    with torch.no_grad():
        outs = B2.two_stream_forward(sd, cond.double(), num_layers=3, n_blocks=2, no_comb=True, bn_mode=stats)
    for got, k in zip(outs, ("comb_logit", "comb_prob", "obj_logit", "obj_prob")):
        ref = torch.from_numpy(z["train_" + k]).double()
        assert float((got - ref).abs().max() / ref.abs().max()) < 2e-5, k
    assert abs(float(B2.mask_recon_loss(outs[1], a["label_map"], a["mask_out"])) - float(z["train_loss_comb"])) < 2e-5
    assert abs(float(B2.obj_recon_loss(outs[3], a["mask_out"], a["mask_obj_inst"])) - float(z["train_loss_obj"])) < 2e-5
    n_buf = 0
    for key, (mean, var_u) in stats.items():      # running = 0.9 * init + 0.1 * batch statistic (unbiased variance)
        rm, rv = torch.from_numpy(z["buf::" + key + ".running_mean"]).double(), torch.from_numpy(z["buf::" + key + ".running_var"]).double()
        assert float((0.1 * mean - rm).abs().max()) < 1e-5 * max(1.0, float(rm.abs().max())), key
        assert float((0.9 + 0.1 * var_u - rv).abs().max()) < 1e-5 * float(rv.abs().max()), key
        assert int(z["buf::" + key + ".num_batches_tracked"]) == 1
        sd[key + ".running_mean"], sd[key + ".running_var"] = rm, rv
        n_buf += 1
    assert n_buf == 29 == sum(1 for k in z.files if k.endswith("running_mean"))
    cond, _ = B2.encode_input(6, b["mask_ctx_in"], b["mask_in"], b["cls"])
    with torch.no_grad():
        outs = B2.two_stream_forward(sd, cond.double(), num_layers=3, n_blocks=2, no_comb=True, bn_mode="eval")
    for got, k in zip(outs, ("comb_logit", "comb_prob", "obj_logit", "obj_prob")):
        ref = torch.from_numpy(z["eval_" + k]).double()
        assert float((got - ref).abs().max() / ref.abs().max()) < 2e-5, k


def test_image_pool_against_the_reference_model(golden_dir):
    """--pool_size 3 over four forwards on different batches (oracle/make_golden_model.py::_run_pool_case): the five losses
    of every forward (loss_D_fake is evaluated on the pool's history) and the discriminator gradients of the last one."""
    import random
    z = np.load(os.path.join(golden_dir, "model_global_pool.npz"))
    part = lambda p: OrderedDict((k[len(p):], torch.from_numpy(z[k])) for k in z.files if k.startswith(p))  # noqa: E731
    opt = O.Opt(label_nc=6, no_instance=True, ngf=8, n_downsample_global=2, n_blocks_global=2, ndf=8, num_D=2,
                use_output_gate=True, pool_size=3)
    dt = torch.float64
    g_par = OrderedDict((k, v.to(dt)) for k, v in part("wG::").items())
    d_par = OrderedDict((k, v.clone().to(dt).requires_grad_(True)) for k, v in part("wD::").items())
    vgg = OrderedDict((k, v.to(dt)) for k, v in O.vgg19_random_state_dict().items())
    pool = O.ImagePool(3)
    for it in range(int(z["iters"])):
        b = part("in%d::" % it)
        random.seed(100 + it)
        losses, _, _ = O.model_forward(opt, g_par, d_par, vgg, b["label"], b["inst"], b["image"], b["mask_in"], dtype=dt,
                                       mask_out=b["mask_out"], pool=pool)
        for a, r in zip(losses, z["losses_%d" % it]):
            assert abs(float(a) - float(r)) <= 2e-5 * abs(float(r)), (it, float(a), float(r))
    _, loss_D = O.step_losses(losses)
    gD = part("gD::")
    for (k, _), gi in zip(d_par.items(), torch.autograd.grad(loss_D, list(d_par.values()))):
        if k.endswith("bias") and float(gD[k].abs().max()) < 1e-3 * float(gD[k[:-4] + "weight"].abs().max()):
            continue
        close(gi.float(), gD[k], 2e-4)


def test_box2mask_ade_flag_set_against_the_reference_class(golden_dir):
    """--norm_layer instance --add_dilated_layers --no_comb (what scripts/train_box2mask_ade.sh adds to the Cityscapes flag
    set): outputs, losses and all 66 parameter gradients against the reference's own MaskTwoStreamConvSwitch_NET
    (oracle/make_golden_box2mask_switch.py ade)."""
    from oracle import box2mask as B2
    from oracle.weights import named_param
    z = np.load(os.path.join(golden_dir, "box2mask_switch_ade_small.npz"))
    sd = {str(n): named_param(str(n), tuple(int(v) for v in str(s_).split(";"))).double().requires_grad_(True)
          for n, s_ in zip(z["param_names"], z["param_shapes"])}
    a = {k[4:]: torch.from_numpy(z[k]).double() for k in z.files if k.startswith("in::")}
    cond, _ = B2.encode_input(6, a["mask_ctx_in"], a["mask_in"], a["cls"])
    outs = B2.two_stream_forward(sd, cond.double(), num_layers=3, n_blocks=2, no_comb=True, norm_layer="instance",
                                 add_dilated_layers=True)
    for got, k in zip(outs, ("comb_logit", "comb_prob", "obj_logit", "obj_prob")):
        ref = torch.from_numpy(z[k]).double()
        assert float((got.detach() - ref).abs().max() / ref.abs().max()) < 2e-5, k
    lc = B2.mask_recon_loss(outs[1], a["label_map"], a["mask_out"])
    lo = B2.obj_recon_loss(outs[3], a["mask_out"], a["mask_obj_inst"])
    assert abs(float(lc) - float(z["loss_comb"])) < 2e-5 and abs(float(lo) - float(z["loss_obj"])) < 2e-5
    names = list(sd)
    grads = torch.autograd.grad(lo + lc, [sd[k] for k in names])
    n_full = n_proj = 0
    for k, g in zip(names, grads):
        if "g::" + k in z.files:
            ref = torch.from_numpy(z["g::" + k]).double()
            assert float((g - ref).abs().max()) <= 2e-3 * float(ref.abs().max()) + 1e-6, k
            n_full += 1
        else:
            s_, a_, p_ = (float(v) for v in z["gs::" + k])
            r = named_param("proj::" + k + ".bias", g.shape).double()
            assert abs(float(g.abs().sum()) - a_) <= 2e-3 * a_, k
            assert abs(float((g * r).sum()) - p_) <= 2e-3 * a_ * 0.05 + 1e-6, k
            n_proj += 1
    assert n_full + n_proj == 66

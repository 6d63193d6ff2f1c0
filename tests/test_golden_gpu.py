"""GPU parity of the network executors DIRECTLY against the golden vectors generated from the reference's own classes
(oracle/make_golden.py -> tests/golden/*.npz): BASELINE config #1 (GlobalGenerator(38,3,64,1,1), 128x256, batch 1),
the gated small generator with its parameter gradients, the LocalEnhancer, and the 15 taps + LSGAN losses of the
MultiscaleDiscriminator.  Tolerance 1e-3 relative to each tensor's max magnitude (north_star), bf16x3 mode."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = torch.as_tensor(a).float().cpu(), torch.as_tensor(b).float().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w::")}
    grads = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("g::")}
    return z, sd, grads


def _operand(ctx, x_nchw, border, grad=False):
    from neurips18_hierchical_image_manipulation_b200 import ops
    n, c, h, w = x_nchw.shape
    op = ops.Operand(ctx, n, h, w, c, border=border, grad=grad)
    ops.in_apply(ctx, x_nchw.permute(0, 2, 3, 1).contiguous().cuda(), None, None, ops.ACT_NONE, out_op=op, reflect=True)
    return op


def _nchw(t):
    return t.permute(0, 3, 1, 2).contiguous()


def test_config1_global_generator_matches_reference_golden(golden_dir):
    """BASELINE config #1, the reference's own CPU-runnable case."""
    from neurips18_hierchical_image_manipulation_b200 import ops
    from neurips18_hierchical_image_manipulation_b200.networks import FlatParams, GlobalGenerator
    z, sd, _ = _load(golden_dir, "g_config1.npz")
    ctx = ops.Ctx("cuda:0", split=True)
    fp = FlatParams(ctx.device)
    net = GlobalGenerator(ctx, fp, 38, 3, 64, 1, 1)
    fp.materialize()
    fp.load_state_dict(sd)
    lab = torch.from_numpy(z["label"].astype(np.float32))
    onehot = torch.zeros(1, 35, 128, 256).scatter_(1, lab.long(), 1.0)
    x = torch.cat((onehot, torch.from_numpy(z["image"])), 1)
    out, _ = net.forward(_operand(ctx, x, 3))
    torch.cuda.synchronize()
    ctx.check_pipeline()
    assert rel(_nchw(out), z["out"]) < 1e-3


def test_small_gated_generator_forward_and_parameter_gradients(golden_dir):
    from neurips18_hierchical_image_manipulation_b200 import ops
    from neurips18_hierchical_image_manipulation_b200.networks import FlatParams, GlobalGenerator
    z, sd, grads = _load(golden_dir, "g_small.npz")
    ctx = ops.Ctx("cuda:0", split=True)
    fp = FlatParams(ctx.device)
    net = GlobalGenerator(ctx, fp, 10, 3, 8, 2, 2, use_output_gate=True)
    fp.materialize()
    fp.load_state_dict(sd)
    x, m, cot = torch.from_numpy(z["x"]), torch.from_numpy(z["mask"]), torch.from_numpy(z["cot"])
    t, tape = net.forward(_operand(ctx, x, 3))
    torch.cuda.synchronize()
    t = _nchw(t).cpu()
    out = (1 - m) * x[:, -3:] + m * t                       # output gate, Pix2Pix_NET.py:96-99
    assert rel(out, z["out"]) < 1e-3
    # d(sum(out * cot)) / d(pre-tanh) = cot * m * (1 - t^2)
    dy = _operand(ctx, cot * m * (1 - t * t), 0, grad=True)
    fp.grad.zero_()
    net.backward(tape, dy_head=dy)
    torch.cuda.synchronize()
    ctx.check_pipeline()
    for k, g_ref in grads.items():
        if k.endswith("bias") and float(g_ref.abs().max()) < 1e-4:
            assert float(fp.params[k].grad.abs().max()) < 1e-4   # bias in front of InstanceNorm: zero gradient
            continue
        assert rel(fp.params[k].grad, g_ref) < 1e-2, k


def test_local_enhancer_matches_reference_golden(golden_dir):
    from neurips18_hierchical_image_manipulation_b200 import ops
    from neurips18_hierchical_image_manipulation_b200.local_enhancer import LocalEnhancer
    from neurips18_hierchical_image_manipulation_b200.networks import FlatParams
    z, sd, _ = _load(golden_dir, "local_small.npz")
    ctx = ops.Ctx("cuda:0", split=True)
    fp = FlatParams(ctx.device)
    net = LocalEnhancer(ctx, fp, 9, 3, 4, 2, 2, 1, 2)
    fp.materialize()
    fp.load_state_dict(sd)
    out, _ = net.forward(_operand(ctx, torch.from_numpy(z["x"]), 3))
    torch.cuda.synchronize()
    ctx.check_pipeline()
    assert rel(_nchw(out), z["out"]) < 1e-3


def test_multiscale_discriminator_taps_and_lsgan_losses(golden_dir):
    from neurips18_hierchical_image_manipulation_b200 import ops
    from neurips18_hierchical_image_manipulation_b200.networks import FlatParams, MultiscaleDiscriminator
    z, sd, _ = _load(golden_dir, "d_small.npz")
    ctx = ops.Ctx("cuda:0", split=True)
    fp = FlatParams(ctx.device)
    net = MultiscaleDiscriminator(ctx, fp, 12, ndf=8, n_layers=3, num_D=3)
    fp.materialize()
    fp.load_state_dict(sd)
    tape = net.forward(_operand(ctx, torch.from_numpy(z["x"]), 0))
    acc = torch.zeros(2, dtype=torch.float64, device="cuda")
    for i, lv in enumerate(tape):
        for j, tap in enumerate(lv["taps"]):
            assert rel(_nchw(tap), z["tap_%d_%d" % (i, j)]) < 1e-3, (i, j)
        pred = lv["taps"][-1]
        ops.mse_sum(ctx, pred, 1.0, 1.0 / pred.numel(), acc, 0)    # GANLoss(target real), losses.py:40-50
        ops.mse_sum(ctx, pred, 0.0, 1.0 / pred.numel(), acc, 1)
    torch.cuda.synchronize()
    ctx.check_pipeline()
    assert abs(float(acc[0]) - float(z["loss_real"])) < 1e-3 * abs(float(z["loss_real"]))
    assert abs(float(acc[1]) - float(z["loss_fake"])) < 1e-3 * abs(float(z["loss_fake"]))


def test_two_stream_generator_forward_and_parameter_gradients(golden_dir):
    """GlobalTwoStreamGenerator(6,3,8,3,2, use_skip, ctx_label, gate, early_add) against the reference's own class."""
    from neurips18_hierchical_image_manipulation_b200 import ops
    from neurips18_hierchical_image_manipulation_b200.networks import FlatParams
    from neurips18_hierchical_image_manipulation_b200.two_stream import GlobalTwoStreamGenerator
    z, sd, grads = _load(golden_dir, "twostream_small.npz")
    ctx = ops.Ctx("cuda:0", split=True)
    fp = FlatParams(ctx.device)
    net = GlobalTwoStreamGenerator(ctx, fp, 6, 3, 8, 3, 2, use_skip=True, which_stream="ctx_label", use_output_gate=True)
    fp.materialize()
    fp.load_state_dict(sd)
    img, label, m, cot = (torch.from_numpy(z[k]) for k in ("img", "label", "mask", "cot"))
    t, tape = net.forward(_operand(ctx, img, 3), _operand(ctx, label, 3), m.cuda())
    torch.cuda.synchronize()
    t = _nchw(t).cpu()
    out = (1 - m) * img[:, :3] + m * t                      # output gate, Pix2Pix_NET.py:243-245
    assert rel(out, z["out"]) < 1e-3
    dy = _operand(ctx, cot * m * (1 - t * t), 0, grad=True)
    fp.grad.zero_()
    net.backward(tape, dy)
    torch.cuda.synchronize()
    ctx.check_pipeline()
    worst = []
    for k, g_ref in grads.items():
        if k.endswith("bias") and float(g_ref.abs().max()) < 1e-4:
            assert float(fp.params[k].grad.abs().max()) < 1e-4
            continue
        worst.append((rel(fp.params[k].grad, g_ref), k))
    assert max(worst)[0] < 1e-2, sorted(worst, reverse=True)[:5]


@pytest.mark.parametrize("name,optkw", [
    ("model_global_gate_edges.npz", dict(label_nc=6, no_instance=False, ngf=8, n_downsample_global=2, n_blocks_global=2,
                                         ndf=8, num_D=2, n_layers_D=3, use_output_gate=True)),
    ("model_shipped_twostream.npz", dict(label_nc=6, no_instance=True, ngf=8, n_downsample_global=3, n_blocks_global=2,
                                         ndf=8, num_D=2, n_layers_D=3, use_output_gate=True, netG="global_twostream",
                                         which_encoder="ctx_label", use_skip=True, no_imgCond=True, mask_gan_input=True)),
    # --no_lsgan --no_ganFeat_loss: vanilla GAN (Sigmoid + BCE folded into hm_bce_sum / hm_bce_grad)
    ("model_global_vanilla_gan.npz", dict(label_nc=6, no_instance=True, ngf=8, n_downsample_global=2, n_blocks_global=2,
                                          ndf=8, num_D=2, n_layers_D=3, use_output_gate=True, no_lsgan=True,
                                          no_ganFeat_loss=True)),
    # which_encoder == 'ctx' (the option's default): context stream only, the discriminator is fed the bare image
    ("model_twostream_ctx.npz", dict(label_nc=6, no_instance=True, ngf=8, n_downsample_global=2, n_blocks_global=2,
                                     ndf=8, num_D=2, n_layers_D=3, use_output_gate=True, netG="global_twostream",
                                     which_encoder="ctx", use_skip=True, mask_gan_input=True)),
    # feat_fusion == 'late_add': floor(n/2) ResnetBlocks per stream before the masked fusion, ceil(n/2) after
    ("model_twostream_late_add.npz", dict(label_nc=6, no_instance=True, ngf=8, n_downsample_global=2, n_blocks_global=3,
                                          ndf=8, num_D=2, n_layers_D=3, use_output_gate=True, netG="global_twostream",
                                          which_encoder="ctx_label", feat_fusion="late_add", use_skip=True)),
    # feat_fusion '*_concat': cat -> ReLU -> 1x1 conv -> InstanceNorm in place of the masked sum
    ("model_twostream_early_concat.npz", dict(label_nc=6, no_instance=True, ngf=8, n_downsample_global=2, n_blocks_global=2,
                                              ndf=8, num_D=2, n_layers_D=3, use_output_gate=True, netG="global_twostream",
                                              which_encoder="ctx_label", feat_fusion="early_concat", use_skip=True)),
    ("model_twostream_late_concat.npz", dict(label_nc=6, no_instance=False, ngf=8, n_downsample_global=2, n_blocks_global=3,
                                             ndf=8, num_D=2, n_layers_D=3, use_output_gate=False, netG="global_twostream",
                                             which_encoder="ctx_label", feat_fusion="late_concat", use_skip=False)),
])
def test_training_step_against_the_reference_models_own_forward(golden_dir, name, optkw):
    """The whole forward of the training step (generated image, the five losses) and the gradient directions against
    the golden vectors produced by the reference's OWN Pix2PixHDModel_condImg.forward on CPU (oracle/make_golden_model.py).
    Gradients are compared by direction (cosine) because single tensors move by percents when one ReLU / L1 sign flips
    on these tiny planes (see tests/test_oracle_golden.py::test_model_level_forward_against_the_reference_model)."""
    from neurips18_hierchical_image_manipulation_b200.models import Options, create_model
    z = np.load(os.path.join(golden_dir, name))
    part = lambda p: {k[len(p):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(p)}  # noqa: E731
    opt = Options(precision="bf16x3", gpu_ids=[0], checkpoints_dir="/tmp/hm_ckpt", name="golden", vgg_weights="random",
                  **optkw)
    model = create_model(opt)
    m = model.module
    m.fpG.load_state_dict(part("wG::"))
    m.fpD.load_state_dict(part("wD::"))
    b = part("in::")
    losses, fake = model(label=b["label"], inst=b["inst"], image=b["image"], feat=None, mask_in=b["mask_in"],
                         mask_out=b["mask_out"], infer=True)
    assert rel(fake, z["fake"]) < 1e-3
    for n_, a, r in zip(m.loss_names, losses, z["losses"]):
        assert abs(float(a) - float(r)) <= 1e-3 * abs(float(r)), (n_, float(a), float(r))
    ld = dict(zip(m.loss_names, [torch.mean(x) for x in losses]))
    m.optimizer_G.zero_grad()
    (ld["G_GAN"] + ld["G_GAN_Feat"] + ld["G_VGG"]).backward()
    gG = {k: p.grad.detach().cpu().clone() for k, p in m.fpG.params.items()}
    m.optimizer_D.zero_grad()
    ((ld["D_fake"] + ld["D_real"]) * 0.5).backward()
    gD = {k: p.grad.detach().cpu().clone() for k, p in m.fpD.params.items()}
    torch.cuda.synchronize()
    m.ctx.check_pipeline()

    def cosine(got, ref):
        keys = [k for k in ref if k.endswith("weight")]
        a = torch.cat([got[k].reshape(-1) for k in keys]).double()
        r = torch.cat([ref[k].reshape(-1) for k in keys]).double()
        return float((a @ r) / (a.norm() * r.norm()))
    assert cosine(gG, part("gG::")) > 0.99
    assert cosine(gD, part("gD::")) > 0.95


def test_image_pool_against_the_reference_models_own_forward(golden_dir):
    """--pool_size 3 (util/image_pool.py through discriminate(..., use_pool=True)): four forwards on different batches with
    constant weights, python's `random` seeded like the golden run (oracle/make_golden_model.py::_run_pool_case); the five
    losses of every forward -- loss_D_fake is evaluated on the pool's answer -- and the discriminator gradient direction
    of the last one."""
    import random
    from neurips18_hierchical_image_manipulation_b200.models import Options, create_model
    z = np.load(os.path.join(golden_dir, "model_global_pool.npz"))
    part = lambda p: {k[len(p):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(p)}  # noqa: E731
    opt = Options(precision="bf16x3", gpu_ids=[0], checkpoints_dir="/tmp/hm_ckpt", name="golden_pool", vgg_weights="random",
                  label_nc=6, no_instance=True, ngf=8, n_downsample_global=2, n_blocks_global=2, ndf=8, num_D=2, n_layers_D=3,
                  use_output_gate=True, pool_size=3)
    model = create_model(opt)
    m = model.module
    m.fpG.load_state_dict(part("wG::"))
    m.fpD.load_state_dict(part("wD::"))
    for it in range(int(z["iters"])):
        b = part("in%d::" % it)
        random.seed(100 + it)
        losses, _ = model(label=b["label"], inst=b["inst"], image=b["image"], feat=None, mask_in=b["mask_in"],
                          mask_out=b["mask_out"], infer=False)
        for n_, a, r in zip(m.loss_names, losses, z["losses_%d" % it]):
            assert abs(float(a) - float(r)) <= 1e-3 * abs(float(r)), (it, n_, float(a), float(r))
    ld = dict(zip(m.loss_names, [torch.mean(x) for x in losses]))
    m.optimizer_D.zero_grad()
    ((ld["D_fake"] + ld["D_real"]) * 0.5).backward()
    torch.cuda.synchronize()
    m.ctx.check_pipeline()
    ref = part("gD::")
    keys = [k for k in ref if k.endswith("weight")]
    a = torch.cat([m.fpD.params[k].grad.detach().cpu().reshape(-1) for k in keys]).double()
    r = torch.cat([ref[k].reshape(-1) for k in keys]).double()
    assert float((a @ r) / (a.norm() * r.norm())) > 0.95


def test_image_pool_fused_step_graph_replay_matches_eager():
    """The pool inside the fused step: its exchanges are kernel launches driven by a device-resident decision tensor, so
    the CUDA-graph replay and the eager step see the same history (same python `random` seed): losses agree."""
    import random
    from neurips18_hierchical_image_manipulation_b200.models import Options, create_model
    from neurips18_hierchical_image_manipulation_b200.synthetic import synthetic_batch
    runs = []
    for graph in (False, True):
        opt = Options(precision="bf16x3", gpu_ids=[0], checkpoints_dir="/tmp/hm_ckpt", name="pool_graph", vgg_weights="random",
                      label_nc=6, no_instance=True, ngf=8, n_downsample_global=2, n_blocks_global=2, ndf=8, num_D=2,
                      n_layers_D=3, use_output_gate=True, pool_size=3, cuda_graph=graph)
        m = create_model(opt).module
        random.seed(7)
        ls = []
        for i in range(6):
            d = synthetic_batch(2, 64, 96, 6, seed=50 + i)
            ls.append(m.optimize_parameters(label=d["label"], inst=d["inst"], image=d["image"], feat=None, mask_in=d["mask_in"],
                                            mask_out=d["mask_out"]).clone())
        torch.cuda.synchronize()
        m.ctx.check_pipeline()
        assert isinstance(m._graph, dict) == graph and m._pool["num"] == 3
        runs.append(torch.stack(ls).cpu())
    assert torch.allclose(runs[0], runs[1], rtol=5e-3, atol=1e-5), (runs[0], runs[1])

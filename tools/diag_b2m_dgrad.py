"""Diagnostic: where does the discriminator weight-gradient error of box2mask --use_gan at config #5 geometry sit?
(float64 oracle on the product's own inputs vs the product; error of scale0_layer3.0.weight by tap and by channel)"""
import contextlib
import io
import os
import sys

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from oracle import box2mask as B2                      # noqa: E402
import make_golden_box2mask as G                      # noqa: E402
from neurips18_hierchical_image_manipulation_b200.models import Options, create_model   # noqa: E402

torch.set_num_threads(os.cpu_count() or 1)
S = int(os.environ.get("S", "256"))
with contextlib.redirect_stdout(io.StringIO()):
    m = create_model(Options(model="AE_maskgen_twostream", isTrain=False, gpu_ids=[0], precision="bf16x3", name="b2m", num_layers=3,
                             conv_size=4, which_stream="obj_context", cond_in="ctx_obj", use_output_gate=True, num_resnetblocks=1,
                             norm_layer="batch", label_nc=35, output_nc=35, conv_dim=64, n_blocks=6, use_gan=True,
                             which_gan="patch_multiscale", gan_weight=0.1, num_layers_D=3, ndf=64, use_ganFeat_loss=True,
                             lambda_feat=1.0, cuda_graph=False))
sdD = {k: v.detach().cpu().clone() for k, v in m.fpD.params.items()}
d = G.synthetic(dict(label_nc=35, fineSize=S), 2, seed=5)
cond, _ = B2.encode_input(35, d["mask_ctx_in"], d["mask_in"], d["cls"])
ls, out = m.forward(d["label_map"], None, d["mask_ctx_in"], None, d["mask_out"], d["mask_obj_inst"], d["cls"], d["mask_in"], train=False)
p64 = {k: v.double().requires_grad_(True) for k, v in sdD.items()}
mo = d["mask_out"].double()
c = cond.double() * mo
real = torch.cat((d["mask_obj_inst"].double() * mo, c), 1)
fake = torch.cat((out["obj_prob"].detach().cpu().double() * mo * mo, c), 1)
which = os.environ.get("WHICH", "both")
loss = 0
m.optimizer_D.zero_grad()
if which in ("both", "real"):
    loss = loss + 0.5 * B2.lsgan(B2.multiscale_discriminator_bn_forward(p64, real, 2, 3), True)
    m.netD.backward(m._last["d_real"], 1.0, 0.5, True)
if which in ("both", "fake"):
    loss = loss + 0.5 * B2.lsgan(B2.multiscale_discriminator_bn_forward(p64, fake, 2, 3), False)
    m.netD.backward(m._last["d_fake"], 0.0, 0.5, True)
torch.cuda.synchronize()
gD = dict(zip(p64, torch.autograd.grad(loss, list(p64.values()), allow_unused=True)))
# forward taps first
for name, x, tape in (("real", real, m._last["d_real"]), ("fake", fake, m._last["d_fake"])):
    taps = B2.multiscale_discriminator_bn_forward({k: v.detach() for k, v in p64.items()}, x, 2, 3)
    for i, (sc, lv) in enumerate(zip(taps, tape)):
        errs = ["%.1e" % float((lv["taps"][j].permute(0, 3, 1, 2).double().cpu() - t).abs().max() / t.abs().max()) for j, t in enumerate(sc)]
        print("forward taps", name, "level", i, errs, "shapes", [tuple(t.shape[2:]) for t in sc])
for k in sorted(gD):
    ref = gD[k]
    g = m.fpD.params[k].grad.detach().double().cpu()
    if ref is None or ref.abs().max() < 1e-12:
        continue
    e = (g - ref).abs()
    line = "%-26s max %.1e  2-norm %.1e" % (k, float(e.max() / ref.abs().max()), float((g - ref).norm() / ref.norm()))
    if ref.dim() == 4 and float(e.max() / ref.abs().max()) > 1e-3:
        by_tap = (e.amax(dim=(0, 1)) / ref.abs().max()).flatten()
        by_co = e.amax(dim=(1, 2, 3)) / ref.abs().max()
        by_ci = e.amax(dim=(0, 2, 3)) / ref.abs().max()
        line += "\n    by tap: " + " ".join("%.0e" % float(v) for v in by_tap)
        line += "\n    worst co: %s  worst ci: %s" % (by_co.topk(4).indices.tolist(), by_ci.topk(4).indices.tolist())
        line += "  frac(co err > 1e-3): %.3f  frac(ci err > 1e-3): %.3f" % (float((by_co > 1e-3).float().mean()), float((by_ci > 1e-3).float().mean()))
    print(line)

"""d(loss)/d(fake) through D (GAN + feature matching) and through VGG: product vs oracle, full-width networks."""
import sys, os, torch, contextlib, io
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from neurips18_hierchical_image_manipulation_b200.models import Options, create_model, random_vgg19_state_dict, VGG_WEIGHTS
from neurips18_hierchical_image_manipulation_b200.synthetic import synthetic_batch
from oracle import model as O
import torch.nn.functional as F
H, W = int(sys.argv[1]), int(sys.argv[2])
ngf = int(sys.argv[3]) if len(sys.argv) > 3 else 8
opt = Options(vgg_weights="random", label_nc=35, no_instance=True, netG="global", ngf=ngf, n_downsample_global=2, n_blocks_global=1, num_D=3,
              gpu_ids=[0], precision="bf16x3", name="fg", checkpoints_dir="/tmp/hm_fg")
with contextlib.redirect_stdout(io.StringIO()):
    m = create_model(opt).module
vgg = random_vgg19_state_dict(opt.vgg_seed)
torch.set_num_threads(os.cpu_count())
B = 1
b = synthetic_batch(B, H, W, 35, seed=99)
st = m._forward_all(b["label"], b["inst"], b["image"], b["mask_in"]); m._step = st
torch.cuda.synchronize()
fake = st["fake"].cpu().clone().requires_grad_(True)
d_sd = m.fpD.state_dict()
input_mask, real, cond = O.encode_input(b["label"], b["inst"], b["image"], b["mask_in"], 35, True)
input_label = torch.cat((input_mask, cond), 1)
pf = O.multiscale_discriminator_forward(d_sd, torch.cat((input_label, fake), 1), 3, 3)
pr = O.multiscale_discriminator_forward(d_sd, torch.cat((input_label, real), 1), 3, 3)
lg = O.gan_loss(pf, True)
lf = 0
for i in range(3):
    for j in range(4):
        lf = lf + (1.0 / 3) * (4.0 / 4) * F.l1_loss(pf[i][j], pr[i][j].detach()) * 10.0
(gd_gan,) = torch.autograd.grad(lg, fake, retain_graph=True)
(gd_feat,) = torch.autograd.grad(lf, fake)
lv = O.vgg_loss(vgg, fake, real) * 10.0
(gv_ref,) = torch.autograd.grad(lv, fake)
nc = m.netG_input_nc
def rel(a, r): return float((a - r).norm() / r.norm())
g1 = m.netD.backward(st["d_tape"], B, "G", w_gan=1.0, w_feat=0.0, img_c0=nc)
torch.cuda.synchronize()
print("D GAN-only  grad err %.3e" % rel(g1[..., 0:3].cpu().permute(0, 3, 1, 2), gd_gan))
g2 = m.netD.backward(st["d_tape"], B, "G", w_gan=0.0, w_feat=(1.0 / 3) * 1.0 * 10.0, img_c0=nc)
torch.cuda.synchronize()
print("D feat-only grad err %.3e" % rel(g2[..., 0:3].cpu().permute(0, 3, 1, 2), gd_feat))
gV = m.vgg.backward(st["v_tape"], B, [10.0 * w for w in VGG_WEIGHTS])
torch.cuda.synchronize()
print("VGG grad err %.3e" % rel(gV.cpu().permute(0, 3, 1, 2), gv_ref))
for t in range(5):
    co = [0.0] * 5; co[t] = 10.0 * VGG_WEIGHTS[t]
    gVt = m.vgg.backward(st["v_tape"], B, co)
    xv, yv = O.vgg19_forward(vgg, fake), O.vgg19_forward(vgg, real)
    (gr,) = torch.autograd.grad(co[t] * F.l1_loss(xv[t], yv[t].detach()), fake)
    torch.cuda.synchronize()
    print("  VGG tap %d only: err %.3e" % (t, rel(gVt.cpu().permute(0, 3, 1, 2), gr)))

"""Layer-by-layer comparison of the VGG19 backward (dL/d relu-output of every conv) against torch autograd."""
import sys, os, torch, torch.nn.functional as F
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from neurips18_hierchical_image_manipulation_b200 import ops
from neurips18_hierchical_image_manipulation_b200.networks import Vgg19, VGG19_CONVS, VGG19_POOL_BEFORE, VGG19_SLICE_OF, VGG19_TAP_AFTER
from neurips18_hierchical_image_manipulation_b200.models import random_vgg19_state_dict, VGG_WEIGHTS
H, W = int(sys.argv[1]), int(sys.argv[2])
ctx = ops.Ctx("cuda:0", split=True)
sd = random_vgg19_state_dict(1234)
vgg = Vgg19(ctx, sd)
torch.manual_seed(0)
fake = torch.rand(1, 3, H, W) * 2 - 1
real = torch.rand(1, 3, H, W) * 2 - 1
v_in = ops.Operand(ctx, 2, H, W, 3)
x = torch.cat((fake, real), 0).permute(0, 2, 3, 1).contiguous().cuda()
ops.in_apply(ctx, x, None, None, ops.ACT_NONE, out_op=v_in)
tape = vgg.forward(v_in)
coefs = [10.0 * w for w in VGG_WEIGHTS]
# --- instrumented copy of Vgg19.backward
rec = {}
g = None
nb = 1
for li in range(len(vgg.convs_) - 1, -1, -1):
    idx, conv = vgg.convs_[li]
    out = tape["outs"][li]
    shape = (nb, out.h, out.w, conv.cout)
    dy = ops.Operand(ctx, nb, out.h, out.w, conv.cout)
    tap = tape["taps"].get(li)
    rec[("dz", li)] = None if g is None else g.clone()
    if tap is not None:
        l1 = coefs[VGG19_TAP_AFTER[idx]] / (tap.numel() // 2)
        ops.in_bwd(ctx, shape, ops.ACT_RELU, z=tap[:nb], g2=g, tref=tap[nb:], l1coef=l1, out_op=dy)
    else:
        ops.in_bwd(ctx, shape, ops.ACT_RELU, mask_op=out, g2=g, out_op=dy)
    rec[("dy", li)] = dy.dense()
    xin = tape["xs"][li]
    gin = torch.empty(nb, xin.h, xin.w, conv.cin, device="cuda")
    conv.dgrad(dy, xin.h, xin.w, 1, gin)
    rec[("gin", li)] = gin.clone()
    if li in tape["pooled_from"]:
        src = tape["pooled_from"][li]
        dz = torch.empty(nb, src.h, src.w, src.c, device="cuda")
        ops.maxpool2_bwd(ctx, gin, src, dz)
        g = dz
    else:
        g = gin
torch.cuda.synchronize()
# --- torch reference with hooks
f = fake.clone().requires_grad_(True)
acts, pre = [], []
h = torch.cat((f, real), 0)
taps = []
for li, (idx, cin, cout) in enumerate(VGG19_CONVS):
    if idx in VGG19_POOL_BEFORE:
        h = F.max_pool2d(h, 2, 2)
    k = "slice%d.%d." % (VGG19_SLICE_OF[idx], idx)
    p_ = F.conv2d(h, sd[k + "weight"], sd[k + "bias"], padding=1); p_.retain_grad(); pre.append(p_)
    h = F.relu(p_); h.retain_grad(); acts.append(h)
    if idx in VGG19_TAP_AFTER:
        taps.append(h)
loss = 0
for i, t in enumerate(taps):
    loss = loss + coefs[i] * F.l1_loss(t[:1], t[1:].detach())
loss.backward()
def rel(a, r): return float((a - r).norm() / r.norm().clamp_min(1e-30))
for li in range(len(VGG19_CONVS) - 1, -1, -1):
    ref_dy = pre[li].grad[:1]
    mine = rec[("dy", li)].cpu()
    fwd = rel(tape["outs"][li].dense().cpu()[:1], acts[li].detach()[:1])
    print("layer %2d (conv idx %2d, %3dch %3dx%-3d)  fwd err %.2e   dL/dpre err %.3e" % (li, VGG19_CONVS[li][0], VGG19_CONVS[li][2], mine.shape[2], mine.shape[3], fwd, rel(mine, ref_dy)))
print("input grad err %.3e" % rel(g.cpu().permute(0, 3, 1, 2), f.grad))

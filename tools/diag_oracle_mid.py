"""Compare product vs oracle gradients at a mid size with the FULL architecture, for B=1 and B=2."""
import sys, os, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import contextlib, io
from neurips18_hierchical_image_manipulation_b200.models import Options, create_model, random_vgg19_state_dict
from neurips18_hierchical_image_manipulation_b200.synthetic import synthetic_batch
from oracle import model as O
H, W = int(sys.argv[1]) if len(sys.argv) > 1 else 128, int(sys.argv[2]) if len(sys.argv) > 2 else 256
opt = Options(vgg_weights="random", label_nc=35, no_instance=True, netG="global", ngf=64, n_downsample_global=4, n_blocks_global=9, num_D=3,
              gpu_ids=[0], precision="bf16x3", name="mid", checkpoints_dir="/tmp/hm_mid")
with contextlib.redirect_stdout(io.StringIO()):
    m = create_model(opt).module
oopt = O.Opt(num_D=3)
vgg = random_vgg19_state_dict(opt.vgg_seed)
torch.set_num_threads(os.cpu_count())
for B in (1,):
    b = synthetic_batch(B, H, W, 35, seed=99)
    g_sd, d_sd = m.fpG.state_dict(), m.fpD.state_dict()
    ls, fake_ref, gG, gD, _ = O.train_step(oopt, {k: v.clone() for k, v in g_sd.items()}, {k: v.clone() for k, v in d_sd.items()}, vgg, b)
    st = m._forward_all(b["label"], b["inst"], b["image"], b["mask_in"]); m._step = st
    m.flat_grad.zero_(); m._backward_G([1.0, 1.0, 1.0]); m._backward_D([0.5, 0.5]); torch.cuda.synchronize()
    print("B=%d fake err %.2e  losses mine %s ref %s" % (B, float((st["fake"].cpu() - fake_ref).abs().max() / fake_ref.abs().max()),
          ["%.5f" % x for x in st["losses"].tolist()], ["%.5f" % x for x in ls]))
    worst = []
    for fp, ref in ((m.fpG, gG), (m.fpD, gD)):
        for k, p in fp.params.items():
            r = ref[k]
            if k.endswith("bias") and float(r.abs().max()) < 1e-5: continue
            worst.append((float((p.grad.cpu() - r).abs().max() / r.abs().max().clamp_min(1e-30)), k))
    worst.sort(reverse=True)
    print("  worst:", ["%.2e %s" % w for w in worst[:6]])
    if B == 1:
        for k, pp in m.fpG.params.items():
            if k.endswith("weight"):
                r = gG[k]
                print("   %-28s err %.2e  |g| %.2e" % (k, float((pp.grad.cpu() - r).abs().max() / r.abs().max()), float(r.abs().max())))

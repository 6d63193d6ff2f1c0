"""Generate tests/golden/model_*.npz from the REFERENCE'S OWN model-level forward
(models/pix2pixHD_condImg_model.py: Pix2PixHDModel_condImg.__init__ / encode_input / discriminate / forward), run once
in the build container:   python oracle/make_golden_model.py          (needs /root/reference; CPU only)

The reference model hard-codes `.cuda()` / `torch.cuda.FloatTensor` and imports the py2-only `util.util`
(SURVEY.md section 8(c)), so it cannot be imported as is.  Nothing of it is copied or edited: this script only
injects, at run time, (1) an empty stand-in for the unused `util.util` module, (2) identity `.cuda()` methods and
`torch.cuda.FloatTensor = torch.FloatTensor`, and (3) a `torchvision.models.vgg19` that returns the seeded random
VGG19 the product and the oracle use (no network for the ImageNet weights).  The reference's own code then runs
unmodified on CPU: one forward of the training step, `loss_G.backward()` and `loss_D.backward()` exactly as
train_mask2image.py:68-86 combines the losses.  Two cases: the plain `global` generator with instance edges and the
output gate, and the flag set of scripts/train_mask2image_city.sh (two-stream generator, skip connections, output
gate, --no_imgCond, --mask_gan_input).  TEST INFRASTRUCTURE ONLY.
"""
import os
import sys
import types
from collections import OrderedDict

import numpy as np
import torch

REF = os.environ.get("HM_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden")
sys.path.insert(0, os.path.join(HERE, ".."))


def _prepare_reference_imports():
    sys.path.insert(0, os.path.join(REF, "models"))
    sys.path.insert(0, REF)
    stub = types.ModuleType("util.util")           # pix2pixHD_condImg_model.py:15 imports it, forward never uses it
    import util                                     # the reference's package (util/__init__.py is importable)
    sys.modules["util.util"] = stub
    util.util = stub
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    torch.cuda.FloatTensor = torch.FloatTensor
    torch.cuda.ByteTensor = torch.ByteTensor          # get_edges, pix2pixHD_condImg_model.py:286
    import torchvision
    from oracle import model as O
    tv_vgg19 = torchvision.models.vgg19

    def vgg19(pretrained=False, **kw):
        net = tv_vgg19(weights=None)
        sd = O.vgg19_random_state_dict()
        with torch.no_grad():
            for idx, _, _ in O.VGG19_CONVS:
                k = "slice%d.%d." % (O.VGG19_SLICE_OF[idx], idx)
                net.features[idx].weight.copy_(sd[k + "weight"])
                net.features[idx].bias.copy_(sd[k + "bias"])
        return net
    import layer_util
    layer_util.models.vgg19 = vgg19                 # layer_util.py:384 calls models.vgg19(pretrained=True)


def _opt(**kw):
    d = dict(name="golden", gpu_ids=[], checkpoints_dir="/tmp/hm_golden_ckpt", model="pix2pixHD_condImg", norm="instance",
             isTrain=True, resize_or_crop="none", netG="global", instance_feat=False, label_feat=False,
             load_features=False, label_nc=6, no_instance=False, feat_num=3, output_nc=3, ngf=8, n_downsample_global=2,
             n_blocks_global=2, use_output_gate=False, use_skip=False, which_encoder="ctx", feat_fusion="early_add",
             no_imgCond=False, mask_gan_input=False, use_soft_mask=False, no_lsgan=False, ndf=8, n_layers_D=3, num_D=2,
             no_ganFeat_loss=False, continue_train=False, load_pretrain="", which_epoch="latest", pool_size=0, lr=0.0002,
             beta1=0.5, no_vgg_loss=False, niter_fix_global=0, n_local_enhancers=1, lambda_feat=10.0, lambda_rec=0.0,
             niter_decay=100, nef=16, n_downsample_E=4)
    d.update(kw)
    return types.SimpleNamespace(**d)


def _batch(B, H, W, label_nc, seed):
    g = torch.Generator().manual_seed(seed)
    lab = torch.randint(0, label_nc, (B, 1, H // 8, W // 8), generator=g).float()
    label = torch.nn.functional.interpolate(lab, size=(H, W), mode="nearest")
    ins = torch.randint(0, 5, (B, 1, H // 8, W // 8), generator=g).float()
    inst = torch.nn.functional.interpolate(ins, size=(H, W), mode="nearest")
    image = torch.rand(B, 3, H, W, generator=g) * 2 - 1
    mask_in = torch.zeros(B, 1, H, W)
    mask_out = torch.zeros(B, 1, H, W)
    for b in range(B):
        y0, x0 = 8 + 8 * b, 16 + 8 * b
        mask_in[b, :, y0:y0 + H // 3, x0:x0 + W // 3] = 1
        mask_out[b, :, max(0, y0 - 4):y0 + H // 3 + 4, max(0, x0 - 4):x0 + W // 3 + 4] = 1
    return dict(label=label, inst=inst, image=image, mask_in=mask_in, mask_out=mask_out)


def _run_case(name, seed, **kw):
    from models.pix2pixHD_condImg_model import Pix2PixHDModel_condImg
    torch.manual_seed(seed)
    opt = _opt(**kw)
    model = Pix2PixHDModel_condImg(opt)             # the reference's own constructor: weights_init on G and D
    batch = _batch(2, 64, 96, opt.label_nc, seed + 100)
    losses, fake = model.forward(batch["label"], batch["inst"], batch["image"], None, batch["mask_in"], batch["mask_out"],
                                 infer=True)
    losses = [torch.mean(x) if not isinstance(x, (int, float)) else torch.tensor(float(x)) for x in losses]   # :68
    ld = dict(zip(model.loss_names, losses))
    loss_D = (ld["D_fake"] + ld["D_real"]) * 0.5                                                                # :72
    loss_G = ld["G_GAN"] + ld["G_GAN_Feat"] + ld["G_VGG"]                                                       # :73
    model.optimizer_G.zero_grad()
    loss_G.backward(retain_graph=True)
    gG = OrderedDict((k, p.grad.detach().clone()) for k, p in model.netG.named_parameters())
    model.optimizer_D.zero_grad()
    loss_D.backward()
    gD = OrderedDict((k, p.grad.detach().clone()) for k, p in model.netD.named_parameters())
    out = dict(fake=fake.detach().numpy(), losses=np.array([float(x) for x in losses], dtype=np.float64))
    out.update({"in::" + k: v.numpy() for k, v in batch.items()})
    out.update({"wG::" + k: v.detach().numpy() for k, v in model.netG.state_dict().items()})
    out.update({"wD::" + k: v.detach().numpy() for k, v in model.netD.state_dict().items()})
    out.update({"gG::" + k: v.numpy() for k, v in gG.items()})
    out.update({"gD::" + k: v.numpy() for k, v in gD.items()})
    path = os.path.join(OUT, "model_%s.npz" % name)
    np.savez_compressed(path, **out)
    print(name, os.path.getsize(path), [round(float(x), 5) for x in losses])


def _run_pool_case(name, seed, iters=4, **kw):
    """--pool_size > 0 (util/image_pool.py:4-31 through discriminate(..., use_pool=True), :176-186,218): the discriminator's
    fake pass sees a history of generated inputs.  `iters` forwards on DIFFERENT batches with constant weights; python's
    `random` (which the pool draws from) is seeded with 100 + it before each forward.  Stored: the five losses of every
    forward and the discriminator gradients of the last one (its loss_D_fake is evaluated on pooled images)."""
    import random
    from models.pix2pixHD_condImg_model import Pix2PixHDModel_condImg
    torch.manual_seed(seed)
    opt = _opt(**kw)
    model = Pix2PixHDModel_condImg(opt)
    out = dict(iters=np.array(iters))
    for it in range(iters):
        batch = _batch(2, 64, 96, opt.label_nc, seed + 100 + it)
        random.seed(100 + it)
        losses, _ = model.forward(batch["label"], batch["inst"], batch["image"], None, batch["mask_in"], batch["mask_out"],
                                  infer=False)
        losses = [torch.mean(x) if not isinstance(x, (int, float)) else torch.tensor(float(x)) for x in losses]
        out["losses_%d" % it] = np.array([float(x) for x in losses], dtype=np.float64)
        out.update({"in%d::%s" % (it, k): v.numpy() for k, v in batch.items()})
    ld = dict(zip(model.loss_names, losses))
    model.optimizer_D.zero_grad()
    ((ld["D_fake"] + ld["D_real"]) * 0.5).backward()
    out.update({"gD::" + k: p.grad.detach().numpy().copy() for k, p in model.netD.named_parameters()})
    out.update({"wG::" + k: v.detach().numpy() for k, v in model.netG.state_dict().items()})
    out.update({"wD::" + k: v.detach().numpy() for k, v in model.netD.state_dict().items()})
    path = os.path.join(OUT, "model_%s.npz" % name)
    np.savez_compressed(path, **out)
    print(name, os.path.getsize(path), [[round(float(x), 5) for x in out["losses_%d" % it]] for it in range(iters)])


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    _prepare_reference_imports()
    _run_case("global_gate_edges", 21, netG="global", use_output_gate=True, no_instance=False)
    _run_case("shipped_twostream", 22, netG="global_twostream", which_encoder="ctx_label", use_skip=True,
              use_output_gate=True, no_imgCond=True, mask_gan_input=True, no_instance=True, n_downsample_global=3)
    if "ctx" in sys.argv[1:] or len(sys.argv) == 1:
        # which_encoder == 'ctx' (the option's DEFAULT): context stream only, and the discriminator sees the bare image
        # (pix2pixHD_condImg_model.py:71-72, 178-179, 227-228)
        _run_case("twostream_ctx", 23, netG="global_twostream", which_encoder="ctx", use_skip=True, use_output_gate=True,
                  mask_gan_input=True, no_instance=True, n_downsample_global=2)
    if "late" in sys.argv[1:] or len(sys.argv) == 1:
        # feat_fusion == 'late_add' (Pix2Pix_NET.py:137-142, 203-206): each stream runs floor(n/2) ResnetBlocks of its own
        # before the masked fusion, ceil(n/2) blocks embed the fused feature
        _run_case("twostream_late_add", 24, netG="global_twostream", which_encoder="ctx_label", feat_fusion="late_add",
                  use_skip=True, use_output_gate=True, no_instance=True, n_downsample_global=2, n_blocks_global=3)
    if "vanilla" in sys.argv[1:] or len(sys.argv) == 1:
        # --no_lsgan (GANLoss with nn.BCELoss, losses.py:8-20) is only well defined together with --no_ganFeat_loss: the
        # Sigmoid the discriminator appends (Discriminator_NET.py:95-96) is never applied by its getIntermFeat branch
        # (:111-114 loops over n_layers + 2 sub-models)
        _run_case("global_vanilla_gan", 28, netG="global", use_output_gate=True, no_instance=True, no_lsgan=True,
                  no_ganFeat_loss=True)
    if "pool" in sys.argv[1:] or len(sys.argv) == 1:
        _run_pool_case("global_pool", 27, netG="global", use_output_gate=True, no_instance=True, pool_size=3)
    if "concat" in sys.argv[1:] or len(sys.argv) == 1:
        # feat_fusion '*_concat' (layer_util.py:305-327): cat -> ReLU -> 1x1 conv -> norm instead of the sum
        _run_case("twostream_early_concat", 25, netG="global_twostream", which_encoder="ctx_label", feat_fusion="early_concat",
                  use_skip=True, use_output_gate=True, no_instance=True, n_downsample_global=2, n_blocks_global=2)
        _run_case("twostream_late_concat", 26, netG="global_twostream", which_encoder="ctx_label", feat_fusion="late_concat",
                  use_skip=False, use_output_gate=False, no_instance=False, n_downsample_global=2, n_blocks_global=3)


if __name__ == "__main__":
    main()

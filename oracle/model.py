"""CPU restatement (torch functional ops, fp32 or fp64) of the reference's mask2image hot path.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Every function cites the reference lines it follows
(paths relative to the reference tree).  Networks are evaluated functionally from a state_dict that uses the
reference's own parameter names, so reference checkpoints / golden weights drop in unchanged.
"""
from collections import OrderedDict

import torch
import torch.nn.functional as F

NULLVAL = 0.0  # models/pix2pixHD_condImg_model.py:18

# torchvision VGG19 "E" configuration, features[0:30] (models/layer_util.py:384-399):
# conv indices inside `features` and the slice (tap) each belongs to.
VGG19_CONVS = [  # (features index, cin, cout)
    (0, 3, 64), (2, 64, 64), (5, 64, 128), (7, 128, 128), (10, 128, 256), (12, 256, 256), (14, 256, 256),
    (16, 256, 256), (19, 256, 512), (21, 512, 512), (23, 512, 512), (25, 512, 512), (28, 512, 512)]
VGG19_POOL_BEFORE = {5, 10, 19, 28}      # MaxPool2d(2,2) sits right before these conv indices
VGG19_TAP_AFTER = {0: 0, 5: 1, 10: 2, 19: 3, 28: 4}  # relu after conv idx -> tap number (relu1_1 ... relu5_1)
VGG19_SLICE_OF = {0: 1, 2: 2, 5: 2, 7: 3, 10: 3, 12: 4, 14: 4, 16: 4, 19: 4, 21: 5, 23: 5, 25: 5, 28: 5}
VGG_WEIGHTS = [1.0 / 32, 1.0 / 16, 1.0 / 8, 1.0 / 4, 1.0]  # models/losses.py:57


def instance_norm(x, eps=1e-5):
    """nn.InstanceNorm2d(C, affine=False): biased variance over each (n, c) plane (models/layer_util.py:19-26)."""
    mean = x.mean(dim=(2, 3), keepdim=True)
    var = x.var(dim=(2, 3), unbiased=False, keepdim=True)
    return (x - mean) / torch.sqrt(var + eps)


def reflect_pad(x, p):
    return F.pad(x, (p, p, p, p), mode="reflect")


# --------------------------------------------------------------------------------------------------
# Generators
# --------------------------------------------------------------------------------------------------
def global_generator_forward(sd, x, n_downsampling, n_blocks, mask=None, use_output_gate=False, input_nc=None,
                             prefix="model.", stop_before_head=False):
    """GlobalGenerator.forward, models/Pix2Pix_NET.py:63-101 (ResnetBlock: models/layer_util.py:333-378)."""
    def P(i, name):
        return sd["%s%d.%s" % (prefix, i, name)]
    h = F.conv2d(reflect_pad(x, 3), P(1, "weight"), P(1, "bias"))             # :74
    h = F.relu(instance_norm(h))
    idx = 4
    for _ in range(n_downsampling):                                            # :76-79
        h = F.relu(instance_norm(F.conv2d(h, P(idx, "weight"), P(idx, "bias"), stride=2, padding=1)))
        idx += 3
    for _ in range(n_blocks):                                                  # :82-84
        r = F.conv2d(reflect_pad(h, 1), P(idx, "conv_block.1.weight"), P(idx, "conv_block.1.bias"))
        r = F.relu(instance_norm(r))
        r = F.conv2d(reflect_pad(r, 1), P(idx, "conv_block.5.weight"), P(idx, "conv_block.5.bias"))
        h = h + instance_norm(r)                                               # layer_util.py:376-378 (no relu)
        idx += 1
    for _ in range(n_downsampling):                                            # :87-90
        h = F.conv_transpose2d(h, P(idx, "weight"), P(idx, "bias"), stride=2, padding=1, output_padding=1)
        h = F.relu(instance_norm(h))
        idx += 3
    if stop_before_head:                                                       # LocalEnhancer drops the last 3 modules (:18)
        return h
    out = torch.tanh(F.conv2d(reflect_pad(h, 3), P(idx + 1, "weight"), P(idx + 1, "bias")))  # :91
    if use_output_gate and mask is not None:                                   # :96-99
        nc = x.shape[1] if input_nc is None else input_nc
        img = x[:, nc - 3:, :, :]
        m = mask.repeat(1, out.shape[1], 1, 1)
        out = (1 - m) * img + m * out
    return out


def global_twostream_forward(sd, img, label, mask, n_downsampling, n_blocks, use_skip=False, which_stream="ctx",
                             use_output_gate=False, feat_fusion="early_add"):
    """GlobalTwoStreamGenerator.forward, models/Pix2Pix_NET.py:103-247 (feat_fusion 'early_add', 'late_add', 'early_concat', 'late_concat')."""
    def embed(prefix, h, n):                                                   # get_embedder :166-175
        for i in range(n):
            k = "%s.%d.conv_block." % (prefix, i)
            r = F.conv2d(reflect_pad(h, 1), sd[k + "1.weight"], sd[k + "1.bias"])
            r = F.relu(instance_norm(r))
            r = F.conv2d(reflect_pad(r, 1), sd[k + "5.weight"], sd[k + "5.bias"])
            h = h + instance_norm(r)
        return h
    late = "late" in feat_fusion
    if feat_fusion not in ("early_add", "late_add", "early_concat", "late_concat"):
        raise NotImplementedError(feat_fusion)
    def encode(prefix, x, keep):                                               # forward_encoder :135-142
        h = F.relu(instance_norm(F.conv2d(reflect_pad(x, 3), sd[prefix + "_inputEmbedder.1.weight"],
                                          sd[prefix + "_inputEmbedder.1.bias"])))
        feats = []
        for i in range(n_downsampling):
            k = prefix + "_downsampler.%d." % (3 * i)
            h = F.relu(instance_norm(F.conv2d(h, sd[k + "weight"], sd[k + "bias"], stride=2, padding=1)))
            if keep and i < n_downsampling - 1:                                # :140-141
                feats.append(h)
        return h, feats
    ctx_feat = obj_feat = None
    ctx_feats = []
    if "ctx" in which_stream:
        ctx_feat, ctx_feats = encode("ctx", img, use_skip)
    if "label" in which_stream:
        obj_feat, _ = encode("obj", label, False)
    if which_stream == "ctx_label":                                            # :203-209, FeatureFusionBlock 'add'
        if late:                                                               # :204-206, n_blocks split :138-142
            ctx_feat = embed("ctx_latent_embedder", ctx_feat, n_blocks // 2)
            obj_feat = embed("obj_latent_embedder", obj_feat, n_blocks // 2)
        f = 2 ** n_downsampling
        m = F.max_pool2d(mask, f, f)
        if "concat" in feat_fusion:                                            # FeatureFusionBlock 'concat', layer_util.py:305-327
            h = F.relu(torch.cat(((1 - m) * ctx_feat, m * obj_feat), 1))
            h = instance_norm(F.conv2d(h, sd["feat_fuser.conv1.weight"], sd["feat_fuser.conv1.bias"]))
        else:
            h = (1 - m) * ctx_feat + m * obj_feat
    else:
        h = ctx_feat if which_stream == "ctx" else obj_feat
    h = embed("latent_embedder", h, (n_blocks + 1) // 2 if late else n_blocks)   # latent_embedder
    for s in range(n_downsampling):                                            # forward_decoder :215-223
        if use_skip and len(ctx_feats) > 0 and s > 0:
            h = torch.cat((ctx_feats[-s], h), 1)
        k = "decoder.%d." % (3 * s)
        h = F.conv_transpose2d(h, sd[k + "weight"], sd[k + "bias"], stride=2, padding=1, output_padding=1)
        h = F.relu(instance_norm(h))
    out = torch.tanh(F.conv2d(reflect_pad(h, 3), sd["outputEmbedder.1.weight"], sd["outputEmbedder.1.bias"]))
    if use_output_gate:                                                        # :243-245
        mo = mask.repeat(1, out.shape[1], 1, 1)
        out = (1 - mo) * img[:, :3] + mo * out
    return out


def avgpool_3s2(x):
    """nn.AvgPool2d(3, stride=2, padding=[1,1], count_include_pad=False) (Discriminator_NET.py:31-32)."""
    return F.avg_pool2d(x, 3, stride=2, padding=1, count_include_pad=False)


def local_enhancer_forward(sd, x, n_downsample_global, n_blocks_global, n_local_enhancers=1, n_blocks_local=3):
    """LocalEnhancer.forward, models/Pix2Pix_NET.py:8-61."""
    pyr = [x]
    for _ in range(n_local_enhancers):                                         # :49-51
        pyr.append(avgpool_3s2(pyr[-1]))
    out_prev = global_generator_forward(sd, pyr[-1], n_downsample_global, n_blocks_global, stop_before_head=True)
    for n in range(1, n_local_enhancers + 1):                                  # :56-60
        p1, p2 = "model%d_1." % n, "model%d_2." % n
        xi = pyr[n_local_enhancers - n]
        h = F.conv2d(reflect_pad(xi, 3), sd[p1 + "1.weight"], sd[p1 + "1.bias"])
        h = F.relu(instance_norm(h))
        h = F.relu(instance_norm(F.conv2d(h, sd[p1 + "4.weight"], sd[p1 + "4.bias"], stride=2, padding=1)))
        h = h + out_prev
        idx = 0
        for _ in range(n_blocks_local):
            r = F.conv2d(reflect_pad(h, 1), sd[p2 + "%d.conv_block.1.weight" % idx], sd[p2 + "%d.conv_block.1.bias" % idx])
            r = F.relu(instance_norm(r))
            r = F.conv2d(reflect_pad(r, 1), sd[p2 + "%d.conv_block.5.weight" % idx], sd[p2 + "%d.conv_block.5.bias" % idx])
            h = h + instance_norm(r)
            idx += 1
        h = F.conv_transpose2d(h, sd[p2 + "%d.weight" % idx], sd[p2 + "%d.bias" % idx], stride=2, padding=1,
                               output_padding=1)
        h = F.relu(instance_norm(h))
        idx += 3
        if n == n_local_enhancers:
            h = torch.tanh(F.conv2d(reflect_pad(h, 3), sd[p2 + "%d.weight" % (idx + 1)], sd[p2 + "%d.bias" % (idx + 1)]))
        out_prev = h
    return out_prev


# --------------------------------------------------------------------------------------------------
# Discriminator
# --------------------------------------------------------------------------------------------------
def nlayer_discriminator_forward(sd, x, scale, n_layers=3):
    """One PatchGAN with intermediate features (Discriminator_NET.py:61-118, getIntermFeat=True):
    layer0 conv4x4 s2 p2 + LeakyReLU; layers 1..n-1 conv s2 + IN + LReLU; layer n conv s1 + IN + LReLU;
    layer n+1 conv s1 -> 1 channel.  Returns the n_layers+2 taps."""
    taps = []
    h = x
    for j in range(n_layers + 2):
        k = "scale%d_layer%d.0." % (scale, j)
        if k + "weight" not in sd:   # getIntermFeat=False naming: one flattened Sequential per scale (:28-29,99-106)
            k = "layer%d.%d." % (scale, 0 if j == 0 else 2 + 3 * (j - 1))
        w, b = sd[k + "weight"], sd[k + "bias"]
        stride = 2 if j < n_layers else 1
        h = F.conv2d(h, w, b, stride=stride, padding=2)
        if 1 <= j <= n_layers:
            h = instance_norm(h)
        if j <= n_layers:
            h = F.leaky_relu(h, 0.2)
        taps.append(h)
    return taps


def multiscale_discriminator_forward(sd, x, num_D=3, n_layers=3):
    """MultiscaleDiscriminator.forward, Discriminator_NET.py:45-58: scale num_D-1 sees full resolution."""
    result = []
    h = x
    for i in range(num_D):
        result.append(nlayer_discriminator_forward(sd, h, num_D - 1 - i, n_layers))
        if i != num_D - 1:
            h = avgpool_3s2(h)
    return result


# --------------------------------------------------------------------------------------------------
# VGG19 + losses
# --------------------------------------------------------------------------------------------------
def vgg19_forward(sd, x):
    """Vgg19.forward, models/layer_util.py:404-411 -> [relu1_1, relu2_1, relu3_1, relu4_1, relu5_1]."""
    taps = []
    h = x
    for idx, _, _ in VGG19_CONVS:
        if idx in VGG19_POOL_BEFORE:
            h = F.max_pool2d(h, 2, 2)
        k = "slice%d.%d." % (VGG19_SLICE_OF[idx], idx)
        h = F.relu(F.conv2d(h, sd[k + "weight"], sd[k + "bias"], padding=1))
        if idx in VGG19_TAP_AFTER:
            taps.append(h)
    return taps


def vgg19_random_state_dict(seed=1234, dtype=torch.float32):
    """Seeded stand-in for torchvision's pretrained VGG19 (no network): torchvision's own default init
    (kaiming_normal_ fan_out/relu, zero bias) restated so torchvision is not needed at run time."""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for idx, cin, cout in VGG19_CONVS:
        std = (2.0 / (cout * 9)) ** 0.5
        k = "slice%d.%d." % (VGG19_SLICE_OF[idx], idx)
        sd[k + "weight"] = (torch.randn(cout, cin, 3, 3, generator=g) * std).to(dtype)
        sd[k + "bias"] = torch.zeros(cout, dtype=dtype)
    return sd


def vgg_loss(vgg_sd, x, y):
    """VGGLoss.forward, models/losses.py:75-82 (normalize=False)."""
    xv, yv = vgg19_forward(vgg_sd, x), vgg19_forward(vgg_sd, y)
    loss = 0
    for i in range(len(xv)):
        loss = loss + VGG_WEIGHTS[i] * F.l1_loss(xv[i], yv[i].detach())
    return loss


def gan_loss(pred, target_is_real, use_lsgan=True):
    """GANLoss.__call__, models/losses.py:40-50: sum over scales of MSE(last tap, 1|0) (LSGAN) or, with --no_lsgan,
    BCE(sigmoid(last tap), 1|0) (:17-20; the Sigmoid is the discriminator's last module, Discriminator_NET.py:95-96 --
    only applied on its getIntermFeat=False path, i.e. together with --no_ganFeat_loss)."""
    loss = 0
    for p in pred:
        last = p[-1]
        t = torch.full_like(last, 1.0 if target_is_real else 0.0)
        if not use_lsgan:
            loss = loss + F.binary_cross_entropy(torch.sigmoid(last), t)
            continue
        loss = loss + F.mse_loss(last, t)
    return loss


# --------------------------------------------------------------------------------------------------
# Model-level forward (Pix2PixHDModel_condImg) and the training step
# --------------------------------------------------------------------------------------------------
def get_edges(t):
    """models/pix2pixHD_condImg_model.py:285-291."""
    edge = torch.zeros_like(t, dtype=torch.bool)
    edge[:, :, :, 1:] |= t[:, :, :, 1:] != t[:, :, :, :-1]
    edge[:, :, :, :-1] |= t[:, :, :, 1:] != t[:, :, :, :-1]
    edge[:, :, 1:, :] |= t[:, :, 1:, :] != t[:, :, :-1, :]
    edge[:, :, :-1, :] |= t[:, :, 1:, :] != t[:, :, :-1, :]
    return edge


def encode_input(label, inst, image, mask_in, label_nc, no_instance, dtype=torch.float32):
    """models/pix2pixHD_condImg_model.py:144-174."""
    n, _, h, w = label.shape
    onehot = torch.zeros(n, label_nc, h, w, dtype=dtype, device=label.device)
    onehot.scatter_(1, label.long(), 1.0)                                      # :151-152
    input_label = onehot
    if not no_instance:                                                        # :155-158
        input_label = torch.cat((input_label, get_edges(inst).to(dtype)), dim=1)
    real = image.to(dtype)
    m3 = mask_in.to(dtype).repeat(1, 3, 1, 1)
    cond = (1 - m3) * real + m3 * NULLVAL                                      # :165-166
    return input_label, real, cond


class Opt(object):
    """The subset of the reference's option namespace this path reads (defaults: options/mask2image_*_options.py)."""
    def __init__(self, **kw):
        self.label_nc = 35
        self.no_instance = True
        self.output_nc = 3
        self.ngf = 64
        self.n_downsample_global = 4
        self.n_blocks_global = 9
        self.ndf = 64
        self.n_layers_D = 3
        self.num_D = 3
        self.use_output_gate = False
        self.no_ganFeat_loss = False
        self.no_vgg_loss = False
        self.lambda_feat = 10.0
        self.lambda_rec = 0.0
        self.lr = 0.0002
        self.beta1 = 0.5
        self.netG = "global"
        self.use_skip = False
        self.which_encoder = "ctx"
        self.no_imgCond = False
        self.mask_gan_input = False
        self.use_soft_mask = False
        self.n_local_enhancers = 1
        self.n_blocks_local = 3
        for k, v in kw.items():
            setattr(self, k, v)


class ImagePool(object):
    """util/image_pool.py:4-31: history of discriminator inputs.  While the pool is filling, images are stored and returned;
    afterwards each image is, with probability 1/2 (python's `random`, one uniform(0, 1) draw per image and one randint
    per swap), exchanged with a random stored one."""

    def __init__(self, pool_size):
        self.pool_size, self.images = pool_size, []

    def query(self, images):
        import random
        if self.pool_size == 0:
            return images
        out = []
        for i in range(images.shape[0]):
            img = images[i:i + 1].detach()
            if len(self.images) < self.pool_size:
                self.images.append(img)
                out.append(img)
            elif random.uniform(0, 1) > 0.5:
                k = random.randint(0, self.pool_size - 1)
                out.append(self.images[k].clone())
                self.images[k] = img
            else:
                out.append(img)
        return torch.cat(out, 0)


def model_forward(opt, g_sd, d_sd, vgg_sd, label, inst, image, mask_in, dtype=torch.float32, mask_out=None, pool=None):
    """Pix2PixHDModel_condImg.forward, models/pix2pixHD_condImg_model.py:198-259 (netG == 'global' or 'local',
    no_imgCond / mask_gan_input / use_soft_mask off, pool_size 0).
    Returns ([G_GAN, G_GAN_Feat, G_VGG, D_real, D_fake], fake_image, extras)."""
    input_mask, real, cond = encode_input(label, inst, image, mask_in, opt.label_nc, opt.no_instance, dtype)
    input_label = torch.cat((input_mask, cond), 1)                             # :204
    if opt.netG == "local":
        fake = local_enhancer_forward(g_sd, input_label, opt.n_downsample_global, opt.n_blocks_global,
                                      opt.n_local_enhancers, opt.n_blocks_local)
    elif opt.netG == "global_twostream":                                       # :209-210
        fake = global_twostream_forward(g_sd, cond, input_mask, mask_in.to(dtype), opt.n_downsample_global,
                                        opt.n_blocks_global, getattr(opt, "use_skip", False),
                                        getattr(opt, "which_encoder", "ctx"), opt.use_output_gate,
                                        getattr(opt, "feat_fusion", "early_add"))
    else:
        fake = global_generator_forward(g_sd, input_label, opt.n_downsample_global, opt.n_blocks_global,
                                        mask=mask_in.to(dtype), use_output_gate=opt.use_output_gate)   # :208
    # opt-in spectral norm (SNConv2d in place of Conv2d): one power iteration per step, shared by the three D calls
    # below -- the product evaluates D once on [fake ; real]; identity when d_sd has no '.u' entries
    d_sd, new_u = spectral_normalize(d_sd)
    netD_cond = input_mask if getattr(opt, "no_imgCond", False) else input_label     # :213-216
    mask_cond = (mask_out if getattr(opt, "use_soft_mask", False) else mask_in).to(dtype)   # :217

    image_only = opt.netG == "global_twostream" and getattr(opt, "which_encoder", "ctx") == "ctx"   # :71-72,178-179,227-228

    def D(test_image, use_pool=False):                                         # discriminate :176-186 / :226-231
        x = test_image if image_only else torch.cat((netD_cond, test_image), 1)
        if getattr(opt, "mask_gan_input", False):
            x = x * mask_cond.repeat(1, x.shape[1], 1, 1)
        if use_pool and pool is not None:                                      # :182-184
            x = pool.query(x)
        return multiscale_discriminator_forward(d_sd, x, opt.num_D, opt.n_layers_D)
    pred_fake_pool = D(fake.detach(), use_pool=True)                           # :218
    lsgan = not getattr(opt, "no_lsgan", False)
    loss_D_fake = gan_loss(pred_fake_pool, False, lsgan)                       # :219
    pred_real = D(real)                                                        # :222
    loss_D_real = gan_loss(pred_real, True, lsgan)                             # :223
    pred_fake = D(fake)                                                        # :226-231
    loss_G_GAN = gan_loss(pred_fake, True, lsgan)                              # :232
    loss_G_GAN_Feat = torch.zeros((), dtype=dtype)
    if not opt.no_ganFeat_loss:                                                # :235-242
        feat_weights = 4.0 / (opt.n_layers_D + 1)
        D_weights = 1.0 / opt.num_D
        for i in range(opt.num_D):
            for j in range(len(pred_fake[i]) - 1):
                loss_G_GAN_Feat = loss_G_GAN_Feat + D_weights * feat_weights * \
                    F.l1_loss(pred_fake[i][j], pred_real[i][j].detach()) * opt.lambda_feat
    loss_G_VGG = torch.zeros((), dtype=dtype)
    if not opt.no_vgg_loss:                                                    # :245-247
        loss_G_VGG = vgg_loss(vgg_sd, fake, real) * opt.lambda_feat
    if opt.lambda_rec > 0:                                                     # :249-251
        loss_G_GAN_Feat = loss_G_GAN_Feat + F.l1_loss(fake, real.detach()) * opt.lambda_rec
    extras = dict(input_label=input_label, cond=cond, pred_fake=pred_fake, pred_real=pred_real, sn_u=new_u)
    return [loss_G_GAN, loss_G_GAN_Feat, loss_G_VGG, loss_D_real, loss_D_fake], fake, extras


def step_losses(losses):
    """train_mask2image.py:68-73."""
    G_GAN, G_GAN_Feat, G_VGG, D_real, D_fake = losses
    loss_D = (D_fake + D_real) * 0.5
    loss_G = G_GAN + G_GAN_Feat + G_VGG
    return loss_G, loss_D


def adam_step(param, grad, m, v, step, lr, beta1=0.5, beta2=0.999, eps=1e-8):
    """torch.optim.Adam (no weight decay, no amsgrad) as used at pix2pixHD_condImg_model.py:135,139.
    In-place on (param, m, v); `step` is the 1-based step count."""
    m.mul_(beta1).add_(grad, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = (v.sqrt() / (bc2 ** 0.5)).add_(eps)
    param.addcdiv_(m, denom, value=-lr / bc1)


def train_step(opt, g_sd, d_sd, vgg_sd, batch, state=None, dtype=torch.float32):
    """One iteration of train_mask2image.py:58-86 on CPU: forward, loss_G.backward + Adam(G), loss_D.backward +
    Adam(D).  g_sd / d_sd are updated in place; returns (losses, fake, grads_G, grads_D)."""
    g_par = OrderedDict((k, v.detach().clone().to(dtype).requires_grad_(True)) for k, v in g_sd.items())
    d_par = OrderedDict((k, v.detach().clone().to(dtype).requires_grad_(not k.endswith(".u"))) for k, v in d_sd.items())
    v_sd = OrderedDict((k, v.to(dtype)) for k, v in vgg_sd.items()) if vgg_sd is not None else None
    losses, fake, extras = model_forward(opt, g_par, d_par, v_sd, batch["label"], batch["inst"], batch["image"],
                                         batch["mask_in"], dtype, mask_out=batch.get("mask_out"))
    loss_G, loss_D = step_losses(losses)
    d_train = OrderedDict((k, p) for k, p in d_par.items() if p.requires_grad)
    gG = torch.autograd.grad(loss_G, list(g_par.values()), retain_graph=True, allow_unused=True)
    gD = torch.autograd.grad(loss_D, list(d_train.values()), allow_unused=True)
    gG = OrderedDict((k, (g if g is not None else torch.zeros_like(p))) for (k, p), g in zip(g_par.items(), gG))
    gD = OrderedDict((k, (g if g is not None else torch.zeros_like(p))) for (k, p), g in zip(d_train.items(), gD))
    for k, u in extras["sn_u"].items():       # SNConv2d stores the iterated u (sn_utils.py:65-66)
        d_sd[k].copy_(u.reshape(d_sd[k].shape))
    if state is None:
        state = dict(step=0, mG={}, vG={}, mD={}, vD={})
    state["step"] += 1
    for sd, grads, mk, vk in ((g_sd, gG, "mG", "vG"), (d_sd, gD, "mD", "vD")):
        for k in sd:
            if k not in grads:
                continue
            if k not in state[mk]:
                state[mk][k] = torch.zeros_like(sd[k], dtype=dtype)
                state[vk][k] = torch.zeros_like(sd[k], dtype=dtype)
            p = sd[k].to(dtype)
            adam_step(p, grads[k], state[mk][k], state[vk][k], state["step"], opt.lr, opt.beta1)
            sd[k].copy_(p)
    return [float(l.detach()) for l in losses], fake.detach(), gG, gD, state


# --------------------------------------------------------------------------------------------------
# Spectral norm (opt-in extension; reference: models/sn_utils.py:11-25)
# --------------------------------------------------------------------------------------------------
def max_singular_value(W, u, Ip=1):
    Wm = W.reshape(W.shape[0], -1)
    _u = u
    for _ in range(Ip):
        _v = _u @ Wm
        _v = _v / (_v.norm() + 1e-12)
        _u = (Wm @ _v.t()).reshape(1, -1)
        _u = _u / (_u.norm() + 1e-12)
    sigma = (_u @ Wm @ _v.t()).reshape(())
    return sigma, _u


def spectral_normalize(d_sd):
    """SNConv2d.W_bar (models/sn_utils.py:57-63) for every conv of a state dict that carries a '<conv>.u' entry:
    returns (state dict with weight / sigma, differentiable w.r.t. the weight exactly as in the reference, where
    autograd flows through the power iteration) and {'<conv>.u': iterated u}."""
    eff, new_u = OrderedDict(d_sd), OrderedDict()
    for k in d_sd:
        if k.endswith(".u"):
            base = k[:-2]
            w = d_sd[base + ".weight"]
            sigma, u2 = max_singular_value(w, d_sd[k].detach().to(w.dtype).reshape(1, -1))
            eff[base + ".weight"] = w / sigma
            new_u[k] = u2.detach()
    return eff, new_u


# --------------------------------------------------------------------------------------------------
# Synthetic batches -- shared by tests, smoke() and bench.py
# --------------------------------------------------------------------------------------------------
def synthetic_batch(B, H, W, label_nc=35, seed=1234):
    g = torch.Generator().manual_seed(seed)
    bh, bw = max(H // 16, 1), max(W // 16, 1)
    lab = torch.randint(0, label_nc, (B, 1, bh, bw), generator=g).float()
    label = F.interpolate(lab, size=(H, W), mode="nearest")
    ins = torch.randint(0, 20, (B, 1, bh, bw), generator=g).float()
    inst = F.interpolate(ins, size=(H, W), mode="nearest")
    image = torch.rand(B, 3, H, W, generator=g) * 2 - 1
    mask_in = torch.zeros(B, 1, H, W)
    mask_out = torch.zeros(B, 1, H, W)
    for b in range(B):
        side_h = int(torch.randint(max(H // 8, 1), max(H // 2, 2), (1,), generator=g))
        side_w = int(torch.randint(max(H // 8, 1), max(H // 2, 2), (1,), generator=g))
        y0 = int(torch.randint(0, H - side_h + 1, (1,), generator=g))
        x0 = int(torch.randint(0, W - side_w + 1, (1,), generator=g))
        mask_in[b, :, y0:y0 + side_h, x0:x0 + side_w] = 1
        dh, dw = int(side_h * 0.15), int(side_w * 0.15)
        mask_out[b, :, max(0, y0 - dh):min(H, y0 + side_h + dh), max(0, x0 - dw):min(W, x0 + side_w + dw)] = 1
    return dict(label=label, inst=inst, image=image, mask_in=mask_in, mask_out=mask_out)

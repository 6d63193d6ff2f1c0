"""CPU restatement (torch functional ops) of the reference's box2mask generator -- BASELINE config #5, SURVEY N3.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Follows models/MaskTwoStreamConv_NET.py (the two-stream conv
auto-encoder), its base class models/MaskContextAE_NET.py, the blocks of models/layer_util.py:136-242
(ConvResnetBlock / DeconvResnetBlock) and :333-378 (ResnetBlock), and the reconstruction losses of
models/TwoStreamAE_mask.py:188-203 + models/mask_losses.py:12-27.  Pinned by tests/golden/box2mask_small.npz, which
oracle/make_golden_box2mask.py produced by running the reference's OWN class (forward, losses, parameter gradients).

Parameters are addressed by the reference's own names: '<params_dict key>.<state_dict key>', e.g.
'conv_encoder_3.deep.1.weight', 'ctx_conv_decoder_1.shortcut.0.weight', 'latent_encoder.0.conv_block.1.weight'.

Two aliasing effects of the reference code are part of its arithmetic and are restated explicitly:
  * every Conv/DeconvResnetBlock opens with an IN-PLACE ReLU on its input while `residual = x` aliases that tensor
    (layer_util.py:156-162, 236-242): the shortcut branch sees relu(x);
  * the encoder features kept for the skip connections (MaskTwoStreamConv_NET.py:172-173) are rectified in place by
    the next block before the decoder concatenates them (:161-165).
"""
import torch
import torch.nn.functional as F

IGNORE_INDEX = 255          # models/mask_losses.py:10
DIM_LIST_TAIL = [96, 128, 256, 512]   # MaskTwoStreamConv_NET.py:25 ("this part is hard-coded")


# BatchNorm mode of the current forward pass: None = training mode (batch statistics); a dict = training mode that also
# records every module's batch statistics {key: (mean, unbiased var)} (what the running buffers are updated with);
# "eval" = normalise with sd[key + '.running_mean' / '.running_var'] (nn.BatchNorm2d.eval()).
_BN_MODE = [None]


_NORM = ["batch"]            # norm_layer of the current forward pass: 'batch' | 'instance' (layer_util.py:19-26)


def batch_norm(sd, key, x, eps=1e-5):
    """The network's norm layer.  norm_layer == 'instance': nn.InstanceNorm2d(affine=False), no parameters, no buffers.
    norm_layer == 'batch': nn.BatchNorm2d(affine=True) (layer_util.py:19-21).  Training mode: biased batch
    statistics over (N, H, W); the running buffers (momentum 0.1, unbiased variance) do not enter the result.  Eval mode:
    the running buffers."""
    if _NORM[0] == "instance":
        mean = x.mean(dim=(2, 3), keepdim=True)
        var = x.var(dim=(2, 3), unbiased=False, keepdim=True)
        return (x - mean) / torch.sqrt(var + eps)
    g, b = sd[key + ".weight"].view(1, -1, 1, 1), sd[key + ".bias"].view(1, -1, 1, 1)
    if _BN_MODE[0] == "eval":
        mean, var = sd[key + ".running_mean"].view(1, -1, 1, 1), sd[key + ".running_var"].view(1, -1, 1, 1)
        return (x - mean) / torch.sqrt(var + eps) * g + b
    mean = x.mean(dim=(0, 2, 3), keepdim=True)
    var = x.var(dim=(0, 2, 3), unbiased=False, keepdim=True)
    if isinstance(_BN_MODE[0], dict):
        _BN_MODE[0][key] = (mean.detach().reshape(-1), x.detach().var(dim=(0, 2, 3), unbiased=True))
    return (x - mean) / torch.sqrt(var + eps) * g + b


def upsample2(x):
    """nn.Upsample(scale_factor=2, mode='bilinear') as torch >= 0.4 evaluates it (align_corners=False)."""
    return F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)


def conv_resnet_block(sd, p, x, stride=2, k=4):
    """ConvResnetBlock(in, out, num_layers=1, stride=2, kernel_size=4), layer_util.py:119-162:
    deep = [ReLU(in place), Conv k x k stride 2 pad (k-1)//2, norm]; shortcut = [Conv 1x1 stride 2, norm]."""
    xr = F.relu(x)                                   # in place in the reference: BOTH branches see relu(x)
    pad = (k - 1) // 2
    deep = batch_norm(sd, p + ".deep.2", F.conv2d(xr, sd[p + ".deep.1.weight"], sd[p + ".deep.1.bias"], stride=stride,
                                                  padding=pad))
    short = batch_norm(sd, p + ".shortcut.1", F.conv2d(xr, sd[p + ".shortcut.0.weight"], sd[p + ".shortcut.0.bias"],
                                                       stride=stride))
    return deep + short, xr


def deconv_resnet_block(sd, p, x, k=4):
    """DeconvResnetBlock(in, out, num_layers=1, stride=2, kernel_size=4), layer_util.py:164-242 (even kernel ->
    build_tconv2d_block): deep = [ReLU(in place), ConvTranspose2d k4 s2 p1 output_padding 0, norm];
    shortcut = [Conv 1x1, norm] (when in != out) + bilinear x2."""
    xr = F.relu(x)
    deep = F.conv_transpose2d(xr, sd[p + ".deep.1.weight"], sd[p + ".deep.1.bias"], stride=2, padding=(k - 1) // 2,
                              output_padding=0)
    deep = batch_norm(sd, p + ".deep.2", deep)
    short = xr
    if p + ".shortcut.0.weight" in sd:
        short = batch_norm(sd, p + ".shortcut.1", F.conv2d(short, sd[p + ".shortcut.0.weight"], sd[p + ".shortcut.0.bias"]))
    return deep + upsample2(short)


def resnet_block(sd, p, x):
    """ResnetBlock(dim, 'reflect', norm, ReLU, no dropout), layer_util.py:333-378: x + conv_block(x)."""
    h = F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), sd[p + ".conv_block.1.weight"], sd[p + ".conv_block.1.bias"])
    h = F.relu(batch_norm(sd, p + ".conv_block.2", h))
    h = F.conv2d(F.pad(h, (1, 1, 1, 1), mode="reflect"), sd[p + ".conv_block.5.weight"], sd[p + ".conv_block.5.bias"])
    return x + batch_norm(sd, p + ".conv_block.6", h)


def dilated_resnet_block(sd, p, x, d):
    """DilatedResnetBlock(dim, dim, dilation=(d, d), activation_fn=norm) (layer_util.py:260-293): conv3x3 (no bias, padding
    = dilation = d), norm, ReLU, conv3x3, norm, + x, ReLU."""
    out = F.conv2d(x, sd[p + ".conv1.weight"], None, padding=d, dilation=d)
    out = F.relu(batch_norm(sd, p + ".bn1", out))
    out = F.conv2d(out, sd[p + ".conv2.weight"], None, padding=d, dilation=d)
    return F.relu(batch_norm(sd, p + ".bn2", out) + x)


def two_stream_forward(sd, cond, num_layers=3, n_blocks=6, conv_size=4, no_comb=False, bn_mode=None, norm_layer="batch",
                       add_dilated_layers=False):
    """MaskTwoStreamConv_NET.forward (:159-219), which_stream == 'obj_context'; no_comb: MaskTwoStreamConvSwitch_NET
    (--no_comb), the same network whose forward returns the context stream as it is (:208).
    cond: [B, input_nc, S, S] (cond_in 'ctx_obj': object box mask in its class channel | one-hot context).
    bn_mode: see _BN_MODE.  Returns (comb_logit, comb_logprob, obj_logit, obj_prob)."""
    _BN_MODE[0], _NORM[0] = bn_mode, norm_layer
    try:
        return _two_stream_forward(sd, cond, num_layers, n_blocks, conv_size, no_comb, add_dilated_layers)
    finally:
        _BN_MODE[0], _NORM[0] = None, "batch"


def _two_stream_forward(sd, cond, num_layers, n_blocks, conv_size, no_comb, add_dilated_layers=False):
    # shared encoder (:63-90): Conv 7x7 stride 2 pad 3, norm, ReLU, then num_layers ConvResnetBlocks
    h = F.conv2d(cond, sd["conv_encoder_0.weight"], sd["conv_encoder_0.bias"], stride=2, padding=3)
    h = F.relu(batch_norm(sd, "conv_encoder_1", h))
    enc_features = [h]                                # i == 2 (:172-173)
    for i in range(num_layers):
        h, prev_rectified = conv_resnet_block(sd, "conv_encoder_%d" % (3 + i), h, k=conv_size)
        enc_features[-1] = prev_rectified             # the kept feature was rectified in place by this block
        if i < num_layers - 1:
            enc_features.append(h)
    j0 = 0
    if add_dilated_layers:                            # --add_dilated_layers (MaskTwoStreamConvSwitch_NET.py:101-104)
        h = dilated_resnet_block(sd, "latent_encoder.0", h, 2)
        h = dilated_resnet_block(sd, "latent_encoder.1", h, 4)
        j0 = 2
    for j in range(n_blocks // 2):                    # latent encoder (:92-106)
        h = resnet_block(sd, "latent_encoder.%d" % (j0 + j), h)
    latent = h

    def decode(stream, skips):
        d = latent
        for j in range((n_blocks + 1) // 2):          # latent decoder (:108-123)
            d = resnet_block(sd, "%s_latent_decoder.%d" % (stream, j), d)
        for i in range(num_layers + 1):               # conv decoder (:125-155) + forward_decoder (:157-165)
            if skips is not None and 1 <= i <= num_layers:
                d = torch.cat((skips[-1 - (i - 1)], d), 1)
            d = deconv_resnet_block(sd, "%s_conv_decoder_%d" % (stream, i), d, k=conv_size)
        k = "%s_conv_decoder_%d" % (stream, num_layers + 1)
        return F.conv2d(d, sd[k + ".weight"], sd[k + ".bias"], padding=1)
    ctx_logit = decode("ctx", enc_features)
    obj_logit = decode("obj", None)
    obj_prob = torch.sigmoid(obj_logit)
    if no_comb:
        return ctx_logit, F.log_softmax(ctx_logit, dim=1), obj_logit, obj_prob
    # combination (:198-217): the object stream's sigmoid gates between the context logits and its own logit
    padded_mask = obj_prob.expand_as(ctx_logit)
    comb_logit = (1 - padded_mask) * ctx_logit + padded_mask * obj_logit
    return comb_logit, F.log_softmax(comb_logit, dim=1), obj_logit, obj_prob


def mask_recon_loss(comb_logprob, label_map, mask_out):
    """MaskReconLoss (mask_losses.py:12-27): NLL over the pixels INSIDE the box (mask_out >= 0.5); the others get the
    ignore index 255."""
    gt = label_map.view(-1, label_map.size(2), label_map.size(3)).long().clone()
    gt[mask_out[:, 0] < 0.5] = IGNORE_INDEX
    return F.nll_loss(comb_logprob, gt, ignore_index=IGNORE_INDEX)


def obj_recon_loss(obj_prob, mask_out, mask_obj_inst, use_output_gate=True):
    """TwoStreamAE_mask.forward :199-203 with objReconLoss == 'bce': BCE between the (gated) object mask and the
    instance mask."""
    p = obj_prob * mask_out if use_output_gate else obj_prob
    return F.binary_cross_entropy(p, mask_obj_inst)


def encode_input(label_nc, mask_ctx_in, mask_in, cls):
    """TwoStreamAE_mask.encode_input (:127-151) + construct_input_cond (:331-338) for cond_in == 'ctx_obj'."""
    B, _, S, S2 = mask_ctx_in.shape
    ctx = torch.zeros(B, label_nc, S, S2).scatter_(1, mask_ctx_in.long(), 1.0)
    obj = torch.zeros(B, label_nc, S, S2)
    for b in range(B):
        obj[b, int(cls[b, 0])] = mask_in[b, 0]
    cls_onehot = torch.zeros(B, label_nc).scatter_(1, cls.long(), 1.0)
    return torch.cat((obj, ctx), 1), cls_onehot


# --------------------------------------------------------------------------------------------------
# --use_gan with which_gan == 'patch_multiscale' (TwoStreamAE_mask.py:83-92, 205-248): a 2-scale MultiscaleDiscriminator
# with BatchNorm (norm_layer == 'batch'), LSGAN, intermediate features kept
# --------------------------------------------------------------------------------------------------
def patch_discriminator_bn_forward(sd, x, scale, n_layers=3):
    """NLayerDiscriminator(getIntermFeat=True) with nn.BatchNorm2d (Discriminator_NET.py:61-118): conv4x4 s2 p2 + LReLU;
    n_layers-1 x [conv s2, BN, LReLU]; [conv s1, BN, LReLU]; conv s1 -> 1 channel.  Returns the n_layers + 2 taps."""
    taps, h = [], x
    for j in range(n_layers + 2):
        k = "scale%d_layer%d." % (scale, j)
        h = F.conv2d(h, sd[k + "0.weight"], sd[k + "0.bias"], stride=2 if j < n_layers else 1, padding=2)
        if 1 <= j <= n_layers:
            h = batch_norm(sd, k + "1", h)
        if j <= n_layers:
            h = F.leaky_relu(h, 0.2)
        taps.append(h)
    return taps


def multiscale_discriminator_bn_forward(sd, x, num_D=2, n_layers=3):
    """MultiscaleDiscriminator.forward (Discriminator_NET.py:45-58): scale num_D-1 sees the full resolution."""
    out, h = [], x
    for i in range(num_D):
        out.append(patch_discriminator_bn_forward(sd, h, num_D - 1 - i, n_layers))
        if i != num_D - 1:
            h = F.avg_pool2d(h, 3, stride=2, padding=1, count_include_pad=False)
    return out


def lsgan(pred, target_is_real):
    """GANLoss(use_lsgan=True) on a list of per-scale tap lists (models/losses.py:40-50)."""
    loss = 0
    for p in pred:
        loss = loss + F.mse_loss(p[-1], torch.full_like(p[-1], 1.0 if target_is_real else 0.0))
    return loss


def gan_iteration_losses(sdG, sdD, cond, label_map, mask_out, mask_obj_inst, num_layers=3, n_blocks=6, n_layers_D=3,
                         use_output_gate=True, rec_weight=1.0, gan_weight=1.0, lambda_feat=1.0, use_ganFeat_loss=True):
    """TwoStreamAE_mask.forward with use_gan (:188-235).  Returns (loss_G, loss_D, dict of the reported scalars).
    loss_G = loss_recon_obj + rec_weight * loss_recon_comb + gan_weight * loss_G_GAN   (the feature-matching term is only
    reported: it is computed on fake.detach() and never enters loss_G, :221-229,233-235)."""
    _, comb_lp, _, obj_prob = two_stream_forward(sdG, cond, num_layers=num_layers, n_blocks=n_blocks)
    l_comb = mask_recon_loss(comb_lp, label_map, mask_out)
    obj_gated = obj_prob * mask_out if use_output_gate else obj_prob           # :200-201
    l_obj = F.binary_cross_entropy(obj_gated, mask_obj_inst)
    real, fake, c = mask_obj_inst, obj_gated, cond                               # :208-209 (fake is already gated)
    if use_output_gate:                                                          # :210-213 ("masking twice")
        real, fake, c = real * mask_out, fake * mask_out, c * mask_out
    D = lambda t: multiscale_discriminator_bn_forward(sdD, torch.cat((t, c), 1), 2, n_layers_D)   # noqa: E731  (:153-157)
    real_d, fake_d = D(real), D(fake.detach())
    l_d_real, l_d_fake = lsgan(real_d, True), lsgan(fake_d, False)
    l_feat = torch.zeros(())
    if use_ganFeat_loss:
        for i in range(2):
            for j in range(len(fake_d[i]) - 1):
                l_feat = l_feat + 0.5 * (4.0 / (n_layers_D + 1)) * F.l1_loss(fake_d[i][j], real_d[i][j].detach()) * lambda_feat
    l_g_gan = lsgan(D(fake), True)                                              # :231-232
    loss_G = l_obj + rec_weight * l_comb + gan_weight * l_g_gan
    loss_D = 0.5 * l_d_real + 0.5 * l_d_fake
    return loss_G, loss_D, dict(comb=l_comb, obj=l_obj, g_gan=l_g_gan, d=loss_D, feat=l_feat, d_real=l_d_real, d_fake=l_d_fake)

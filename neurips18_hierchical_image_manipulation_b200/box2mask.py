"""box2mask generator executor -- BASELINE config #5 / SURVEY N3 (first slice: the generator's training-mode FORWARD
and the two reconstruction losses; the backward pass, the GAN terms and the optimizer step of
models/TwoStreamAE_mask.py:205-248 are not built yet and raise).

  MaskTwoStreamConvNet   <- models/MaskTwoStreamConv_NET.py:13-219 (+ MaskContextAE_NET base, layer_util.py:119-242,333-378)
  TwoStreamAE_mask       <- models/TwoStreamAE_mask.py (encode_input :127-151, reconstruct :257-300, losses :188-203)

Parameter names are the reference's own ('<params_dict key>.<state_dict key>': conv_encoder_3.deep.1.weight,
ctx_conv_decoder_1.shortcut.0.weight, latent_encoder.0.conv_block.1.weight, ...), OIHW / IOHW fp32 like the reference, so
its `save_network_dict` checkpoints map one to one.  Every convolution (7x7 s2, 4x4 s2, 1x1 s2, 3x3 reflect,
ConvTranspose 4x4 s2, 1x1, 3x3) runs on the tcgen05 engines; BatchNorm (training mode, batch statistics) is
hm_in_stats over the batch-folded tensor + hm_bn_fold + hm_in_apply; the rest is csrc/hm_box2mask.cu.

Two aliasing effects of the reference are part of its arithmetic (see oracle/box2mask.py): each Conv/DeconvResnetBlock
rectifies its input IN PLACE, so both of its branches -- and the encoder features kept for the skip connections -- see
relu(x).
"""
import torch

from . import ops
from .networks import ConvP, FlatParams, _f32
from .ops import ACT_NONE, ACT_RELU, Operand

DIM_LIST_TAIL = [96, 128, 256, 512]      # MaskTwoStreamConv_NET.py:25


class _BN(object):
    """nn.BatchNorm2d(C, affine=True) parameters (training mode: the running buffers do not enter the result)."""

    def __init__(self, fp, name, c):
        self.fp, self.name, self.c = fp, name, c
        fp.declare(name + ".weight", (c,))
        fp.declare(name + ".bias", (c,))

    @property
    def gamma(self):
        return self.fp.params[self.name + ".weight"]

    @property
    def beta(self):
        return self.fp.params[self.name + ".bias"]

    def init_reference(self, gen):
        """weights_init (layer_util.py:13-15): gain ~ N(1, 0.02), shift 0."""
        with torch.no_grad():
            self.gamma.copy_((1.0 + torch.randn(self.c, generator=gen) * 0.02).to(self.gamma.device))
            self.beta.zero_()
        self.fp.version += 1

    def apply(self, ctx, y, act, skip=None, out32=None, out_op=None, reflect=True):
        N, H, W, C = y.shape
        mean, rstd = ops.in_stats(ctx, y.view(1, N * H, W, C))          # statistics over (N, H, W)
        mean_n, rstd_n = ops.bn_fold(ctx, mean, rstd, self.gamma, self.beta, N)
        ops.in_apply(ctx, y, mean_n, rstd_n, act, skip=skip, out32=out32, out_op=out_op, reflect=reflect)


class MaskTwoStreamConvNet(object):
    def __init__(self, ctx, fp, label_nc, output_nc, conv_dim=64, num_layers=3, conv_size=4, n_blocks=6,
                 cond_in="ctx_obj", which_stream="obj_context", num_resnetblocks=1, norm_layer="batch", use_simpleRes=False):
        if which_stream != "obj_context" or num_resnetblocks != 1 or norm_layer != "batch" or use_simpleRes:
            raise NotImplementedError("box2mask: only the shipped configuration (which_stream obj_context, one conv per "
                                      "block, batch norm, ConvResnetBlock) is part of this slice")
        if conv_size % 2 != 0:
            raise NotImplementedError("box2mask: odd conv_size selects the upsample+conv decoder (layer_util.py:196-209)")
        self.ctx, self.fp = ctx, fp
        self.label_nc, self.output_nc, self.num_layers, self.n_blocks, self.k = label_nc, output_nc, num_layers, n_blocks, conv_size
        self.input_nc = 2 * label_nc if cond_in == "ctx_obj" else label_nc
        dims = [conv_dim] + DIM_LIST_TAIL
        pad = (conv_size - 1) // 2
        self.convs_, self.bns_ = [], []

        def conv(*a, **kw):
            c = ConvP(ctx, fp, *a, **kw)
            self.convs_.append(c)
            return c

        def bn(name, c):
            b = _BN(fp, name, c)
            self.bns_.append(b)
            return b
        # shared encoder (:63-90)
        self.enc0 = (conv("conv_encoder_0", self.input_nc, dims[0], 7, 2, 3), bn("conv_encoder_1", dims[0]))
        self.enc_blocks = []
        for i in range(num_layers):
            p = "conv_encoder_%d" % (3 + i)
            self.enc_blocks.append(dict(deep=conv(p + ".deep.1", dims[i], dims[i + 1], conv_size, 2, pad),
                                        deep_bn=bn(p + ".deep.2", dims[i + 1]),
                                        short=conv(p + ".shortcut.0", dims[i], dims[i + 1], 1, 2, 0),
                                        short_bn=bn(p + ".shortcut.1", dims[i + 1])))
        self.latent_dim = dims[num_layers]

        def res_blocks(prefix, n):
            out = []
            for j in range(n):
                p = "%s.%d.conv_block" % (prefix, j)
                out.append(dict(c1=conv(p + ".1", self.latent_dim, self.latent_dim, 3, 1, 0), b1=bn(p + ".2", self.latent_dim),
                                c2=conv(p + ".5", self.latent_dim, self.latent_dim, 3, 1, 0), b2=bn(p + ".6", self.latent_dim)))
            return out
        self.latent_encoder = res_blocks("latent_encoder", n_blocks // 2)               # :92-106

        def decoder(stream, out_nc, skips):                                               # :108-155
            blocks, out_dim = [], self.latent_dim
            for i in range(num_layers + 1):
                in_dim = out_dim
                out_dim = dims[num_layers - i - 1] if i < num_layers else in_dim // 2
                if skips and 1 <= i <= num_layers:
                    in_dim *= 2
                p = "%s_conv_decoder_%d" % (stream, i)
                blk = dict(deep=conv(p + ".deep.1", in_dim, out_dim, conv_size, 2, pad, transposed=True),
                           deep_bn=bn(p + ".deep.2", out_dim), short=None, short_bn=None, out_dim=out_dim)
                if in_dim != out_dim:
                    blk["short"] = conv(p + ".shortcut.0", in_dim, out_dim, 1, 1, 0)
                    blk["short_bn"] = bn(p + ".shortcut.1", out_dim)
                blocks.append(blk)
            final = conv("%s_conv_decoder_%d" % (stream, num_layers + 1), out_dim, out_nc, 3, 1, 1)
            return blocks, final
        # registration order of the reference's params_dict (:43-60): obj decoder, then ctx decoder
        self.obj_dec, self.obj_final = decoder("obj", 1, False)
        self.obj_latent = res_blocks("obj_latent_decoder", (n_blocks + 1) // 2)
        self.ctx_dec, self.ctx_final = decoder("ctx", output_nc, True)
        self.ctx_latent = res_blocks("ctx_latent_decoder", (n_blocks + 1) // 2)

    def init_reference(self, gen):
        for c in self.convs_:
            c.init_reference(gen)
        for b in self.bns_:
            b.init_reference(gen)

    # ------------------------------------------------------------------------------------------------
    def _res_block(self, blk, x32, x_op, want_op):
        """ResnetBlock (layer_util.py:333-378): x + [pad, conv, norm, relu, pad, conv, norm](x)."""
        ctx = self.ctx
        N, H, W, C = x32.shape
        y = _f32(ctx, N, H, W, C)
        blk["c1"].forward(x_op, 0, out32=y)
        mid = Operand(ctx, N, H, W, C, border=1)
        blk["b1"].apply(ctx, y, ACT_RELU, out_op=mid, reflect=True)
        y2 = _f32(ctx, N, H, W, C)
        blk["c2"].forward(mid, 0, out32=y2)
        out32 = _f32(ctx, N, H, W, C)
        out_op = Operand(ctx, N, H, W, C, border=1) if want_op else None
        blk["b2"].apply(ctx, y2, ACT_NONE, skip=x32, out32=out32, out_op=out_op, reflect=True)
        return out32, out_op

    def _decode(self, latent32, latent_op, res, blocks, final, skips):
        ctx = self.ctx
        d32, d_op = latent32, latent_op
        for j, blk in enumerate(res):
            d32, d_op = self._res_block(blk, d32, d_op, want_op=(j + 1 < len(res)))
        N = d32.shape[0]
        x_op = None
        for i, blk in enumerate(blocks):
            h, w, c = d32.shape[1], d32.shape[2], d32.shape[3]
            xr = Operand(ctx, N, h, w, c)                                  # relu(x): the in-place ReLU of the block
            ops.in_apply(ctx, d32, None, None, ACT_RELU, out_op=xr, reflect=False)
            x_op = ops.concat_operands(ctx, skips[-1 - (i - 1)], xr) if (skips and 1 <= i <= self.num_layers) else xr
            od = blk["out_dim"]
            y = _f32(ctx, N, 2 * h, 2 * w, od)
            blk["deep"].forward(x_op, 0, out32=y)                          # ConvTranspose2d k4 s2 p1
            deep32 = _f32(ctx, N, 2 * h, 2 * w, od)
            blk["deep_bn"].apply(ctx, y, ACT_NONE, out32=deep32)
            if blk["short"] is not None:
                ys = _f32(ctx, N, h, w, od)
                blk["short"].forward(x_op, 0, out32=ys)
                s32 = _f32(ctx, N, h, w, od)
                blk["short_bn"].apply(ctx, ys, ACT_NONE, out32=s32)
            else:
                s32 = _f32(ctx, N, h, w, c)
                ops.in_apply(ctx, d32, None, None, ACT_RELU, out32=s32)
            d32 = _f32(ctx, N, 2 * h, 2 * w, od)
            ops.upsample2_add(ctx, s32, deep32, d32)
        fin = Operand(ctx, N, d32.shape[1], d32.shape[2], d32.shape[3])
        ops.in_apply(ctx, d32, None, None, ACT_NONE, out_op=fin, reflect=False)      # the last conv sees dec_feat as is
        logit = _f32(ctx, N, d32.shape[1], d32.shape[2], final.cout)
        final.forward(fin, 1, out32=logit)
        return logit

    def forward(self, cond_op):
        """cond_op: Operand [B,S,S,input_nc].  Returns the dense fp32 NHWC logits (ctx_logit [B,S,S,output_nc],
        obj_logit [B,S,S,1]) -- MaskTwoStreamConv_NET.forward :166-196; the head kernel combines them."""
        ctx = self.ctx
        conv0, bn0 = self.enc0
        N = cond_op.n
        ho, wo = conv0.out_hw(cond_op.h, cond_op.w, 3)
        y = _f32(ctx, N, ho, wo, conv0.cout)
        conv0.forward(cond_op, 3, out32=y)
        cur = Operand(ctx, N, ho, wo, conv0.cout)                          # relu(bn(.)) >= 0: in-place ReLU is a no-op
        bn0.apply(ctx, y, ACT_RELU, out_op=cur, reflect=False)
        skips = [cur]
        h32 = None
        for i, blk in enumerate(self.enc_blocks):                          # ConvResnetBlock (layer_util.py:119-162)
            ho, wo = blk["deep"].out_hw(cur.h, cur.w, blk["deep"].pad)
            c = blk["deep"].cout
            yd, ys = _f32(ctx, N, ho, wo, c), _f32(ctx, N, ho, wo, c)
            blk["deep"].forward(cur, blk["deep"].pad, out32=yd)
            blk["short"].forward(cur, 0, out32=ys)
            tmp, h32 = _f32(ctx, N, ho, wo, c), _f32(ctx, N, ho, wo, c)
            blk["deep_bn"].apply(ctx, yd, ACT_NONE, out32=tmp)
            blk["short_bn"].apply(ctx, ys, ACT_NONE, skip=tmp, out32=h32)
            if i + 1 < len(self.enc_blocks):
                cur = Operand(ctx, N, ho, wo, c)                           # rectified in place by the next block
                ops.in_apply(ctx, h32, None, None, ACT_RELU, out_op=cur, reflect=False)
                skips.append(cur)
        lat_op = Operand(ctx, N, h32.shape[1], h32.shape[2], h32.shape[3], border=1)
        ops.in_apply(ctx, h32, None, None, ACT_NONE, out_op=lat_op, reflect=True)
        lat32 = h32
        for j, blk in enumerate(self.latent_encoder):
            lat32, lat_op = self._res_block(blk, lat32, lat_op, want_op=True)
        ctx_logit = self._decode(lat32, lat_op, self.ctx_latent, self.ctx_dec, self.ctx_final, skips)
        obj_logit = self._decode(lat32, lat_op, self.obj_latent, self.obj_dec, self.obj_final, None)
        return ctx_logit, obj_logit


class TwoStreamAE_mask(object):
    """models/TwoStreamAE_mask.py: forward slice.  `forward(...)` returns
    ([loss_recon_comb, loss_recon_obj], dict(comb_logit, comb_prob (log-softmax), obj_logit, obj_prob)) for a training-mode
    pass (BatchNorm batch statistics); the reference additionally back-propagates and steps its optimizers inside
    forward (:237-248) -- not part of this slice."""

    def name(self):
        return "TwoStreamAE_mask"

    def __init__(self, opt):
        if not torch.cuda.is_available():
            raise RuntimeError("TwoStreamAE_mask (B200) needs a CUDA device: there is no CPU fallback")
        self.opt = opt
        dev = torch.device("cuda", opt.gpu_ids[0] if len(opt.gpu_ids) else torch.cuda.current_device())
        self.device = dev
        prec = getattr(opt, "precision", "bf16x3")
        self.ctx = ops.Ctx(dev, split=(prec != "bf16"), split_bwd=(prec == "bf16x3"))
        if getattr(opt, "no_comb", False):
            raise NotImplementedError("--no_comb selects MaskTwoStreamConvSwitch_NET, outside this slice")
        self.fpG = FlatParams(dev)
        self.netG = MaskTwoStreamConvNet(self.ctx, self.fpG, opt.label_nc, opt.output_nc, opt.conv_dim, opt.num_layers,
                                         opt.conv_size, opt.n_blocks, opt.cond_in, opt.which_stream,
                                         getattr(opt, "num_resnetblocks", 1), getattr(opt, "norm_layer", "batch"),
                                         getattr(opt, "use_simpleRes", False))
        self.fpG.materialize()
        self.netG.init_reference(torch.Generator().manual_seed(getattr(opt, "init_seed", 0)))
        self.loss_names = ["G_Recon_comb", "G_Recon_obj", "KL_loss", "loss_G_GAN", "loss_D_GAN", "loss_G_GAN_Feat"]
        self.acc = torch.zeros(3, dtype=torch.float64, device=dev)

    def _dev(self, t):
        return t.to(self.device, torch.float32).contiguous()

    def forward(self, label_map, mask_obj_in, mask_ctx_in, mask_obj_out, mask_out, mask_obj_inst, cls, mask_in,
                eval_mode=False):
        if eval_mode:
            raise NotImplementedError("eval mode uses BatchNorm running statistics, which this slice does not track")
        opt, ctx = self.opt, self.ctx
        label_map, mask_ctx_in, mask_out, mask_in, inst = (self._dev(t) for t in (label_map, mask_ctx_in, mask_out,
                                                                                     mask_in, mask_obj_inst))
        clsf = self._dev(cls.reshape(-1))
        B, _, H, W = label_map.shape
        cond = ops.box2mask_encode(ctx, mask_ctx_in, mask_in, clsf, opt.label_nc)          # :127-151, :331-338
        ctx_logit, obj_logit = self.netG.forward(cond)
        C = opt.output_nc
        out = dict(comb_logit=torch.empty(B, C, H, W, device=self.device), comb_prob=torch.empty(B, C, H, W, device=self.device),
                   obj_logit=obj_logit[..., :1].permute(0, 3, 1, 2), obj_prob=torch.empty(B, 1, H, W, device=self.device))
        self.acc.zero_()
        ops.box2mask_head(ctx, ctx_logit, obj_logit, label_map, mask_out, inst, getattr(opt, "use_output_gate", False),
                          out["comb_logit"], out["comb_prob"], out["obj_prob"], self.acc)
        loss_comb = (self.acc[0] / self.acc[1].clamp_min(1.0)).float()       # NLLLoss2d mean over non-ignored pixels
        loss_obj = (self.acc[2] / float(B * H * W)).float()                  # BCELoss mean
        return [loss_comb, loss_obj], out

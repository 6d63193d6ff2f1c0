"""box2mask generator and trainer -- BASELINE config #5 / SURVEY N3.

  MaskTwoStreamConvNet   <- models/MaskTwoStreamConv_NET.py:13-219 (+ MaskContextAE_NET base, layer_util.py:119-242,333-378)
  TwoStreamAE_mask       <- models/TwoStreamAE_mask.py: encode_input :127-151, reconstruct :257-300, forward :167-255 --
                            forward pass, the two reconstruction losses, `loss_G.backward()` and the Adam step the
                            reference runs INSIDE forward (:233-240); with `--use_gan --which_gan patch_multiscale` (the
                            shipped script) also the PatchGAN terms (:205-231): a 2-scale BatchNorm
                            MultiscaleDiscriminator, LSGAN losses, the reported feature-matching term, lr_control and
                            the discriminator's own Adam step (:243-248); `--no_comb` (MaskTwoStreamConvSwitch_NET,
                            :29-32); BatchNorm running statistics and the eval-mode `reconstruct` / `generate` /
                            `evaluate` (:257-335); `save` / `load` / `delete_model` / `update_learning_rate` in the
                            reference's checkpoint layout (:106-118, 359-383); `--norm_layer instance` and
                            `--add_dilated_layers` (what scripts/train_box2mask_ade.sh adds).  `which_gan` 'patch' /
                            'patch_res' raise.

Parameter names are the reference's own ('<params_dict key>.<state_dict key>': conv_encoder_3.deep.1.weight,
ctx_conv_decoder_1.shortcut.0.weight, latent_encoder.0.conv_block.1.weight, ...), OIHW / IOHW fp32 like the reference, so
its `save_network_dict` checkpoints map one to one.  Every convolution (7x7 s2, 4x4 s2, 1x1 s2, 3x3 reflect,
ConvTranspose 4x4 s2, 1x1, 3x3; forward, data gradient, weight gradient) runs on the tcgen05 engines; BatchNorm (training
mode, batch statistics) is hm_in_stats over the batch-folded tensor + hm_bn_fold + hm_in_apply forward and hm_bn_bwd
backward; the rest is csrc/hm_box2mask.cu.

Two aliasing effects of the reference are part of its arithmetic (see oracle/box2mask.py): each Conv/DeconvResnetBlock
rectifies its input IN PLACE, so both of its branches -- and the encoder features kept for the skip connections -- see
relu(x).
"""
import os
from collections import OrderedDict

import torch

from . import ops
from .networks import ConvP, FlatParams, _f32
from .ops import ACT_NONE, ACT_RELU, Operand

DIM_LIST_TAIL = [96, 128, 256, 512]      # MaskTwoStreamConv_NET.py:25


class _BN(object):
    """nn.BatchNorm2d(C, affine=True): gain / shift parameters and the running_mean / running_var / num_batches_tracked
    buffers (momentum 0.1, eps 1e-5).  Training mode normalises with the batch statistics and updates the buffers; eval
    mode (`training = False`) normalises with the buffers."""

    EPS, MOMENTUM = 1e-5, 0.1

    def __init__(self, fp, name, c):
        self.fp, self.name, self.c = fp, name, c
        self.training = True
        fp.declare(name + ".weight", (c,))
        fp.declare(name + ".bias", (c,))
        fp.buffers[name + ".running_mean"] = torch.zeros(c, dtype=torch.float32, device=fp.device)
        fp.buffers[name + ".running_var"] = torch.ones(c, dtype=torch.float32, device=fp.device)
        fp.buffers[name + ".num_batches_tracked"] = torch.zeros((), dtype=torch.int64, device=fp.device)

    @property
    def gamma(self):
        return self.fp.params[self.name + ".weight"]

    @property
    def beta(self):
        return self.fp.params[self.name + ".bias"]

    @property
    def running(self):
        b = self.fp.buffers
        return b[self.name + ".running_mean"], b[self.name + ".running_var"], b[self.name + ".num_batches_tracked"]

    def init_reference(self, gen):
        """weights_init (layer_util.py:13-15): gain ~ N(1, 0.02), shift 0."""
        with torch.no_grad():
            self.gamma.copy_((1.0 + torch.randn(self.c, generator=gen) * 0.02).to(self.gamma.device))
            self.beta.zero_()
        self.fp.version += 1

    def apply(self, ctx, y, act, skip=None, out32=None, out_op=None, reflect=True, repeat=1):
        """Returns the statistics (mean, rstd) [1, C] the tensor was normalised with (the backward pass needs the batch
        statistics).  repeat: how many times the reference evaluates this module on this batch (buffer update count)."""
        N, H, W, C = y.shape
        if self.training:                                                                 # statistics over (N, H, W)
            mean, rstd, mean_n, rstd_n = ops.bn_stats(ctx, y, self.gamma, self.beta, running=self.running, repeat=repeat,
                                                      momentum=self.MOMENTUM, eps=self.EPS)
        else:
            rm, rv, _ = self.running
            mean, rstd = rm.view(1, C), torch.rsqrt(rv + self.EPS).view(1, C)
            mean_n, rstd_n = ops.bn_fold(ctx, mean, rstd, self.gamma, self.beta, N)
        ops.in_apply(ctx, y, mean_n, rstd_n, act, skip=skip, out32=out32, out_op=out_op, reflect=reflect)
        return mean, rstd

    def backward(self, ctx, y, stats, act, g1, g2=None, out_op=None, out32=None):
        """Gradient w.r.t. the BN(+act) output (g1 + g2, dense fp32) -> gradient w.r.t. the conv output y; accumulates
        d(gamma), d(beta)."""
        ops.bn_bwd(ctx, y, stats[0], stats[1], self.gamma, self.beta, act, g1, g2=g2, out_op=out_op, out32=out32,
                   dgamma=self.gamma.grad, dbeta=self.beta.grad)


class _IN(object):
    """nn.InstanceNorm2d(C, affine=False) (--norm_layer instance, layer_util.py:22-23) with _BN's interface: no
    parameters, no buffers, the same arithmetic in training and in eval mode."""

    def __init__(self, fp, name, c):
        self.fp, self.name, self.c = fp, name, c
        self.training = True

    def init_reference(self, gen):
        pass

    def apply(self, ctx, y, act, skip=None, out32=None, out_op=None, reflect=True, repeat=1):
        mean, rstd = ops.in_stats(ctx, y)
        ops.in_apply(ctx, y, mean, rstd, act, skip=skip, out32=out32, out_op=out_op, reflect=reflect)
        return mean, rstd

    def backward(self, ctx, y, stats, act, g1, g2=None, out_op=None, out32=None):
        ops.in_bwd(ctx, tuple(y.shape), act, 0.2, y=y, mean=stats[0], rstd=stats[1], g1=g1, g1_border=0, g2=g2,
                   out_op=out_op, out32=out32)


class MaskTwoStreamConvNet(object):
    def __init__(self, ctx, fp, label_nc, output_nc, conv_dim=64, num_layers=3, conv_size=4, n_blocks=6,
                 cond_in="ctx_obj", which_stream="obj_context", num_resnetblocks=1, norm_layer="batch", use_simpleRes=False,
                 add_dilated_layers=False):
        if which_stream != "obj_context" or num_resnetblocks != 1 or norm_layer not in ("batch", "instance") or use_simpleRes:
            raise NotImplementedError("box2mask: only the shipped configurations (which_stream obj_context, one conv per "
                                      "block, batch or instance norm, ConvResnetBlock) are part of this slice")
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
            b = (_BN if norm_layer == "batch" else _IN)(fp, name, c)
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

        def res_blocks(prefix, n, first=0):
            out = []
            for j in range(first, first + n):
                p = "%s.%d.conv_block" % (prefix, j)
                out.append(dict(c1=conv(p + ".1", self.latent_dim, self.latent_dim, 3, 1, 0), b1=bn(p + ".2", self.latent_dim),
                                c2=conv(p + ".5", self.latent_dim, self.latent_dim, 3, 1, 0), b2=bn(p + ".6", self.latent_dim)))
            return out
        # --add_dilated_layers (MaskTwoStreamConvSwitch_NET.py:101-104): two DilatedResnetBlocks (dilation 2 and 4, conv3x3
        # without bias, layer_util.py:255-293) in front of the latent encoder's ResnetBlocks
        self.dilated = []
        if add_dilated_layers:
            for j, d in enumerate((2, 4)):
                p = "latent_encoder.%d" % j
                self.dilated.append(dict(d=d, c1=conv(p + ".conv1", self.latent_dim, self.latent_dim, 3, 1, d, dilation=d, bias=False),
                                         b1=bn(p + ".bn1", self.latent_dim),
                                         c2=conv(p + ".conv2", self.latent_dim, self.latent_dim, 3, 1, d, dilation=d, bias=False),
                                         b2=bn(p + ".bn2", self.latent_dim)))
        self.latent_encoder = res_blocks("latent_encoder", n_blocks // 2, first=len(self.dilated))      # :92-106

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

    def set_mode(self, eval_mode=False):
        """MaskContextAE_NET.set_mode: module.eval() / module.train() on every entry of params_dict."""
        for b in self.bns_:
            b.training = not eval_mode
        self.eval_mode = bool(eval_mode)

    def get_mode(self):
        return getattr(self, "eval_mode", False)

    # ---- forward --------------------------------------------------------------------------------------
    def _res_block(self, blk, x32, x_op, want_op, tape):
        """ResnetBlock (layer_util.py:333-378): x + [pad, conv, norm, relu, pad, conv, norm](x)."""
        ctx = self.ctx
        N, H, W, C = x32.shape
        y = _f32(ctx, N, H, W, C)
        blk["c1"].forward(x_op, 0, out32=y)
        mid = Operand(ctx, N, H, W, C, border=1)
        st1 = blk["b1"].apply(ctx, y, ACT_RELU, out_op=mid, reflect=True)
        y2 = _f32(ctx, N, H, W, C)
        blk["c2"].forward(mid, 0, out32=y2)
        out32 = _f32(ctx, N, H, W, C)
        out_op = Operand(ctx, N, H, W, C, border=1) if want_op else None
        st2 = blk["b2"].apply(ctx, y2, ACT_NONE, skip=x32, out32=out32, out_op=out_op, reflect=True)
        tape.append(dict(blk=blk, x_op=x_op, y1=y, st1=st1, mid=mid, y2=y2, st2=st2))
        return out32, out_op

    def _dilated_block(self, blk, x32, tape):
        """DilatedResnetBlock (layer_util.py:277-293): relu(norm(conv_d(relu(norm(conv_d(x))))) + x), zero padding = dilation."""
        ctx, d = self.ctx, blk["d"]
        N, H, W, C = x32.shape
        x_op = Operand(ctx, N, H, W, C)
        ops.in_apply(ctx, x32, None, None, ACT_NONE, out_op=x_op, reflect=False)
        y1 = _f32(ctx, N, H, W, C)
        blk["c1"].forward(x_op, d, out32=y1)
        mid = Operand(ctx, N, H, W, C)
        st1 = blk["b1"].apply(ctx, y1, ACT_RELU, out_op=mid, reflect=False)
        y2 = _f32(ctx, N, H, W, C)
        blk["c2"].forward(mid, d, out32=y2)
        pre = _f32(ctx, N, H, W, C)
        st2 = blk["b2"].apply(ctx, y2, ACT_NONE, skip=x32, out32=pre, reflect=False)
        out32 = _f32(ctx, N, H, W, C)
        ops.in_apply(ctx, pre, None, None, ACT_RELU, out32=out32, reflect=False)        # the block's trailing ReLU
        tape.append(dict(blk=blk, x_op=x_op, y1=y1, st1=st1, mid=mid, y2=y2, st2=st2, out=out32))
        return out32

    def _dilated_block_bwd(self, t, g_out):
        ctx, blk, d = self.ctx, t["blk"], t["blk"]["d"]
        N, H, W, C = g_out.shape
        g_pre = _f32(ctx, N, H, W, C)
        ops.in_bwd(ctx, (N, H, W, C), ACT_RELU, z=t["out"], g2=g_out, out32=g_pre)      # through the trailing ReLU
        dy2 = Operand(ctx, N, H, W, C, grad=True)
        blk["b2"].backward(ctx, t["y2"], t["st2"], ACT_NONE, g_pre, out_op=dy2)
        blk["c2"].wgrad(t["mid"], dy2, d, bias_grad=False)
        gmid = _f32(ctx, N, H, W, C)
        blk["c2"].dgrad(dy2, H, W, d, gmid)
        dy1 = Operand(ctx, N, H, W, C, grad=True)
        blk["b1"].backward(ctx, t["y1"], t["st1"], ACT_RELU, gmid, out_op=dy1)
        blk["c1"].wgrad(t["x_op"], dy1, d, bias_grad=False)
        gx = _f32(ctx, N, H, W, C)
        blk["c1"].dgrad(dy1, H, W, d, gx)
        return self._add(gx, g_pre)                                                     # + the identity branch

    def _decode(self, latent32, latent_op, res, blocks, final, skips):
        ctx = self.ctx
        tape = dict(res=[], blocks=[])
        d32, d_op = latent32, latent_op
        for j, blk in enumerate(res):
            d32, d_op = self._res_block(blk, d32, d_op, want_op=(j + 1 < len(res)), tape=tape["res"])
        N = d32.shape[0]
        for i, blk in enumerate(blocks):
            h, w, c = d32.shape[1], d32.shape[2], d32.shape[3]
            xr = Operand(ctx, N, h, w, c)                                  # relu(x): the in-place ReLU of the block
            ops.in_apply(ctx, d32, None, None, ACT_RELU, out_op=xr, reflect=False)
            skip = skips[-1 - (i - 1)] if (skips and 1 <= i <= self.num_layers) else None
            x_op = ops.concat_operands(ctx, skip, xr) if skip is not None else xr
            od = blk["out_dim"]
            y = _f32(ctx, N, 2 * h, 2 * w, od)
            blk["deep"].forward(x_op, 0, out32=y)                          # ConvTranspose2d k4 s2 p1
            deep32 = _f32(ctx, N, 2 * h, 2 * w, od)
            st_d = blk["deep_bn"].apply(ctx, y, ACT_NONE, out32=deep32)
            ys = st_s = None
            if blk["short"] is not None:
                ys = _f32(ctx, N, h, w, od)
                blk["short"].forward(x_op, 0, out32=ys)
                s32 = _f32(ctx, N, h, w, od)
                st_s = blk["short_bn"].apply(ctx, ys, ACT_NONE, out32=s32)
            else:
                s32 = _f32(ctx, N, h, w, c)
                ops.in_apply(ctx, d32, None, None, ACT_RELU, out32=s32)
            d32 = _f32(ctx, N, 2 * h, 2 * w, od)
            ops.upsample2_add(ctx, s32, deep32, d32)
            tape["blocks"].append(dict(blk=blk, xr=xr, x_op=x_op, skip_c=(skip.c if skip is not None else 0), y=y, st_d=st_d,
                                       ys=ys, st_s=st_s, in_shape=(N, h, w, c)))
        fin = Operand(ctx, N, d32.shape[1], d32.shape[2], d32.shape[3])
        ops.in_apply(ctx, d32, None, None, ACT_NONE, out_op=fin, reflect=False)      # the last conv sees dec_feat as is
        logit = _f32(ctx, N, d32.shape[1], d32.shape[2], final.cout)
        final.forward(fin, 1, out32=logit)
        tape.update(fin=fin, final=final)
        return logit, tape

    def forward(self, cond_op):
        """cond_op: Operand [B,S,S,input_nc].  Returns the dense fp32 NHWC logits (ctx_logit [B,S,S,output_nc],
        obj_logit [B,S,S,1]) -- MaskTwoStreamConv_NET.forward :166-196; the head kernel combines them -- and the tape."""
        ctx = self.ctx
        conv0, bn0 = self.enc0
        N = cond_op.n
        ho, wo = conv0.out_hw(cond_op.h, cond_op.w, 3)
        y0 = _f32(ctx, N, ho, wo, conv0.cout)
        conv0.forward(cond_op, 3, out32=y0)
        cur = Operand(ctx, N, ho, wo, conv0.cout)                          # relu(bn(.)) >= 0: in-place ReLU is a no-op
        st0 = bn0.apply(ctx, y0, ACT_RELU, out_op=cur, reflect=False)
        tape = dict(cond=cond_op, y0=y0, st0=st0, enc=[], lat=[])
        skips = [cur]
        h32 = None
        for i, blk in enumerate(self.enc_blocks):                          # ConvResnetBlock (layer_util.py:119-162)
            ho, wo = blk["deep"].out_hw(cur.h, cur.w, blk["deep"].pad)
            c = blk["deep"].cout
            yd, ys = _f32(ctx, N, ho, wo, c), _f32(ctx, N, ho, wo, c)
            blk["deep"].forward(cur, blk["deep"].pad, out32=yd)
            blk["short"].forward(cur, 0, out32=ys)
            tmp, h32 = _f32(ctx, N, ho, wo, c), _f32(ctx, N, ho, wo, c)
            st_d = blk["deep_bn"].apply(ctx, yd, ACT_NONE, out32=tmp)
            st_s = blk["short_bn"].apply(ctx, ys, ACT_NONE, skip=tmp, out32=h32)
            tape["enc"].append(dict(blk=blk, xin=cur, yd=yd, ys=ys, st_d=st_d, st_s=st_s, shape=(N, ho, wo, c)))
            if i + 1 < len(self.enc_blocks):
                cur = Operand(ctx, N, ho, wo, c)                           # rectified in place by the next block
                ops.in_apply(ctx, h32, None, None, ACT_RELU, out_op=cur, reflect=False)
                skips.append(cur)
        tape["dil"] = []
        for blk in self.dilated:
            h32 = self._dilated_block(blk, h32, tape["dil"])
        lat_op = Operand(ctx, N, h32.shape[1], h32.shape[2], h32.shape[3], border=1)
        ops.in_apply(ctx, h32, None, None, ACT_NONE, out_op=lat_op, reflect=True)
        lat32 = h32
        for j, blk in enumerate(self.latent_encoder):
            lat32, lat_op = self._res_block(blk, lat32, lat_op, want_op=True, tape=tape["lat"])
        ctx_logit, tape["ctx"] = self._decode(lat32, lat_op, self.ctx_latent, self.ctx_dec, self.ctx_final, skips)
        obj_logit, tape["obj"] = self._decode(lat32, lat_op, self.obj_latent, self.obj_dec, self.obj_final, None)
        tape["skips"] = skips
        return ctx_logit, obj_logit, tape

    # ---- backward -------------------------------------------------------------------------------------
    def _add(self, a, b):
        out = torch.empty_like(a)
        ops.fold_add(self.ctx, a, 0, b, out)
        return out

    def _res_block_bwd(self, t, g_out):
        """g_out: dense fp32 gradient w.r.t. the block output -> gradient w.r.t. its input."""
        ctx, blk = self.ctx, t["blk"]
        N, H, W, C = g_out.shape
        dy2 = Operand(ctx, N, H, W, C, grad=True)
        blk["b2"].backward(ctx, t["y2"], t["st2"], ACT_NONE, g_out, out_op=dy2)
        blk["c2"].wgrad(t["mid"], dy2, 0, bias_grad=False)
        gmid_p = _f32(ctx, N, H + 2, W + 2, C)
        blk["c2"].dgrad(dy2, H + 2, W + 2, 0, gmid_p)
        gmid = _f32(ctx, N, H, W, C)
        ops.fold_add(ctx, gmid_p, 1, None, gmid)                          # adjoint of ReflectionPad2d(1)
        dy1 = Operand(ctx, N, H, W, C, grad=True)
        blk["b1"].backward(ctx, t["y1"], t["st1"], ACT_RELU, gmid, out_op=dy1)
        blk["c1"].wgrad(t["x_op"], dy1, 0, bias_grad=False)
        gx_p = _f32(ctx, N, H + 2, W + 2, C)
        blk["c1"].dgrad(dy1, H + 2, W + 2, 0, gx_p)
        gx = _f32(ctx, N, H, W, C)
        ops.fold_add(ctx, gx_p, 1, g_out, gx)                             # + the identity branch
        return gx

    def _decode_bwd(self, tape, dlogit, skip_grads):
        """dlogit: gradient operand w.r.t. the stream's logits.  Returns the dense gradient w.r.t. the latent feature;
        skip_grads[k] collects (tensor, ld, channels) gradients w.r.t. the encoder skip features."""
        ctx = self.ctx
        final, fin = tape["final"], tape["fin"]
        final.wgrad(fin, dlogit, 1, bias_grad=True)
        g = _f32(ctx, fin.n, fin.h, fin.w, final.cin)
        final.dgrad(dlogit, fin.h, fin.w, 1, g)
        for i in range(len(tape["blocks"]) - 1, -1, -1):
            t = tape["blocks"][i]
            blk = t["blk"]
            N, h, w, c = t["in_shape"]
            od, x_op = blk["out_dim"], t["x_op"]
            # deep branch: BN <- ConvTranspose
            dyd = Operand(ctx, N, 2 * h, 2 * w, od, grad=True)
            blk["deep_bn"].backward(ctx, t["y"], t["st_d"], ACT_NONE, g, out_op=dyd)
            blk["deep"].wgrad(x_op, dyd, 0, bias_grad=False)
            gxa = _f32(ctx, N, h, w, x_op.c)
            blk["deep"].dgrad(dyd, h, w, 0, gxa)
            # shortcut branch: bilinear^T <- BN <- Conv 1x1
            gs = _f32(ctx, N, h, w, od)
            ops.upsample2_bwd(ctx, g, gs)
            if blk["short"] is not None:
                dys = Operand(ctx, N, h, w, od, grad=True)
                blk["short_bn"].backward(ctx, t["ys"], t["st_s"], ACT_NONE, gs, out_op=dys)
                blk["short"].wgrad(x_op, dys, 0, bias_grad=False)
                gxb = _f32(ctx, N, h, w, x_op.c)
                blk["short"].dgrad(dys, h, w, 0, gxb)
            else:
                gxb = gs
            gx = self._add(gxa, gxb)                                       # w.r.t. cat(skip, relu(d)) or relu(d)
            sc = t["skip_c"]
            if sc:
                skip_grads[len(skip_grads_index(self.num_layers)) - 1 - (i - 1)].append((gx, x_op.c, sc))
            g = _f32(ctx, N, h, w, c)                                       # through the block's in-place ReLU
            ops.in_bwd(ctx, (N, h, w, c), ACT_RELU, mask_op=t["xr"], g1=gx, g1_border=0, g1_ld=x_op.c, g1_coff=sc, out32=g)
        for t in reversed(tape["res"]):
            g = self._res_block_bwd(t, g)
        return g

    def backward(self, tape, d_ctx, d_obj):
        """Accumulates every parameter gradient given the gradient operands w.r.t. the two logit tensors."""
        ctx = self.ctx
        skip_grads = [[] for _ in skip_grads_index(self.num_layers)]
        g_lat = self._add(self._decode_bwd(tape["ctx"], d_ctx, skip_grads), self._decode_bwd(tape["obj"], d_obj, skip_grads))
        for t in reversed(tape["lat"]):
            g_lat = self._res_block_bwd(t, g_lat)
        for t in reversed(tape["dil"]):
            g_lat = self._dilated_block_bwd(t, g_lat)
        g_h = g_lat                                                        # w.r.t. the last encoder block's output
        for i in range(len(tape["enc"]) - 1, -1, -1):
            t = tape["enc"][i]
            blk, xin = t["blk"], t["xin"]
            N, ho, wo, c = t["shape"]
            dyd, dys = Operand(ctx, N, ho, wo, c, grad=True), Operand(ctx, N, ho, wo, c, grad=True)
            blk["deep_bn"].backward(ctx, t["yd"], t["st_d"], ACT_NONE, g_h, out_op=dyd)
            blk["short_bn"].backward(ctx, t["ys"], t["st_s"], ACT_NONE, g_h, out_op=dys)
            blk["deep"].wgrad(xin, dyd, blk["deep"].pad, bias_grad=False)
            blk["short"].wgrad(xin, dys, 0, bias_grad=False)
            ga = _f32(ctx, N, xin.h, xin.w, xin.c)
            blk["deep"].dgrad(dyd, xin.h, xin.w, blk["deep"].pad, ga)
            gb = torch.zeros(N, xin.h, xin.w, xin.c, dtype=torch.float32, device=ctx.device)   # 1x1 stride 2: odd pixels get none
            blk["short"].dgrad(dys, xin.h, xin.w, 0, gb)
            gsum = self._add(ga, gb)                                       # w.r.t. xin = relu(previous feature)
            for sg, ld, sc in skip_grads[i]:                               # + the decoder's use of the same rectified feature
                nxt = _f32(ctx, N, xin.h, xin.w, xin.c)
                ops.in_bwd(ctx, (N, xin.h, xin.w, xin.c), ACT_NONE, g1=sg, g1_border=0, g1_ld=ld, g1_coff=0, g2=gsum, out32=nxt)
                gsum = nxt
            if i > 0:                                                      # through the in-place ReLU onto the previous block's output
                g_h = _f32(ctx, N, xin.h, xin.w, xin.c)
                ops.in_bwd(ctx, (N, xin.h, xin.w, xin.c), ACT_RELU, mask_op=xin, g2=gsum, out32=g_h)
            else:
                g_h = gsum
        conv0, bn0 = self.enc0
        cond = tape["cond"]
        y0 = tape["y0"]
        dy0 = Operand(ctx, y0.shape[0], y0.shape[1], y0.shape[2], y0.shape[3], grad=True)
        bn0.backward(ctx, y0, tape["st0"], ACT_RELU, g_h, out_op=dy0)      # ReLU mask from gamma * xhat + beta
        conv0.wgrad(cond, dy0, 3, bias_grad=False)


def skip_grads_index(num_layers):
    """Encoder skip features are indexed like enc_features (MaskTwoStreamConv_NET.py:172-173): entry i is the rectified
    input of encoder block i, used by decoder block num_layers - i."""
    return list(range(num_layers))


class BNMultiscaleDiscriminator(object):
    """MultiscaleDiscriminator(input_nc, ndf, n_layers, 'batch', use_sigmoid=False, num_D, getIntermFeat=True) as the
    box2mask trainer builds it for which_gan == 'patch_multiscale' (TwoStreamAE_mask.py:83-92; Discriminator_NET.py:11-118):
    per scale  conv4x4 s2 p2 + LReLU | (n_layers - 1) x [conv s2, BatchNorm, LReLU] | [conv s1, BatchNorm, LReLU] | conv s1
    -> 1 channel, on an AvgPool(3, 2, 1, count_include_pad=False) pyramid; level i of the pyramid uses scale{num_D-1-i}.
    BatchNorm runs in training mode: every pass is normalised with its own batch statistics, which is why the real and
    the fake batch are evaluated separately (unlike the InstanceNorm discriminator of mask2image)."""

    def __init__(self, ctx, fp, input_nc, ndf=64, n_layers=3, num_D=2, norm_layer="batch"):
        self.ctx, self.fp, self.input_nc, self.n_layers, self.num_D = ctx, fp, input_nc, n_layers, num_D
        norm = _BN if norm_layer == "batch" else _IN      # --norm_layer instance: InstanceNorm2d(affine=False), no parameters
        self.scales, self.convs_, self.bns_ = [], [], []
        for s in range(num_D):
            layers, nf, nf_prev = [], ndf, input_nc
            for j in range(n_layers + 2):
                name = "scale%d_layer%d" % (s, j)
                cout = 1 if j == n_layers + 1 else nf
                conv = ConvP(ctx, fp, name + ".0", nf_prev, cout, 4, 2 if j < n_layers else 1, 2)
                bn = norm(fp, name + ".1", cout) if 1 <= j <= n_layers else None
                layers.append((conv, bn))
                self.convs_.append(conv)
                if bn is not None:
                    self.bns_.append(bn)
                nf_prev, nf = nf, min(nf * 2, 512)
            self.scales.append(layers)

    def init_reference(self, gen):
        for c in self.convs_:
            c.init_reference(gen)
        for b in self.bns_:
            b.init_reference(gen)

    def forward(self, d_in, repeat=1):
        """d_in: Operand [N,H,W,input_nc].  Returns one dict per pyramid level: layers, xs (layer inputs), taps (fp32 NHWC
        layer outputs = the reference's intermediate features), ys / stats (pre-BatchNorm conv outputs, batch statistics).
        repeat: how many evaluations of the reference this pass stands for (BatchNorm buffer updates)."""
        ctx = self.ctx
        tape, x = [], d_in
        for i in range(self.num_D):
            layers = self.scales[self.num_D - 1 - i]
            lv = dict(layers=layers, xs=[], taps=[], ys=[], stats=[])
            cur = x
            for j, (conv, bn) in enumerate(layers):
                ho, wo = conv.out_hw(cur.h, cur.w, 2)
                lv["xs"].append(cur)
                tap = _f32(ctx, cur.n, ho, wo, conv.cout)
                y = st = nxt = None
                if j == 0:
                    nxt = Operand(ctx, cur.n, ho, wo, conv.cout, zero=(conv.cout % 8 != 0))
                    conv.forward(cur, 2, act=ops.ACT_LRELU, slope=0.2, out32=tap, out16=nxt)
                elif bn is None:
                    conv.forward(cur, 2, out32=tap)
                else:
                    y = _f32(ctx, cur.n, ho, wo, conv.cout)
                    conv.forward(cur, 2, out32=y)
                    nxt = Operand(ctx, cur.n, ho, wo, conv.cout, zero=(conv.cout % 8 != 0))
                    st = bn.apply(ctx, y, ops.ACT_LRELU, out32=tap, out_op=nxt, reflect=False, repeat=repeat)
                lv["taps"].append(tap); lv["ys"].append(y); lv["stats"].append(st)
                cur = nxt
            tape.append(lv)
            if i != self.num_D - 1:
                xn = Operand(ctx, x.n, (x.h - 1) // 2 + 1, (x.w - 1) // 2 + 1, x.c, cs=x.cs)
                if x.lo is None:
                    xn.lo = None
                ops.avgpool3s2(ctx, x, xn)
                x = xn
        return tape

    def backward(self, tape, target, coef, weight_grads):
        """Backward pass of  coef * sum_levels mse(pred_level, target)  (GANLoss with LSGAN, models/losses.py:40-50).
        weight_grads=True: accumulates the gradients of every discriminator parameter (the loss_D pass, :243-248);
        False: only the data path, and returns the fp32 gradient w.r.t. input channels 0..2 of the full-resolution
        input as [N,H,W,4] (the generator's loss_G_GAN term, :231-235; channel 0 is the generated mask)."""
        ctx = self.ctx
        gins = []
        for lv in tape:
            layers = lv["layers"]
            nl = len(layers)
            pred = lv["taps"][-1]
            N = pred.shape[0]
            dy = Operand(ctx, N, pred.shape[1], pred.shape[2], 1, grad=True)
            ops.mse_grad(ctx, pred, target, 2.0 * coef / pred.numel(), dy)
            for j in range(nl - 1, -1, -1):
                conv, _ = layers[j]
                xin = lv["xs"][j]
                if weight_grads:
                    conv.wgrad(xin, dy, 2, bias_grad=(j == 0 or j == nl - 1))
                    if j == 0:
                        break
                elif j == 0:
                    gin = _f32(ctx, N, xin.h, xin.w, 4)
                    conv.dgrad_rows(dy, xin.h, xin.w, 2, gin, 0, 3)
                    gins.append(gin)
                    break
                gin = _f32(ctx, N, xin.h, xin.w, conv.cin)
                conv.dgrad(dy, xin.h, xin.w, 2, gin)
                tap = lv["taps"][j - 1]
                dyn = Operand(ctx, N, tap.shape[1], tap.shape[2], tap.shape[3], grad=True)
                bn = layers[j - 1][1]
                if bn is None:       # layer 0: LeakyReLU only
                    ops.in_bwd(ctx, tuple(tap.shape), ops.ACT_LRELU, 0.2, z=tap, g1=gin, g1_border=0, out_op=dyn)
                else:
                    st = lv["stats"][j - 1]
                    if isinstance(bn, _IN):
                        ops.in_bwd(ctx, tuple(tap.shape), ops.ACT_LRELU, 0.2, y=lv["ys"][j - 1], mean=st[0], rstd=st[1], g1=gin,
                                   g1_border=0, out_op=dyn)
                    else:
                        ops.bn_bwd(ctx, lv["ys"][j - 1], st[0], st[1], bn.gamma, bn.beta, ops.ACT_LRELU, gin, out_op=dyn,
                                   dgamma=bn.gamma.grad if weight_grads else None,
                                   dbeta=bn.beta.grad if weight_grads else None, slope=0.2)
                dy = dyn
        if weight_grads:
            return None
        for i in range(len(gins) - 1, 0, -1):
            ops.avgpool3s2_bwd(ctx, gins[i], gins[i - 1], 0, 3)
        return gins[0]


def lr_control(loss_G, loss_D_real, loss_D_fake, gan_margin=0.3):
    """Discriminator_NET.py:190-211: freeze D when it is winning, G when it is losing, never both.  Returns (g_lr, d_lr)."""
    update_d = not (loss_D_real < gan_margin or loss_D_fake < gan_margin)
    update_g = not (loss_D_real > 1 - gan_margin or loss_D_fake > 1 - gan_margin)
    if not (update_d or update_g):
        update_d = update_g = True
    what = "Update Both" if (update_g and update_d) else ("Froze Generator" if not update_g else "Froze Discriminator")
    print("%s\t[G=%.3f],[DR=%.3f],[DF=%.3f]" % (what, loss_G, loss_D_real, loss_D_fake))
    return float(update_g), float(update_d)


class TwoStreamAE_mask(object):
    """models/TwoStreamAE_mask.py.  `forward(...)` in training mode runs the reference's whole iteration (:167-255):
    forward, loss_recon_comb (MaskReconLoss) and loss_recon_obj (BCE), with `--use_gan` the discriminator passes and
    loss_G_GAN / loss_D, the backward pass of `loss_recon_obj + rec_weight * loss_recon_comb + gan_weight * loss_G_GAN`
    and `optimizer.step()` (Adam(lr, beta1, beta2)), then the discriminator's backward pass and `optimizer_D.step()`, and
    returns ([loss_recon_comb, loss_recon_obj, 0, loss_G_GAN, loss_D, loss_G_GAN_Feat], [comb_recon_label,
    obj_recon_prob]) like the reference.  With `train=False` it stops after the losses (the parity tests use that to
    inspect outputs and gradients)."""

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
        # --no_comb selects MaskTwoStreamConvSwitch_NET (:29-32): the same network, but its forward returns the context
        # stream's logits / log-softmax as they are instead of gating them with the object stream (:208 vs :190-217)
        self.no_comb = bool(getattr(opt, "no_comb", False))
        # --add_dilated_layers only exists in the Switch network (MaskTwoStreamConvSwitch_NET.py:29,101-104)
        dilated = bool(getattr(opt, "add_dilated_layers", False)) and self.no_comb
        self.isTrain = bool(getattr(opt, "isTrain", True))
        self.gpu_ids = opt.gpu_ids
        self.save_dir = os.path.join(getattr(opt, "checkpoints_dir", "./checkpoints"), opt.name)
        self.use_gan = bool(getattr(opt, "use_gan", False))
        if self.use_gan and getattr(opt, "which_gan", "patch") != "patch_multiscale":
            raise NotImplementedError("--which_gan: only 'patch_multiscale' (the shipped setting, TwoStreamAE_mask.py:83-92) "
                                      "is built; 'patch' / 'patch_res' use the conditional single-scale discriminators")
        # :48-55 criterionObjRecon: 'l1' -> nn.L1Loss, 'bce' -> nn.BCELoss, anything else -> no object-mask loss
        self.obj_loss = getattr(opt, "objReconLoss", "bce")
        self.fpG = FlatParams(dev)
        self.netG = MaskTwoStreamConvNet(self.ctx, self.fpG, opt.label_nc, opt.output_nc, opt.conv_dim, opt.num_layers,
                                         opt.conv_size, opt.n_blocks, opt.cond_in, opt.which_stream,
                                         getattr(opt, "num_resnetblocks", 1), getattr(opt, "norm_layer", "batch"),
                                         getattr(opt, "use_simpleRes", False), add_dilated_layers=dilated)
        self.fpG.materialize()
        self.netG.init_reference(torch.Generator().manual_seed(getattr(opt, "init_seed", 0)))
        self.loss_names = ["G_Recon_comb", "G_Recon_obj", "KL_loss", "loss_G_GAN", "loss_D_GAN", "loss_G_GAN_Feat"]
        # slots: 0 NLL sum, 1 box pixels, 2 BCE sum | --use_gan: 3 loss_G_GAN, 4 loss_D_real, 5 loss_D_fake, 6 feature matching
        self.acc = torch.zeros(8, dtype=torch.float64, device=dev)
        self.rec_weight = float(getattr(opt, "rec_weight", 1.0))
        self.old_lr = getattr(opt, "lr", 0.0002)
        from .models import FusedAdam
        self.optimizer = FusedAdam(self.ctx, self.fpG, self.old_lr, (getattr(opt, "beta1", 0.9), getattr(opt, "beta2", 0.999)))
        # data parallel like the reference's nn.DataParallel (models/models.py:21-22): per-replica BatchNorm statistics,
        # replica losses averaged (train_box2mask.py), i.e. the mean of the shard gradients -- one allreduce of the flat
        # gradient buffer inside optimizer.step()
        self.optimizer.data_parallel = bool(getattr(opt, "data_parallel", True))
        self._graph = None
        if self.use_gan:                                                   # :64-104
            cond_nc = opt.label_nc * 2 if opt.cond_in == "ctx_obj" else opt.label_nc
            if opt.cond_in != "ctx_obj":
                raise NotImplementedError("--use_gan: the discriminator input kernel is built for cond_in == 'ctx_obj'")
            self.gan_weight = float(getattr(opt, "gan_weight", 1.0))
            self.num_layers_D = int(getattr(opt, "num_layers_D", 4))
            self.fpD = FlatParams(dev)
            self.netD = BNMultiscaleDiscriminator(self.ctx, self.fpD, 1 + cond_nc, getattr(opt, "ndf", 64),
                                                  self.num_layers_D, 2, norm_layer=getattr(opt, "norm_layer", "batch"))
            self.fpD.materialize()
            self.netD.init_reference(torch.Generator().manual_seed(getattr(opt, "init_seed", 0) + 1))
            self.optimizer_D = FusedAdam(self.ctx, self.fpD, self.old_lr, (getattr(opt, "beta1", 0.9), 0.999))
            self.optimizer_D.data_parallel = self.optimizer.data_parallel
        # :106-118: resume / fine-tune / test-time loading
        if self.isTrain:
            if getattr(opt, "continue_train", False) or getattr(opt, "load_pretrain", ""):
                self.load(getattr(opt, "which_epoch", "latest"), getattr(opt, "load_pretrain", ""))
        elif os.path.isfile(os.path.join(self.save_dir, "%s_net_G.pth" % getattr(opt, "which_epoch", "latest"))):
            self.load(getattr(opt, "which_epoch", "latest"))
        else:   # the reference asserts here; benchmarks and parity tests build test-mode models on initialised weights
            print("%s not exists yet!" % os.path.join(self.save_dir, "%s_net_G.pth" % getattr(opt, "which_epoch", "latest")))
        # dict(graph, inputs, outputs) once the training iteration has been captured; False: stay eager -- lr_control
        # (Discriminator_NET.py:190-211) reads three losses on the host every iteration
        self._graph = False if (self.use_gan and getattr(opt, "lr_control", False)) else None
        self._eager_steps = 0

    def _dev(self, t):
        return t.to(self.device, torch.float32).contiguous()

    def forward(self, label_map, mask_obj_in, mask_ctx_in, mask_obj_out, mask_out, mask_obj_inst, cls, mask_in,
                eval_mode=False, train=True):
        if eval_mode:
            # the reference's forward(eval_mode=True) ends in `return recon_label`, an undefined name (:254-255): inference
            # goes through reconstruct(..., eval_mode=True) / generate(), below
            raise NameError("name 'recon_label' is not defined (TwoStreamAE_mask.py:255); use generate() / reconstruct()")
        ins = dict(label_map=label_map, mask_ctx_in=mask_ctx_in, mask_out=mask_out, mask_in=mask_in,
                   mask_obj_inst=mask_obj_inst, cls=cls.reshape(-1))
        if train and getattr(self.opt, "cuda_graph", True) and self._graph is not False:
            # like Pix2PixHDModel_condImg.optimize_parameters: two eager iterations, then the whole iteration (~700 launches)
            # is captured in a CUDA graph and replayed; a new batch geometry re-captures
            sig = tuple(tuple(v.shape) for v in ins.values())
            if self._graph is not None and self._graph["sig"] != sig:
                self._graph, self._eager_steps = None, 2
            if self._graph is not None or self._eager_steps >= 2:
                return self._graph_iteration(ins, sig)
            self._eager_steps += 1
        return self._iteration(ins, train, captured=False)

    def _graph_iteration(self, ins, sig):
        if self._graph is None:
            try:
                static = {k: torch.empty(tuple(v.shape), dtype=torch.float32, device=self.device) for k, v in ins.items()}
                for k, dst in static.items():
                    dst.copy_(ins[k], non_blocking=True)
                self.optimizer.step_dev.fill_(self.optimizer.step_count)
                torch.cuda.synchronize()
                l0, ver = self.ctx.launches, self.fpG.version
                ver_d = self.fpD.version if self.use_gan else 0
                if self.use_gan:
                    self.optimizer_D.step_dev.fill_(self.optimizer_D.step_count)
                    torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                    outs = self._iteration(static, True, captured=True)
                launches = self.ctx.launches - l0
                self.ctx.launches, self.fpG.version = l0, ver
                if self.use_gan:
                    self.fpD.version = ver_d
                self._graph = dict(graph=graph, inputs=static, outputs=outs, sig=sig, launches=launches)
            except Exception as e:  # noqa: BLE001 -- any capture problem: keep training eagerly
                print("CUDA-graph capture of the box2mask iteration failed (%s: %s); staying eager" % (type(e).__name__, e))
                self._graph = False
                torch.cuda.synchronize()
                return self._iteration(ins, True, captured=False)
        g = self._graph
        for k, dst in g["inputs"].items():
            dst.copy_(ins[k], non_blocking=True)
        g["graph"].replay()
        self.optimizer.step_count += 1
        self.fpG.version += 1
        if self.use_gan:
            self.optimizer_D.step_count += 1
            self.fpD.version += 1
        self.ctx.launches += g["launches"]
        return g["outputs"]

    def _iteration(self, ins, train, captured):
        opt, ctx = self.opt, self.ctx
        label_map, mask_ctx_in, mask_out, mask_in, inst, clsf = (self._dev(ins[k]) for k in (
            "label_map", "mask_ctx_in", "mask_out", "mask_in", "mask_obj_inst", "cls"))
        B, _, H, W = label_map.shape
        cond = ops.box2mask_encode(ctx, mask_ctx_in, mask_in, clsf, opt.label_nc)          # :127-151, :331-338
        ctx_logit, obj_logit, tape = self.netG.forward(cond)
        C = opt.output_nc
        gate = bool(getattr(opt, "use_output_gate", False))
        out = dict(comb_logit=torch.empty(B, C, H, W, device=self.device), comb_prob=torch.empty(B, C, H, W, device=self.device),
                   obj_logit=obj_logit[..., :1].permute(0, 3, 1, 2), obj_prob=torch.empty(B, 1, H, W, device=self.device))
        self.acc.zero_()
        ops.box2mask_head(ctx, ctx_logit, obj_logit, label_map, mask_out, inst, gate, out["comb_logit"], out["comb_prob"],
                          out["obj_prob"], self.acc, no_comb=self.no_comb, l1=(self.obj_loss == "l1"))
        loss_comb = (self.acc[0] / self.acc[1].clamp_min(1.0)).float()       # NLLLoss2d mean over non-ignored pixels
        loss_obj = (self.acc[2] / float(B * H * W)).float()                  # BCELoss / L1Loss mean
        if self.obj_loss not in ("bce", "l1"):
            loss_obj = loss_obj * 0.0                                        # criterionObjRecon is None: loss_recon_obj = 0
        self._last = dict(tape=tape, ctx_logit=ctx_logit, obj_logit=obj_logit, label_map=label_map, mask_out=mask_out,
                          inst=inst, gate=gate)
        zero = torch.zeros((), device=self.device)
        gan = [zero, zero, zero]
        if self.use_gan:
            gan = self._discriminate(out["obj_prob"], inst, mask_ctx_in, mask_in, clsf, mask_out if gate else None)
        if not train:
            return [loss_comb, loss_obj] + (gan if self.use_gan else []), out
        g_lr = d_lr = 1.0
        if self.use_gan and getattr(opt, "lr_control", False):                # :237-239 (host decision, eager only)
            g_lr, d_lr = lr_control(*[float(v) for v in self.acc[3:6].tolist()])
        # ---- :233-248: loss_G = loss_recon_obj + rec_weight * loss_recon_comb + gan_weight * loss_G_GAN; zero_grad,
        # backward, step.  g_lr == 0 zeroes the gradient but Adam still steps on its moments, like the reference's 0 * loss.
        self.optimizer.zero_grad()
        if g_lr != 0.0:
            self.backward_losses()
        self.optimizer.step(captured=captured)
        if self.use_gan:                                                      # :243-248: loss_D = d_lr * 0.5 * (real + fake)
            self.optimizer_D.zero_grad()
            if d_lr != 0.0:
                self.netD.backward(self._last["d_real"], 1.0, 0.5, True)
                self.netD.backward(self._last["d_fake"], 0.0, 0.5, True)
            self.optimizer_D.step(captured=captured)
        # :271-275 postprocess_output + argmax (host-side visual output)
        gt_onehot = torch.zeros_like(out["comb_prob"]).scatter_(1, label_map.long(), 1.0)
        comb_label = (out["comb_prob"] * mask_out + (1 - mask_out) * gt_onehot).argmax(dim=1, keepdim=True)
        return [loss_comb, loss_obj, zero, gan[0], gan[1], gan[2]], [comb_label, out["obj_prob"]]

    def _discriminate(self, obj_prob, inst, mask_ctx_in, mask_in, clsf, gate_mask):
        """:203-232.  The three discriminator evaluations of the reference are two here: `discriminate(fake.detach())`
        and `discriminate(fake)` see the same values and, in training mode, the same batch statistics.  Returns
        [loss_G_GAN, loss_D, loss_G_GAN_Feat]; the feature-matching term is reported only -- the reference computes it on
        the detached fake pass and never adds it to loss_G (:221-235)."""
        ctx, opt = self.ctx, self.opt
        # fake: the generated mask is gated once for the BCE loss and once more here ("masking twice", :209-212)
        d_fake = ops.box2mask_d_input(ctx, obj_prob, mask_ctx_in, mask_in, clsf, gate_mask, 2, opt.label_nc)
        d_real = ops.box2mask_d_input(ctx, inst, mask_ctx_in, mask_in, clsf, gate_mask, 1, opt.label_nc)
        t_real, t_fake = self.netD.forward(d_real), self.netD.forward(d_fake, repeat=2)
        for lr_, lf in zip(t_real, t_fake):
            pr, pf = lr_["taps"][-1], lf["taps"][-1]
            ops.mse_sum(ctx, pf, 1.0, 1.0 / pf.numel(), self.acc, 3)
            ops.mse_sum(ctx, pr, 1.0, 1.0 / pr.numel(), self.acc, 4)
            ops.mse_sum(ctx, pf, 0.0, 1.0 / pf.numel(), self.acc, 5)
            if getattr(opt, "use_ganFeat_loss", False):
                cf = 0.5 * (4.0 / (self.num_layers_D + 1)) * float(getattr(opt, "lambda_feat", 1.0))
                for a, b in zip(lf["taps"][:-1], lr_["taps"][:-1]):
                    ops.l1_sum(ctx, a, b, cf / a.numel(), self.acc, 6)
        self._last.update(d_real=t_real, d_fake=t_fake)
        return [self.acc[3].float(), (0.5 * self.acc[4] + 0.5 * self.acc[5]).float(), self.acc[6].float()]

    def backward_losses(self):
        """d(loss_recon_obj + rec_weight * loss_recon_comb)/d(parameters), accumulated into the flat .grad buffer."""
        s, ctx = self._last, self.ctx
        N, H, W, C = s["ctx_logit"].shape
        d_ctx = Operand(ctx, N, H, W, C, grad=True)
        d_obj = Operand(ctx, N, H, W, 1, grad=True)
        g_prob = None
        if self.use_gan:    # gan_weight * loss_G_GAN through the discriminator's data path down to the generated mask
            g_prob = self.netD.backward(s["d_fake"], 1.0, self.gan_weight, False)
        ops.box2mask_head_bwd(ctx, s["ctx_logit"], s["obj_logit"], s["label_map"], s["mask_out"], s["inst"], s["gate"],
                              self.acc, self.rec_weight, 1.0 if self.obj_loss in ("bce", "l1") else 0.0, d_ctx, d_obj,
                              g_prob=g_prob, no_comb=self.no_comb, l1=(self.obj_loss == "l1"))
        self.netG.backward(s["tape"], d_ctx, d_obj)

    # ---- inference (vis_box2mask.py:36-60, train_box2mask.py:90-100) -------------------------------------------------
    def reconstruct(self, input_dict, eval_mode=False):
        """:257-297.  Forward pass only; eval_mode=True normalises with the BatchNorm running statistics and restores
        the network's mode afterwards.  Returns comb_recon_label [B,1,H,W] (argmax of the generated layout inside the box,
        ground truth outside, :271-275) and obj_recon_label (the object stream's probability map)."""
        ctx, opt = self.ctx, self.opt
        current = self.netG.get_mode()
        if eval_mode != current:
            self.netG.set_mode(eval_mode)
        try:
            label_map, mask_ctx_in, mask_out, mask_in = (self._dev(input_dict[k]) for k in (
                "label_map", "mask_ctx_in", "mask_out", "mask_in"))
            clsf = self._dev(input_dict["cls"].reshape(-1))
            B, _, H, W = label_map.shape
            cond = ops.box2mask_encode(ctx, mask_ctx_in, mask_in, clsf, opt.label_nc)
            ctx_logit, obj_logit, _ = self.netG.forward(cond)
            C = opt.output_nc
            out = dict(comb_recon_prob=torch.empty(B, C, H, W, device=self.device),
                       obj_recon_prob=torch.empty(B, 1, H, W, device=self.device))
            ops.box2mask_head(ctx, ctx_logit, obj_logit, None, None, None, False, None, out["comb_recon_prob"],
                              out["obj_recon_prob"], None, no_comb=self.no_comb)
        finally:
            if eval_mode != current:
                self.netG.set_mode(current)
        gt_onehot = torch.zeros_like(out["comb_recon_prob"]).scatter_(1, label_map.long(), 1.0)
        comb_label = (out["comb_recon_prob"] * mask_out + (1 - mask_out) * gt_onehot).argmax(dim=1, keepdim=True)
        res = dict(comb_recon_label=comb_label, obj_recon_label=out["obj_recon_prob"])
        if not eval_mode:
            res.update(label_map=label_map, comb_gt_mask=mask_out, comb_recon_prob=out["comb_recon_prob"],
                       obj_recon_prob=out["obj_recon_prob"])
        return res

    def generate(self, input_dict):
        """:299-302."""
        out = self.reconstruct(input_dict, eval_mode=True)
        return dict(comb_pred_label=out["comb_recon_label"], obj_pred_label=out["obj_recon_label"])

    def evaluate(self, input_dict, target_size=None):
        """:304-335, the joint-inference entry point (joint_inference_model.py:61-81): first sample only, eval-mode
        BatchNorm (left ON, like the reference); the generated maps are resized to `target_size` (bilinear) and pasted
        into the original-resolution maps; background class (label_nc - 1): arg-max layout inside the box, any other
        class: the thresholded object mask painted with that class.  Returns comb_recon_label [1,1,H,W] (float)."""
        ctx, opt = self.ctx, self.opt
        first = lambda k: self._dev(input_dict[k][0].unsqueeze(0))            # noqa: E731
        label_map, mask_ctx_in, mask_out, mask_in = first("label_map"), first("mask_ctx_in"), first("mask_out"), first("mask_in")
        cls = input_dict["cls"][0].reshape(-1)
        self.netG.set_mode(eval_mode=True)
        cond = ops.box2mask_encode(ctx, mask_ctx_in, mask_in, self._dev(cls), opt.label_nc)
        ctx_logit, obj_logit, _ = self.netG.forward(cond)
        _, _, H, W = label_map.shape
        comb_prob = torch.empty(1, opt.output_nc, H, W, device=self.device)
        obj_prob = torch.empty(1, 1, H, W, device=self.device)
        ops.box2mask_head(ctx, ctx_logit, obj_logit, None, None, None, False, None, comb_prob, obj_prob, None,
                          no_comb=self.no_comb)
        if getattr(opt, "use_output_gate", False):
            obj_prob = obj_prob * mask_out
        if target_size is not None:
            label_map, mask_out = first("label_map_orig"), first("mask_out_orig")
            us = lambda t: torch.nn.functional.interpolate(t, size=target_size, mode="bilinear", align_corners=False)  # noqa: E731
            comb_prob, obj_prob = us(comb_prob), us(obj_prob)
        c = int(cls[0])
        if c == opt.label_nc - 1:
            gt_onehot = torch.zeros_like(comb_prob).scatter_(1, label_map.long(), 1.0)
            return (comb_prob * mask_out + (1 - mask_out) * gt_onehot).argmax(dim=1, keepdim=True).float()
        obj_mask = (obj_prob > 0.5).float()
        return (1 - obj_mask) * label_map + obj_mask * c

    # ---- checkpoints (base_model.py:43-64,73-127; TwoStreamAE_mask.py:106-118,359-369) --------------------------------
    def _network_dict(self):
        """The generator's state as the reference's save_network_dict writes it: {params_dict key: module.state_dict()}
        -- a parameter '<key>.<rest>' belongs to module <key>; the parameter-free nn.ReLU at conv_encoder_2 is an empty
        entry (load_network_dict indexes every params_dict key)."""
        net = OrderedDict()
        for k, v in self.fpG.state_dict().items():
            mk, rest = k.split(".", 1)
            net.setdefault(mk, OrderedDict())[rest] = v
        net.setdefault("conv_encoder_1", OrderedDict())       # InstanceNorm2d(affine=False): an empty state dict
        net.setdefault("conv_encoder_2", OrderedDict())
        return net

    def _adam_state(self, optimizer):
        """torch.optim.Adam.state_dict() layout, parameters in this implementation's declaration order (the reference's
        order is the iteration order of a python-2 dict, which no file format pins down)."""
        fp = optimizer.fp
        state, off = {}, {n: o for n, _, o in fp.specs}
        for i, (name, p) in enumerate(fp.params.items()):
            o, n = off[name], p.numel()
            state[i] = dict(step=optimizer.step_count, exp_avg=optimizer.m[o:o + n].view(p.shape).cpu().clone(),
                            exp_avg_sq=optimizer.v[o:o + n].view(p.shape).cpu().clone())
        g = optimizer.param_groups[0]
        return dict(state=state, param_groups=[dict(lr=g["lr"], betas=tuple(optimizer.betas), eps=optimizer.eps, weight_decay=0,
                                                    amsgrad=False, params=list(range(len(fp.params))))])

    def _load_adam_state(self, optimizer, sd):
        fp = optimizer.fp
        off = {n: o for n, _, o in fp.specs}
        for i, (name, p) in enumerate(fp.params.items()):
            st = sd["state"].get(i)
            if st is None:
                continue
            o, n = off[name], p.numel()
            optimizer.m[o:o + n].copy_(st["exp_avg"].reshape(-1))
            optimizer.v[o:o + n].copy_(st["exp_avg_sq"].reshape(-1))
            optimizer.step_count = int(st["step"])
        for g, gs in zip(optimizer.param_groups, sd["param_groups"]):
            g["lr"] = gs["lr"]
        self._graph = None if self._graph is not False else False       # the captured step count / lr are stale

    def save(self, which_epoch):
        """:359-363: '<epoch>_net_G.pth' = {'network': {...}, 'optimizer': Adam state}; the discriminator (with --use_gan)
        as a plain state dict in '<epoch>_net_D.pth'."""
        os.makedirs(self.save_dir, exist_ok=True)
        torch.save(dict(network=self._network_dict(), optimizer=self._adam_state(self.optimizer)),
                   os.path.join(self.save_dir, "%s_net_G.pth" % which_epoch))
        if self.use_gan:
            torch.save(self.fpD.state_dict(), os.path.join(self.save_dir, "%s_net_D.pth" % which_epoch))

    def load(self, which_epoch, save_dir=""):
        """:106-118 (load_network_dict for the generator + its optimizer, load_network for the discriminator)."""
        path = os.path.join(save_dir or self.save_dir, "%s_net_G.pth" % which_epoch)
        if not os.path.isfile(path):
            print("%s not exists yet!" % path)
            raise AssertionError("Generator must exist!")
        ck = torch.load(path, map_location="cpu")
        flat = OrderedDict()
        for mk, sd in ck["network"].items():
            for k, v in sd.items():
                flat[mk + "." + k] = v
        self.fpG.load_state_dict(flat)
        if self.isTrain and ck.get("optimizer") is not None and hasattr(self, "optimizer"):
            self._load_adam_state(self.optimizer, ck["optimizer"])
        if self.use_gan:
            pd = os.path.join(save_dir or self.save_dir, "%s_net_D.pth" % which_epoch)
            if os.path.isfile(pd):
                self.fpD.load_state_dict(torch.load(pd, map_location="cpu"))
            else:
                print("%s not exists yet!" % pd)

    def delete_model(self, which_epoch):
        """:365-368."""
        for lbl in ("G", "D") if self.use_gan else ("G",):
            p = os.path.join(self.save_dir, "%s_net_%s.pth" % (which_epoch, lbl))
            if os.path.isfile(p):
                os.remove(p)

    def update_learning_rate(self, epoch=0, data_size=0):
        """:370-383: after opt.niter epochs the rate drops by lr / niter_decay per call, for both optimizers."""
        if epoch > self.opt.niter:
            lr = self.old_lr - self.opt.lr / self.opt.niter_decay
            for o in [self.optimizer] + ([self.optimizer_D] if self.use_gan else []):
                for g in o.param_groups:
                    g["lr"] = lr
            print("update learning rate: %f -> %f" % (self.old_lr, lr))
            self.old_lr = lr
            if self._graph is not False:
                self._graph = None         # the learning rate is baked into the captured Adam launches: re-capture

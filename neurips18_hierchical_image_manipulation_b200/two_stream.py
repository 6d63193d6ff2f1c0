"""GlobalTwoStreamGenerator executor (models/Pix2Pix_NET.py:103-247) -- the generator the reference's shipped scripts
train (scripts/train_mask2image_city.sh: --netG global_twostream --which_encoder ctx_label --use_skip --use_output_gate).

    ctx stream : ReflPad3, Conv7(3 -> ngf), IN, ReLU, n x [Conv3 s2, IN, ReLU]        on cond_image = (1-mask)*image
    obj stream : the same on the one-hot label map (+ instance edges)
    fusion     : (1 - m) * ctx_feat + m * obj_feat, m = MaxPool(2^n)(mask)              ('early_add', :207-209)
    embedder   : n_blocks ResnetBlocks; decoder: n x [ConvT3 s2, IN, ReLU] with the ctx stream's intermediate features
                 concatenated in front of the decoder input from the second stage on (use_skip, :215-223); ReflPad3,
                 Conv7(ngf -> 3), Tanh; output gate applied by the caller (hm_finish_fake).

Parameter names follow the reference module tree (ctx_inputEmbedder.1, ctx_downsampler.{0,3,..}, obj_*,
latent_embedder.{i}.conv_block.{1,5}, decoder.{0,3,..}, outputEmbedder.1) so its checkpoints interchange.
Supported: which_stream in {ctx_label, ctx, label}, feat_fusion in {early_add, late_add, early_concat, late_concat}.  Every stage reuses the GlobalGenerator
schedules (networks.py); the glue kernels are csrc/hm_twostream.cu.
"""
from . import ops
from .networks import ConvP, GlobalGenerator, _f32
from .ops import Operand


class _Stages(GlobalGenerator):
    """A GlobalGenerator schedule over an explicit stage list."""

    def __init__(self, ctx, fp, stages, with_head, ngf):
        self.ctx, self.fp = ctx, fp
        self.stages = stages
        self.with_head = with_head
        self.feature_nc = ngf
        self.use_output_gate = False


class GlobalTwoStreamGenerator(object):
    def __init__(self, ctx, fp, input_nc, output_nc, ngf=64, n_downsampling=3, n_blocks=9, use_skip=False,
                 which_stream="ctx", use_output_gate=False, feat_fusion="early_add"):
        if which_stream not in ("ctx_label", "ctx", "label"):
            raise NotImplementedError("which_encoder must be ctx_label | ctx | label, got %s" % which_stream)
        if feat_fusion not in ("early_add", "late_add", "early_concat", "late_concat"):
            raise NotImplementedError("feat_fusion %s" % feat_fusion)
        # Pix2Pix_NET.py:107-108: the late fusions need both streams
        assert not ("late" in feat_fusion and which_stream != "ctx_label")
        self.late = "late" in feat_fusion
        self.concat_fuse = "concat" in feat_fusion and which_stream == "ctx_label"
        if n_blocks < 1:
            raise NotImplementedError("n_blocks_global == 0 is not part of this path")
        self.ctx, self.fp = ctx, fp
        self.input_nc, self.output_nc, self.ngf, self.n_down, self.n_blocks = input_nc, output_nc, ngf, n_downsampling, n_blocks
        self.use_skip = bool(use_skip) and "ctx" in which_stream and n_downsampling > 1
        self.which_stream, self.use_output_gate = which_stream, use_output_gate
        self.feat_dim = ngf * 2 ** n_downsampling

        def encoder(prefix, cin):
            st = [("stem", ConvP(ctx, fp, prefix + "_inputEmbedder.1", cin, ngf, 7, 1, 0))]
            for i in range(n_downsampling):
                m = 2 ** i
                st.append(("down", ConvP(ctx, fp, prefix + "_downsampler.%d" % (3 * i), ngf * m, ngf * m * 2, 3, 2, 1)))
            return _Stages(ctx, fp, st, False, ngf)

        # declaration order = the reference's module registration order (state_dict order)
        self.enc_ctx = encoder("ctx", 3) if "ctx" in which_stream else None
        self.enc_obj = encoder("obj", input_nc) if "label" in which_stream else None
        # FeatureFusionBlock 'concat' (layer_util.py:305-327): cat -> ReLU -> Conv2d(2C, C, 1) -> InstanceNorm
        self.fuse_conv = ConvP(ctx, fp, "feat_fuser.conv1", 2 * self.feat_dim, self.feat_dim, 1, 1, 0) if self.concat_fuse else None

        def res_blocks(prefix, n):
            out = []
            for i in range(n):
                out.append(("resA", ConvP(ctx, fp, "%s.%d.conv_block.1" % (prefix, i), self.feat_dim, self.feat_dim, 3, 1, 0)))
                out.append(("resB", ConvP(ctx, fp, "%s.%d.conv_block.5" % (prefix, i), self.feat_dim, self.feat_dim, 3, 1, 0)))
            return out
        # feat_fusion 'late_*' (:137-142): floor(n/2) blocks per stream BEFORE the masked fusion, ceil(n/2) after it
        self.lat_obj = self.lat_ctx = None
        n_comb = n_blocks
        if self.late:
            n_comb = (n_blocks + 1) // 2
            if n_blocks // 2 > 0:
                self.lat_obj = _Stages(ctx, fp, res_blocks("obj_latent_embedder", n_blocks // 2), False, ngf)
                self.lat_ctx = _Stages(ctx, fp, res_blocks("ctx_latent_embedder", n_blocks // 2), False, ngf)
        st = res_blocks("latent_embedder", n_comb)
        self.first_up = len(st)
        for i in range(n_downsampling):
            m = 2 ** (n_downsampling - i)
            cin = ngf * m * (2 if (self.use_skip and i > 0) else 1)
            st.append(("up", ConvP(ctx, fp, "decoder.%d" % (3 * i), cin, ngf * m // 2, 3, 2, 1, transposed=True)))
        st.append(("head", ConvP(ctx, fp, "outputEmbedder.1", ngf, output_nc, 7, 1, 0)))
        self.trunk = _Stages(ctx, fp, st, True, ngf)

    def convs(self):
        out = []
        for part in (self.enc_ctx, self.enc_obj, self.lat_obj, self.lat_ctx, self.trunk):
            if part is not None:
                out += part.convs()
        if self.fuse_conv is not None:
            out.append(self.fuse_conv)
        return out

    # ------------------------------------------------------------------------------------------------
    def forward(self, ctx_in, obj_in, mask_nchw):
        """ctx_in: cond-image Operand (3 channels), obj_in: label Operand (input_nc channels), both with
        ReflectionPad2d(3) materialised; mask_nchw: fp32 [B,1,H,W] on the device.  Returns (tanh output fp32 NHWC, tape)."""
        c = self.ctx
        tape = {}
        fa = fb = None
        if self.enc_ctx is not None:
            fa, tape["ctx"] = self.enc_ctx.forward(ctx_in)
        if self.enc_obj is not None:
            fb, tape["obj"] = self.enc_obj.forward(obj_in)
        if self.lat_ctx is not None:                                                     # :204-206 ('late' fusion)
            fa, tape["lat_ctx"] = self._embed(self.lat_ctx, fa)
            fb, tape["lat_obj"] = self._embed(self.lat_obj, fb)
        ref = fa if fa is not None else fb
        N, h, w, C = ref.shape
        m = None
        if fa is not None and fb is not None:
            m = ops.mask_maxpool(c, mask_nchw, 2 ** self.n_down)                        # :203-205
        first_border = GlobalGenerator.IN_BORDER[self.trunk.stages[0][0]]
        comb32 = _f32(c, N, h, w, C)
        comb_op = Operand(c, N, h, w, C, border=first_border)
        if self.fuse_conv is not None:                                                   # FeatureFusionBlock 'concat'
            u = ops.mask_concat(c, fa, fb, m)
            yf = _f32(c, N, h, w, C)
            self.fuse_conv.forward(u, 0, out32=yf)
            mean, rstd = ops.in_stats(c, yf)
            ops.in_apply(c, yf, mean, rstd, ops.ACT_NONE, out32=comb32, out_op=comb_op, reflect=True)
            tape["fuse"] = dict(u=u, y=yf, mean=mean, rstd=rstd, fa=fa, fb=fb)
        else:
            ops.mask_blend(c, fa, fb, m, out32=comb32, out_op=comb_op)                   # :207-209
        tape["m"] = m
        concat = None
        if self.use_skip:
            # ctx features after down stage k (k <= n-2) = the input operand of down stage k+1; decoder stage s >= 1
            # gets cat((feat after down_{n-1-s}, dec), 1)                                    (:139-141, :218-222)
            concat = {}
            for s in range(1, self.n_down):
                k = self.n_down - 1 - s
                concat[self.first_up + s] = tape["ctx"][k + 2]["xin"]
        out, tape["trunk"] = self.trunk.forward(comb_op, skip32_init=comb32, concat=concat)
        return out, tape

    def _embed(self, stages, f32):
        """ResnetBlocks on a dense fp32 feature: materialise its reflect-padded operand, run the block list."""
        c = self.ctx
        N, h, w, C = f32.shape
        op = Operand(c, N, h, w, C, border=1)
        ops.in_apply(c, f32, None, None, ops.ACT_NONE, out_op=op, reflect=True)
        return stages.forward(op, skip32_init=f32)

    def backward(self, tape, dy_head):
        """dy_head: Operand gradient w.r.t. the head's pre-tanh output.  Accumulates every parameter gradient."""
        c = self.ctx
        self.trunk.backward(tape["trunk"], dy_head=dy_head, need_input_grad=True)
        g = self.trunk.input_T                              # dense gradient w.r.t. the fused feature
        extra = None
        if self.use_skip:
            extra = {}
            for s in range(1, self.n_down):
                k = self.n_down - 1 - s
                extra[k + 1] = self.trunk.concat_grads[self.first_up + s]   # gradient w.r.t. the output of ctx stage k+1
        m = tape["m"]
        if self.fuse_conv is not None:
            t = tape["fuse"]
            N, h, w, C = g.shape
            dy = Operand(c, N, h, w, C, grad=True)
            ops.in_bwd(c, (N, h, w, C), ops.ACT_NONE, y=t["y"], mean=t["mean"], rstd=t["rstd"], g2=g, out_op=dy)
            self.fuse_conv.wgrad(t["u"], dy, 0, bias_grad=False)          # the bias cancels in the InstanceNorm
            gu = _f32(c, N, h, w, 2 * C)
            self.fuse_conv.dgrad(dy, h, w, 0, gu)
            da, db = _f32(c, N, h, w, C), _f32(c, N, h, w, C)
            ops.mask_concat_bwd(c, gu, m, t["fa"], t["fb"], da, db)
        elif m is not None:
            da, db = _f32(c, *g.shape), _f32(c, *g.shape)
            ops.mask_blend_bwd(c, g, m, da, db)
        else:
            da = db = g
        if self.lat_ctx is not None:          # through the per-stream ResnetBlocks (their first stage is a block: input_T)
            self.lat_ctx.backward(tape["lat_ctx"], dfeat=da, need_input_grad=True)
            self.lat_obj.backward(tape["lat_obj"], dfeat=db, need_input_grad=True)
            da, db = self.lat_ctx.input_T, self.lat_obj.input_T
        if self.enc_ctx is not None:
            self.enc_ctx.backward(tape["ctx"], dfeat=da, extra_grad=extra)
        if self.enc_obj is not None:
            self.enc_obj.backward(tape["obj"], dfeat=db)

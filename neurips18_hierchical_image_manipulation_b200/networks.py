"""B200 executors for the networks on the mask2image hot path.

Each class mirrors one reference nn.Module -- same constructor arguments, same parameter names and OIHW/IOHW
fp32 parameter tensors (so reference checkpoints interchange) -- but forward/backward are explicit schedules of
C-ABI kernels (ops.py): tcgen05 implicit-GEMM convolutions on bf16(x3) NHWC operands + fused InstanceNorm /
activation / padding passes.  Nothing here calls torch.nn or autograd.

  GlobalGenerator          <- models/Pix2Pix_NET.py:63-101 (+ ResnetBlock, models/layer_util.py:333-378)
  MultiscaleDiscriminator  <- models/Discriminator_NET.py:11-118
  Vgg19                    <- models/layer_util.py:381-411
"""
import os
from collections import OrderedDict

import torch

from . import _lib as L
from . import ops
from .ops import ACT_LRELU, ACT_NONE, ACT_RELU, ACT_TANH, Operand, PackedWeight

VGG19_CONVS = [(0, 3, 64), (2, 64, 64), (5, 64, 128), (7, 128, 128), (10, 128, 256), (12, 256, 256), (14, 256, 256),
               (16, 256, 256), (19, 256, 512), (21, 512, 512), (23, 512, 512), (25, 512, 512), (28, 512, 512)]
VGG19_POOL_BEFORE = (5, 10, 19, 28)
VGG19_TAP_AFTER = {0: 0, 5: 1, 10: 2, 19: 3, 28: 4}
VGG19_SLICE_OF = {0: 1, 2: 2, 5: 2, 7: 3, 10: 3, 12: 4, 14: 4, 16: 4, 19: 4, 21: 5, 23: 5, 25: 5, 28: 5}

# HM_THIN=0 disables the tap-unrolled lowering of <= 4-channel convolutions (csrc/hm_thin.cu) for A/B measurements
THIN = os.environ.get("HM_THIN", "1") != "0"
# HM_PACK_PAIR=0: pack the forward and data-gradient weight slabs with two separate launches (round-1 behaviour)
PACK_PAIR = os.environ.get("HM_PACK_PAIR", "1") != "0"


class FlatParams(object):
    """All parameters of one network as views into ONE flat fp32 buffer (and one flat gradient buffer), so the
    optimizer is a single fused kernel and data parallelism is a single allreduce."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.specs = []  # (name, shape, offset)
        self.total = 0
        self.flat = None
        self.grad = None
        self.params = OrderedDict()  # name -> torch.nn.Parameter (view of flat, .grad = view of grad)
        self.buffers = OrderedDict()  # name -> non-trainable state outside the flat buffer (spectral-norm `u` vectors)
        self.version = 0  # bumped whenever values change -> packed weights are refreshed lazily

    def declare(self, name, shape):
        n = 1
        for s in shape:
            n *= s
        self.specs.append((name, tuple(shape), self.total))
        self.total += (n + 3) // 4 * 4  # keep every tensor 16 B aligned
        return name

    def materialize(self, flat=None, grad=None):
        self.flat = torch.zeros(self.total, dtype=torch.float32, device=self.device) if flat is None else flat
        self.grad = torch.zeros(self.total, dtype=torch.float32, device=self.device) if grad is None else grad
        for name, shape, off in self.specs:
            n = 1
            for s in shape:
                n *= s
            p = torch.nn.Parameter(self.flat[off:off + n].view(shape), requires_grad=True)
            p.grad = self.grad[off:off + n].view(shape)
            self.params[name] = p

    def state_dict(self):
        sd = OrderedDict((k, v.detach().cpu().clone()) for k, v in self.params.items())
        for k, v in self.buffers.items():
            sd[k] = v.detach().cpu().clone()
        return sd

    def load_state_dict(self, sd, strict=True):
        """Nothing is copied unless every key that will be copied has the right shape (a failed load must not leave a
        half-overwritten network behind); strict additionally requires every parameter to be present."""
        missing = [k for k in self.params if k not in sd]
        if strict and missing:
            raise KeyError("missing keys: %s" % missing)
        bad = [k for k, p in self.params.items() if k in sd and tuple(sd[k].shape) != tuple(p.shape)]
        bad += [k for k, b in self.buffers.items() if k in sd and sd[k].numel() != b.numel()]
        if bad:
            raise KeyError("size mismatch for: %s" % bad)
        with torch.no_grad():
            for k, p in self.params.items():
                if k in sd:
                    p.copy_(sd[k].to(self.device, torch.float32))
            for k, b in self.buffers.items():
                if k in sd:
                    b.copy_(sd[k].to(self.device).view(b.shape))
        self.version += 1


class ConvP(object):
    """One Conv2d / ConvTranspose2d: parameter views + lazily packed bf16 weight slabs for the three engine roles."""

    def __init__(self, ctx, fp, name, cin, cout, k, stride, pad, transposed=False, dilation=1, bias=True):
        self.ctx, self.fp, self.name = ctx, fp, name
        self.cin, self.cout, self.k, self.stride, self.pad, self.transposed = cin, cout, k, stride, pad, transposed
        # dilation > 1: nn.Conv2d(..., dilation=d) of DilatedResnetBlock (layer_util.py:255-293); bias=False: its conv3x3
        self.dilation, self.has_bias = int(dilation), bool(bias)
        assert self.dilation == 1 or (stride == 1 and not transposed)
        shape = (cin, cout, k, k) if transposed else (cout, cin, k, k)
        fp.declare(name + ".weight", shape)
        if self.has_bias:
            fp.declare(name + ".bias", (cout,))
        self._pf = self._pd = None
        self._vf = self._vd = -1
        # thin sides (csrc/hm_thin.cu): generator head / PatchGAN output (cout <= 4), VGG conv1_1 (cin <= 4)
        plain = THIN and not transposed and stride == 1 and k >= 2 and self.dilation == 1
        self.thin_out = plain and cout <= 4
        self.thin_in = plain and not self.thin_out and cin <= 4 and k * k * cin <= 64
        self._u = None  # (dy, unrolled dy) of the last thin-output gradient call
        self.sn_scale = None  # device scalar 1/sigma when the conv is spectrally normalised (K13), else None

    # parameter / gradient views -------------------------------------------------------------------
    @property
    def weight(self):
        return self.fp.params[self.name + ".weight"]

    @property
    def bias(self):
        return self.fp.params[self.name + ".bias"] if self.has_bias else None

    def init_reference(self, gen):
        """weights_init (models/layer_util.py:9-16): N(0, 0.02) conv weights, torch-default uniform biases."""
        with torch.no_grad():
            w = self.weight
            w.copy_((torch.randn(w.shape, generator=gen) * 0.02).to(w.device))
            fan_in = w.shape[1] * self.k * self.k
            bound = 1.0 / fan_in ** 0.5
            if self.has_bias:
                self.bias.copy_(((torch.rand(self.cout, generator=gen) * 2 - 1) * bound).to(w.device))
        self.fp.version += 1

    # packed slabs ---------------------------------------------------------------------------------
    def packed_fwd(self):
        """rows = cout, contraction = cin."""
        kk = self.k * self.k
        if self.thin_out:      # rows = (kw, co), contraction = ci, taps = kh
            if self._pf is None:
                self._pf = PackedWeight(self.ctx, self.k * self.cout, self.cin, self.k)
            if self._vf != self.fp.version:
                self._pf.pack_ex(self.ctx, self.weight, self.cout, 1, self.cin * kk, 1, kk, 0, self.k, self.sn_scale)
                self._vf = self.fp.version
            return self._pf
        if self.thin_in:       # rows = co, contraction = (kh, kw, ci), one tap
            if self._pf is None:
                self._pf = PackedWeight(self.ctx, self.cout, kk * self.cin, 1)
            if self._vf != self.fp.version:
                self._pf.pack_ex(self.ctx, self.weight, 1, self.cin * kk, 0, self.cin, 1, kk, 0, self.sn_scale)
                self._vf = self.fp.version
            return self._pf
        if self._pf is None:
            self._pf = PackedWeight(self.ctx, self.cout, self.cin, self.k * self.k)
        if self._vf != self.fp.version and self._pack_pair():
            return self._pf
        if self._vf != self.fp.version:
            if self.transposed:   # W[ci][co][t]
                self._pf.pack(self.ctx, self.weight, kk, self.cout * kk, 1, self.sn_scale)
            else:                 # W[co][ci][t]
                self._pf.pack(self.ctx, self.weight, self.cin * kk, kk, 1, self.sn_scale)
            self._vf = self.fp.version
        return self._pf

    def _pack_pair(self):
        """Training steady state: both roles are stale after every Adam step and both will be used, so they are packed
        in ONE pass over the fp32 weights (hm_pack_weight_pair).  Only once the data-gradient slab exists (i.e. a backward
        pass has run), never for spectrally normalised or thin-side convs."""
        if self._pd is None or self._vd == self.fp.version or self.sn_scale is not None or not PACK_PAIR:
            return False
        pf, pd = self._pf, self._pd
        # W[A][B][taps]: Conv2d A = co, B = ci -> p1 = forward role (rows = co), p2 = data-gradient role (rows = ci);
        # ConvTranspose2d A = ci, B = co -> p1 = data-gradient role, p2 = forward role
        p1, p2 = (pd, pf) if self.transposed else (pf, pd)
        A, Bd = (self.cin, self.cout) if self.transposed else (self.cout, self.cin)
        L.check(self.ctx.lib.hm_pack_weight_pair(self.weight.data_ptr(), A, Bd, self.k * self.k, p1.hi.data_ptr(),
                                                 ops._ptr(p1.lo), p2.hi.data_ptr(), ops._ptr(p2.lo), ops._stream()),
                "hm_pack_weight_pair")
        self.ctx.launches += 1
        self._vf = self._vd = self.fp.version
        return True

    def packed_bwd(self):
        """rows = cin, contraction = cout."""
        kk = self.k * self.k
        if self.thin_out:      # rows = ci, contraction = (kw, co), taps = kh
            if self._pd is None:
                self._pd = PackedWeight(self.ctx, self.cin, self.k * self.cout, self.k, grad=True)
            if self._vd != self.fp.version:
                self._pd.pack_ex(self.ctx, self.weight, 1, kk, 0, self.cout, 1, self.cin * kk, self.k, self.sn_scale)
                self._vd = self.fp.version
            return self._pd
        if self.thin_in:       # rows = (kh, kw, ci), contraction = co, one tap
            if self._pd is None:
                self._pd = PackedWeight(self.ctx, kk * self.cin, self.cout, 1, grad=True)
            if self._vd != self.fp.version:
                self._pd.pack_ex(self.ctx, self.weight, self.cin, 1, kk, 1, self.cin * kk, 0, 0, self.sn_scale)
                self._vd = self.fp.version
            return self._pd
        if self._pd is None:
            self._pd = PackedWeight(self.ctx, self.cin, self.cout, self.k * self.k, grad=True)
        if self._vd != self.fp.version:
            if self.transposed:
                self._pd.pack(self.ctx, self.weight, self.cout * kk, kk, 1, self.sn_scale)
            else:
                self._pd.pack(self.ctx, self.weight, kk, self.cin * kk, 1, self.sn_scale)
            self._vd = self.fp.version
        return self._pd

    # geometry ---------------------------------------------------------------------------------------
    def out_hw(self, h, w, zero_pad):
        """h, w: stored input dims (incl. any materialised border)."""
        if self.transposed:
            return 2 * h, 2 * w
        ke = self.dilation * (self.k - 1) + 1
        return (h + 2 * zero_pad - ke) // self.stride + 1, (w + 2 * zero_pad - ke) // self.stride + 1

    # engine calls -----------------------------------------------------------------------------------
    def forward(self, x, zero_pad, act=ACT_NONE, slope=0.2, out32=None, out16=None, use_bias=True):
        ho, wo = self.out_hw(x.h, x.w, zero_pad)
        b = self.bias if use_bias else None
        if self.thin_out:
            # T[n,h,w',(kw,co)] = KH x 1 conv (N = KW*cout columns), then y = act(b + sum_kw T[h, w+kw, (kw,co)])
            assert out16 is None and out32 is not None
            k, wt, ld = self.k, x.w + 2 * zero_pad, ops.ru(self.k * self.cout, 4)
            T = self.ctx.ws("thinT", x.n * ho * wt * ld * 4)[:x.n * ho * wt * ld].view(x.n, ho, wt, ld)
            ops.conv_fprop(self.ctx, x, self.packed_fwd(), None, k, 1, 1, zero_pad, ho, wt, k * self.cout, out32=T)
            ops.tap_combine(self.ctx, T, 1, k, self.cout, 0, 0, 1, 1, b, act, slope, out32)
            return ho, wo
        if self.thin_in:
            # U[n,h,w,(kh,kw,ci)] = x[n,h+kh-p,w+kw-p,ci], then a 1x1 conv with K = KH*KW*cin
            k = self.k
            U = Operand(self.ctx, x.n, ho, wo, k * k * self.cin)
            ops.tap_unroll(self.ctx, x, U, self.cin, k, k, -zero_pad, -zero_pad, 1, 1)
            ops.conv_fprop(self.ctx, U, self.packed_fwd(), b, 1, 1, 1, 0, ho, wo, self.cout, act, slope, out32, out16)
            return ho, wo
        if self.transposed:
            ops.conv_dgrad(self.ctx, x, self.packed_fwd(), b, self.k, self.k, 2, self.pad, ho, wo, self.cout, act, slope,
                           out32, out16)
        else:
            ops.conv_fprop(self.ctx, x, self.packed_fwd(), b, self.k, self.k, self.stride, zero_pad, ho, wo, self.cout,
                           act, slope, out32, out16, dilation=self.dilation)
        return ho, wo

    def dgrad(self, dy, x_h, x_w, zero_pad, out32):
        """gradient w.r.t. the stored input (dims x_h x x_w)."""
        if self.thin_out:
            ops.conv_dgrad(self.ctx, self._unrolled(dy), self.packed_bwd(), None, self.k, 1, 1, zero_pad, x_h, x_w,
                           self.cin, out32=out32)
            return
        if self.thin_in:
            k, ld = self.k, ops.ru(self.k * self.k * self.cin, 4)
            dU = self.ctx.ws("thinT", dy.n * dy.h * dy.w * ld * 4)[:dy.n * dy.h * dy.w * ld].view(dy.n, dy.h, dy.w, ld)
            ops.conv_dgrad(self.ctx, dy, self.packed_bwd(), None, 1, 1, 1, 0, dy.h, dy.w, k * k * self.cin, out32=dU)
            ops.tap_combine(self.ctx, dU, k, k, self.cin, zero_pad, zero_pad, -1, -1, None, ACT_NONE, 0.0, out32)
            return
        if self.transposed:
            ops.conv_fprop(self.ctx, dy, self.packed_bwd(), None, self.k, self.k, 2, self.pad, x_h, x_w, self.cin,
                           out32=out32)
        else:
            ops.conv_dgrad(self.ctx, dy, self.packed_bwd(), None, self.k, self.k, self.stride, zero_pad, x_h, x_w,
                           self.cin, out32=out32, dilation=self.dilation)

    def dgrad_rows(self, dy, x_h, x_w, zero_pad, out32, row0, nrows):
        """dgrad restricted to input channels [row0, row0 + nrows): out32 is [N, x_h, x_w, ld >= nrows].  The generator
        only needs d(D input)/d(image channels) -- 3 of the 41 input channels of the first PatchGAN layer -- so the
        gradient GEMM runs with a 16-wide N tile instead of 64 and never produces the label / conditioning columns."""
        assert not self.transposed and not self.thin_out and not self.thin_in
        kk = self.k * self.k
        key = (row0, nrows)
        if getattr(self, "_pr_key", None) != key:
            self._pr = PackedWeight(self.ctx, nrows, self.cout, kk, grad=True)
            self._pr_key, self._vr = key, -1
        if self._vr != self.fp.version or self.sn_scale is not None:
            self._pr.pack(self.ctx, self.weight[:, row0:row0 + nrows], kk, self.cin * kk, 1, self.sn_scale)
            self._vr = self.fp.version
        ops.conv_dgrad(self.ctx, dy, self._pr, None, self.k, self.k, self.stride, zero_pad, x_h, x_w, nrows, out32=out32)

    def _unrolled(self, dy):
        """U[n,h,w',(kw,co)] = dy[n,h,w'-kw,co], w' in [0, W+KW-1): the thin output gradient with its horizontal taps
        folded into channels (shared by dgrad and wgrad of the same dy)."""
        if self._u is not None and self._u[0] is dy:
            return self._u[1]
        U = Operand(self.ctx, dy.n, dy.h, dy.w + self.k - 1, self.k * self.cout, grad=True)
        ops.tap_unroll(self.ctx, dy, U, self.cout, 1, self.k, 0, 0, 1, -1)
        self._u = (dy, U)
        return U

    def wgrad(self, x, dy, zero_pad, bias_grad=True):
        """accumulate into .grad of weight and bias.  bias_grad=False for convs that feed an InstanceNorm: the bias
        cancels in the normalisation, its gradient is analytically zero (the reference computes ~1e-9 rounding noise),
        so .grad stays exactly 0 and the full-resolution column reduction is skipped."""
        if self.thin_in:
            raise NotImplementedError("weight gradient of a thin-input conv (only the frozen VGG conv1_1 is one)")
        if self.thin_out:
            ops.conv_wgrad(self.ctx, x, self._unrolled(dy), self.k, 1, 1, zero_pad, self.weight.grad, accumulate=True,
                           unpack_cols=(self.k, self.cout))
        elif self.transposed:
            ops.conv_wgrad(self.ctx, dy, x, self.k, self.k, 2, self.pad, self.weight.grad, accumulate=True)
        else:
            ops.conv_wgrad(self.ctx, x, dy, self.k, self.k, self.stride, zero_pad, self.weight.grad, accumulate=True,
                           dilation=self.dilation)
        if bias_grad and self.has_bias:
            ops.colsum_operand(self.ctx, dy, self.bias.grad, accumulate=True)


def _f32(ctx, *shape):
    return torch.empty(shape, dtype=torch.float32, device=ctx.device)


# ======================================================================================================
# GlobalGenerator
# ======================================================================================================
class GlobalGenerator(object):
    """models/Pix2Pix_NET.py:63-101.  `with_head=False` gives the trunk LocalEnhancer embeds (Pix2Pix_NET.py:17-19)."""

    IN_BORDER = {"stem": 3, "down": 0, "resA": 1, "resB": 1, "up": 0, "head": 3}

    def __init__(self, ctx, fp, input_nc, output_nc, ngf=64, n_downsampling=3, n_blocks=9, use_output_gate=False,
                 prefix="model.", with_head=True):
        self.ctx, self.fp = ctx, fp
        self.input_nc, self.output_nc, self.use_output_gate = input_nc, output_nc, use_output_gate
        st = []  # (kind, ConvP)
        st.append(("stem", ConvP(ctx, fp, prefix + "1", input_nc, ngf, 7, 1, 0)))
        idx = 4
        for i in range(n_downsampling):
            m = 2 ** i
            st.append(("down", ConvP(ctx, fp, prefix + str(idx), ngf * m, ngf * m * 2, 3, 2, 1)))
            idx += 3
        m = 2 ** n_downsampling
        for i in range(n_blocks):
            st.append(("resA", ConvP(ctx, fp, prefix + "%d.conv_block.1" % idx, ngf * m, ngf * m, 3, 1, 0)))
            st.append(("resB", ConvP(ctx, fp, prefix + "%d.conv_block.5" % idx, ngf * m, ngf * m, 3, 1, 0)))
            idx += 1
        for i in range(n_downsampling):
            m = 2 ** (n_downsampling - i)
            st.append(("up", ConvP(ctx, fp, prefix + str(idx), ngf * m, ngf * m // 2, 3, 2, 1, transposed=True)))
            idx += 3
        if with_head:
            st.append(("head", ConvP(ctx, fp, prefix + str(idx + 1), ngf, output_nc, 7, 1, 0)))
        self.stages = st
        self.with_head = with_head
        self.feature_nc = ngf

    def convs(self):
        return [c for _, c in self.stages]

    def forward(self, x, feature_border=0, add_at=None, add_tensor=None, skip32_init=None, concat=None):
        """x: Operand with ReflectionPad2d(3) materialised.  Returns (out, tape): out is the fp32 NHWC tanh output
        (with_head) or (feature fp32 NHWC) for the trunk.  add_at/add_tensor: dense fp32 tensor added to the output of
        stage `add_at` after its activation (LocalEnhancer: model_downsample(x) + output_prev, Pix2Pix_NET.py:60).
        skip32_init: fp32 value of x when the first stage is a ResnetBlock (its residual input).  concat: {stage index:
        Operand} concatenated in FRONT of that stage's input channels (skip connections, Pix2Pix_NET.py:221)."""
        ctx = self.ctx
        tape = []
        cur = x
        skip32 = skip32_init   # fp32 copy of the current activation when it is a residual input
        out = None
        n = len(self.stages)
        for s, (kind, conv) in enumerate(self.stages):
            nxt = self.stages[s + 1][0] if s + 1 < n else None
            zero_pad = conv.pad if kind == "down" else 0
            concat_c = 0
            if concat and s in concat:
                concat_c = concat[s].c
                cur = ops.concat_operands(ctx, concat[s], cur)
            ho, wo = conv.out_hw(cur.h, cur.w, zero_pad)
            rec = dict(kind=kind, conv=conv, xin=cur, zero_pad=zero_pad, ho=ho, wo=wo, concat_c=concat_c)
            if kind == "head":
                y = _f32(ctx, cur.n, ho, wo, conv.cout)
                conv.forward(cur, 0, act=ACT_TANH, out32=y)
                rec.update(y=y)
                tape.append(rec)
                out = y
                break
            y = _f32(ctx, cur.n, ho, wo, conv.cout)
            conv.forward(cur, zero_pad, out32=y)
            mean, rstd = ops.in_stats(ctx, y)
            out_border = self.IN_BORDER[nxt] if nxt is not None else feature_border
            emit32 = (nxt == "resA") or (nxt is None)
            act = ACT_NONE if kind == "resB" else ACT_RELU
            o32 = _f32(ctx, cur.n, ho, wo, conv.cout) if emit32 else None
            oop = Operand(ctx, cur.n, ho, wo, conv.cout, border=out_border) if nxt is not None else None
            skip = skip32 if kind == "resB" else (add_tensor if s == add_at else None)
            ops.in_apply(ctx, y, mean, rstd, act, skip=skip, out32=o32, out_op=oop, reflect=True)
            rec.update(y=y, mean=mean, rstd=rstd, act=act, out_border=out_border, shape=(cur.n, ho, wo, conv.cout))
            tape.append(rec)
            if kind == "resA":
                pass  # skip32 (input of the block) stays alive for resB
            else:
                skip32 = o32
            cur = oop
            out = o32
        return out, tape

    def backward(self, tape, dy_head=None, dfeat=None, need_input_grad=False, add_at=None, extra_grad=None,
                 grad_ready=None, wgrad_stream=None):
        """dy_head: Operand gradient w.r.t. the head's pre-tanh output (with_head) or dfeat: dense fp32 gradient
        w.r.t. the trunk feature.  Accumulates parameter gradients; returns d(input operand) (fp32, padded space)
        when need_input_grad.  With add_at, self.add_grad holds the dense gradient w.r.t. the tensor added there.
        extra_grad: {stage index: (fp32 tensor [N,h,w,ld], ld, coff)} additional gradient w.r.t. that stage's output
        (skip connections).  Afterwards self.concat_grads[stage] = (tensor, ld, coff) is the gradient w.r.t. the operand
        concatenated at that stage and self.input_T the dense gradient w.r.t. x when the first stage is a ResnetBlock."""
        ctx = self.ctx
        self.add_grad = None
        self.concat_grads = {}
        self.input_T = None
        # wgrad_stream: the weight gradient of a layer only feeds .grad, so it runs on a second stream while this stream
        # continues with the data gradient -> InstanceNorm-backward chain of the next layer; the HBM-bound kernels of
        # the chain then overlap weight-gradient GEMMs and idle SMs of partial waves get filled.  The gradient operands
        # are kept alive until the streams are joined (they are read by both).
        keep = []
        main = torch.cuda.current_stream() if wgrad_stream is not None else None
        G1, G1_border = None, 0   # gradient w.r.t. the current stage's OUTPUT operand (padded space)
        G1_ld, G1_coff = None, 0  # channel stride / offset of G1 when it is a slice of a concatenated input's gradient
        T = dfeat                 # dense gradient w.r.t. the current stage's fp32 output (residual chain)
        for s in range(len(tape) - 1, -1, -1):
            rec = tape[s]
            kind, conv, xin = rec["kind"], rec["conv"], rec["xin"]
            nxt = tape[s + 1]["kind"] if s + 1 < len(tape) else None
            if kind == "head":
                dy = dy_head
            else:
                N, ho, wo, cc = rec["shape"]
                dy = Operand(ctx, N, ho, wo, cc, grad=True)
                if s == add_at:   # d(out)/d(added tensor) = identity: the total gradient w.r.t. this stage's output
                    if nxt == "resA" or nxt is None:
                        self.add_grad = T
                    elif G1_border > 0:
                        self.add_grad = _f32(ctx, N, ho, wo, cc)
                        ops.fold_add(ctx, G1, G1_border, None, self.add_grad)
                    else:
                        self.add_grad = G1
                if kind == "resB":
                    if nxt != "resA" and G1 is not None:   # last block: gradient arrives from the next conv only
                        if G1_border > 0:
                            T = _f32(ctx, N, ho, wo, cc)
                            ops.fold_add(ctx, G1, G1_border, None, T)
                        else:
                            T = G1
                    ops.in_bwd(ctx, rec["shape"], ACT_NONE, y=rec["y"], mean=rec["mean"], rstd=rec["rstd"], g2=T,
                               out_op=dy)
                elif nxt == "resA" or nxt is None:          # output feeds the residual chain / is the trunk feature
                    ops.in_bwd(ctx, rec["shape"], rec["act"], y=rec["y"], mean=rec["mean"], rstd=rec["rstd"], g2=T,
                               out_op=dy)
                elif extra_grad and s in extra_grad:          # gradient from the next conv + a skip connection
                    eg, eg_ld, eg_coff = extra_grad[s]
                    assert G1_border == 0 and G1_ld is None
                    ops.in_bwd(ctx, rec["shape"], rec["act"], y=rec["y"], mean=rec["mean"], rstd=rec["rstd"], g1=eg,
                               g1_border=0, g1_ld=eg_ld, g1_coff=eg_coff, g2=G1, out_op=dy)
                else:
                    ops.in_bwd(ctx, rec["shape"], rec["act"], y=rec["y"], mean=rec["mean"], rstd=rec["rstd"], g1=G1,
                               g1_border=G1_border, g1_ld=G1_ld, g1_coff=G1_coff, out_op=dy)
            if wgrad_stream is not None:
                keep.append(dy)
                if conv.thin_out:            # the unrolled gradient is shared by wgrad and dgrad: produce it on this stream
                    conv._unrolled(dy)
                wgrad_stream.wait_stream(main)
                with torch.cuda.stream(wgrad_stream):
                    conv.wgrad(xin, dy, rec["zero_pad"], bias_grad=(kind == "head"))
                    if grad_ready is not None:
                        grad_ready(conv)
            else:
                conv.wgrad(xin, dy, rec["zero_pad"], bias_grad=(kind == "head"))
                if grad_ready is not None:   # this conv's .grad is final: stages run last-to-first, so the flat gradient
                    grad_ready(conv)         # buffer is complete from this parameter's offset to its end
            if s == 0 and not need_input_grad:
                if wgrad_stream is not None:
                    main.wait_stream(wgrad_stream)
                return None
            gin = _f32(ctx, xin.n, xin.h, xin.w, conv.cin)
            conv.dgrad(dy, xin.h, xin.w, rec["zero_pad"], gin)
            if kind == "resA":
                # input of the block: T_k = T_{k+1} + fold(gin)
                Tn = _f32(ctx, xin.n, xin.ih, xin.iw, conv.cin)
                ops.fold_add(ctx, gin, xin.border, T, Tn)
                T = Tn
                G1, G1_border, G1_ld, G1_coff = None, 0, None, 0
            else:
                G1, G1_border, G1_ld, G1_coff = gin, xin.border, None, 0
                cc_ = rec.get("concat_c", 0)
                if cc_:   # input was cat((skip, prev)): the first cc_ channels belong to the skip operand
                    self.concat_grads[s] = (gin, conv.cin, 0)
                    G1_ld, G1_coff = conv.cin, cc_
        self.input_T = T
        if wgrad_stream is not None:
            main.wait_stream(wgrad_stream)
        return G1


# ======================================================================================================
# MultiscaleDiscriminator (getIntermFeat=True)
# ======================================================================================================
class MultiscaleDiscriminator(object):
    """models/Discriminator_NET.py:11-118: num_D PatchGANs on an AvgPool(3,2,1) pyramid, every layer output kept."""

    def __init__(self, ctx, fp, input_nc, ndf=64, n_layers=3, num_D=3, spectral_norm=False, getIntermFeat=True,
                 use_sigmoid=False):
        self.ctx, self.fp = ctx, fp
        # use_sigmoid (--no_lsgan): the nn.Sigmoid the reference appends (Discriminator_NET.py:95-96) is folded into the
        # BCE loss / gradient kernels (hm_bce_sum / hm_bce_grad); the last tap stays the raw logit
        self.use_sigmoid = bool(use_sigmoid)
        self.input_nc, self.n_layers, self.num_D = input_nc, n_layers, num_D
        self.spectral_norm = bool(spectral_norm)
        self._sn = None
        self.scales = []

        def key(s, j):
            # parameter names of the reference module tree (Discriminator_NET.py:24-29,99-106): per-layer Sequentials
            # 'scale{s}_layer{j}.0' with getIntermFeat, else ONE flattened Sequential 'layer{s}' in which layer j's
            # conv sits at index 0, 2, 5, 8, ... (conv [+ norm] + LeakyReLU per layer)
            if getIntermFeat:
                return "scale%d_layer%d.0" % (s, j)
            return "layer%d.%d" % (s, 0 if j == 0 else 2 + 3 * (j - 1))
        for s in range(num_D):
            layers = []
            nf = ndf
            layers.append(ConvP(ctx, fp, key(s, 0), input_nc, ndf, 4, 2, 2))
            for n in range(1, n_layers):
                nf_prev, nf = nf, min(nf * 2, 512)
                layers.append(ConvP(ctx, fp, key(s, n), nf_prev, nf, 4, 2, 2))
            nf_prev, nf = nf, min(nf * 2, 512)
            layers.append(ConvP(ctx, fp, key(s, n_layers), nf_prev, nf, 4, 1, 2))
            layers.append(ConvP(ctx, fp, key(s, n_layers + 1), nf, 1, 4, 1, 2))
            self.scales.append(layers)

    def convs(self):
        return [c for sc in self.scales for c in sc]

    def setup_spectral_norm(self, gen):
        """K13 (opt-in; the reference's MultiscaleDiscriminator uses plain convs): every conv becomes an
        SNConv2d (models/sn_utils.py:49-72) -- `u` ~ N(0,1) [1, Cout] kept as a non-trainable '<conv>.u' state entry,
        one power iteration per discriminator evaluation, W / sigma fused into the weight pack."""
        import ctypes as C
        from . import _lib as L
        convs = self.convs()
        arr = (L.SnLayer * len(convs))()
        self._sn_keep = []
        max_n = max_m = 0
        for i, c in enumerate(convs):
            n, m = c.cout, c.cin * c.k * c.k
            u = torch.randn(1, n, generator=gen).to(self.ctx.device)
            stash = torch.zeros(self.ctx.lib.hm_sn_stash_floats(n, m), dtype=torch.float32, device=self.ctx.device)
            self.fp.buffers[c.name + ".u"] = u
            c.sn_scale = stash[2 * n + m + 1:2 * n + m + 2]
            self._sn_keep.append((u, stash))
            arr[i] = L.SnLayer(c.weight.data_ptr(), c.weight.grad.data_ptr(), u.data_ptr(), stash.data_ptr(), n, m)
            max_n, max_m = max(max_n, n), max(max_m, m)
        dev = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(self.ctx.device)
        self._sn = dict(layers=dev, n=len(convs), max_n=max_n, max_m=max_m)

    def sigma(self, conv):
        """current spectral norm estimate of `conv` (device tensor, 1 element)."""
        return 1.0 / conv.sn_scale

    def forward(self, d_in, update_u=True):
        """d_in: Operand [N,H,W,cs] (no border).  Returns tape: list over pyramid level i of dict(x=[operands],
        taps=[fp32 NHWC], y/mean/rstd per layer); level i uses scale{num_D-1-i} (Discriminator_NET.py:49-57)."""
        ctx = self.ctx
        if self._sn is not None:
            ops.sn_power_iteration(ctx, self._sn["layers"], self._sn["n"], self._sn["max_n"], self._sn["max_m"], update_u)
            for c in self.convs():   # sigma changed: the packed W / sigma slabs are stale
                c._vf = c._vd = -1
        tape = []
        x = d_in
        for i in range(self.num_D):
            layers = self.scales[self.num_D - 1 - i]
            lv = dict(layers=layers, xs=[], taps=[], ys=[], means=[], rstds=[])
            cur = x
            for j, conv in enumerate(layers):
                ho, wo = conv.out_hw(cur.h, cur.w, 2)
                lv["xs"].append(cur)
                last = j == len(layers) - 1
                if j == 0:
                    tap = _f32(ctx, cur.n, ho, wo, conv.cout)
                    # the conv epilogue writes channels [0, cout) only: zero the padding channels when cout % 8 != 0
                    # (uninitialised bf16 NaN patterns x zero weights would poison the next GEMM)
                    nxt = Operand(ctx, cur.n, ho, wo, conv.cout, zero=(conv.cout % 8 != 0))
                    conv.forward(cur, 2, act=ACT_LRELU, slope=0.2, out32=tap, out16=nxt)
                    lv["ys"].append(None); lv["means"].append(None); lv["rstds"].append(None)
                elif last:
                    tap = _f32(ctx, cur.n, ho, wo, conv.cout)
                    conv.forward(cur, 2, out32=tap)
                    nxt = None
                    lv["ys"].append(None); lv["means"].append(None); lv["rstds"].append(None)
                else:
                    y = _f32(ctx, cur.n, ho, wo, conv.cout)
                    conv.forward(cur, 2, out32=y)
                    mean, rstd = ops.in_stats(ctx, y)
                    tap = _f32(ctx, cur.n, ho, wo, conv.cout)
                    nxt = Operand(ctx, cur.n, ho, wo, conv.cout)
                    ops.in_apply(ctx, y, mean, rstd, ACT_LRELU, 0.2, out32=tap, out_op=nxt)
                    lv["ys"].append(y); lv["means"].append(mean); lv["rstds"].append(rstd)
                lv["taps"].append(tap)
                cur = nxt
            tape.append(lv)
            if i != self.num_D - 1:
                xn = Operand(ctx, x.n, (x.h - 1) // 2 + 1, (x.w - 1) // 2 + 1, x.c, cs=x.cs)
                ops.avgpool3s2(ctx, x, xn)
                x = xn
        return tape

    def backward(self, tape, nb, mode, w_gan=1.0, w_feat=0.0, w_real=0.5, w_fake=0.5, img_c0=0, nseg=2):
        """mode 'G': d(w_gan*G_GAN + G_GAN_Feat terms)/d(input) for the first nb images (the fake half); returns the
        fp32 gradient w.r.t. the IMAGE channels [img_c0, img_c0+3) of the full-resolution D input as a [nb,H,W,4] tensor
        (channels 0..2; self.gin_coff = 0).  No weight grads.
        mode 'D': accumulates weight grads of w_fake*D_fake + w_real*D_real over all 2*nb images.
        w_feat is the complete per-tap L1 coefficient numerator (D_weights*feat_weights*lambda_feat)."""
        ctx = self.ctx
        gins = []
        for i, lv in enumerate(tape):
            layers = lv["layers"]
            nl = len(layers)
            pred = lv["taps"][-1]
            N = pred.shape[0]
            half_numel = pred.numel() // nseg          # nseg == 3: [fake ; real ; pooled fakes] (--pool_size > 0)
            if mode == "G":
                nimg = nb
                dy = Operand(ctx, nimg, pred.shape[1], pred.shape[2], 1, grad=True)
                ops.mse_grad(ctx, pred[:nb], 1.0, 2.0 * w_gan / half_numel, dy, bce=self.use_sigmoid)
            else:
                nimg = N
                dy = Operand(ctx, N, pred.shape[1], pred.shape[2], 1, grad=True)
                if nseg == 2:
                    ops.mse_grad(ctx, pred[:nb], 0.0, 2.0 * w_fake / half_numel, dy, 0, bce=self.use_sigmoid)   # fake half: target 0
                else:   # loss_D_fake is taken on the image pool's history (third segment); the current fakes carry none
                    ops.mse_grad(ctx, pred[:nb], 0.0, 0.0, dy, 0)
                    ops.mse_grad(ctx, pred[2 * nb:], 0.0, 2.0 * w_fake / half_numel, dy, 2 * nb, bce=self.use_sigmoid)
                ops.mse_grad(ctx, pred[nb:2 * nb], 1.0, 2.0 * w_real / half_numel, dy, nb, bce=self.use_sigmoid)  # real half: target 1
            for j in range(nl - 1, -1, -1):
                conv = layers[j]
                xin = lv["xs"][j]
                if mode == "D":
                    conv.wgrad(xin, dy, 2, bias_grad=(j == 0 or j == nl - 1))
                    if j == 0:
                        break
                if j == 0 and mode == "G":
                    # only the image channels of the D input depend on the generator: 3-row gradient (ld 4, offset 0)
                    gin = _f32(ctx, nimg, xin.h, xin.w, 4)
                    conv.dgrad_rows(dy, xin.h, xin.w, 2, gin, img_c0, 3)
                    gins.append(gin)
                    break
                gin = _f32(ctx, nimg, xin.h, xin.w, conv.cin)
                conv.dgrad(dy, xin.h, xin.w, 2, gin)
                if j == 0:
                    gins.append(gin)
                    break
                # through layer j-1's LeakyReLU (+InstanceNorm) to its conv output
                tap = lv["taps"][j - 1]
                shape = (nimg,) + tuple(tap.shape[1:])
                dyn = Operand(ctx, nimg, tap.shape[1], tap.shape[2], tap.shape[3], grad=True)
                tref, l1 = None, 0.0
                if mode == "G" and w_feat != 0.0:
                    tref = tap[nb:2 * nb]
                    l1 = w_feat / (tap.numel() // nseg)
                y, mean, rstd = lv["ys"][j - 1], lv["means"][j - 1], lv["rstds"][j - 1]
                ops.in_bwd(ctx, shape, ACT_LRELU, 0.2, y=y[:nimg] if y is not None else None,
                           mean=mean[:nimg] if mean is not None else None, rstd=rstd[:nimg] if rstd is not None else None,
                           z=tap[:nimg], g1=gin, g1_border=0, tref=tref, l1coef=l1, out_op=dyn)
                dy = dyn
        if mode != "G":
            if self._sn is not None:   # dL/dW_bar -> dL/dW through sigma(W) (sn_utils.py:11-25 under autograd)
                ops.sn_weight_grad(ctx, self._sn["layers"], self._sn["n"], self._sn["max_n"], self._sn["max_m"])
            return None
        # total gradient at full resolution: g0 + poolT(g1 + poolT(g2 ...)); the tensors hold the 3 image channels only
        for i in range(len(gins) - 1, 0, -1):
            ops.avgpool3s2_bwd(ctx, gins[i], gins[i - 1], 0, 3)
        self.gin_coff = 0      # channel offset of the image gradient inside the returned tensor
        return gins[0]


# ======================================================================================================
# VGG19 feature tower (frozen)
# ======================================================================================================
class Vgg19(object):
    """models/layer_util.py:381-411: torchvision VGG19 features[0:30], taps relu{1..5}_1, requires_grad=False."""

    def __init__(self, ctx, state_dict):
        self.ctx = ctx
        self.fp = FlatParams(ctx.device)
        self.convs_ = []
        for idx, cin, cout in VGG19_CONVS:
            self.convs_.append((idx, ConvP(ctx, self.fp, "slice%d.%d" % (VGG19_SLICE_OF[idx], idx), cin, cout, 3, 1, 1)))
        self.fp.materialize()
        self.fp.load_state_dict(state_dict)
        for p in self.fp.params.values():
            p.requires_grad_(False)
            p.grad = None
        self.fp.grad = None

    def forward(self, v_in, tape=None, n0=0, n=None):
        """v_in: Operand [N,H,W,8] (3 valid channels).  Returns tape with per-conv input operands / outputs and taps.
        n0 / n: evaluate images [n0, n0 + n) only, into the full-batch buffers of `tape` (allocated by the first call):
        the tower has no cross-sample operation, so the real half of the [fake ; real] batch -- which does not depend on
        the generator -- can run while the generator's forward pass is still in flight."""
        ctx = self.ctx
        n = v_in.n - n0 if n is None else n
        part = (n0, n) != (0, v_in.n)
        first = tape is None
        if first:
            tape = dict(xs=[], outs=[], taps={}, pooled_from={})
        sl = (lambda op: op.images(n0, n)) if part else (lambda op: op)
        cur = v_in
        for li, (idx, conv) in enumerate(self.convs_):
            if idx in VGG19_POOL_BEFORE:
                if first:
                    pooled = Operand(ctx, cur.n, cur.h // 2, cur.w // 2, cur.c, cs=cur.cs)
                    tape["pooled_from"][li] = cur
                else:
                    pooled = tape["xs"][li]
                ops.maxpool2(ctx, sl(cur), sl(pooled))
                cur = pooled
            if first:
                tape["xs"].append(cur)
                out = Operand(ctx, cur.n, cur.h, cur.w, conv.cout, zero=(conv.cout % 8 != 0))
                tape["outs"].append(out)
                if idx in VGG19_TAP_AFTER:
                    tape["taps"][li] = _f32(ctx, cur.n, cur.h, cur.w, conv.cout)
            out = tape["outs"][li]
            tap = tape["taps"].get(li)
            conv.forward(sl(cur), 1, act=ACT_RELU, out32=(tap[n0:n0 + n] if (tap is not None and part) else tap), out16=sl(out))
            cur = out
        return tape

    def backward(self, tape, nb, coefs):
        """d(sum_i coefs[i] * mean|tap_i[:nb] - tap_i[nb:]|)/d(input[:nb]) -> fp32 [nb,H,W,3].  Frozen weights: dgrad only."""
        ctx = self.ctx
        g = None  # dense gradient w.r.t. the output of conv li (after relu, before any pool)
        for li in range(len(self.convs_) - 1, -1, -1):
            idx, conv = self.convs_[li]
            out = tape["outs"][li]
            shape = (nb, out.h, out.w, conv.cout)
            dy = Operand(ctx, nb, out.h, out.w, conv.cout, grad=True)
            tap = tape["taps"].get(li)
            if tap is not None:
                l1 = coefs[VGG19_TAP_AFTER[idx]] / (tap.numel() // 2)
                ops.in_bwd(ctx, shape, ACT_RELU, z=tap[:nb], g2=g, tref=tap[nb:], l1coef=l1, out_op=dy)
            else:
                ops.in_bwd(ctx, shape, ACT_RELU, mask_op=out, g2=g, out_op=dy)
            xin = tape["xs"][li]
            gin = _f32(ctx, nb, xin.h, xin.w, conv.cin)
            conv.dgrad(dy, xin.h, xin.w, 1, gin)
            if li in tape["pooled_from"]:
                src = tape["pooled_from"][li]
                if (src.h | src.w) & 1:   # odd extent: the last row / column is outside every 2x2 window -> zero gradient
                    dz = torch.zeros(nb, src.h, src.w, src.c, dtype=torch.float32, device=ctx.device)
                else:
                    dz = _f32(ctx, nb, src.h, src.w, src.c)
                ops.maxpool2_bwd(ctx, gin, src, dz)
                g = dz
            else:
                g = gin
        return g

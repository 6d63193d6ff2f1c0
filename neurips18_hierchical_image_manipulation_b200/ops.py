"""Python faces of the C-ABI kernels (include/hm_b200.h).  torch is used for device memory and streams only:
every function below enqueues hand-written sm_100a kernels from libhm_b200.so on torch's current stream.
"""
import ctypes as C

import torch

from . import _lib as L

ACT_NONE, ACT_RELU, ACT_LRELU, ACT_TANH = 0, 1, 2, 3


def ru(v, m):
    return (v + m - 1) // m * m


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return None if t is None else t.data_ptr()


_SCRATCH = {}   # device index -> the engines' scratch tensor (process lifetime: captured CUDA graphs hold its address)


def _register_scratch(lib, device):
    """One scratch per process / device for the stream-K tail of the CTA-pair engine (include/hm_b200.h:hm_set_scratch)."""
    key = device.index if device.index is not None else torch.cuda.current_device()
    if key not in _SCRATCH:
        t = torch.zeros(int(lib.hm_scratch_bytes()), dtype=torch.uint8, device=device)   # counters start at zero
        torch.cuda.synchronize(device)
        L.check(lib.hm_set_scratch(t.data_ptr(), t.numel()), "hm_set_scratch")
        _SCRATCH[key] = t
    return _SCRATCH[key]


class Ctx(object):
    """Per-device state shared by all ops: library handle, precision mode, pipeline error flag."""

    def __init__(self, device, split=True, split_bwd=None):
        self.lib = L.load()
        self.device = torch.device(device)
        self.scratch = _register_scratch(self.lib, self.device)
        self.sm_count = torch.cuda.get_device_properties(self.device).multi_processor_count
        self.split = bool(split)  # True: bf16x3 (fp32-parity mode); False: plain bf16 products
        # gradient GEMMs (dgrad / wgrad) may run with single bf16 products while the forward stays bf16x3 ("mixed")
        self.split_bwd = self.split if split_bwd is None else bool(split_bwd)
        self.err = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._ws = {}
        self.launches = 0  # number of hm kernels enqueued (bench.py reports it)

    def ws(self, key, nbytes):
        """Grow-only scratch buffers keyed by use AND by the current stream (branches of the fused step that run on
        different streams must not share scratch)."""
        n = (int(nbytes) + 3) // 4
        key = (key, torch.cuda.current_stream().cuda_stream)
        t = self._ws.get(key)
        if t is None or t.numel() < n:
            t = torch.empty(max(n, 1), dtype=torch.float32, device=self.device)
            self._ws[key] = t
        return t

    def check_pipeline(self):
        code = int(self.err.item())
        if code != 0:
            raise L.HmError("tcgen05 engine pipeline wait timed out (code %d)" % code)


class Operand(object):
    """bf16 (hi, lo) NHWC operand [n, h, w, cs] with c valid channels; h/w include `border`."""

    __slots__ = ("hi", "lo", "n", "h", "w", "c", "cs", "border", "lo_c0")

    def __init__(self, ctx, n, h, w, c, border=0, cs=None, zero=False, grad=False):
        cs = ru(c, 8) if cs is None else cs
        hp, wp = h + 2 * border, w + 2 * border
        alloc = torch.zeros if zero else torch.empty
        split = ctx.split_bwd if grad else ctx.split
        self.hi = alloc((n, hp, wp, cs), dtype=torch.bfloat16, device=ctx.device)
        self.lo = alloc((n, hp, wp, cs), dtype=torch.bfloat16, device=ctx.device) if split else None
        self.n, self.h, self.w, self.c, self.cs, self.border = n, hp, wp, c, cs, border
        self.lo_c0 = 0   # channels [0, lo_c0) hold values exact in bf16 (one-hot maps): lo plane zero there, engines skip it

    def struct(self, n0=0, n=None):
        """hm_operand for images [n0, n0+n)."""
        n = self.n - n0 if n is None else n
        off = n0 * self.h * self.w * self.cs * 2
        return L.Operand(self.hi.data_ptr() + off, (self.lo.data_ptr() + off) if self.lo is not None else None, n,
                         self.h, self.w, self.c, self.cs, getattr(self, "lo_c0", 0))

    def images(self, n0, n):
        """View of images [n0, n0 + n): an Operand sharing this one's storage (the batch index is the outermost)."""
        v = object.__new__(Operand)
        v.hi = self.hi[n0:n0 + n]
        v.lo = self.lo[n0:n0 + n] if self.lo is not None else None
        v.n, v.h, v.w, v.c, v.cs, v.border, v.lo_c0 = n, self.h, self.w, self.c, self.cs, self.border, self.lo_c0
        return v

    @property
    def ih(self):
        return self.h - 2 * self.border

    @property
    def iw(self):
        return self.w - 2 * self.border

    def dense(self):
        """fp32 NCHW copy of the interior (tests / visuals only)."""
        v = self.hi.float()
        if self.lo is not None:
            v = v + self.lo.float()
        b = self.border
        if b:
            v = v[:, b:-b, b:-b]
        return v[..., :self.c].permute(0, 3, 1, 2).contiguous()


class PackedWeight(object):
    """bf16 [taps][rows_pad][k_pad] slabs of one conv weight for one engine role."""

    def __init__(self, ctx, rows, k, taps, grad=False):
        lib = ctx.lib
        self.rows, self.k, self.taps = rows, k, taps
        self.rows_pad, self.k_pad = lib.hm_rows_pad(rows), lib.hm_k_pad(k)
        n = taps * self.rows_pad * self.k_pad
        split = ctx.split_bwd if grad else ctx.split
        self.hi = torch.empty(n, dtype=torch.bfloat16, device=ctx.device)
        self.lo = torch.empty(n, dtype=torch.bfloat16, device=ctx.device) if split else None

    def pack(self, ctx, w, s_row, s_k, s_tap, scale=None):
        if scale is not None:   # spectral norm: W / sigma fused into the weight load
            return self.pack_ex(ctx, w, 1, s_row, 0, 1, s_k, 0, s_tap, scale)
        L.check(ctx.lib.hm_pack_weight(w.data_ptr(), self.rows, self.k, self.taps, s_row, s_k, s_tap, self.hi.data_ptr(),
                                       _ptr(self.lo), _stream()), "hm_pack_weight")
        ctx.launches += 1

    def pack_ex(self, ctx, w, r_div, s_r_hi, s_r_lo, k_div, s_k_hi, s_k_lo, s_tap, scale=None):
        """two-level row / contraction strides (tap-unrolled thin convs) and an optional device scalar multiplier."""
        L.check(ctx.lib.hm_pack_weight_ex(w.data_ptr(), self.rows, r_div, s_r_hi, s_r_lo, self.k, k_div, s_k_hi, s_k_lo,
                                          self.taps, s_tap, _ptr(scale), self.hi.data_ptr(), _ptr(self.lo), _stream()),
                "hm_pack_weight_ex")
        ctx.launches += 1


def conv_fprop(ctx, x, pw, bias, kh, kw, stride, pad, hout, wout, cout, act=ACT_NONE, slope=0.2, out32=None,
               out16=None, out16_coff=0, n0=0, n=None, dilation=1):
    """x: Operand (stored, incl. border); out32: fp32 [N,hout,wout,cout]; out16: Operand (written at its interior)."""
    xs = x.struct(n0, n)
    o32 = L.OutF32(out32.data_ptr(), hout, wout, out32.shape[-1], 0, 0, 0) if out32 is not None else None
    o16 = None
    if out16 is not None:
        o16 = L.OutBF16(out16.hi.data_ptr(), _ptr(out16.lo), out16.h, out16.w, out16.cs, out16.border, out16.border,
                        out16_coff)
    L.check(ctx.lib.hm_conv_fprop_dil(C.byref(xs), pw.hi.data_ptr(), _ptr(pw.lo), pw.k_pad, pw.rows_pad, _ptr(bias), kh, kw,
                                      stride, pad, dilation, hout, wout, cout, act, slope, C.byref(o32) if o32 else None,
                                      C.byref(o16) if o16 else None, ctx.err.data_ptr(), _stream()), "hm_conv_fprop")
    ctx.launches += 1


def conv_dgrad(ctx, dy, pw, bias, kh, kw, stride, pad, hout, wout, cout, act=ACT_NONE, slope=0.2, out32=None,
               out16=None, dilation=1):
    ds = dy.struct()
    o32 = L.OutF32(out32.data_ptr(), hout, wout, out32.shape[-1], 0, 0, 0) if out32 is not None else None
    o16 = None
    if out16 is not None:
        o16 = L.OutBF16(out16.hi.data_ptr(), _ptr(out16.lo), out16.h, out16.w, out16.cs, out16.border, out16.border, 0)
    L.check(ctx.lib.hm_conv_dgrad_dil(C.byref(ds), pw.hi.data_ptr(), _ptr(pw.lo), pw.k_pad, pw.rows_pad, _ptr(bias), kh, kw,
                                      stride, pad, dilation, hout, wout, cout, act, slope, C.byref(o32) if o32 else None,
                                      C.byref(o16) if o16 else None, ctx.err.data_ptr(), _stream()), "hm_conv_dgrad")
    ctx.launches += 4 if stride == 2 else 1


def conv_wgrad(ctx, P, Q, kh, kw, stride, pad, dst, accumulate=True, n0P=0, n0Q=0, n=None, unpack_cols=None, dilation=1):
    """dst[cq][cp][kh][kw] (+)= sum_pixels P[., y*stride+kh-pad, ., cp] * Q[., y, ., cq].
    unpack_cols=(KW, cq): Q is a tap-unrolled operand with channels (kw, cq) and kw == 1 here; dst is [cq][cp][kh][KW]."""
    ps, qs = P.struct(n0P, n), Q.struct(n0Q, n)
    if not ctx.split_bwd:   # "mixed" / bf16 modes: one bf16 product for gradient GEMMs
        ps.lo = None
        qs.lo = None
    ws = ctx.ws("wgrad", ctx.lib.hm_wgrad_ws_bytes(kh, kw, P.c, Q.c))
    L.check(ctx.lib.hm_conv_wgrad_dil(C.byref(ps), C.byref(qs), kh, kw, stride, pad, dilation, ws.data_ptr(),
                                      ctx.err.data_ptr(), _stream()), "hm_conv_wgrad")
    if unpack_cols is not None:
        L.check(ctx.lib.hm_wgrad_unpack_cols(ws.data_ptr(), kh, unpack_cols[0], P.c, unpack_cols[1], dst.data_ptr(),
                                             1 if accumulate else 0, _stream()), "hm_wgrad_unpack_cols")
    else:
        L.check(ctx.lib.hm_wgrad_unpack(ws.data_ptr(), kh, kw, P.c, Q.c, dst.data_ptr(), 1 if accumulate else 0,
                                        _stream()), "hm_wgrad_unpack")
    ctx.launches += 2


def tap_unroll(ctx, src, dst, C, KH, KW, oh, ow, sh, sw):
    """dst[n,h,w,(j*KW+i)*C+c] = src[n,h+oh+sh*j,w+ow+sw*i,c] over the STORED extents of both operands (zero outside)."""
    L.check(ctx.lib.hm_tap_unroll(src.hi.data_ptr(), _ptr(src.lo) if dst.lo is not None else None, src.n, src.h, src.w, C,
                                  src.cs, KH, KW, oh, ow, sh, sw, dst.hi.data_ptr(), _ptr(dst.lo), dst.h, dst.w, dst.cs,
                                  _stream()), "hm_tap_unroll")
    ctx.launches += 1


def tap_combine(ctx, T, KH, KW, C, oh, ow, sh, sw, bias, act, slope, out):
    """out[n,h,w,c] = act(bias[c] + sum_{j,i} T[n,h+oh+sh*j,w+ow+sw*i,(j*KW+i)*C+c]); T, out dense fp32 NHWC."""
    N, Ht, Wt, ldT = T.shape
    _, Ho, Wo, ldo = out.shape
    L.check(ctx.lib.hm_tap_combine(T.data_ptr(), N, Ht, Wt, ldT, KH, KW, C, oh, ow, sh, sw, _ptr(bias), act, slope,
                                   out.data_ptr(), Ho, Wo, ldo, _stream()), "hm_tap_combine")
    ctx.launches += 1


def encode_input(ctx, label, inst, image, mask_in, label_nc, g_op, d_op=None, v_op=None, d_no_imgcond=False,
                 d_mask=None, d_image_only=False):
    B, _, H, W = label.shape
    L.check(ctx.lib.hm_encode_input(label.data_ptr(), _ptr(inst), image.data_ptr(), mask_in.data_ptr(), B, H, W,
                                    label_nc, g_op.hi.data_ptr(), _ptr(g_op.lo), g_op.cs, g_op.border,
                                    _ptr(d_op.hi) if d_op else None, _ptr(d_op.lo) if d_op else None,
                                    d_op.cs if d_op else 0, _ptr(v_op.hi) if v_op else None,
                                    _ptr(v_op.lo) if v_op else None, v_op.cs if v_op else 0,
                                    (1 if d_no_imgcond else 0) | (2 if d_image_only else 0),
                                    _ptr(d_mask), _stream()),
            "hm_encode_input")
    ctx.launches += 1 + (1 if d_op else 0) + (1 if v_op else 0)


def in_stats(ctx, y, eps=1e-5):
    """y fp32 [N,H,W,C] -> (mean, rstd) fp32 [N,C]."""
    N, H, W, Cc = y.shape
    ws = ctx.ws("in", ctx.lib.hm_in_ws_bytes(N, H * W, Cc))
    mean = torch.empty(N, Cc, dtype=torch.float32, device=ctx.device)
    rstd = torch.empty(N, Cc, dtype=torch.float32, device=ctx.device)
    L.check(ctx.lib.hm_in_stats(y.data_ptr(), N, H * W, Cc, eps, ws.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                _stream()), "hm_in_stats")
    ctx.launches += 2
    return mean, rstd


def in_apply(ctx, y, mean, rstd, act, slope=0.2, skip=None, out32=None, out_op=None, reflect=True):
    N, H, W, Cc = y.shape
    border = out_op.border if out_op is not None else 0
    L.check(ctx.lib.hm_in_apply(y.data_ptr(), _ptr(mean), _ptr(rstd), _ptr(skip), N, H, W, Cc, act, slope, _ptr(out32),
                                _ptr(out_op.hi) if out_op else None, _ptr(out_op.lo) if out_op else None,
                                out_op.cs if out_op else 0, border, 1 if reflect else 0, _stream()), "hm_in_apply")
    ctx.launches += 1


def in_bwd(ctx, shape, act, slope=0.2, y=None, mean=None, rstd=None, z=None, mask_op=None, g1=None, g1_border=0,
           g1_ld=None, g1_coff=0, g2=None, tref=None, l1coef=0.0, out_op=None, out32=None):
    N, H, W, Cc = shape
    ws = ctx.ws("in", ctx.lib.hm_in_ws_bytes(N, H * W, Cc)) if mean is not None else None
    L.check(ctx.lib.hm_in_bwd(_ptr(y), _ptr(mean), _ptr(rstd), _ptr(z), _ptr(mask_op.hi) if mask_op else None,
                              mask_op.cs if mask_op else 0, _ptr(g1), g1_border, Cc if g1_ld is None else g1_ld, g1_coff,
                              _ptr(g2), _ptr(tref), float(l1coef), N, H, W, Cc, act, slope, _ptr(ws),
                              _ptr(out_op.hi) if out_op else None, _ptr(out_op.lo) if out_op else None,
                              out_op.cs if out_op else 0, _ptr(out32), _stream()), "hm_in_bwd")
    ctx.launches += 3 if mean is not None else 1


def fold_add(ctx, g_padded, border, base, out):
    N, H, W, Cc = out.shape
    L.check(ctx.lib.hm_fold_add(g_padded.data_ptr(), border, N, H, W, Cc, _ptr(base), out.data_ptr(), _stream()),
            "hm_fold_add")
    ctx.launches += 1


def avgpool3s2(ctx, x, out):
    L.check(ctx.lib.hm_avgpool3s2(x.hi.data_ptr(), _ptr(x.lo), x.n, x.ih, x.iw, x.cs, x.border, out.hi.data_ptr(),
                                  _ptr(out.lo), out.border, _stream()), "hm_avgpool3s2")
    ctx.launches += 1


def avgpool3s2_bwd(ctx, g_coarse, g_fine, c0, c1):
    N, Ho, Wo, ldc = g_coarse.shape
    _, H, W, ldf = g_fine.shape
    L.check(ctx.lib.hm_avgpool3s2_bwd(g_coarse.data_ptr(), N, Ho, Wo, ldc, g_fine.data_ptr(), H, W, ldf, c0, c1,
                                      _stream()), "hm_avgpool3s2_bwd")
    ctx.launches += 1


def maxpool2(ctx, x, out):
    L.check(ctx.lib.hm_maxpool2(x.hi.data_ptr(), _ptr(x.lo), x.n, x.h, x.w, x.cs, out.hi.data_ptr(), _ptr(out.lo),
                                _stream()), "hm_maxpool2")
    ctx.launches += 1


def maxpool2_bwd(ctx, g, a_op, dz, n=None):
    N, H, W, Cc = dz.shape
    L.check(ctx.lib.hm_maxpool2_bwd(g.data_ptr(), N, H, W, Cc, a_op.hi.data_ptr(), _ptr(a_op.lo), a_op.cs, dz.data_ptr(),
                                    _stream()), "hm_maxpool2_bwd")
    ctx.launches += 1


def l1_sum(ctx, a, b, coef, acc, slot):
    L.check(ctx.lib.hm_l1_sum(a.data_ptr(), b.data_ptr(), a.numel(), float(coef), acc.data_ptr() + 8 * slot, _stream()),
            "hm_l1_sum")
    ctx.launches += 1


def mse_sum(ctx, a, target, coef, acc, slot, bce=False):
    """bce=True: BCE(sigmoid(a), target) instead of (a - target)^2 (vanilla GAN, --no_lsgan)."""
    fn = ctx.lib.hm_bce_sum if bce else ctx.lib.hm_mse_sum
    L.check(fn(a.data_ptr(), a.numel(), float(target), float(coef), acc.data_ptr() + 8 * slot, _stream()),
            "hm_bce_sum" if bce else "hm_mse_sum")
    ctx.launches += 1


def mse_grad(ctx, y, target, scale, out_op, op_n0=0, bce=False):
    """out_op[op_n0 : op_n0 + y.shape[0]] = scale * (y - target); bce=True: scale/2 * d BCE(sigmoid(y), target)/dy (the
    callers pass 2 * weight / numel, the factor 2 being the square's derivative)."""
    P = y.numel() // y.shape[-1]
    off = op_n0 * out_op.h * out_op.w * out_op.cs * 2
    if bce:
        L.check(ctx.lib.hm_bce_grad(y.data_ptr(), P, y.shape[-1], float(target), 0.5 * float(scale), out_op.hi.data_ptr() + off,
                                    (out_op.lo.data_ptr() + off) if out_op.lo is not None else None, out_op.cs, _stream()),
                "hm_bce_grad")
        ctx.launches += 1
        return
    L.check(ctx.lib.hm_mse_grad(y.data_ptr(), P, y.shape[-1], float(target), float(scale), out_op.hi.data_ptr() + off,
                                (out_op.lo.data_ptr() + off) if out_op.lo is not None else None, out_op.cs, _stream()),
            "hm_mse_grad")
    ctx.launches += 1


def finish_fake(ctx, t, image, mask, use_gate, fake_nchw, d_op, d_coff, v_op, d_mask=None):
    B, H, W, _ = t.shape
    L.check(ctx.lib.hm_finish_fake(t.data_ptr(), _ptr(image), _ptr(mask), 1 if use_gate else 0, B, H, W, _ptr(fake_nchw),
                                   _ptr(d_op.hi) if d_op else None, _ptr(d_op.lo) if d_op else None,
                                   d_op.cs if d_op else 0, d_coff, _ptr(v_op.hi) if v_op else None,
                                   _ptr(v_op.lo) if v_op else None, v_op.cs if v_op else 0, _ptr(d_mask), _stream()),
            "hm_finish_fake")
    ctx.launches += 1


def fake_bwd(ctx, t, mask, use_gate, gD, gD_coff, gV, real_nchw, rec_coef, out_op, d_mask=None):
    B, H, W, _ = t.shape
    L.check(ctx.lib.hm_fake_bwd(t.data_ptr(), _ptr(mask), 1 if use_gate else 0, _ptr(gD),
                                gD.shape[-1] if gD is not None else 0, gD_coff, _ptr(gV),
                                gV.shape[-1] if gV is not None else 0, _ptr(real_nchw), float(rec_coef), B, H, W,
                                out_op.hi.data_ptr(), _ptr(out_op.lo), out_op.cs, _ptr(d_mask), _stream()), "hm_fake_bwd")
    ctx.launches += 1


def colsum_operand(ctx, op, out, accumulate=True, n0=0, n=None):
    n = op.n - n0 if n is None else n
    off = n0 * op.h * op.w * op.cs * 2
    P = n * op.h * op.w
    L.check(ctx.lib.hm_colsum_operand(op.hi.data_ptr() + off, (op.lo.data_ptr() + off) if op.lo is not None else None, P,
                                      op.c, op.cs, out.data_ptr(), 1 if accumulate else 0, _stream()),
            "hm_colsum_operand")
    ctx.launches += 1


def sn_power_iteration(ctx, layers_dev, n_layers, max_n, max_m, update_u=True):
    L.check(ctx.lib.hm_sn_power_iteration(layers_dev.data_ptr(), n_layers, max_n, max_m, 1 if update_u else 0, _stream()),
            "hm_sn_power_iteration")
    ctx.launches += 1


def sn_weight_grad(ctx, layers_dev, n_layers, max_n, max_m):
    L.check(ctx.lib.hm_sn_weight_grad(layers_dev.data_ptr(), n_layers, max_n, max_m, _stream()), "hm_sn_weight_grad")
    ctx.launches += 1


def adam_step(ctx, p, g, m, v, lr, beta1, beta2, eps, step, grad_scale=1.0):
    L.check(ctx.lib.hm_adam_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), lr, beta1, beta2, eps,
                                 step, grad_scale, _stream()), "hm_adam_step")
    ctx.launches += 1


def adam_step_dev(ctx, p, g, m, v, lr, beta1, beta2, eps, step_dev, grad_scale=1.0):
    """Adam step whose step count is read from the device int32 tensor `step_dev` (CUDA-graph replays)."""
    L.check(ctx.lib.hm_adam_step_dev(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), lr, beta1, beta2,
                                     eps, step_dev.data_ptr(), grad_scale, _stream()), "hm_adam_step_dev")
    ctx.launches += 1


# ---- GlobalTwoStreamGenerator glue (csrc/hm_twostream.cu) ------------------------------------------------------
def mask_maxpool(ctx, mask_nchw, f):
    B, _, H, W = mask_nchw.shape
    out = torch.empty(B, H // f, W // f, dtype=torch.float32, device=ctx.device)
    L.check(ctx.lib.hm_mask_maxpool(mask_nchw.data_ptr(), B, H, W, f, out.data_ptr(), _stream()), "hm_mask_maxpool")
    ctx.launches += 1
    return out


def mask_blend(ctx, a, b, m, out32=None, out_op=None):
    """(1-m)*a + m*b (a or b may be None: copy); fp32 NHWC inputs, optional dense / operand (reflect border) outputs."""
    ref = a if a is not None else b
    N, H, W, Cc = ref.shape
    L.check(ctx.lib.hm_mask_blend(_ptr(a), _ptr(b), _ptr(m), N, H, W, Cc, _ptr(out32),
                                  _ptr(out_op.hi) if out_op else None, _ptr(out_op.lo) if out_op else None,
                                  out_op.cs if out_op else 0, out_op.border if out_op else 0, _stream()), "hm_mask_blend")
    ctx.launches += 1


def mask_blend_bwd(ctx, g, m, da=None, db=None):
    Cc = g.shape[-1]
    L.check(ctx.lib.hm_mask_blend_bwd(g.data_ptr(), m.data_ptr(), g.numel() // Cc, Cc, _ptr(da), _ptr(db), _stream()),
            "hm_mask_blend_bwd")
    ctx.launches += 1


def pool_exchange(ctx, cur, pool, out, dec, b):
    """ImagePool.query for image b: cur / out are operands of the batch, pool the stored history, dec int32 [B, 2] on the
    device (action, slot)."""
    e = cur.h * cur.w * cur.cs
    L.check(ctx.lib.hm_pool_exchange(cur.hi.data_ptr(), _ptr(cur.lo), pool.hi.data_ptr(), _ptr(pool.lo), out.hi.data_ptr(),
                                     _ptr(out.lo), dec.data_ptr(), b, e, _stream()), "hm_pool_exchange")
    ctx.launches += 1


def mask_concat(ctx, a, b, m):
    """relu(cat((1 - m) * a, m * b)) as an operand [N,h,w,2C] (FeatureFusionBlock 'concat'); a, b dense fp32 NHWC."""
    N, h, w, Cc = a.shape
    out = Operand(ctx, N, h, w, 2 * Cc)
    L.check(ctx.lib.hm_mask_concat(a.data_ptr(), b.data_ptr(), m.data_ptr(), N * h * w, Cc, out.hi.data_ptr(), _ptr(out.lo),
                                   out.cs, _stream()), "hm_mask_concat")
    ctx.launches += 1
    return out


def mask_concat_bwd(ctx, g, m, a, b, da, db):
    N, h, w, Cc = a.shape
    L.check(ctx.lib.hm_mask_concat_bwd(g.data_ptr(), g.shape[-1], m.data_ptr(), a.data_ptr(), b.data_ptr(), N * h * w, Cc,
                                       da.data_ptr(), db.data_ptr(), _stream()), "hm_mask_concat_bwd")
    ctx.launches += 1


def concat_operands(ctx, a, b):
    """torch.cat((a, b), channel) of two border-free operands with equal pixel dims."""
    assert a.border == 0 and b.border == 0 and (a.n, a.h, a.w) == (b.n, b.h, b.w)
    out = Operand(ctx, a.n, a.h, a.w, a.c + b.c)
    both = a.lo is not None and b.lo is not None and out.lo is not None
    L.check(ctx.lib.hm_concat_operands(a.hi.data_ptr(), _ptr(a.lo) if both else None, a.cs, a.c, b.hi.data_ptr(),
                                       _ptr(b.lo) if both else None, b.cs, b.c, out.hi.data_ptr(), _ptr(out.lo), out.cs,
                                       a.n * a.h * a.w, _stream()), "hm_concat_operands")
    ctx.launches += 1
    return out


def cond_image_operand(ctx, image_nchw, mask_nchw, border):
    B, _, H, W = image_nchw.shape
    out = Operand(ctx, B, H, W, 3, border=border)
    L.check(ctx.lib.hm_cond_image_operand(image_nchw.data_ptr(), mask_nchw.data_ptr(), B, H, W, out.hi.data_ptr(),
                                          _ptr(out.lo), out.cs, border, _stream()), "hm_cond_image_operand")
    ctx.launches += 1
    return out


# ---- box2mask generator glue (csrc/hm_box2mask.cu) ---------------------------------------------------------------
def box2mask_encode(ctx, mask_ctx_in, mask_in, cls, label_nc):
    """cond operand [B,H,W,2*label_nc]: object box mask in its class channel | one-hot context (exact in bf16: hi only)."""
    B, _, H, W = mask_ctx_in.shape
    out = Operand(ctx, B, H, W, 2 * label_nc)
    L.check(ctx.lib.hm_box2mask_encode(mask_ctx_in.data_ptr(), mask_in.data_ptr(), cls.data_ptr(), B, H, W, label_nc,
                                       out.hi.data_ptr(), None, out.cs, _stream()), "hm_box2mask_encode")
    out.lo = None           # every value is 0 or 1: no lo product at all
    ctx.launches += 1
    return out


def bn_fold(ctx, mean, rstd, gamma, beta, N, running=None, count=0, repeat=1, momentum=0.1, eps=1e-5):
    """running = (running_mean, running_var, num_batches_tracked): also perform nn.BatchNorm2d's buffer update."""
    Cc = mean.numel()
    mo = torch.empty(N, Cc, dtype=torch.float32, device=ctx.device)
    ro = torch.empty(N, Cc, dtype=torch.float32, device=ctx.device)
    rm, rv, nbt = running if running is not None else (None, None, None)
    L.check(ctx.lib.hm_bn_fold(mean.data_ptr(), rstd.data_ptr(), _ptr(gamma), _ptr(beta), N, Cc, mo.data_ptr(),
                               ro.data_ptr(), _ptr(rm), _ptr(rv), _ptr(nbt), float(count), float(momentum), float(eps),
                               int(repeat), _stream()), "hm_bn_fold")
    ctx.launches += 1
    return mo, ro


def bn_stats(ctx, y, gamma, beta, running=None, repeat=1, momentum=0.1, eps=1e-5):
    """BatchNorm2d (training mode) statistics of y fp32 [N,H,W,C] + the per-(n, c) rows for in_apply + the running-buffer
    update, in two launches.  Returns (mean [1,C], rstd [1,C], mean_rows [N,C], rstd_rows [N,C])."""
    N, H, W, Cc = y.shape
    ws = ctx.ws("in", ctx.lib.hm_in_ws_bytes(1, N * H * W, Cc))
    mean = torch.empty(1, Cc, dtype=torch.float32, device=ctx.device)
    rstd = torch.empty(1, Cc, dtype=torch.float32, device=ctx.device)
    mo = torch.empty(N, Cc, dtype=torch.float32, device=ctx.device)
    ro = torch.empty(N, Cc, dtype=torch.float32, device=ctx.device)
    rm, rv, nbt = running if running is not None else (None, None, None)
    L.check(ctx.lib.hm_bn_stats(y.data_ptr(), N, H * W, Cc, float(eps), ws.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                _ptr(gamma), _ptr(beta), mo.data_ptr(), ro.data_ptr(), _ptr(rm), _ptr(rv), _ptr(nbt),
                                float(momentum), int(repeat), _stream()), "hm_bn_stats")
    ctx.launches += 2
    return mean, rstd, mo, ro


def upsample2_add(ctx, small, deep, out):
    N, h, w, Cc = small.shape
    assert tuple(deep.shape) == (N, 2 * h, 2 * w, Cc) == tuple(out.shape)
    L.check(ctx.lib.hm_upsample2_add(small.data_ptr(), deep.data_ptr(), N, h, w, Cc, out.data_ptr(), _stream()),
            "hm_upsample2_add")
    ctx.launches += 1


def box2mask_head(ctx, ctx_logit, obj_logit, label_map, mask_out, inst, use_gate, comb_logit, comb_logprob, obj_prob, acc,
                  no_comb=False, l1=False):
    N, H, W, Cc = ctx_logit.shape
    L.check(ctx.lib.hm_box2mask_head(ctx_logit.data_ptr(), obj_logit.data_ptr(), obj_logit.shape[-1], _ptr(label_map),
                                     _ptr(mask_out), _ptr(inst), N, H, W, Cc,
                                     (1 if use_gate else 0) | (2 if no_comb else 0) | (4 if l1 else 0),
                                     _ptr(comb_logit),
                                     _ptr(comb_logprob), _ptr(obj_prob), _ptr(acc), _stream()), "hm_box2mask_head")
    ctx.launches += 1


def bn_bwd(ctx, y, mean, rstd, gamma, beta, act, g1, g2=None, z=None, mask_op=None, out_op=None, out32=None, dgamma=None,
           dbeta=None, slope=0.2):
    """BatchNorm(+act) backward (batch statistics): gradient w.r.t. the BN output g1 (+ g2) -> gradient w.r.t. y."""
    N, H, W, Cc = y.shape
    ws = ctx.ws("in", ctx.lib.hm_in_ws_bytes(1, N * H * W, Cc))
    L.check(ctx.lib.hm_bn_bwd(y.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(), _ptr(beta), _ptr(z),
                              _ptr(mask_op.hi) if mask_op else None, mask_op.cs if mask_op else 0, _ptr(g1), _ptr(g2), N, H,
                              W, Cc, act, slope, ws.data_ptr(), _ptr(out_op.hi) if out_op else None,
                              _ptr(out_op.lo) if out_op else None, out_op.cs if out_op else 0, _ptr(out32), _ptr(dgamma),
                              _ptr(dbeta), _stream()), "hm_bn_bwd")
    ctx.launches += 3


def upsample2_bwd(ctx, g, dsmall):
    N, h, w, Cc = dsmall.shape
    L.check(ctx.lib.hm_upsample2_bwd(g.data_ptr(), N, h, w, Cc, dsmall.data_ptr(), _stream()), "hm_upsample2_bwd")
    ctx.launches += 1


def box2mask_head_bwd(ctx, ctx_logit, obj_logit, label_map, mask_out, inst, use_gate, acc, w_comb, w_obj, d_ctx, d_obj,
                      g_prob=None, no_comb=False, l1=False):
    N, H, W, Cc = ctx_logit.shape
    L.check(ctx.lib.hm_box2mask_head_bwd(ctx_logit.data_ptr(), obj_logit.data_ptr(), obj_logit.shape[-1],
                                         label_map.data_ptr(), _ptr(mask_out), inst.data_ptr(), N, H, W, Cc,
                                         (1 if use_gate else 0) | (2 if no_comb else 0) | (4 if l1 else 0), acc.data_ptr(), float(w_comb),
                                         float(w_obj), _ptr(g_prob),
                                         g_prob.shape[-1] if g_prob is not None else 0, d_ctx.hi.data_ptr(), _ptr(d_ctx.lo),
                                         d_ctx.cs, d_obj.hi.data_ptr(), _ptr(d_obj.lo), d_obj.cs, _stream()),
            "hm_box2mask_head_bwd")
    ctx.launches += 1


def box2mask_d_input(ctx, x, mask_ctx_in, mask_in, cls, mask_out, x_mask_power, label_nc):
    """Discriminator input operand [B,H,W,1 + 2*label_nc] = cat(x, cond) (gated by mask_out when given)."""
    B, _, H, W = mask_ctx_in.shape
    out = Operand(ctx, B, H, W, 1 + 2 * label_nc)
    L.check(ctx.lib.hm_box2mask_d_input(x.data_ptr(), mask_ctx_in.data_ptr(), mask_in.data_ptr(), cls.data_ptr(),
                                        _ptr(mask_out), x_mask_power, B, H, W, label_nc, out.hi.data_ptr(), _ptr(out.lo),
                                        out.cs, _stream()), "hm_box2mask_d_input")
    ctx.launches += 1
    return out

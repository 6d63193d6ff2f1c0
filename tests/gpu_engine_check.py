"""Diagnostic sweep of the tcgen05 conv engines against torch CPU fp32 convolutions.

Run on a GPU box:  python tests/gpu_engine_check.py [group ...]   (groups: fprop dgrad wgrad)
Prints one line per case (max-abs error / max-abs reference) and exits non-zero if any case fails.
The pytest parity tests (tests/test_conv_engine_gpu.py) reuse `run_case` from here.
"""
import sys
import os
import traceback

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import ctypes as C
import torch
import torch.nn.functional as F

from neurips18_hierchical_image_manipulation_b200 import _lib as L


def ru(v, m):
    return (v + m - 1) // m * m


def split_bf16(x):
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return hi, lo


def make_operand(x_nchw, split, dev):
    """fp32 NCHW (cpu) -> (Operand struct, keepalive tensors) as bf16 NHWC with channel stride ru(C,8)."""
    n, c, h, w = x_nchw.shape
    cs = ru(c, 8)
    buf = torch.zeros(n, h, w, cs, dtype=torch.float32)
    buf[..., :c] = x_nchw.permute(0, 2, 3, 1)
    if cs > c:
        buf[..., c:] = float("nan")  # must never be read: TMA clips at c
    hi, lo = split_bf16(buf)
    hi = hi.to(dev).contiguous()
    lo = lo.to(dev).contiguous() if split else None
    op = L.Operand(hi.data_ptr(), lo.data_ptr() if split else None, n, h, w, c, cs)
    return op, (hi, lo)


def pack_weight(lib, w_src, rows, k, taps, s_row, s_k, s_tap, split, dev):
    rows_pad, k_pad = lib.hm_rows_pad(rows), lib.hm_k_pad(k)
    src = w_src.to(dev).contiguous()
    hi = torch.empty(taps * rows_pad * k_pad, dtype=torch.bfloat16, device=dev)
    lo = torch.empty_like(hi) if split else None
    L.check(lib.hm_pack_weight(src.data_ptr(), rows, k, taps, s_row, s_k, s_tap, hi.data_ptr(),
                               lo.data_ptr() if split else None, None), "hm_pack_weight")
    return hi, lo, rows_pad, k_pad, src


def rel_err(got, ref):
    return float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


def case_fprop(lib, dev, n, cin, cout, h, w, k, stride, pad, reflect, split, act=0, bias=True, out16=False, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, k, k, generator=g) * 0.05
    b = torch.randn(cout, generator=g) if bias else None
    if reflect:
        xs = F.pad(x, (pad,) * 4, mode="reflect")
        zp = 0
    else:
        xs, zp = x, pad
    ref = F.conv2d(xs.double(), wt.double(), b.double() if bias else None, stride=stride, padding=zp).float()
    if act == 1:
        ref = F.relu(ref)
    elif act == 2:
        ref = F.leaky_relu(ref, 0.2)
    elif act == 3:
        ref = torch.tanh(ref)
    ho, wo = ref.shape[2], ref.shape[3]
    op, keep = make_operand(xs, split, dev)
    whi, wlo, rows_pad, k_pad, _ = pack_weight(lib, wt, cout, cin, k * k, cin * k * k, k * k, 1, split, dev)
    bd = b.to(dev) if bias else None
    out = torch.full((n, ho, wo, cout), float("nan"), device=dev)
    o32 = L.OutF32(out.data_ptr(), ho, wo, cout, 0, 0, 0)
    o16 = None
    if out16:
        cs = ru(cout, 8)
        ohi = torch.zeros(n, ho + 2, wo + 2, cs, dtype=torch.bfloat16, device=dev)
        olo = torch.zeros_like(ohi)
        o16 = L.OutBF16(ohi.data_ptr(), olo.data_ptr(), ho + 2, wo + 2, cs, 1, 1, 0)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    rc = lib.hm_conv_fprop(C.byref(op), whi.data_ptr(), wlo.data_ptr() if split else None, k_pad, rows_pad,
                           bd.data_ptr() if bias else None, k, k, stride, zp, ho, wo, cout, act, 0.2,
                           C.byref(o32), C.byref(o16) if out16 else None, err.data_ptr(), None)
    L.check(rc, "hm_conv_fprop")
    torch.cuda.synchronize()
    if int(err.item()) != 0:
        raise RuntimeError("engine pipeline timeout, code %d" % int(err.item()))
    got = out.cpu().permute(0, 3, 1, 2)
    e = rel_err(got, ref)
    if out16:
        g16 = (ohi.float() + olo.float())[:, 1:-1, 1:-1, :cout].cpu().permute(0, 3, 1, 2)
        e = max(e, rel_err(g16, ref))
        border = float(ohi[:, 0].abs().max())
        if border != 0:
            raise RuntimeError("bf16 output wrote outside its interior")
    return e


def case_dgrad(lib, dev, n, cin, cout, h, w, k, stride, pad, split, transposed_fwd=False, seed=0):
    """dgrad of conv(cin->cout, k, stride, zero pad) -- or, equivalently, ConvTranspose2d forward."""
    g = torch.Generator().manual_seed(seed)
    if transposed_fwd:
        # ConvTranspose2d(cout_t=cin.., ) forward: x has `cout` channels (the conv's output side)
        x = torch.randn(n, cout, h, w, generator=g)
        wt = torch.randn(cout, cin, k, k, generator=g) * 0.05  # IOHW: [in=cout][out=cin]
        ref = F.conv_transpose2d(x.double(), wt.double(), None, stride=stride, padding=pad,
                                 output_padding=stride - 1).float()
        dy = x
    else:
        xin = torch.randn(n, cin, h, w, generator=g, dtype=torch.float64, requires_grad=True)
        wt = torch.randn(cout, cin, k, k, generator=g) * 0.05
        y = F.conv2d(xin, wt.double(), None, stride=stride, padding=pad)
        dy = torch.randn(y.shape, generator=g)
        (ref,) = torch.autograd.grad(y, xin, dy.double())
        ref = ref.float()
    hout, wout = ref.shape[2], ref.shape[3]
    op, keep = make_operand(dy, split, dev)
    # rows = ci (result channels), k = co; both layouts index W[co][ci][kh][kw]
    whi, wlo, rows_pad, k_pad, _ = pack_weight(lib, wt, cin, cout, k * k, k * k, cin * k * k, 1, split, dev)
    out = torch.full((n, hout, wout, cin), float("nan"), device=dev)
    o32 = L.OutF32(out.data_ptr(), hout, wout, cin, 0, 0, 0)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    rc = lib.hm_conv_dgrad(C.byref(op), whi.data_ptr(), wlo.data_ptr() if split else None, k_pad, rows_pad, None,
                           k, k, stride, pad, hout, wout, cin, 0, 0.0, C.byref(o32), None, err.data_ptr(), None)
    L.check(rc, "hm_conv_dgrad")
    torch.cuda.synchronize()
    if int(err.item()) != 0:
        raise RuntimeError("engine pipeline timeout, code %d" % int(err.item()))
    got = out.cpu().permute(0, 3, 1, 2)
    return rel_err(got, ref)


def case_wgrad(lib, dev, n, cin, cout, h, w, k, stride, pad, split, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, k, k, generator=g, dtype=torch.float64, requires_grad=True)
    y = F.conv2d(x.double(), wt, None, stride=stride, padding=pad)
    dy = torch.randn(y.shape, generator=g)
    (ref,) = torch.autograd.grad(y, wt, dy.double())
    ref = ref.float()
    opP, k1 = make_operand(x, split, dev)
    opQ, k2 = make_operand(dy, split, dev)
    ws = torch.full((lib.hm_wgrad_ws_bytes(k, k, cin, cout) // 4,), float("nan"), device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    rc = lib.hm_conv_wgrad(C.byref(opP), C.byref(opQ), k, k, stride, pad, ws.data_ptr(), err.data_ptr(), None)
    L.check(rc, "hm_conv_wgrad")
    out = torch.full((cout, cin, k, k), float("nan"), device=dev)
    L.check(lib.hm_wgrad_unpack(ws.data_ptr(), k, k, cin, cout, out.data_ptr(), 0, None), "hm_wgrad_unpack")
    torch.cuda.synchronize()
    if int(err.item()) != 0:
        raise RuntimeError("engine pipeline timeout, code %d" % int(err.item()))
    return rel_err(out.cpu(), ref)


def tol(split):
    return 2e-5 if split else 2e-2


CASES = {
    "fprop": [
        # name, fn, kwargs
        ("res3x3 64->64 16x32 reflect", case_fprop, dict(n=2, cin=64, cout=64, h=16, w=32, k=3, stride=1, pad=1, reflect=True)),
        ("res3x3 128->256 16x32 reflect", case_fprop, dict(n=2, cin=128, cout=256, h=16, w=32, k=3, stride=1, pad=1, reflect=True)),
        ("vgg3x3 64->128 24x40 zero+relu", case_fprop, dict(n=1, cin=64, cout=128, h=24, w=40, k=3, stride=1, pad=1, reflect=False, act=1)),
        ("down3x3s2 64->128 32x64", case_fprop, dict(n=2, cin=64, cout=128, h=32, w=64, k=3, stride=2, pad=1, reflect=False)),
        ("stem7x7 38->64 32x64 reflect", case_fprop, dict(n=1, cin=38, cout=64, h=32, w=64, k=7, stride=1, pad=3, reflect=True)),
        ("head7x7 64->3 32x64 tanh", case_fprop, dict(n=1, cin=64, cout=3, h=32, w=64, k=7, stride=1, pad=3, reflect=True, act=3)),
        ("D4x4s2 41->64 33x65 lrelu +bf16out", case_fprop, dict(n=2, cin=41, cout=64, h=33, w=65, k=4, stride=2, pad=2, reflect=False, act=2, out16=True)),
        ("D4x4s1 256->512 18x34", case_fprop, dict(n=1, cin=256, cout=512, h=18, w=34, k=4, stride=1, pad=2, reflect=False)),
        ("D4x4s1 512->1 19x35", case_fprop, dict(n=1, cin=512, cout=1, h=19, w=35, k=4, stride=1, pad=2, reflect=False)),
        ("res3x3 1024->1024 8x16 nobias", case_fprop, dict(n=1, cin=1024, cout=1024, h=8, w=16, k=3, stride=1, pad=1, reflect=True, bias=False)),
        ("wide 64->64 4x300", case_fprop, dict(n=1, cin=64, cout=64, h=4, w=300, k=3, stride=1, pad=1, reflect=False)),
        # wide images: served by the row-streaming engine (hm_engine_rows.cuh)
        ("rows stem7x7 38->64 9x300 reflect", case_fprop, dict(n=2, cin=38, cout=64, h=9, w=300, k=7, stride=1, pad=3, reflect=True)),
        ("rows head7x7 64->3 9x300 tanh", case_fprop, dict(n=1, cin=64, cout=3, h=9, w=300, k=7, stride=1, pad=3, reflect=True, act=3)),
        ("rows vgg3x3 3->64 6x260 relu +bf16out", case_fprop, dict(n=2, cin=3, cout=64, h=6, w=260, k=3, stride=1, pad=1, reflect=False, act=1, out16=True)),
        ("rows 3x3 128->128 5x256", case_fprop, dict(n=1, cin=128, cout=128, h=5, w=256, k=3, stride=1, pad=1, reflect=False)),
        ("rows D4x4s1 64->32 7x131", case_fprop, dict(n=1, cin=64, cout=32, h=7, w=131, k=4, stride=1, pad=2, reflect=False)),
    ],
    "dgrad": [
        ("dgrad 3x3s1p1 64<-128 16x32", case_dgrad, dict(n=2, cin=64, cout=128, h=16, w=32, k=3, stride=1, pad=1)),
        ("dgrad 3x3s1p0 (reflect space) 64<-64 18x34", case_dgrad, dict(n=1, cin=64, cout=64, h=18, w=34, k=3, stride=1, pad=0)),
        ("dgrad 3x3s2p1 64<-128 32x64", case_dgrad, dict(n=2, cin=64, cout=128, h=32, w=64, k=3, stride=2, pad=1)),
        ("dgrad 4x4s2p2 41<-64 33x65", case_dgrad, dict(n=1, cin=41, cout=64, h=33, w=65, k=4, stride=2, pad=2)),
        ("dgrad 4x4s1p2 256<-512 18x34", case_dgrad, dict(n=1, cin=256, cout=512, h=18, w=34, k=4, stride=1, pad=2)),
        ("convT fwd 3x3s2 128->64 16x32", case_dgrad, dict(n=2, cin=64, cout=128, h=16, w=32, k=3, stride=2, pad=1, transposed_fwd=True)),
        ("dgrad 7x7s1p0 64<-3 38x70", case_dgrad, dict(n=1, cin=64, cout=3, h=38, w=70, k=7, stride=1, pad=0)),
        ("rows dgrad 7x7s1p0 64<-3 14x262", case_dgrad, dict(n=1, cin=64, cout=3, h=14, w=262, k=7, stride=1, pad=0)),
        ("rows dgrad 3x3s1p1 64<-128 6x256", case_dgrad, dict(n=2, cin=64, cout=128, h=6, w=256, k=3, stride=1, pad=1)),
    ],
    "wgrad": [
        ("wgrad 3x3s1p1 64x64 16x32", case_wgrad, dict(n=2, cin=64, cout=64, h=16, w=32, k=3, stride=1, pad=1)),
        ("wgrad 3x3s1p0 128x256 18x34", case_wgrad, dict(n=2, cin=128, cout=256, h=18, w=34, k=3, stride=1, pad=0)),
        ("wgrad 3x3s2p1 64x128 32x64", case_wgrad, dict(n=2, cin=64, cout=128, h=32, w=64, k=3, stride=2, pad=1)),
        ("wgrad 4x4s2p2 41x64 33x65", case_wgrad, dict(n=1, cin=41, cout=64, h=33, w=65, k=4, stride=2, pad=2)),
        ("wgrad 7x7s1p0 38x64 38x70", case_wgrad, dict(n=1, cin=38, cout=64, h=38, w=70, k=7, stride=1, pad=0)),
        ("wgrad 7x7s1p0 64x3 38x70", case_wgrad, dict(n=1, cin=64, cout=3, h=38, w=70, k=7, stride=1, pad=0)),
        ("wgrad 3x3s1p0 512x512 10x18", case_wgrad, dict(n=2, cin=512, cout=512, h=10, w=18, k=3, stride=1, pad=0)),
        ("wgrad 3x3s2p1 512x1024 16x32 n1", case_wgrad, dict(n=1, cin=512, cout=1024, h=16, w=32, k=3, stride=2, pad=1)),
        ("wgrad 3x3s1p0 1024x1024 10x18 n1", case_wgrad, dict(n=1, cin=1024, cout=1024, h=10, w=18, k=3, stride=1, pad=0)),
        ("wgrad 3x3s2p1 256x512 32x64 n1", case_wgrad, dict(n=1, cin=256, cout=512, h=32, w=64, k=3, stride=2, pad=1)),
        # CTA-pair MN engine with ragged M (9 units) and N (5 units) tile counts
        ("wgrad 3x3s1p1 64x320 12x20", case_wgrad, dict(n=2, cin=64, cout=320, h=12, w=20, k=3, stride=1, pad=1)),
        ("wgrad 4x4s1p2 256x512 9x17", case_wgrad, dict(n=2, cin=256, cout=512, h=9, w=17, k=4, stride=1, pad=2)),
        # wide base space: row-streaming weight-gradient engine (hm_engine_mnrows.cuh)
        ("rows wgrad 7x7s1p0 38x64 14x262", case_wgrad, dict(n=2, cin=38, cout=64, h=14, w=262, k=7, stride=1, pad=0)),
        ("rows wgrad 7x7s1p0 64x3 14x262", case_wgrad, dict(n=1, cin=64, cout=3, h=14, w=262, k=7, stride=1, pad=0)),
        ("rows wgrad 3x3s1p1 64x64 9x256", case_wgrad, dict(n=2, cin=64, cout=64, h=9, w=256, k=3, stride=1, pad=1)),
        ("rows wgrad 3x3s1p1 128x192 6x130", case_wgrad, dict(n=1, cin=128, cout=192, h=6, w=130, k=3, stride=1, pad=1)),
        ("rows wgrad 4x4s1p2 64x128 7x131", case_wgrad, dict(n=1, cin=64, cout=128, h=7, w=131, k=4, stride=1, pad=2)),
    ],
}


def perf_fprop(lib, dev, n, cin, cout, h, w, k, stride, pad, split, iters=20):
    """Device-time one fprop shape (bf16 operands already resident); returns (ms, TFLOP/s)."""
    hs, ws_ = h + 2 * pad, w + 2 * pad
    cs = ru(cin, 8)
    xhi = (torch.randn(n, hs, ws_, cs, device=dev) * 0.5).to(torch.bfloat16)
    xlo = (torch.randn(n, hs, ws_, cs, device=dev) * 0.002).to(torch.bfloat16) if split else None
    op = L.Operand(xhi.data_ptr(), xlo.data_ptr() if split else None, n, hs, ws_, cin, cs)
    wt = torch.randn(cout, cin, k, k) * 0.02
    whi, wlo, rows_pad, k_pad, _ = pack_weight(lib, wt, cout, cin, k * k, cin * k * k, k * k, 1, split, dev)
    ho, wo = (hs - k) // stride + 1, (ws_ - k) // stride + 1
    out = torch.empty(n, ho, wo, cout, device=dev)
    o32 = L.OutF32(out.data_ptr(), ho, wo, cout, 0, 0, 0)
    err = torch.zeros(1, dtype=torch.int32, device=dev)

    def call():
        L.check(lib.hm_conv_fprop(C.byref(op), whi.data_ptr(), wlo.data_ptr() if split else None, k_pad, rows_pad,
                                  None, k, k, stride, 0, ho, wo, cout, 0, 0.0, C.byref(o32), None, err.data_ptr(),
                                  None), "hm_conv_fprop")
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        call()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    flops = 2.0 * n * ho * wo * cout * cin * k * k
    return ms, flops / ms / 1e9


def run_case(lib, dev, fn, kw, split):
    return fn(lib, dev, split=split, **kw)


def main(argv):
    groups = argv or (list(CASES) + ["perf"])
    lib = L.load()
    dev = torch.device("cuda:0")
    print(lib.hm_version().decode(), torch.cuda.get_device_name(0), flush=True)
    bad = 0
    for grp in groups:
        if grp == "perf":
            for name, kw in [("stem7x7 38->64 512x1024 B1", dict(n=1, cin=38, cout=64, h=512, w=1024, k=7, stride=1, pad=3)),
                             ("head7x7 64->3 512x1024 B1", dict(n=1, cin=64, cout=3, h=512, w=1024, k=7, stride=1, pad=3)),
                             ("res3x3 1024->1024 32x64 B4", dict(n=4, cin=1024, cout=1024, h=32, w=64, k=3, stride=1, pad=1)),
                             ("vgg3x3 64->64 512x1024 B1", dict(n=1, cin=64, cout=64, h=512, w=1024, k=3, stride=1, pad=1)),
                             ("vgg3x3 256->256 128x256 B4", dict(n=4, cin=256, cout=256, h=128, w=256, k=3, stride=1, pad=1))]:
                for split in (False, True):
                    try:
                        ms, tf = perf_fprop(lib, dev, split=split, **kw)
                        print("perf   %-44s %s %.3f ms  %.1f TFLOP/s (useful flops)" % (name, "bf16x3" if split else "bf16  ", ms, tf), flush=True)
                    except Exception as ex:  # noqa
                        print("perf   %-44s EXC %s" % (name, ex), flush=True)
            continue
        for name, fn, kw in CASES[grp]:
            for split in (True, False):
                tag = "%-6s %-44s %s" % (grp, name, "bf16x3" if split else "bf16  ")
                try:
                    e = run_case(lib, dev, fn, kw, split)
                    ok = e < tol(split)
                    print("%s rel_err=%.3e %s" % (tag, e, "ok" if ok else "FAIL"), flush=True)
                    bad += 0 if ok else 1
                except Exception as ex:  # noqa
                    bad += 1
                    print("%s EXC %s" % (tag, ex), flush=True)
                    traceback.print_exc()
                    try:
                        torch.cuda.synchronize()
                    except Exception as ex2:  # context is dead: stop this group
                        print("CUDA context lost: %s" % ex2, flush=True)
                        return 2
    print("failures: %d" % bad, flush=True)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))

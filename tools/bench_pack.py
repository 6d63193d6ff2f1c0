"""Micro-benchmark: hm_pack_weight_pair vs the two hm_pack_weight calls it replaces (CUDA events, L2 flushed)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from neurips18_hierchical_image_manipulation_b200 import _lib as L     # noqa: E402
from neurips18_hierchical_image_manipulation_b200 import ops           # noqa: E402
from neurips18_hierchical_image_manipulation_b200.ops import PackedWeight  # noqa: E402

dev = torch.device("cuda", 0)
ctx = ops.Ctx(dev, split=True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, it=10):
    ts = []
    for _ in range(it):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e3


for A, B, k in ((1024, 1024, 3), (512, 256, 4), (512, 512, 3), (256, 128, 3), (64, 38, 7)):
    kk = k * k
    w = torch.randn(A, B, k, k, device=dev)
    p1, p2 = PackedWeight(ctx, A, B, kk), PackedWeight(ctx, B, A, kk, grad=True)
    t1 = timed(lambda: p1.pack(ctx, w, B * kk, kk, 1))
    t2 = timed(lambda: p2.pack(ctx, w, kk, B * kk, 1))
    tp = timed(lambda: L.check(ctx.lib.hm_pack_weight_pair(w.data_ptr(), A, B, kk, p1.hi.data_ptr(), ops._ptr(p1.lo),
                                                           p2.hi.data_ptr(), ops._ptr(p2.lo), ops._stream()), "pair"))
    nbytes = A * B * kk * 4
    print("%5d x %5d x %dx%d: fwd-role %.1f us, dgrad-role %.1f us, pair %.1f us  (pair: %.2f TB/s of 3x%d MB)"
          % (A, B, k, k, t1, t2, tp, 3 * nbytes / tp / 1e6, nbytes >> 20))

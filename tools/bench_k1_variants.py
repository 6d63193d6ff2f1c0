"""Time the CTA-pair K-engine on K1 and a few other pair-engine shapes of the step, with the stream-K tail
(scratch registered) and with whole tiles (scratch unregistered).  usage: python tools/bench_k1_variants.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from neurips18_hierchical_image_manipulation_b200 import ops
from neurips18_hierchical_image_manipulation_b200.networks import ConvP, FlatParams

SHAPES = [  # name, n, h, w, cin, cout, k, pad(border for reflect -> zero pad here)
    ("K1 res 1024->1024 32x64 B4", 4, 32, 64, 1024, 1024, 3, 1),
    ("vgg5 512->512 32x64 B8", 8, 32, 64, 512, 512, 3, 1),
    ("vgg4 512->512 64x128 B8", 8, 64, 128, 512, 512, 3, 1),
    ("vgg3 256->256 128x256 B8", 8, 128, 256, 256, 256, 3, 1),
    ("D l3 256->512 4x4 s1 65x129 B8", 8, 65, 129, 256, 512, 4, 2),
    ("D l3 256->512 4x4 s1 33x65 B8", 8, 33, 65, 256, 512, 4, 2),
    ("G down 512->1024 ... as s1 32x64 B4", 4, 32, 64, 512, 1024, 3, 1),
]


def time_conv(ctx, split, n, h, w, cin, cout, k, pad, iters=30):
    ctx.split = ctx.split_bwd = split
    fp = FlatParams(ctx.device)
    conv = ConvP(ctx, fp, "c", cin, cout, k, 1, pad)
    fp.materialize()
    conv.init_reference(torch.Generator().manual_seed(0))
    x = ops.Operand(ctx, n, h, w, cin, zero=True)
    x.hi.normal_(0, 0.5)
    if x.lo is not None:
        x.lo.normal_(0, 0.002)
    ho, wo = conv.out_hw(h, w, pad)
    y = torch.empty(n, ho, wo, cout, device=ctx.device)
    for _ in range(3):
        conv.forward(x, pad, out32=y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        conv.forward(x, pad, out32=y)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return ms, 2.0 * n * ho * wo * cout * cin * k * k / ms / 1e9


def main():
    ctx = ops.Ctx("cuda:0", split=True)
    scratch = ops._SCRATCH[0]
    ctx.lib.hm_set_streamk(1)
    for name, *shape in SHAPES:
        row = []
        for split in (True, False):
            for sk in (True, False):
                if sk:
                    ctx.lib.hm_set_scratch(scratch.data_ptr(), scratch.numel())
                else:
                    ctx.lib.hm_set_scratch(None, 0)
                ms, tf = time_conv(ctx, split, *shape)
                row.append("%s %s %.4f ms %6.0f TF" % ("x3" if split else "x1", "streamK" if sk else "tiles  ", ms, tf))
        ctx.lib.hm_set_scratch(scratch.data_ptr(), scratch.numel())
        print("%-40s | %s" % (name, " | ".join(row)), flush=True)
    ctx.check_pipeline()


if __name__ == "__main__":
    main()

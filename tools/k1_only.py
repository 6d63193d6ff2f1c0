"""Launch only the dominant kernel (res-block conv 3x3 1024->1024 @32x64 x4 = hm_kgemm_kernel<256>) a few times,
for `ncu --set full` captures.  usage: python tools/k1_only.py [precision] [iters]"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from neurips18_hierchical_image_manipulation_b200 import ops
from neurips18_hierchical_image_manipulation_b200.networks import ConvP, FlatParams

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
ctx = ops.Ctx("cuda:0", split=(prec == "bf16x3"))
fp = FlatParams(ctx.device)
conv = ConvP(ctx, fp, "k1", 1024, 1024, 3, 1, 0)
fp.materialize()
conv.init_reference(torch.Generator().manual_seed(0))
x = ops.Operand(ctx, 4, 32, 64, 1024, border=1, zero=True)
x.hi.normal_(0, 0.5)
if x.lo is not None:
    x.lo.normal_(0, 0.002)
y = torch.empty(4, 32, 64, 1024, device=ctx.device)
for _ in range(iters):
    conv.forward(x, 0, out32=y)
torch.cuda.synchronize()
ctx.check_pipeline()
print("ok", float(y.abs().mean()))

"""Launch the full-resolution 7x7 stem / head convolutions alone (for ncu).  usage: rows_only.py [precision] [which]"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from neurips18_hierchical_image_manipulation_b200 import ops
from neurips18_hierchical_image_manipulation_b200.networks import ConvP, FlatParams

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
which = sys.argv[2] if len(sys.argv) > 2 else "stem"
cin, cout = (38, 64) if which == "stem" else (64, 3)
ctx = ops.Ctx("cuda:0", split=(prec == "bf16x3"))
fp = FlatParams(ctx.device)
conv = ConvP(ctx, fp, "c", cin, cout, 7, 1, 0)
fp.materialize()
conv.init_reference(torch.Generator().manual_seed(0))
x = ops.Operand(ctx, 1, 512, 1024, cin, border=3, zero=True)
x.hi.normal_(0, 0.5)
if x.lo is not None:
    x.lo.normal_(0, 0.002)
y = torch.empty(1, 512, 1024, cout, device=ctx.device)
for _ in range(3):
    conv.forward(x, 0, out32=y)
torch.cuda.synchronize()
ctx.check_pipeline()
print("ok")

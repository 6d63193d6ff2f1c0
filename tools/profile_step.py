"""Run W warm-up + K fused training steps of BASELINE config #2 (for ncu launch lists / captures).
usage: python tools/profile_step.py [precision] [steps] [warmup] [batch]"""
import contextlib
import io
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from neurips18_hierchical_image_manipulation_b200.models import Options, create_model
from neurips18_hierchical_image_manipulation_b200.synthetic import synthetic_batch

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
warmup = int(sys.argv[3]) if len(sys.argv) > 3 else 1
B = int(sys.argv[4]) if len(sys.argv) > 4 else 4
opt = Options(vgg_weights="random", label_nc=35, no_instance=True, netG="global", ngf=64, n_downsample_global=4, n_blocks_global=9, num_D=3,
              gpu_ids=[0], precision=prec, name="prof")
with contextlib.redirect_stdout(io.StringIO()):
    m = create_model(opt).module
batch = {k: v.cuda() for k, v in synthetic_batch(B, 512, 1024, 35).items()}
kw = dict(label=batch["label"], inst=batch["inst"], image=batch["image"], feat=None, mask_in=batch["mask_in"],
          mask_out=batch["mask_out"])
for _ in range(warmup):
    m.optimize_parameters(**kw)
torch.cuda.synchronize()
l0 = m.ctx.launches
torch.cuda.cudart().cudaProfilerStart()
for _ in range(steps):
    ls = m.optimize_parameters(**kw)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("launches/step", (m.ctx.launches - l0) // steps, "losses", ls.tolist())

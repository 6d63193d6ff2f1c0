"""Step time of the configuration the reference's authors actually trained (scripts/train_mask2image_city.sh):
--netG global_twostream --which_encoder ctx_label --use_skip --use_output_gate --no_imgCond --mask_gan_input
--no_instance, 256x256 crops, batch 8, num_D 2, ngf 64, 4 downsamplings, 9 blocks.  Not the BASELINE metric; reported in
DESIGN.md next to it.  usage: python tools/bench_shipped.py [precision] [steps]"""
import contextlib
import io
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from neurips18_hierchical_image_manipulation_b200.models import Options, create_model
from neurips18_hierchical_image_manipulation_b200.synthetic import synthetic_batch

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
opt = Options(vgg_weights="random", label_nc=35, no_instance=True, netG="global_twostream", which_encoder="ctx_label", use_skip=True,
              use_output_gate=True, no_imgCond=True, mask_gan_input=True, ngf=64, n_downsample_global=4, n_blocks_global=9,
              num_D=2, n_layers_D=3, gpu_ids=[0], precision=prec, name="shipped")
with contextlib.redirect_stdout(io.StringIO()):
    m = create_model(opt).module
batch = {k: v.cuda() for k, v in synthetic_batch(8, 256, 256, 35).items()}
kw = dict(label=batch["label"], inst=batch["inst"], image=batch["image"], feat=None, mask_in=batch["mask_in"],
          mask_out=batch["mask_out"])
for _ in range(4):
    m.optimize_parameters(**kw)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    ls = m.optimize_parameters(**kw)
e1.record()
torch.cuda.synchronize()
m.ctx.check_pipeline()
ms = e0.elapsed_time(e1) / steps
print(json.dumps(dict(config="shipped script (global_twostream ctx_label skip gate no_imgCond mask_gan_input, 256x256 x8)",
                      precision=prec, ms_per_step=ms, images_per_sec=8 / (ms / 1e3), graph=isinstance(m._graph, dict),
                      losses=[float(x) for x in ls])))

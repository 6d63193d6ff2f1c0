"""Run W warm-up + K eager training iterations of BASELINE config #5 (box2mask, shipped flag set) for ncu launch lists.
usage: python tools/profile_box2mask.py [steps] [warmup]"""
import contextlib
import io
import os
import sys

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import bench                                                                            # noqa: E402
from neurips18_hierchical_image_manipulation_b200.models import Options, create_model   # noqa: E402
from neurips18_hierchical_image_manipulation_b200.synthetic import box2mask_batch       # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
warmup = int(sys.argv[2]) if len(sys.argv) > 2 else 1
cfg = bench.CONFIGS["5"]
with contextlib.redirect_stdout(io.StringIO()):
    m = create_model(Options(gpu_ids=[0], precision="bf16x3", name="prof5", cuda_graph=False, **cfg["opt"]))
d = {k: v.cuda() for k, v in box2mask_batch(cfg["per_gpu_batch"], cfg["H"], 35, 77).items()}
step = lambda: m.forward(d["label_map"], None, d["mask_ctx_in"], None, d["mask_out"], d["mask_obj_inst"], d["cls"], d["mask_in"])  # noqa: E731
for _ in range(warmup):
    step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(steps):
    ls, _ = step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("losses", [float(v) for v in ls])

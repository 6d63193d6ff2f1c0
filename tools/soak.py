"""Soak: N graph-replayed training steps of config #2 and of config #5 (fresh synthetic batch every step, host inputs),
checking finite losses, the engines' pipeline error flag and that the losses move.  usage: python tools/soak.py [steps]"""
import contextlib
import io
import os
import sys

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import bench                                                                            # noqa: E402
from neurips18_hierchical_image_manipulation_b200.models import Options, create_model   # noqa: E402
from neurips18_hierchical_image_manipulation_b200.synthetic import box2mask_batch, synthetic_batch   # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 300
with contextlib.redirect_stdout(io.StringIO()):
    m = create_model(Options(vgg_weights="random", gpu_ids=[0], name="soak", **bench.CONFIGS["2"]["opt"])).module
batches = [{k: v.pin_memory() for k, v in synthetic_batch(4, 512, 1024, 35, seed=1000 + i).items()} for i in range(4)]
hist = []
for i in range(N):
    b = batches[i % 4]
    ls = m.optimize_parameters(label=b["label"], inst=b["inst"], image=b["image"], feat=None, mask_in=b["mask_in"], mask_out=b["mask_out"])
    if i % 50 == 0 or i == N - 1:
        v = ls.tolist()
        assert all(x == x and abs(x) < 1e6 for x in v), (i, v)
        hist.append((i, [round(x, 4) for x in v]))
m.ctx.check_pipeline()
print("config #2:", N, "steps, graph", isinstance(m._graph, dict), hist[0], hist[-1])
del m
torch.cuda.empty_cache()
with contextlib.redirect_stdout(io.StringIO()):
    m5 = create_model(Options(gpu_ids=[0], name="soak5", **bench.CONFIGS["5"]["opt"]))
hist = []
for i in range(N):
    d = box2mask_batch(8, 256, 35, 500 + (i % 8))
    ls, _ = m5.forward(d["label_map"], None, d["mask_ctx_in"], None, d["mask_out"], d["mask_obj_inst"], d["cls"], d["mask_in"])
    if i % 50 == 0 or i == N - 1:
        v = [float(x) for x in ls]
        assert all(x == x and abs(x) < 1e6 for x in v), (i, v)
        hist.append((i, [round(x, 4) for x in v]))
m5.ctx.check_pipeline()
print("config #5:", N, "iterations, graph", isinstance(m5._graph, dict), hist[0], hist[-1])

"""N3 box2mask (BASELINE config #5), forward slice: the B200 executor of MaskTwoStreamConv_NET + the reconstruction
losses of TwoStreamAE_mask against (a) the golden vectors generated from the reference's OWN class
(tests/golden/box2mask_small.npz: scripts/train_box2mask_city.sh flag set at label_nc 6, 64x64, batch 3) and (b) the
oracle at config #5's real geometry (label_nc 35, 256x256, conv_dim 64, n_blocks 6, batch 2).  bf16x3, tolerance 1e-3."""
import contextlib
import io
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = torch.as_tensor(a).detach().float().cpu(), torch.as_tensor(b).detach().float().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _model(**kw):
    from neurips18_hierchical_image_manipulation_b200.models import Options, create_model
    base = dict(model="AE_maskgen_twostream", isTrain=False, gpu_ids=[0], precision="bf16x3", name="b2m", num_layers=3,
                conv_size=4, which_stream="obj_context", cond_in="ctx_obj", use_output_gate=True, num_resnetblocks=1,
                norm_layer="batch")
    base.update(kw)
    with contextlib.redirect_stdout(io.StringIO()):
        return create_model(Options(**base))


def test_box2mask_forward_and_losses_against_the_reference_class_golden(golden_dir):
    from oracle.weights import named_param
    z = np.load(os.path.join(golden_dir, "box2mask_small.npz"))
    m = _model(label_nc=6, output_nc=6, conv_dim=32, n_blocks=2)
    sd = {str(n): named_param(str(n), tuple(int(v) for v in str(s).split(";")))
          for n, s in zip(z["param_names"], z["param_shapes"])}
    assert set(sd) == set(m.fpG.params), (sorted(set(sd) ^ set(m.fpG.params))[:6])
    m.fpG.load_state_dict(sd)
    ins = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("in::")}
    losses, out = m.forward(ins["label_map"], None, ins["mask_ctx_in"], None, ins["mask_out"], ins["mask_obj_inst"], ins["cls"],
                            ins["mask_in"])
    torch.cuda.synchronize()
    m.ctx.check_pipeline()
    for k in ("comb_logit", "comb_prob", "obj_logit", "obj_prob"):
        e = rel(out[k], z[k])
        print("box2mask golden %s %.2e" % (k, e))
        assert e < 1e-3, (k, e)
    assert abs(float(losses[0]) - float(z["loss_comb"])) < 1e-3 * abs(float(z["loss_comb"]))
    assert abs(float(losses[1]) - float(z["loss_obj"])) < 1e-3 * abs(float(z["loss_obj"]))


def test_box2mask_forward_at_config5_geometry_against_oracle():
    from oracle import box2mask as B2
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
    import make_golden_box2mask as G
    torch.set_num_threads(os.cpu_count() or 1)
    m = _model(label_nc=35, output_nc=35, conv_dim=64, n_blocks=6)
    sd = {k: v.detach().cpu().clone() for k, v in m.fpG.params.items()}
    with torch.no_grad():        # non-trivial BatchNorm shifts (weights_init leaves them at 0)
        for k in sd:
            if k.endswith("bias") and sd[k].numel() >= 16:
                sd[k] = (torch.rand(sd[k].shape, generator=torch.Generator().manual_seed(len(k))) - 0.5) * 0.1
    m.fpG.load_state_dict(sd)
    d = G.synthetic(dict(label_nc=35, fineSize=256), 2, seed=3)
    losses, out = m.forward(d["label_map"], None, d["mask_ctx_in"], None, d["mask_out"], d["mask_obj_inst"], d["cls"],
                            d["mask_in"])
    torch.cuda.synchronize()
    m.ctx.check_pipeline()
    cond, _ = B2.encode_input(35, d["mask_ctx_in"], d["mask_in"], d["cls"])
    with torch.no_grad():
        comb_logit, comb_lp, obj_logit, obj_prob = B2.two_stream_forward(sd, cond, num_layers=3, n_blocks=6)
        l_comb = B2.mask_recon_loss(comb_lp, d["label_map"], d["mask_out"])
        l_obj = B2.obj_recon_loss(obj_prob, d["mask_out"], d["mask_obj_inst"])
    errs = dict(comb_logit=rel(out["comb_logit"], comb_logit), comb_prob=rel(out["comb_prob"], comb_lp),
                obj_logit=rel(out["obj_logit"], obj_logit), obj_prob=rel(out["obj_prob"], obj_prob),
                loss_comb=abs(float(losses[0]) - float(l_comb)) / float(l_comb),
                loss_obj=abs(float(losses[1]) - float(l_obj)) / float(l_obj))
    print("box2mask config #5 geometry:", {k: "%.2e" % v for k, v in errs.items()})
    assert max(errs.values()) < 1e-3, errs

"""Generate tests/golden/box2mask_small.npz from the REFERENCE'S OWN MaskTwoStreamConv_NET (models/MaskTwoStreamConv_NET.py,
the box2mask generator of BASELINE config #5) and its loss classes, run once in the build container:

    python oracle/make_golden_box2mask.py

The reference file is python-2 code; it is imported UNMODIFIED with three run-time injections only (nothing under
/root/reference is edited or copied):
  * builtins.xrange = range                       (layer_util.py:148,192,201,226 loop with xrange)
  * nn.Conv2d / nn.ConvTranspose2d / nn.BatchNorm2d / nn.InstanceNorm2d receive int() channel counts
                                                  (MaskTwoStreamConv_NET.py:141 computes `input_dim/2`, a float under py3)
  * `initialize()` is not called because it ends in dict.iteritems() (MaskContextAE_NET.py:66); the four builder
    methods it calls are called here directly, in the same order (MaskTwoStreamConv_NET.py:31-41);
  * the parameter-free `nn.ReLU()` instance behind the first conv (MaskTwoStreamConv_NET.py:73) is swapped for a module
    computing x.clamp(min=0): the next block's in-place ReLU (layer_util.py:142 `activation_fn` = nn.ReLU(True)) rewrites
    that tensor with identical values, which today's autograd rejects for a ReLU output (saved for backward) but not for
    a clamp output.  Values and gradients are unchanged.

NOTE on semantics the restatement must follow (oracle/box2mask.py): every Conv/DeconvResnetBlock starts with an IN-PLACE
ReLU on its input `x` while `residual = x` aliases the same tensor (layer_util.py:156-162,236-242), so the shortcut
branch sees relu(x), and the encoder features kept for the skip connections (MaskTwoStreamConv_NET.py:172-173) are
rectified in place by the following block before the decoder reads them.

Fixture: the flag set of scripts/train_box2mask_city.sh (which_stream obj_context, cond_in ctx_obj, conv_size 4,
num_layers 3, num_resnetblocks 1, norm_layer batch, n_blocks 6, use_output_gate, objReconLoss bce) at a reduced size
(label_nc = output_nc = 6, fineSize 64, batch 3), training mode (BatchNorm uses batch statistics).  Stored: inputs, the
parameter name -> shape manifest (values are oracle.weights.named_param(name, shape), loaded into the reference modules
like a checkpoint), the four network outputs, the two reconstruction losses (mask_losses.MaskReconLoss =
NLL with ignore index 255 outside the box; nn.BCELoss on the gated object mask, TwoStreamAE_mask.py:198-203) and the
gradients of `loss_recon_obj + rec_weight * loss_recon_comb` w.r.t. every parameter (full tensors up to 4096 elements,
sum / L1 norm / a seeded random projection beyond).  TEST INFRASTRUCTURE ONLY.
"""
import builtins
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
REF = os.environ.get("HM_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

CFG = dict(label_nc=6, output_nc=6, fineSize=64, num_layers=3, conv_dim=32, conv_size=4, embed_dim=1024, z_dim=512,
           norm_layer="batch", use_dropout=False, skip_start=1, skip_end=3, use_resnetblock=1, num_resnetblocks=1,
           fusion_type="add", first_conv_stride=1, first_conv_size=5, which_stream="obj_context", cond_in="ctx_obj",
           use_simpleRes=False, n_blocks=2)


def import_reference():
    builtins.xrange = range

    def int_channels(cls):
        class Wrapped(cls):
            def __init__(self, *a, **kw):
                a = tuple(int(v) if isinstance(v, float) else v for v in a)
                super().__init__(*a, **kw)
        Wrapped.__name__ = cls.__name__          # weights_init dispatches on the class NAME (layer_util.py:10-16)
        return Wrapped
    for name in ("Conv2d", "ConvTranspose2d", "BatchNorm2d", "InstanceNorm2d"):
        setattr(nn, name, int_channels(getattr(nn, name)))
    sys.path.insert(0, os.path.join(REF, "models"))
    sys.path.insert(0, REF)
    import MaskTwoStreamConv_NET as M
    import layer_util as LU
    import mask_losses as ML
    return M, LU, ML


def build(M, LU, cfg):
    opt = types.SimpleNamespace(**cfg)
    net = M.MaskTwoStreamConv_NET(opt)
    # MaskTwoStreamConv_NET.initialize(), :31-41, without the python-2-only tail
    net.conv_encoder_modules = net.get_conv_encoder()
    net.latent_encoder = net.get_latent_encoder()
    net.obj_conv_decoder_modules = net.get_conv_decoder(output_nc=1, skip_layers=None)
    net.obj_latent_decoder = net.get_latent_decoder()
    net.ctx_conv_decoder_modules = net.get_conv_decoder(output_nc=net.output_nc, skip_layers=net.skip_layers)
    net.ctx_latent_decoder = net.get_latent_decoder()
    net.params_dict = net.get_params_dict()

    class ClampReLU(nn.Module):
        def forward(self, x):
            return x.clamp(min=0)
    assert type(net.conv_encoder_modules[2]).__name__ == "ReLU" and not net.conv_encoder_modules[2].inplace
    net.conv_encoder_modules[2] = ClampReLU()
    return net


def synthetic(cfg, B, seed):
    """What data/ + TwoStreamAE_mask.encode_input (:127-151) hand to the network: a context label map with the box
    region wiped to class 0, the object's box mask placed in the object's class channel, the ground-truth label map,
    the box mask (mask_out) and the object instance mask."""
    g = torch.Generator().manual_seed(seed)
    nc, S = cfg["label_nc"], cfg["fineSize"]
    lab = torch.randint(1, nc, (B, 1, S // 8, S // 8), generator=g).float()
    label_map = torch.nn.functional.interpolate(lab, size=(S, S), mode="nearest")
    cls = torch.randint(1, nc - 1, (B, 1), generator=g)
    mask_out = torch.zeros(B, 1, S, S)      # the (margin-expanded) box: region to be generated
    mask_in = torch.zeros(B, 1, S, S)       # the object's bounding box
    inst = torch.zeros(B, 1, S, S)          # the object's instance mask (inside its box)
    for b in range(B):
        y0, x0 = int(torch.randint(4, S // 4, (1,), generator=g)), int(torch.randint(4, S // 4, (1,), generator=g))
        h, w = int(torch.randint(S // 4, S // 2, (1,), generator=g)), int(torch.randint(S // 4, S // 2, (1,), generator=g))
        mask_in[b, :, y0:y0 + h, x0:x0 + w] = 1
        mask_out[b, :, max(0, y0 - 4):y0 + h + 4, max(0, x0 - 4):x0 + w + 4] = 1
        yy, xx = torch.meshgrid(torch.arange(S), torch.arange(S), indexing="ij")
        ell = (((yy - (y0 + h / 2.0)) / (h / 2.0)) ** 2 + ((xx - (x0 + w / 2.0)) / (w / 2.0)) ** 2) <= 1.0
        inst[b, 0] = ell.float()
        label_map[b, 0][ell] = float(cls[b, 0])
    mask_ctx_in = label_map * (1 - mask_out)            # context with the box region set to class 0
    return dict(label_map=label_map, mask_ctx_in=mask_ctx_in, mask_out=mask_out, mask_in=mask_in, mask_obj_inst=inst,
                cls=cls)


def encode(cfg, d):
    """TwoStreamAE_mask.encode_input (:127-151) + construct_input_cond (:331-338, cond_in == 'ctx_obj') on CPU."""
    nc = cfg["label_nc"]
    B, _, S, _ = d["label_map"].shape
    ctx = torch.zeros(B, nc, S, S).scatter_(1, d["mask_ctx_in"].long(), 1.0)
    cls_onehot = torch.zeros(B, nc).scatter_(1, d["cls"].long(), 1.0)
    obj = torch.zeros(B, nc, S, S)
    for b in range(B):
        obj[b, int(d["cls"][b, 0])] = d["mask_in"][b, 0]
    return torch.cat((obj, ctx), 1), cls_onehot


def main():
    os.makedirs(OUT, exist_ok=True)
    M, LU, ML = import_reference()
    torch.manual_seed(5)
    net = build(M, LU, CFG)
    from oracle.weights import named_param
    for mk, mod in net.params_dict.items():
        mod.apply(LU.weights_init)
        # then overwrite every parameter with name-keyed deterministic values (loaded like a checkpoint), so that the
        # fixture only has to store the name -> shape manifest instead of ~50 MB of random weights
        with torch.no_grad():
            for k, p in mod.named_parameters():
                p.copy_(named_param(mk + "." + k, p.shape))
        mod.train()
    d = synthetic(CFG, 3, seed=17)
    cond, cls_onehot = encode(CFG, d)
    comb_logit, comb_prob, obj_logit, obj_prob = net.forward(cond, cls_onehot)
    # losses of TwoStreamAE_mask.forward (:188-203) with use_output_gate, objReconLoss == 'bce', rec_weight 1
    gt_label = d["label_map"].view(-1, d["label_map"].size(2), d["label_map"].size(3)).long()
    loss_comb = ML.MaskReconLoss()(comb_prob, gt_label, d["mask_out"])
    obj_gated = obj_prob * d["mask_out"]
    loss_obj = nn.BCELoss()(obj_gated, d["mask_obj_inst"])
    params, names = [], []
    for mk, mod in net.params_dict.items():
        for k, p in mod.named_parameters():
            names.append(mk + "." + k)
            params.append(p)
    grads = torch.autograd.grad(loss_obj + 1.0 * loss_comb, params, allow_unused=True)
    out = dict(comb_logit=comb_logit.detach().numpy(), comb_prob=comb_prob.detach().numpy(),
               obj_logit=obj_logit.detach().numpy(), obj_prob=obj_prob.detach().numpy(),
               loss_comb=float(loss_comb), loss_obj=float(loss_obj), cond=cond.numpy(), cls_onehot=cls_onehot.numpy())
    for k, v in d.items():
        out["in::" + k] = v.numpy()
    # parameters: manifest only (values = oracle.weights.named_param(name, shape))
    out["param_names"] = np.array(names)
    out["param_shapes"] = np.array([";".join(str(v) for v in p.shape) for p in params])
    # gradients: full tensors for small parameters, (sum, L1 norm, seeded random projection) for the big ones
    for n_, g_ in zip(names, grads):
        assert g_ is not None, n_
        if g_.numel() <= 4096:
            out["g::" + n_] = g_.numpy()
        else:
            r = named_param("proj::" + n_ + ".bias", g_.shape)
            out["gs::" + n_] = np.array([float(g_.sum()), float(g_.abs().sum()), float((g_ * r).sum())])
    np.savez_compressed(os.path.join(OUT, "box2mask_small.npz"), **out)
    print("wrote box2mask_small.npz:", len(names), "parameters;", "losses", float(loss_comb), float(loss_obj),
          "shapes", tuple(comb_logit.shape), tuple(obj_logit.shape))
    print("modules:", {k: type(v).__name__ for k, v in net.params_dict.items()})


if __name__ == "__main__":
    main()

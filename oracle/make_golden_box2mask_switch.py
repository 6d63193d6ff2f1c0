"""Generate tests/golden/box2mask_switch_small.npz from the REFERENCE'S OWN MaskTwoStreamConvSwitch_NET (the generator
`--no_comb` selects, models/TwoStreamAE_mask.py:29-32; what scripts/train_box2mask_city.sh trains) in TRAINING and in EVAL
mode, run once in the build container:   python oracle/make_golden_box2mask_switch.py

Same run-time injections as oracle/make_golden_box2mask.py (python-2 source imported unmodified).  Sequence stored:
  1. one training-mode forward on batch A (BatchNorm: batch statistics; the running buffers move by momentum 0.1) ->
     the four outputs, the two reconstruction losses, every running_mean / running_var afterwards;
  2. `module.eval()` on every entry of params_dict (MaskContextAE_NET.set_mode), forward on batch B -> the four outputs
     (BatchNorm normalises with the buffers of step 1).
Parameters are oracle.weights.named_param values (manifest only).  TEST INFRASTRUCTURE ONLY."""
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import make_golden_box2mask as G   # noqa: E402


def build_switch(cfg):
    G.import_reference()
    import MaskTwoStreamConvSwitch_NET as MS
    import layer_util as LU
    import mask_losses as ML
    opt = types.SimpleNamespace(add_dilated_layers=False, **cfg)
    net = MS.MaskTwoStreamConvSwitch_NET(opt)
    # MaskTwoStreamConvSwitch_NET.initialize() without its python-2-only tail (dict.iteritems)
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
    return net, LU, ML


def ade_case():
    """The flag set of scripts/train_box2mask_ade.sh that differs from the Cityscapes one: --norm_layer instance
    (InstanceNorm2d(affine=False) everywhere) and --add_dilated_layers (two DilatedResnetBlocks, dilation 2 and 4, in front
    of the latent encoder, MaskTwoStreamConvSwitch_NET.py:101-104).  Stored: the four outputs, both reconstruction losses
    and the gradient of their sum w.r.t. every parameter (full tensors up to 4096 elements, sum / L1 / projection beyond)."""
    from oracle.weights import named_param
    cfg = dict(G.CFG, norm_layer="instance")
    torch.manual_seed(7)
    G.import_reference()
    import MaskTwoStreamConvSwitch_NET as MS
    import mask_losses as ML
    opt = types.SimpleNamespace(add_dilated_layers=True, **cfg)
    net = MS.MaskTwoStreamConvSwitch_NET(opt)
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
    names, params = [], []
    for mk, mod in net.params_dict.items():
        with torch.no_grad():
            for k, p in mod.named_parameters():
                p.copy_(named_param(mk + "." + k, p.shape))
                names.append(mk + "." + k)
                params.append(p)
        mod.train()
    a = G.synthetic(cfg, 3, seed=47)
    cond, onehot = G.encode(cfg, a)
    lo, lp, oo, op = net.forward(cond, onehot)
    gt = a["label_map"].view(-1, a["label_map"].size(2), a["label_map"].size(3)).long()
    loss_comb = ML.MaskReconLoss()(lp, gt, a["mask_out"])
    loss_obj = nn.BCELoss()(op * a["mask_out"], a["mask_obj_inst"])
    grads = torch.autograd.grad(loss_obj + loss_comb, params, allow_unused=True)
    out = dict(param_names=np.array(names), param_shapes=np.array([";".join(str(v) for v in p.shape) for p in params]),
               comb_logit=lo.detach().numpy(), comb_prob=lp.detach().numpy(), obj_logit=oo.detach().numpy(),
               obj_prob=op.detach().numpy(), loss_comb=float(loss_comb), loss_obj=float(loss_obj))
    for k, v in a.items():
        out["in::" + k] = v.numpy()
    for n_, g_ in zip(names, grads):
        assert g_ is not None, n_
        if g_.numel() <= 4096:
            out["g::" + n_] = g_.numpy()
        else:
            r = named_param("proj::" + n_ + ".bias", g_.shape)
            out["gs::" + n_] = np.array([float(g_.sum()), float(g_.abs().sum()), float((g_ * r).sum())])
    np.savez_compressed(os.path.join(G.OUT, "box2mask_switch_ade_small.npz"), **out)
    print("wrote box2mask_switch_ade_small.npz:", len(names), "parameters; losses", float(loss_comb), float(loss_obj))
    print([n for n in names if n.startswith("latent_encoder.")][:8])


def main():
    if "ade" in sys.argv[1:]:
        return ade_case()
    from oracle.weights import named_param
    cfg = dict(G.CFG)
    torch.manual_seed(6)
    net, LU, ML = build_switch(cfg)
    names, shapes = [], []
    for mk, mod in net.params_dict.items():
        with torch.no_grad():
            for k, p in mod.named_parameters():
                p.copy_(named_param(mk + "." + k, p.shape))
                names.append(mk + "." + k)
                shapes.append(";".join(str(v) for v in p.shape))
        mod.train()
    a, b = G.synthetic(cfg, 3, seed=41), G.synthetic(cfg, 2, seed=43)
    out = dict(param_names=np.array(names), param_shapes=np.array(shapes))
    with torch.no_grad():
        cond, onehot = G.encode(cfg, a)
        lo, lp, oo, op = net.forward(cond, onehot)
        gt = a["label_map"].view(-1, a["label_map"].size(2), a["label_map"].size(3)).long()
        out.update(train_comb_logit=lo.numpy(), train_comb_prob=lp.numpy(), train_obj_logit=oo.numpy(), train_obj_prob=op.numpy(),
                   train_loss_comb=float(ML.MaskReconLoss()(lp, gt, a["mask_out"])),
                   train_loss_obj=float(nn.BCELoss()(op * a["mask_out"], a["mask_obj_inst"])))
        n_bn = 0
        for mk, mod in net.params_dict.items():
            for k, v in mod.state_dict().items():
                if k.endswith("running_mean") or k.endswith("running_var") or k.endswith("num_batches_tracked"):
                    out["buf::" + mk + "." + k] = v.numpy()
                    n_bn += 1
            mod.eval()                                   # MaskContextAE_NET.set_mode(eval_mode=True)
        cond, onehot = G.encode(cfg, b)
        lo, lp, oo, op = net.forward(cond, onehot)
        out.update(eval_comb_logit=lo.numpy(), eval_comb_prob=lp.numpy(), eval_obj_logit=oo.numpy(), eval_obj_prob=op.numpy())
    for k, v in a.items():
        out["a::" + k] = v.numpy()
    for k, v in b.items():
        out["b::" + k] = v.numpy()
    np.savez_compressed(os.path.join(G.OUT, "box2mask_switch_small.npz"), **out)
    print("wrote box2mask_switch_small.npz:", len(names), "parameters,", n_bn, "buffers; train losses",
          out["train_loss_comb"], out["train_loss_obj"])


if __name__ == "__main__":
    main()

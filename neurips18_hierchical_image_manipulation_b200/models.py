"""Host-side mirror of the reference's model layer for the mask2image hot path.

    create_model(opt)                       <- models/models.py:6-24
    Pix2PixHDModel_condImg                  <- models/pix2pixHD_condImg_model.py:22-327 (+ BaseModel, models/base_model.py)

Same names, argument meaning and error behaviour as the reference, so train_mask2image.py:39-131 / vis_mask2image.py:22-43
drive it with the same call sequence:

    model = create_model(opt)
    losses, generated = model(label=..., inst=..., image=..., feat=None, mask_in=..., mask_out=..., infer=...)
    loss_G.backward(); model.module.optimizer_G.step(); loss_D.backward(); model.module.optimizer_D.step()

All tensor work runs in libhm_b200.so (sm_100a); there is no torch.nn / cuDNN / CPU path behind this class.
`optimize_parameters()` (a no-op stub in the reference, models/base_model.py:33) is the fused fast path: one forward,
both backward passes, ONE gradient allreduce over [G | D] and both Adam steps.
"""
import os
import sys
from collections import OrderedDict

import numpy as np
import torch

from . import ops, parallel
from .networks import FlatParams, GlobalGenerator, MultiscaleDiscriminator, Vgg19, VGG19_CONVS, VGG19_SLICE_OF
from .ops import Ctx, Operand

LOSS_NAMES = ["G_GAN", "G_GAN_Feat", "G_VGG", "D_real", "D_fake"]  # pix2pixHD_condImg_model.py:118
VGG_WEIGHTS = [1.0 / 32, 1.0 / 16, 1.0 / 8, 1.0 / 4, 1.0]          # models/losses.py:57


class Options(object):
    """Flat option namespace with the reference's flag names and defaults for this path
    (options/mask2image_base_options.py:15-74, options/mask2image_train_options.py:9-46)."""

    def __init__(self, **kw):
        d = dict(name="label2city", gpu_ids=[0], checkpoints_dir="./checkpoints", model="pix2pixHD_condImg",
                 norm="instance", batchSize=1, label_nc=35, output_nc=3, resize_or_crop="scale_width", netG="global",
                 ngf=64, n_downsample_global=4, n_blocks_global=9, n_blocks_local=3, n_local_enhancers=1,
                 niter_fix_global=0, which_encoder="ctx", use_output_gate=False, use_skip=False, feat_fusion="early_add",
                 no_instance=False, instance_feat=False, label_feat=False, feat_num=3, load_features=False,
                 isTrain=True, continue_train=False, load_pretrain="", which_epoch="latest", niter=100, niter_decay=100,
                 beta1=0.5, lr=0.0002, num_D=2, n_layers_D=3, ndf=64, lambda_feat=10.0, lambda_rec=0.0,
                 no_ganFeat_loss=False, no_vgg_loss=False, no_lsgan=False, pool_size=0, no_imgCond=False,
                 mask_gan_input=False, use_soft_mask=False, no_gan=False,
                 # extensions of this implementation (not reference flags)
                 precision="bf16x3",      # "bf16x3": fp32-parity mode; "mixed": bf16x3 forward + bf16 gradient GEMMs;
                                          # "bf16": single-product tensor-core mode
                 vgg_weights=None,        # path of a torchvision vgg19 state dict ('features.N.*' keys) for VGGLoss
                                          # (layer_util.py:384 uses the ImageNet weights); None: $HM_VGG19_WEIGHTS, then
                                          # torch hub's cache; "random": seeded random stand-in (bench / parity tests)
                 vgg_seed=1234,           # seed of that stand-in
                 cuda_graph=True,         # optimize_parameters(): capture the fused step in a CUDA graph after two eager
                                          # steps and replay it (same kernels, no per-launch host work / launch gaps)
                 overlap_streams=True,    # fused step: run the VGG branch / the discriminator's own backward pass on a
                                          # second CUDA stream so that their HBM-bound kernels overlap the other
                                          # branch's tensor-core kernels (HM_STREAMS=0 disables)
                 data_parallel=True,      # under torch.distributed: allreduce the gradients over all ranks (False: this
                                          # replica trains on its own -- used by the multi-GPU equivalence check)
                 sn_D=False)              # K13: spectral-norm the PatchGAN convs (models/sn_utils.py SNConv2d); the
                                          # reference's MultiscaleDiscriminator uses plain convs, so default off
        d.update(kw)
        for k, v in d.items():
            setattr(self, k, v)


def create_model(opt, data_size=None):
    """models/models.py:6-24."""
    if opt.model == "pix2pixHD_condImg":
        model = Pix2PixHDModel_condImg(opt)
    elif opt.model == "AE_maskgen_twostream":      # box2mask (BASELINE config #5)
        from .box2mask import TwoStreamAE_mask
        model = TwoStreamAE_mask(opt)
    else:
        # pix2pixHD_condImgColor exists in the reference factory but is untrainable as shipped
        raise NotImplementedError("the model is not implemented")
    print("model [%s] was created" % (model.name()))
    if opt.isTrain and len(opt.gpu_ids):
        model = _DataParallelShim(model)
    return model


class _DataParallelShim(object):
    """What train_mask2image.py needs from nn.DataParallel (models/models.py:21-22): `.module` and a kwargs call.
    Data parallelism itself is one process per GPU (torchrun) + an NCCL allreduce inside the optimizers."""

    def __init__(self, module):
        self.module = module

    def __call__(self, *a, **kw):
        return self.module.forward(*a, **kw)


class FusedAdam(object):
    """torch.optim.Adam(params, lr, betas=(beta1, 0.999)) (pix2pixHD_condImg_model.py:135,139) over a FlatParams
    buffer: zero_grad / step / param_groups like the reference's optimizers, one kernel per param group."""

    def __init__(self, ctx, fp, lr, betas=(0.5, 0.999), eps=1e-8, groups=None, dist_group=None):
        self.ctx, self.fp = ctx, fp
        self.betas, self.eps = betas, eps
        self.m = torch.zeros_like(fp.flat)
        self.v = torch.zeros_like(fp.flat)
        self.step_count = 0
        # param_groups: list of dicts with 'lr' and a [begin, end) range of the flat buffer
        self.param_groups = groups if groups is not None else [dict(lr=lr, begin=0, end=fp.total, params=list(fp.params.values()))]
        self.dist_group = dist_group
        self.data_parallel = True
        self.grads_reduced = False
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=fp.flat.device)  # step count for captured steps

    def zero_grad(self):
        self.fp.grad.zero_()
        self.grads_reduced = False

    def allreduce(self):
        if not self.data_parallel:
            return 1.0
        if not self.grads_reduced:
            parallel.allreduce_sum_(self.fp.grad, self.dist_group)
            self.grads_reduced = True
        return 1.0 / parallel.world()[1]

    def step(self, grad_scale=None, captured=False):
        """captured=True: the step is being recorded into a CUDA graph -- the step count is advanced and read on the
        device (hm_adam_step_dev); the host mirror `step_count` is advanced by the caller once per replay."""
        scale = self.allreduce() if grad_scale is None else grad_scale
        if captured:
            self.step_dev.add_(1)
        else:
            self.step_count += 1
        for g in self.param_groups:
            b, e = g["begin"], g["end"]
            if e > b:
                if captured:
                    ops.adam_step_dev(self.ctx, self.fp.flat[b:e], self.fp.grad[b:e], self.m[b:e], self.v[b:e],
                                      float(g["lr"]), self.betas[0], self.betas[1], self.eps, self.step_dev, scale)
                else:
                    ops.adam_step(self.ctx, self.fp.flat[b:e], self.fp.grad[b:e], self.m[b:e], self.v[b:e], float(g["lr"]),
                                  self.betas[0], self.betas[1], self.eps, self.step_count, scale)
        self.fp.version += 1
        self.grads_reduced = False

    def lr_signature(self):
        return tuple(float(g["lr"]) for g in self.param_groups)

    def state_dict(self):
        return dict(step=self.step_count, m=self.m.cpu(), v=self.v.cpu(), lrs=[g["lr"] for g in self.param_groups])

    def load_state_dict(self, sd):
        self.step_count = sd["step"]
        self.m.copy_(sd["m"]); self.v.copy_(sd["v"])
        for g, lr in zip(self.param_groups, sd["lrs"]):
            g["lr"] = lr


class _LossFn(torch.autograd.Function):
    """Autograd anchor: the losses returned to the training script are outputs of this node, and its backward
    launches the hand-written backward schedule (writing straight into the flat .grad buffers)."""

    @staticmethod
    def forward(ctx, anchor, model, which, values):
        ctx.model, ctx.which = model, which
        return values.clone()

    @staticmethod
    def backward(ctx, grad_out):
        w = [float(x) for x in grad_out.tolist()]
        if ctx.which == "G":
            ctx.model._backward_G(w)
        else:
            ctx.model._backward_D(w)
        return None, None, None, None


class Pix2PixHDModel_condImg(object):
    def name(self):
        return "Pix2PixHDModel_condImg"

    def __init__(self, opt):
        self.opt = opt
        self.gpu_ids = opt.gpu_ids
        self.isTrain = opt.isTrain
        self.save_dir = os.path.join(opt.checkpoints_dir, opt.name)
        if not torch.cuda.is_available():
            raise RuntimeError("Pix2PixHDModel_condImg (B200) needs a CUDA device: there is no CPU fallback")
        dev = torch.device("cuda", self.gpu_ids[0] if len(self.gpu_ids) else torch.cuda.current_device())
        self.device = dev
        prec = getattr(opt, "precision", "bf16x3")
        if prec not in ("bf16x3", "mixed", "bf16"):
            raise ValueError("precision must be bf16x3 | mixed | bf16, got %s" % prec)
        # bf16x3: every GEMM as 3 split products (fp32 parity); mixed: forward bf16x3 (outputs and losses keep the fp32
        # tolerance), gradient GEMMs single bf16 products; bf16: single products everywhere
        self.ctx = Ctx(dev, split=(prec != "bf16"), split_bwd=(prec == "bf16x3"))
        self.netG_type = opt.netG
        self.use_features = opt.instance_feat or opt.label_feat
        if self.use_features:
            raise NotImplementedError("instance/label feature encoder (netE) is broken in the reference "
                                      "(Pix2Pix_NET.py:249-255) and not part of this path")
        if opt.norm != "instance":
            raise NotImplementedError("normalization layer [%s] is not found" % opt.norm)
        if opt.label_nc == 0:
            raise NotImplementedError("label_nc == 0 (raw label input) is not part of this path")
        input_nc = opt.label_nc
        netG_input_nc = input_nc + (0 if opt.no_instance else 1)
        # ---- generator (pix2pixHD_condImg_model.py:34-56)
        self.fpG = FlatParams(dev)
        if opt.netG == "global":
            netG_input_nc += 3
            self.netG = GlobalGenerator(self.ctx, self.fpG, netG_input_nc, opt.output_nc, opt.ngf, opt.n_downsample_global,
                                        opt.n_blocks_global, opt.use_output_gate)
        elif opt.netG == "local":
            from .local_enhancer import LocalEnhancer
            netG_input_nc += 3
            self.netG = LocalEnhancer(self.ctx, self.fpG, netG_input_nc, opt.output_nc, opt.ngf, opt.n_downsample_global,
                                      opt.n_blocks_global, opt.n_local_enhancers, opt.n_blocks_local)
        elif opt.netG == "global_twostream":     # :47-50 (what scripts/train_mask2image_city.sh trains)
            from .two_stream import GlobalTwoStreamGenerator
            self.netG = GlobalTwoStreamGenerator(self.ctx, self.fpG, netG_input_nc, opt.output_nc, opt.ngf,
                                                 opt.n_downsample_global, opt.n_blocks_global, opt.use_skip,
                                                 opt.which_encoder, opt.use_output_gate, opt.feat_fusion)
            netG_input_nc += 3    # the encode kernel still lays out [label | edge | cond image] (D conditioning, :216)
        else:
            raise NameError("global generator name is not defined properly: %s" % opt.netG)
        self.netG_input_nc = netG_input_nc
        gen = torch.Generator().manual_seed(getattr(opt, "init_seed", 0))
        # ---- discriminator (:59-81)
        self.netD = None
        if self.isTrain:
            if opt.no_lsgan and not opt.no_ganFeat_loss:
                # the reference's feature-matching discriminator never applies the Sigmoid it appends
                # (Discriminator_NET.py:111-114 loops over n_layers + 2 sub-models), so nn.BCELoss would see raw logits
                raise NotImplementedError("--no_lsgan is only defined together with --no_ganFeat_loss (the reference's "
                                          "getIntermFeat discriminator skips its Sigmoid, Discriminator_NET.py:111-114)")
            # :61-70  --no_imgCond drops the masked image from the D conditioning, --mask_gan_input multiplies the D
            # input by mask_in (mask_out with --use_soft_mask)
            n_lab = input_nc + (0 if opt.no_instance else 1)
            netD_input_nc = n_lab + (0 if opt.no_imgCond else 3) + opt.output_nc
            # :71-72, 178-179, 227-228: the context-only two-stream generator is judged on the bare image
            self.d_image_only = opt.netG == "global_twostream" and opt.which_encoder == "ctx"
            if self.d_image_only:
                netD_input_nc = 3
            self.netD_input_nc = netD_input_nc
            self.d_img_c0 = netD_input_nc - opt.output_nc      # first image channel of the D operand
            self.fpD = FlatParams(dev)
            self.netD = MultiscaleDiscriminator(self.ctx, self.fpD, netD_input_nc, opt.ndf, opt.n_layers_D, opt.num_D,
                                                spectral_norm=getattr(opt, "sn_D", False),
                                                getIntermFeat=not opt.no_ganFeat_loss,   # :75 (state-dict key names)
                                                use_sigmoid=opt.no_lsgan)                # :64
        # one flat buffer [G | D] so data parallelism is a single allreduce
        total = self.fpG.total + (self.fpD.total if self.isTrain else 0)
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.fpG.materialize(self.flat[:self.fpG.total], self.flat_grad[:self.fpG.total])
        for c in self.netG.convs():
            c.init_reference(gen)
        if self.isTrain:
            self.fpD.materialize(self.flat[self.fpG.total:], self.flat_grad[self.fpG.total:])
            for c in self.netD.convs():
                c.init_reference(gen)
            if self.netD.spectral_norm:
                self.netD.setup_spectral_norm(gen)
        print("---------- Networks initialized -------------")
        # ---- load networks (:94-101)
        if not self.isTrain or opt.continue_train or opt.load_pretrain:
            pretrained_path = "" if not self.isTrain else opt.load_pretrain
            self.load_network(self.fpG, "G", opt.which_epoch, pretrained_path)
            if self.isTrain:
                self.load_network(self.fpD, "D", opt.which_epoch, pretrained_path)
        # ---- losses and optimizers (:103-139)
        if self.isTrain:
            # :105-107 ImagePool(opt.pool_size): a history of discriminator inputs for the fake pass of loss_D.  One process
            # per GPU here, so the reference's multi-GPU ban applies to a multi-rank job
            if opt.pool_size > 0 and (len(self.gpu_ids) > 1 or parallel.world()[1] > 1):
                raise NotImplementedError("Fake Pool Not Implemented for MultiGPU")
            self.pool_size = int(opt.pool_size)
            self._pool = None          # dict(op=Operand [pool_size, H, W, cs], num=stored images, dec=int32 [B, 2] on the device)
            self.old_lr = opt.lr
            self.vgg = None
            if not opt.no_vgg_loss:
                self.vgg = Vgg19(self.ctx, load_vgg19_state_dict(opt))
            self.loss_names = list(LOSS_NAMES)
            groups = None
            if opt.niter_fix_global > 0 and opt.netG == "local":
                print("------------- Only training the local enhancer network (for %d epochs) ------------" % opt.niter_fix_global)
                groups = self.netG.param_groups(opt.lr, opt.n_local_enhancers)
            self.optimizer_G = FusedAdam(self.ctx, self.fpG, opt.lr, (opt.beta1, 0.999), groups=groups)
            self.optimizer_D = FusedAdam(self.ctx, self.fpD, opt.lr, (opt.beta1, 0.999))
            self.optimizer_G.data_parallel = self.optimizer_D.data_parallel = getattr(opt, "data_parallel", True)
            self.loss_acc = torch.zeros(5, dtype=torch.float64, device=dev)
            self._anchor = torch.zeros(1, device=dev, requires_grad=True)
        self._step = None
        self._side = None
        self._side2 = None
        self._after_d_dgrad = None
        self._grad_ready = None
        # stream overlap only inside optimize_parameters(); forward() / backward() of the script sequence stay single-stream
        self._overlap = False
        self._pinned = {}
        self._graph = None          # dict(graph, inputs, losses, st, sig) once the fused step has been captured
        self._eager_steps = 0
        self.fake_image = self.real_image = self.input_label = self.input_image = None

    # ------------------------------------------------------------------------------------------------
    def _to_device(self, name, t):
        """host (CPU) fp32 tensor -> device, through a cached pinned staging buffer (async H2D)."""
        if t is None:
            return None
        if t.is_cuda:
            return t.to(self.device, torch.float32).contiguous()
        t = t.detach().to(torch.float32).contiguous()
        ent = self._pinned.get(name)
        if ent is None or ent[0].shape != t.shape:
            ent = (torch.empty(t.shape, dtype=torch.float32, pin_memory=True), torch.cuda.Event())
            self._pinned[name] = ent
        buf, ev = ent
        # the previous step's asynchronous H2D copy out of this staging buffer may still be queued behind ~700 kernel
        # launches: wait for it before overwriting the buffer (otherwise step N would train on step N+1's data)
        ev.synchronize()
        buf.copy_(t)
        out = buf.to(self.device, non_blocking=True)
        ev.record()
        return out

    def encode_input(self, label_map, inst_map=None, real_image=None, mask_in=None, train=True, mask_out=None):
        """pix2pixHD_condImg_model.py:144-174 -> operands for G / D / VGG (K11)."""
        assert real_image is not None and mask_in is not None
        opt, ctx = self.opt, self.ctx
        label = self._to_device("label", label_map)
        inst = None if opt.no_instance else self._to_device("inst", inst_map)
        image = self._to_device("image", real_image)
        mask = self._to_device("mask_in", mask_in)
        B, _, H, W = label.shape
        g_in = Operand(ctx, B, H, W, self.netG_input_nc, border=3)
        d_in = v_in = d_mask = None
        if train:
            d_in = Operand(ctx, self._d_segments() * B, H, W, self.netD_input_nc)   # [fake ; real (; pooled fakes)]
            if self.vgg is not None:
                v_in = Operand(ctx, 2 * B, H, W, 3)
            if opt.mask_gan_input:                                   # :217 mask_cond
                d_mask = self._to_device("mask_out", mask_out) if opt.use_soft_mask else mask
        ops.encode_input(ctx, label, inst, image, mask, opt.label_nc, g_in, d_in, v_in,
                         d_no_imgcond=bool(train and opt.no_imgCond), d_mask=d_mask,
                         d_image_only=bool(train and getattr(self, "d_image_only", False)))
        # the one-hot label map and the 0/1 instance edges are exact in bf16: no lo product over those channels
        # (a soft D-input mask scales them to arbitrary values, so the guarantee does not hold for d_in then)
        n_exact = opt.label_nc + (0 if opt.no_instance else 1)
        g_in.lo_c0 = n_exact
        if d_in is not None and not (d_mask is not None and opt.use_soft_mask) and not getattr(self, "d_image_only", False):
            d_in.lo_c0 = n_exact
        return dict(label=label, inst=inst, image=image, mask=mask, g_in=g_in, d_in=d_in, v_in=v_in, B=B, H=H, W=W,
                    d_mask=d_mask)

    # ------------------------------------------------------------------------------------------------
    def _forward_all(self, label, inst, image, mask_in, mask_out=None):
        """The whole forward of pix2pixHD_condImg_model.py:198-259; returns the step context."""
        opt, ctx = self.opt, self.ctx
        st = self.encode_input(label, inst, image, mask_in, train=True, mask_out=mask_out)
        B, H, W = st["B"], st["H"], st["W"]
        side = self._side_stream() if self.vgg is not None else None
        if side is not None and os.environ.get("HM_VGG_SPLIT", "0") == "1":
            # Opt-in: VGG(real image) does not depend on the generator, so its half of the [fake ; real] batch can start
            # now on the second stream and fill the generator forward's idle tensor time (InstanceNorm passes, partial
            # waves).  Measured A/B on B200 (bf16x3, config #2): 65.2 / 65.3 ms with, 64.8 / 65.3 ms without -- the SM
            # clock drops from 1650-1680 to 1605-1665 MHz: the step sits at the 1000 W power cap, more concurrency buys
            # nothing.  Off by default.
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                st["v_tape"] = self.vgg.forward(st["v_in"], n0=B, n=B)
        t, g_tape = self._run_generator(st)
        fake = torch.empty(B, 3, H, W, dtype=torch.float32, device=self.device)
        ops.finish_fake(ctx, t, st["image"], st["mask"], opt.use_output_gate, fake, st["d_in"], self.d_img_c0, st["v_in"],
                        d_mask=st["d_mask"])
        st.update(t=t, g_tape=g_tape, fake=fake)
        if self._d_segments() == 3:           # discriminate(..., use_pool=True), :182-184: third segment = pool.query(fake half)
            pool, d_in = self._pool, st["d_in"]
            cur, out = d_in.images(0, B), d_in.images(2 * B, B)
            for b in range(B):
                ops.pool_exchange(ctx, cur, pool["op"], out, pool["dec"], b)
        acc = self.loss_acc
        acc.zero_()
        side = self._side_stream() if self.vgg is not None else None
        if side is not None:                                           # VGG branch (:245-247) on the second stream
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._vgg_forward_losses(st, B, acc)
        # D on [fake ; real] (the fake.detach() pass of :218 and the pass of :231 see identical values: computed once)
        st["d_tape"] = self.netD.forward(st["d_in"])
        for lv in st["d_tape"]:
            pred = lv["taps"][-1]
            nseg = self._d_segments()
            half = pred.numel() // nseg
            bce = bool(opt.no_lsgan)                                    # GANLoss: MSE (LSGAN) or BCE on sigmoid(pred), losses.py:17-20
            ops.mse_sum(ctx, pred[:B], 1.0, 1.0 / half, acc, 0, bce=bce)      # G_GAN  (:232)
            ops.mse_sum(ctx, pred[B:2 * B], 1.0, 1.0 / half, acc, 3, bce=bce)  # D_real (:223)
            ops.mse_sum(ctx, pred[2 * B:] if nseg == 3 else pred[:B], 0.0, 1.0 / half, acc, 4, bce=bce)   # D_fake (:219; pool :182-184)
            if not opt.no_ganFeat_loss:                               # :235-242
                cf = (1.0 / opt.num_D) * (4.0 / (opt.n_layers_D + 1)) * opt.lambda_feat
                for tap in lv["taps"][:-1]:
                    ops.l1_sum(ctx, tap[:B], tap[B:2 * B], cf / (tap.numel() // nseg), acc, 1)
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)
        elif self.vgg is not None:                                    # :245-247
            self._vgg_forward_losses(st, B, acc)
        if opt.lambda_rec > 0:                                        # :249-251
            ops.l1_sum(ctx, fake, st["image"], opt.lambda_rec / fake.numel(), acc, 1)
        st["losses"] = acc.to(torch.float32)
        return st

    def _vgg_forward_losses(self, st, B, acc):
        if "v_tape" in st:      # the real half already ran while the generator was busy (see _forward_all)
            self.vgg.forward(st["v_in"], tape=st["v_tape"], n0=0, n=B)
        else:
            st["v_tape"] = self.vgg.forward(st["v_in"])
        for li, tap in st["v_tape"]["taps"].items():
            wi = VGG_WEIGHTS[sorted(st["v_tape"]["taps"]).index(li)]
            ops.l1_sum(self.ctx, tap[:B], tap[B:], self.opt.lambda_feat * wi / (tap.numel() // 2), acc, 2)

    def _wgrad_stream(self):
        """Third stream of the fused step: the generator's weight gradients (None when overlap is off)."""
        if not self._overlap or os.environ.get("HM_WGRAD_STREAM", "1") == "0":
            return None
        if self._side2 is None:
            self._side2 = torch.cuda.Stream(device=self.device)
        return self._side2

    def _side_stream(self):
        """Second stream of the fused step, or None when overlap is off / we are not inside the fused step."""
        if not self._overlap:
            return None
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
        return self._side

    def _run_generator(self, st):
        """:207-210.  'global' / 'local': netG(cat(label, cond_image)); 'global_twostream': netG(cond_image, label, mask)
        -- the label stream reads the first input_nc channels of the same operand (a view with fewer valid channels)."""
        if self.netG_type != "global_twostream":
            return self.netG.forward(st["g_in"])
        g = st["g_in"]
        obj = Operand.__new__(Operand)
        obj.hi, obj.lo, obj.n, obj.h, obj.w, obj.cs, obj.border = g.hi, g.lo, g.n, g.h, g.w, g.cs, g.border
        obj.c = self.netG.input_nc
        obj.lo_c0 = min(g.lo_c0, obj.c)
        ctx_in = ops.cond_image_operand(self.ctx, st["image"], st["mask"], 3)
        return self.netG.forward(ctx_in, obj, st["mask"])

    def forward(self, label, inst, image, feat, mask_in, mask_out, infer=False):
        """pix2pixHD_condImg_model.py:198-259.  Inputs are the reference's CPU NCHW tensors; returns
        [[G_GAN, G_GAN_Feat, G_VGG, D_real, D_fake], fake_image | None] with differentiable scalar losses."""
        if self.isTrain:
            self._pool_decide(label)
        st = self._forward_all(label, inst, image, mask_in, mask_out)
        self._step = st
        lg = _LossFn.apply(self._anchor, self, "G", st["losses"][:3])
        ld = _LossFn.apply(self._anchor, self, "D", st["losses"][3:])
        self._keep_visuals(st)
        return [[lg[0], lg[1], lg[2], ld[0], ld[1]], st["fake"] if infer else None]

    def _keep_visuals(self, st):
        # the reference copies 4 tensors to the host EVERY step (:253-256); here they stay on the device and
        # get_current_visuals() fetches them on demand.
        self._vis = st

    # ---- backward schedules -----------------------------------------------------------------------------
    def _backward_G(self, w):
        """d(w0*G_GAN + w1*G_GAN_Feat + w2*G_VGG)/d(G params): VGG dgrad -> D dgrad (no D wgrad: the reference computes
        one and discards it, train_mask2image.py:79,84) -> head -> generator dgrad + wgrad."""
        st, opt, ctx = self._step, self.opt, self.ctx
        B = st["B"]
        cf = 0.0 if opt.no_ganFeat_loss else w[1] * (1.0 / opt.num_D) * (4.0 / (opt.n_layers_D + 1)) * opt.lambda_feat
        gV = None
        side = self._side_stream() if (self.vgg is not None and w[2] != 0.0) else None
        if side is not None:
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                gV = self.vgg.backward(st["v_tape"], B, [w[2] * opt.lambda_feat * wi for wi in VGG_WEIGHTS])
        gD = self.netD.backward(st["d_tape"], B, "G", w_gan=w[0], w_feat=cf, img_c0=self.d_img_c0, nseg=self._d_segments())
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)
        elif self.vgg is not None and w[2] != 0.0:
            gV = self.vgg.backward(st["v_tape"], B, [w[2] * opt.lambda_feat * wi for wi in VGG_WEIGHTS])
        dy = Operand(ctx, B, st["H"], st["W"], 3, grad=True)
        rec = w[1] * opt.lambda_rec / st["fake"].numel() if opt.lambda_rec > 0 else 0.0
        ops.fake_bwd(ctx, st["t"], st["mask"], opt.use_output_gate, gD, self.netD.gin_coff, gV, st["image"], rec, dy,
                     d_mask=st["d_mask"])
        if self._after_d_dgrad is not None:      # fused step: the discriminator's own backward pass may start now
            self._after_d_dgrad()
        if self.netG_type == "global" and (self._grad_ready is not None or self._overlap):
            self.netG.backward(st["g_tape"], dy_head=dy, grad_ready=self._grad_ready, wgrad_stream=self._wgrad_stream())
        else:
            self.netG.backward(st["g_tape"], dy_head=dy)

    def _backward_D(self, w):
        """d(w0*D_real + w1*D_fake)/d(D params) over the [fake ; real] batch."""
        st = self._step
        self.netD.backward(st["d_tape"], st["B"], "D", w_real=w[0], w_fake=w[1], nseg=self._d_segments())

    # ---- image pool (util/image_pool.py; --pool_size > 0) ---------------------------------------------------------------
    def _d_segments(self):
        """Images per sample the discriminator evaluates in training: [fake ; real], plus the pool's answer when it is on."""
        return 3 if (self.isTrain and getattr(self, "pool_size", 0) > 0) else 2

    def _pool_decide(self, label):
        """ImagePool.query's host half, called once per forward BEFORE anything is enqueued (so it also works when the
        step is a CUDA-graph replay): draws the reference's decisions in the reference's order from python's `random`
        -- store while the pool fills, then per image uniform(0, 1) > 0.5 ? exchange with slot randint(0, size - 1) :
        pass -- and ships them to the device tensor the hm_pool_exchange launches read."""
        if self._d_segments() != 3:
            return
        import random
        B, _, H, W = label.shape
        p = self._pool
        if p is None or p["sig"] != (B, H, W):
            op = Operand(self.ctx, self.pool_size, H, W, self.netD_input_nc, zero=True)
            p = self._pool = dict(op=op, num=0, sig=(B, H, W), dec=torch.zeros(B, 2, dtype=torch.int32, device=self.device),
                                  host=[torch.zeros(B, 2, dtype=torch.int32).pin_memory() for _ in range(8)], turn=0)
        # the H2D copy below is asynchronous: rotate the pinned staging buffers so that a queued copy is never overwritten
        host = p["host"][p["turn"] % len(p["host"])]
        p["turn"] += 1
        for b in range(B):
            if p["num"] < self.pool_size:
                host[b, 0], host[b, 1] = 1, p["num"]
                p["num"] += 1
            elif random.uniform(0, 1) > 0.5:
                host[b, 0], host[b, 1] = 2, random.randint(0, self.pool_size - 1)
            else:
                host[b, 0], host[b, 1] = 0, 0
        p["dec"].copy_(host, non_blocking=True)

    def optimize_parameters(self, label=None, inst=None, image=None, feat=None, mask_in=None, mask_out=None):
        """Fused step: forward, G backward, D backward, one allreduce of [G | D] grads, Adam x2.
        Bit-identical maths to train_mask2image.py:58-86 because loss_D's graph holds no G parameter.
        Returns the 5 losses as a device tensor (no host sync).

        With opt.cuda_graph (default) the step is captured into a CUDA graph on its third call with a given batch
        geometry and replayed afterwards: the ~700 kernel launches of a step then cost no host work and no launch
        gaps.  Inputs are copied into the graph's static input tensors (an H2D copy when they are host tensors)."""
        soft = self.opt.mask_gan_input and self.opt.use_soft_mask
        batch = dict(label=label, inst=None if self.opt.no_instance else inst, image=image, mask_in=mask_in,
                     mask_out=mask_out if soft else None)
        self._pool_decide(label)
        if self._use_graph(batch):
            return self._graph_step(batch)
        self._eager_steps += 1
        return self._fused_step(batch["label"], batch["inst"], batch["image"], batch["mask_in"], captured=False,
                                mask_out=batch["mask_out"])

    def generator_step(self, label=None, inst=None, image=None, feat=None, mask_in=None, mask_out=None):
        """The GENERATOR half of one iteration -- BASELINE's "G-step": encode, G forward, the discriminator and VGG19
        forward passes, `loss_G.backward()` and Adam on G (train_mask2image.py:58-80), with the same stream schedule as
        the fused step but without loss_D's backward pass and the discriminator's Adam step.  Eager (not graphed);
        returns the five losses as a device tensor.  bench.py times it for `g_step_ms`."""
        soft = self.opt.mask_gan_input and self.opt.use_soft_mask
        self._pool_decide(label)
        return self._fused_step(label, None if self.opt.no_instance else inst, image, mask_in, captured=False,
                                mask_out=mask_out if soft else None, d_half=False)

    def _fused_step(self, label, inst, image, mask_in, captured, mask_out=None, d_half=True):
        self._overlap = bool(getattr(self.opt, "overlap_streams", True)) and os.environ.get("HM_STREAMS", "1") != "0"
        try:
            return self._fused_step_impl(label, inst, image, mask_in, captured, mask_out, d_half)
        finally:
            self._overlap = False

    def _fused_step_impl(self, label, inst, image, mask_in, captured, mask_out=None, d_half=True):
        st = self._forward_all(label, inst, image, mask_in, mask_out)
        self._step = st
        self._keep_visuals(st)
        self.flat_grad.zero_()
        nG = self.fpG.total
        dp = getattr(self.opt, "data_parallel", True) and parallel.world()[1] > 1
        scale = 1.0 / parallel.world()[1] if dp else 1.0
        # the two allreduce segments run on the communicator's stream: G's overlaps the D backward pass, D's overlaps
        # the generator's Adam step (the sum over both is the ONE [G | D] allreduce of SURVEY section 8(e))
        side = self._side_stream()
        hD = [None]
        if side is not None and d_half:
            # loss_D's graph holds no generator parameter, so its backward pass only needs the D tape: it runs on the
            # second stream as soon as the generator-side pass through D is enqueued (the packed D weights exist then),
            # overlapping the long generator backward; its gradient segment is allreduced from that stream.
            def start_d():
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    self._backward_D([0.5, 0.5])
                    hD[0] = parallel.allreduce_sum_async_(self.flat_grad[nG:]) if dp else None
                    if hD[0] is not None:
                        hD[0].wait()
            self._after_d_dgrad = start_d
        # Data parallel: the generator's gradient buffer becomes final from its END towards its start while the backward
        # pass walks the layers last-to-first, so it is allreduced in buckets as they complete (like DDP) and only the
        # last small bucket (stem + first down-convs) is exposed.  NCCL's CTAs need SMs of their own: the persistent
        # engines size their grids for (SMs - comm_sms) while buckets are in flight (hm_set_sm_limit).
        handles = []
        bucket = None
        comm_sms = int(os.environ.get("HM_COMM_SMS", "8"))
        use_buckets = dp and self.netG_type == "global" and os.environ.get("HM_BUCKETS", "1") != "0"
        if use_buckets:
            offs = {name: off for name, _, off in self.fpG.specs}
            state = dict(hi=nG)
            min_elems = int(float(os.environ.get("HM_BUCKET_MB", "96")) * (1 << 20) / 4)

            def bucket(conv, flush=False):
                lo = 0 if flush else offs[conv.name + ".weight"]
                if state["hi"] - lo >= (1 if flush else min_elems):
                    if not handles and comm_sms > 0:      # first bucket in flight: leave SMs for NCCL from here on
                        self.ctx.lib.hm_set_sm_limit(self.ctx.sm_count - comm_sms)
                    handles.append(parallel.allreduce_sum_async_(self.flat_grad[lo:state["hi"]]))
                    state["hi"] = lo
        self._grad_ready = bucket
        try:
            self._backward_G([1.0, 1.0, 1.0])
            if use_buckets:
                bucket(None, flush=True)
        finally:
            self._after_d_dgrad = None
            self._grad_ready = None
        if dp and not use_buckets:
            handles.append(parallel.allreduce_sum_async_(self.flat_grad[:nG]))
        if side is None and d_half:
            self._backward_D([0.5, 0.5])
            hD[0] = parallel.allreduce_sum_async_(self.flat_grad[nG:]) if dp else None
        for h in handles:
            h.wait()
        if use_buckets and comm_sms > 0:
            self.ctx.lib.hm_set_sm_limit(0)
        self.optimizer_G.step(grad_scale=scale, captured=captured)
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)
        elif hD[0] is not None:
            hD[0].wait()
        if d_half:
            self.optimizer_D.step(grad_scale=scale, captured=captured)
        return st["losses"]

    # ---- CUDA-graph replay of the fused step ------------------------------------------------------------
    def _graph_signature(self, batch):
        shapes = tuple((k, tuple(v.shape)) for k, v in batch.items() if v is not None)
        return shapes, self.optimizer_G.lr_signature(), self.optimizer_D.lr_signature(), id(self.optimizer_G)

    def _use_graph(self, batch):
        if not getattr(self.opt, "cuda_graph", True) or os.environ.get("HM_CUDA_GRAPH", "1") == "0":
            return False
        if parallel.world()[1] > 1 and os.environ.get("HM_CUDA_GRAPH_DDP", "1") == "0":
            return False     # the NCCL allreduce is captured with the step (verified at 2 and 8 ranks); opt-out switch
        if self._graph is False:   # a capture failed earlier: stay eager
            return False
        if self._graph is not None and self._graph["sig"] != self._graph_signature(batch):
            self._graph = None     # geometry or learning rate changed: re-capture
            self._eager_steps = 2
        return self._graph is not None or self._eager_steps >= 2

    def _graph_step(self, batch):
        if self._graph is None:
            try:
                self._capture(batch)
            except Exception as e:  # noqa: BLE001 -- any capture problem: keep training eagerly
                print("CUDA-graph capture of the fused step failed (%s: %s); staying eager" % (type(e).__name__, e))
                self._graph = False
                torch.cuda.synchronize()
                self._eager_steps += 1
                return self._fused_step(batch["label"], batch["inst"], batch["image"], batch["mask_in"], captured=False,
                                        mask_out=batch["mask_out"])
        g = self._graph
        for k, dst in g["inputs"].items():
            dst.copy_(batch[k], non_blocking=True)
        g["graph"].replay()
        # host mirrors of what the replay did on the device
        self.optimizer_G.step_count += 1
        self.optimizer_D.step_count += 1
        self.fpG.version += 1
        self.fpD.version += 1
        self.ctx.launches += g["launches"]
        self._step = g["st"]
        self._keep_visuals(g["st"])
        return g["losses"]

    def _capture(self, batch):
        dev = self.device
        inputs = {k: torch.empty(tuple(v.shape), dtype=torch.float32, device=dev) for k, v in batch.items() if v is not None}
        for k, dst in inputs.items():
            dst.copy_(batch[k], non_blocking=True)
        self.optimizer_G.step_dev.fill_(self.optimizer_G.step_count)
        self.optimizer_D.step_dev.fill_(self.optimizer_D.step_count)
        torch.cuda.synchronize()
        l0 = self.ctx.launches
        vG, vD = self.fpG.version, self.fpD.version
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, capture_error_mode="thread_local"):
            losses = self._fused_step(inputs["label"], inputs.get("inst"), inputs["image"], inputs["mask_in"], captured=True,
                                      mask_out=inputs.get("mask_out"))
        launches = self.ctx.launches - l0
        self.ctx.launches = l0
        # recording did not execute anything: undo the host-side bookkeeping of the recorded step
        self.fpG.version, self.fpD.version = vG, vD
        self._graph = dict(graph=graph, inputs=inputs, losses=losses, st=self._step, launches=launches,
                           sig=self._graph_signature(batch))

    # ------------------------------------------------------------------------------------------------
    def inference(self, label, inst, image, mask_in, mask_out):
        """pix2pixHD_condImg_model.py:261-283."""
        st = self.encode_input(label, inst, image, mask_in, train=False)
        t, _ = self._run_generator(st)
        fake = torch.empty(st["B"], 3, st["H"], st["W"], dtype=torch.float32, device=self.device)
        ops.finish_fake(self.ctx, t, st["image"], st["mask"], self.opt.use_output_gate, fake, None, 0, None)
        st["fake"] = fake
        self._vis = st
        return fake

    def get_current_visuals(self):
        """:293-299 -> OrderedDict of HxWx3 uint8 arrays (util/util.py:67-99 tensor2im / tensor2label)."""
        st = self._vis
        label = st["label"][0, 0].cpu().numpy().astype(np.int64)
        return OrderedDict([
            ("input_label", colorize_labels(label, self.opt.label_nc)),
            ("input_image", tensor2im(((1 - st["mask"][0]) * st["image"][0]).cpu())),
            ("real_image", tensor2im(st["image"][0].cpu())),
            ("synthesized_image", tensor2im(st["fake"][0].cpu()))])

    # ---- checkpoints (models/base_model.py:46-107) -------------------------------------------------------
    def save_network(self, fp, network_label, epoch_label, gpu_ids=None):
        os.makedirs(self.save_dir, exist_ok=True)
        torch.save(fp.state_dict(), os.path.join(self.save_dir, "%s_net_%s.pth" % (epoch_label, network_label)))

    def load_network(self, fp, network_label, epoch_label, save_dir=""):
        save_filename = "%s_net_%s.pth" % (epoch_label, network_label)
        save_path = os.path.join(save_dir or self.save_dir, save_filename)
        if not os.path.isfile(save_path):
            print("%s not exists yet!" % save_path)
            if network_label == "G":
                raise RuntimeError("Generator must exist!")
            return
        sd = torch.load(save_path, map_location="cpu")
        own = dict(fp.params)
        own.update(fp.buffers)
        try:
            # strict like nn.Module.load_state_dict: no missing, no unexpected, no mis-shaped entry (validated before
            # anything is copied)
            unexpected = [k for k in sd if k not in own]
            if unexpected:
                raise KeyError("unexpected keys: %s" % unexpected)
            fp.load_state_dict(sd, strict=True)
        except KeyError:
            # base_model.py:84-107: first try the entries this network knows ...
            known = {k: v for k, v in sd.items() if k in own}
            if all(k in known for k in fp.params) and all(v.numel() == own[k].numel() and (k in fp.buffers or tuple(
                    v.shape) == tuple(own[k].shape)) for k, v in known.items()):
                fp.load_state_dict(known, strict=True)
                print("Pretrained network %s has excessive layers; Only loading layers that are used" % network_label)
                return
            # ... else every entry whose size matches; the rest keeps its initialisation
            usable = {k: v for k, v in known.items() if tuple(v.shape) == tuple(own[k].shape)}
            not_init = sorted({k.split(".")[0] for k in fp.params if k not in usable})
            print("Pretrained network %s has fewer layers; The following are not initialized:" % network_label)
            print(not_init)
            fp.load_state_dict(usable, strict=False)

    def save(self, which_epoch):
        self.save_network(self.fpG, "G", which_epoch, self.gpu_ids)
        self.save_network(self.fpD, "D", which_epoch, self.gpu_ids)

    def delete_model(self, which_epoch):
        for lbl in ("G", "D"):
            p = os.path.join(self.save_dir, "%s_net_%s.pth" % (which_epoch, lbl))
            if os.path.isfile(p):
                os.remove(p)

    def update_fixed_params(self):
        """:311-317: after niter_fix_global epochs, train the whole generator (fresh Adam state, as the reference)."""
        self.optimizer_G = FusedAdam(self.ctx, self.fpG, self.opt.lr, (self.opt.beta1, 0.999))
        self.optimizer_G.data_parallel = getattr(self.opt, "data_parallel", True)
        print("------------ Now also finetuning global generator -----------")

    def update_learning_rate(self):
        """:319-327."""
        lrd = self.opt.lr / self.opt.niter_decay
        lr = self.old_lr - lrd
        for g in self.optimizer_D.param_groups:
            g["lr"] = lr
        for g in self.optimizer_G.param_groups:
            g["lr"] = lr
        print("update learning rate: %f -> %f" % (self.old_lr, lr))
        self.old_lr = lr


# ------------------------------------------------------------------------------------------------------
def load_vgg19_state_dict(opt):
    """VGG19 weights for VGGLoss (layer_util.py:384: torchvision.models.vgg19(pretrained=True).features[0:30]).
    opt.vgg_weights: path of a torchvision vgg19 state dict, or None to look at $HM_VGG19_WEIGHTS and torch hub's cache
    (where torchvision leaves vgg19-dcbb9e9d.pth), or "random" for the seeded stand-in.  Without the real weights
    G_VGG is computed on random features: fine for throughput and parity work, NOT the reference's training objective
    -- hence the loud warning when the fallback is taken implicitly."""
    path = getattr(opt, "vgg_weights", None)
    seed = getattr(opt, "vgg_seed", 1234)
    if path == "random":
        return random_vgg19_state_dict(seed)
    cands = [path] if path else [os.environ.get("HM_VGG19_WEIGHTS"),
                                 os.path.join(torch.hub.get_dir(), "checkpoints", "vgg19-dcbb9e9d.pth")]
    for c in cands:
        if c and os.path.isfile(c):
            return remap_torchvision_vgg19(torch.load(c, map_location="cpu"))
    if path:
        raise FileNotFoundError("opt.vgg_weights: %s does not exist" % path)
    sys.stderr.write("WARNING: no ImageNet VGG19 weights found (opt.vgg_weights / $HM_VGG19_WEIGHTS / torch hub cache): "
                     "VGGLoss uses a SEEDED RANDOM VGG19 -- G_VGG does not match the reference's training objective. "
                     "Pass vgg_weights='random' to silence this for benchmarks and parity tests.\n")
    return random_vgg19_state_dict(seed)


def remap_torchvision_vgg19(tv_sd):
    """torchvision vgg19 keys 'features.N.{weight,bias}' -> the reference Vgg19 module tree 'slice{K}.N.*'
    (layer_util.py:388-399); classifier / deeper feature entries are dropped like features[0:30] does."""
    sd = OrderedDict()
    for idx, cin, cout in VGG19_CONVS:
        for leaf in ("weight", "bias"):
            src = "features.%d.%s" % (idx, leaf)
            alt = "slice%d.%d.%s" % (VGG19_SLICE_OF[idx], idx, leaf)
            if src in tv_sd:
                t = tv_sd[src]
            elif alt in tv_sd:
                t = tv_sd[alt]
            else:
                raise KeyError("VGG19 state dict lacks %s" % src)
            want = (cout, cin, 3, 3) if leaf == "weight" else (cout,)
            if tuple(t.shape) != want:
                raise KeyError("VGG19 %s has shape %s, expected %s" % (src, tuple(t.shape), want))
            sd[alt] = t.detach().to(torch.float32).clone()
    return sd


def random_vgg19_state_dict(seed=1234):
    """Seeded stand-in for torchvision's ImageNet VGG19 (layer_util.py:384 downloads it; no network here):
    torchvision's own default init (kaiming_normal_ fan_out / relu, zero bias)."""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for idx, cin, cout in VGG19_CONVS:
        std = (2.0 / (cout * 9)) ** 0.5
        k = "slice%d.%d." % (VGG19_SLICE_OF[idx], idx)
        sd[k + "weight"] = torch.randn(cout, cin, 3, 3, generator=g) * std
        sd[k + "bias"] = torch.zeros(cout)
    return sd


def tensor2im(t):
    """util/util.py:67-80: CHW in [-1,1] -> HWC uint8."""
    a = (np.transpose(t.float().numpy(), (1, 2, 0)) + 1) / 2.0 * 255.0
    return np.clip(a, 0, 255).astype(np.uint8)


def colorize_labels(label_hw, n):
    """util/util.py:141-181 (labelcolormap): bit-interleaved colour map for n != 35, Cityscapes palette for 35."""
    if n == 35:
        cmap = np.array([(0, 0, 0), (0, 0, 0), (0, 0, 0), (0, 0, 0), (0, 0, 0), (111, 74, 0), (81, 0, 81), (128, 64, 128),
                         (244, 35, 232), (250, 170, 160), (230, 150, 140), (70, 70, 70), (102, 102, 156), (190, 153, 153),
                         (180, 165, 180), (150, 100, 100), (150, 120, 90), (153, 153, 153), (153, 153, 153),
                         (250, 170, 30), (220, 220, 0), (107, 142, 35), (152, 251, 152), (70, 130, 180), (220, 20, 60),
                         (255, 0, 0), (0, 0, 142), (0, 0, 70), (0, 60, 100), (0, 0, 90), (0, 0, 110), (0, 80, 100),
                         (0, 0, 230), (119, 11, 32), (0, 0, 142)], dtype=np.uint8)
    else:
        cmap = np.zeros((n, 3), dtype=np.uint8)
        for i in range(n):
            r = g = b = 0
            idv = i
            for j in range(7):
                r ^= ((idv >> 0) & 1) << (7 - j)
                g ^= ((idv >> 1) & 1) << (7 - j)
                b ^= ((idv >> 2) & 1) << (7 - j)
                idv >>= 3
            cmap[i] = (r, g, b)
    return cmap[np.clip(label_hw, 0, n - 1)]

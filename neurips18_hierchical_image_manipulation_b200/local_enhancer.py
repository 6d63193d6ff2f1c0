"""LocalEnhancer executor (models/Pix2Pix_NET.py:8-61) -- BASELINE config #4.

The reference defines the class but its model factory never constructs it; `netG == 'local'` is
added here as the data layer anticipates (data/base_dataset.py:260-261).  Parameter names follow the reference
module tree: `model.*` (global trunk without its last three modules), `model{n}_1.*`, `model{n}_2.*`.
"""
from . import ops
from .networks import ConvP, GlobalGenerator
from .ops import Operand


class _Enhancer(GlobalGenerator):
    """One local enhancer level: [pad3, conv7, IN, relu, conv3 s2, IN, relu] (+ prev) -> res blocks -> convT [-> head]."""

    def __init__(self, ctx, fp, input_nc, output_nc, ngf_g, n_blocks_local, n, final, ngf):
        self.ctx, self.fp = ctx, fp
        self.input_nc, self.output_nc, self.use_output_gate = input_nc, output_nc, False
        p1, p2 = "model%d_1." % n, "model%d_2." % n
        st = [("stem", ConvP(ctx, fp, p1 + "1", input_nc, ngf_g, 7, 1, 0)),
              ("down", ConvP(ctx, fp, p1 + "4", ngf_g, ngf_g * 2, 3, 2, 1))]
        for i in range(n_blocks_local):
            st.append(("resA", ConvP(ctx, fp, p2 + "%d.conv_block.1" % i, ngf_g * 2, ngf_g * 2, 3, 1, 0)))
            st.append(("resB", ConvP(ctx, fp, p2 + "%d.conv_block.5" % i, ngf_g * 2, ngf_g * 2, 3, 1, 0)))
        idx = n_blocks_local
        st.append(("up", ConvP(ctx, fp, p2 + str(idx), ngf_g * 2, ngf_g, 3, 2, 1, transposed=True)))
        if final:
            st.append(("head", ConvP(ctx, fp, p2 + str(idx + 4), ngf, output_nc, 7, 1, 0)))
        self.stages = st
        self.with_head = final
        self.feature_nc = ngf_g


class LocalEnhancer(object):
    def __init__(self, ctx, fp, input_nc, output_nc, ngf=32, n_downsample_global=3, n_blocks_global=9,
                 n_local_enhancers=1, n_blocks_local=3):
        self.ctx, self.fp = ctx, fp
        self.n_local = n_local_enhancers
        self.input_nc = input_nc
        self.use_output_gate = False
        self.trunk = GlobalGenerator(ctx, fp, input_nc, output_nc, ngf * (2 ** n_local_enhancers), n_downsample_global,
                                     n_blocks_global, with_head=False)
        self.levels = []
        for n in range(1, n_local_enhancers + 1):
            ngf_g = ngf * (2 ** (n_local_enhancers - n))
            self.levels.append(_Enhancer(ctx, fp, input_nc, output_nc, ngf_g, n_blocks_local, n, n == n_local_enhancers, ngf))

    def convs(self):
        out = self.trunk.convs()
        for lv in self.levels:
            out += lv.convs()
        return out

    def param_groups(self, lr, n_local):
        """niter_fix_global (pix2pixHD_condImg_model.py:122-130): lr for 'model{n_local}*' parameters, 0 for the rest.
        Parameters are laid out trunk-first in the flat buffer, so the groups are contiguous ranges."""
        specs = self.fp.specs
        groups, cur = [], None
        for name, shape, off in specs:
            n = 1
            for s_ in shape:
                n *= s_
            g_lr = lr if name.startswith("model" + str(n_local)) else 0.0
            end = off + (n + 3) // 4 * 4
            if cur is not None and cur["lr"] == g_lr and cur["end"] == off:
                cur["end"] = end
            else:
                cur = dict(lr=g_lr, begin=off, end=end, params=[])
                groups.append(cur)
            cur["params"].append(self.fp.params[name])
        return groups

    def forward(self, x):
        """x: full-resolution input Operand with ReflectionPad2d(3) materialised."""
        ctx = self.ctx
        pyr = [x]
        for _ in range(self.n_local):                       # Pix2Pix_NET.py:49-51
            src = pyr[-1]
            dst = Operand(ctx, src.n, (src.ih - 1) // 2 + 1, (src.iw - 1) // 2 + 1, src.c, border=3, cs=src.cs)
            ops.avgpool3s2(ctx, src, dst)
            pyr.append(dst)
        prev, t_tape = self.trunk.forward(pyr[-1])           # :54
        tapes = [t_tape]
        for i, lv in enumerate(self.levels):                 # :56-60
            xi = pyr[self.n_local - (i + 1)]
            prev, tp = lv.forward(xi, add_at=1, add_tensor=prev)
            tapes.append(tp)
        return prev, tapes

    def backward(self, tapes, dy_head=None):
        g = None
        for i in range(len(self.levels) - 1, -1, -1):
            lv = self.levels[i]
            if lv.with_head:
                lv.backward(tapes[i + 1], dy_head=dy_head, add_at=1)
            else:
                lv.backward(tapes[i + 1], dfeat=g, add_at=1)
            g = lv.add_grad
        self.trunk.backward(tapes[0], dfeat=g)

"""Random state_dicts with the reference's parameter names/shapes (weights ~ N(0, 0.02) like weights_init,
models/layer_util.py:9-16; biases uniform like torch's Conv2d default).  TEST INFRASTRUCTURE (see oracle/__init__.py):
shared by the CPU / GPU tests and by bench.py's CPU and stock-PyTorch baseline legs."""
from collections import OrderedDict

import torch


def _conv(sd, key, cout, cin, k, g, transposed=False):
    shape = (cin, cout, k, k) if transposed else (cout, cin, k, k)
    sd[key + ".weight"] = torch.randn(shape, generator=g) * 0.02
    bound = 1.0 / ((cin if not transposed else cout) * k * k) ** 0.5
    sd[key + ".bias"] = (torch.rand(cout, generator=g) * 2 - 1) * bound


def random_g_sd(input_nc, output_nc, ngf, n_down, n_blocks, seed=0, prefix="model."):
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    _conv(sd, prefix + "1", ngf, input_nc, 7, g)
    idx = 4
    for i in range(n_down):
        m = 2 ** i
        _conv(sd, prefix + str(idx), ngf * m * 2, ngf * m, 3, g)
        idx += 3
    m = 2 ** n_down
    for i in range(n_blocks):
        _conv(sd, prefix + "%d.conv_block.1" % idx, ngf * m, ngf * m, 3, g)
        _conv(sd, prefix + "%d.conv_block.5" % idx, ngf * m, ngf * m, 3, g)
        idx += 1
    for i in range(n_down):
        m = 2 ** (n_down - i)
        _conv(sd, prefix + str(idx), ngf * m // 2, ngf * m, 3, g, transposed=True)
        idx += 3
    _conv(sd, prefix + str(idx + 1), output_nc, ngf, 7, g)
    return sd


def random_d_sd(input_nc, ndf, n_layers, num_D, seed=1):
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for s in range(num_D):
        nf = ndf
        _conv(sd, "scale%d_layer0.0" % s, ndf, input_nc, 4, g)
        for n in range(1, n_layers):
            nf_prev, nf = nf, min(nf * 2, 512)
            _conv(sd, "scale%d_layer%d.0" % (s, n), nf, nf_prev, 4, g)
        nf_prev, nf = nf, min(nf * 2, 512)
        _conv(sd, "scale%d_layer%d.0" % (s, n_layers), nf, nf_prev, 4, g)
        _conv(sd, "scale%d_layer%d.0" % (s, n_layers + 1), 1, nf, 4, g)
    return sd


def named_param(name, shape):
    """Deterministic parameter values keyed by the parameter's NAME (seed = crc32(name)): conv / linear weights
    ~ N(0, 0.02) as weights_init leaves them, norm-layer gains ~ N(1, 0.02) (layer_util.py:14), biases / norm shifts
    ~ U(-0.05, 0.05).  Used where a golden fixture would otherwise have to carry tens of MB of random weights: the
    generating script loads these values into the reference's modules (as a checkpoint would) and the tests regenerate
    them from the stored name -> shape manifest."""
    import zlib
    g = torch.Generator().manual_seed(zlib.crc32(name.encode()) & 0x7FFFFFFF)
    shape = tuple(int(v) for v in shape)
    if name.endswith("weight") and len(shape) >= 2:
        return torch.randn(shape, generator=g) * 0.02
    if name.endswith("weight"):
        return 1.0 + torch.randn(shape, generator=g) * 0.02
    return (torch.rand(shape, generator=g) * 2 - 1) * 0.05

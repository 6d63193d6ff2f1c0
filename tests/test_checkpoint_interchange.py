"""N4 checkpoint interchange, product -> reference direction (models/base_model.py:46-107): a state dict written by
this implementation must load into the reference's own nn.Modules with `load_state_dict(strict=True)`.

CPU tests (the executors only declare parameters at construction; no kernel runs):
  * key set + shapes of every network at BASELINE's real configurations equal the manifest generated from the
    reference's classes (oracle/make_golden_manifest.py -> tests/golden/state_dict_manifest.json);
  * when /root/reference is present (the build container), the strict load is performed for real and the reference
    module, fed the golden weights THROUGH a product state dict, reproduces the golden output.
"""
import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("HM_REFERENCE", "/root/reference")


def _product(name):
    from neurips18_hierchical_image_manipulation_b200.local_enhancer import LocalEnhancer
    from neurips18_hierchical_image_manipulation_b200.networks import FlatParams, GlobalGenerator, MultiscaleDiscriminator
    from neurips18_hierchical_image_manipulation_b200.two_stream import GlobalTwoStreamGenerator
    fp = FlatParams("cpu")
    if name == "GlobalGenerator_config2":
        GlobalGenerator(None, fp, 38, 3, 64, 4, 9, False)
    elif name == "LocalEnhancer_config4":
        LocalEnhancer(None, fp, 39, 3, 32, 4, 9, 1, 3)
    elif name == "GlobalTwoStreamGenerator_shipped":
        GlobalTwoStreamGenerator(None, fp, 35, 3, 64, 3, 9, True, "ctx_label", True, "early_add")
    elif name == "MultiscaleDiscriminator_config2":
        MultiscaleDiscriminator(None, fp, 41, 64, 3, 3, getIntermFeat=True)
    elif name == "MultiscaleDiscriminator_no_ganFeat":
        MultiscaleDiscriminator(None, fp, 41, 64, 3, 3, getIntermFeat=False)
    else:
        raise KeyError(name)
    fp.materialize()
    return fp


def _manifest():
    with open(os.path.join(ROOT, "tests", "golden", "state_dict_manifest.json")) as fh:
        return json.load(fh)


@pytest.mark.parametrize("name", sorted(_manifest()))
def test_state_dict_keys_and_shapes_equal_the_reference_modules(name):
    want = {k: tuple(v) for k, v in _manifest()[name]["keys"].items()}
    got = {k: tuple(v.shape) for k, v in _product(name).state_dict().items()}
    assert set(got) == set(want), (sorted(set(got) - set(want))[:5], sorted(set(want) - set(got))[:5])
    assert got == want
    assert list(got) == list(_manifest()[name]["keys"]) or True   # order is not part of load_state_dict's contract


def _ref_modules():
    sys.path.insert(0, os.path.join(REF, "models"))
    sys.path.insert(0, REF)
    import Discriminator_NET
    import Pix2Pix_NET
    return Pix2Pix_NET, Discriminator_NET


needs_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models")), reason="reference tree not present")


@needs_ref
@pytest.mark.parametrize("name", sorted(_manifest()))
def test_product_state_dict_loads_strictly_into_the_reference_module(name):
    import importlib
    _ref_modules()
    spec = _manifest()[name]
    mod, cls = spec["cls"].split(".")
    net = getattr(importlib.import_module(mod), cls)(*spec["args"])
    fp = _product(name)
    with torch.no_grad():
        fp.flat.normal_(0, 0.02, generator=torch.Generator().manual_seed(1))
    res = net.load_state_dict(fp.state_dict(), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    for k, v in net.state_dict().items():
        assert torch.equal(v, fp.params[k].detach()), k


@needs_ref
def test_reference_generator_fed_through_a_product_state_dict_reproduces_the_golden_output(golden_dir):
    from neurips18_hierchical_image_manipulation_b200.networks import FlatParams, GlobalGenerator
    P, _ = _ref_modules()
    z = np.load(os.path.join(golden_dir, "g_config1.npz"))
    fp = FlatParams("cpu")
    GlobalGenerator(None, fp, 38, 3, 64, 1, 1)
    fp.materialize()
    fp.load_state_dict({k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w::")})
    net = P.GlobalGenerator(38, 3, 64, 1, 1, "instance", "reflect", False)
    net.load_state_dict(fp.state_dict(), strict=True)
    lab = torch.from_numpy(z["label"].astype(np.float32))
    x = torch.cat((torch.zeros(1, 35, 128, 256).scatter_(1, lab.long(), 1.0), torch.from_numpy(z["image"])), 1)
    with torch.no_grad():
        y = net(x)
    assert float((y - torch.from_numpy(z["out"])).abs().max()) < 1e-6


@needs_ref
def test_reference_discriminator_fed_through_a_product_state_dict_reproduces_the_golden_taps(golden_dir):
    from neurips18_hierchical_image_manipulation_b200.networks import FlatParams, MultiscaleDiscriminator
    _, D = _ref_modules()
    z = np.load(os.path.join(golden_dir, "d_small.npz"))
    fp = FlatParams("cpu")
    MultiscaleDiscriminator(None, fp, 12, ndf=8, n_layers=3, num_D=3)
    fp.materialize()
    fp.load_state_dict({k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w::")})
    net = D.MultiscaleDiscriminator(12, 8, 3, "instance", False, 3, True)
    net.load_state_dict(fp.state_dict(), strict=True)
    with torch.no_grad():
        out = net(torch.from_numpy(z["x"]))
    for i, scale in enumerate(out):
        for j, tap in enumerate(scale):
            assert float((tap - torch.from_numpy(z["tap_%d_%d" % (i, j)])).abs().max()) < 1e-5, (i, j)

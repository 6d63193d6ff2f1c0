"""Generate tests/golden/visuals.npz from the REFERENCE'S OWN util/util.py (tensor2im, tensor2label / Colorize /
labelcolormap, :67-181), run once in the build container:   python oracle/make_golden_visuals.py
util/util.py imports the py2-only cStringIO; an empty stand-in module is injected at run time, nothing is copied.
These are the conversions behind Pix2PixHDModel_condImg.get_current_visuals (pix2pixHD_condImg_model.py:293-299).
TEST INFRASTRUCTURE ONLY."""
import io
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("HM_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def main():
    stub = types.ModuleType("cStringIO")
    stub.StringIO = io.BytesIO
    sys.modules["cStringIO"] = stub
    sys.path.insert(0, REF)
    import util.util as U
    g = torch.Generator().manual_seed(31)
    out = {}
    for n in (35, 20, 6):
        out["cmap_%d" % n] = U.labelcolormap(n)
        lab = torch.randint(0, n, (1, 24, 40), generator=g)
        onehot = torch.zeros(n, 24, 40).scatter_(0, lab, 1.0)
        out["label_%d" % n] = lab[0].numpy().astype(np.int64)
        out["color_%d" % n] = U.tensor2label(onehot, n)
    img = torch.rand(3, 24, 40, generator=g) * 2.4 - 1.2
    out["img"] = img.numpy()
    out["img_u8"] = U.tensor2im(img)
    np.savez_compressed(os.path.join(OUT, "visuals.npz"), **out)
    print("visuals.npz", os.path.getsize(os.path.join(OUT, "visuals.npz")))


if __name__ == "__main__":
    main()

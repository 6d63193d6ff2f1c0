"""Write tests/golden/state_dict_manifest.json: the state-dict KEYS and SHAPES of the reference's own network classes
at BASELINE's real configurations (run once in the build container:  python oracle/make_golden_manifest.py).

A product checkpoint loads into the reference's modules with `load_state_dict(strict=True)` exactly when its key set
and shapes equal these (base_model.py:73-107 direction); tests/test_checkpoint_interchange.py checks that for the
product's executors on CPU, and -- when /root/reference is present -- performs the strict load itself.
TEST INFRASTRUCTURE ONLY.
"""
import json
import os
import sys

REF = os.environ.get("HM_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "state_dict_manifest.json")

# name -> (class name, positional constructor arguments as the reference's model code passes them)
CASES = {
    # pix2pixHD_condImg_model.py:44-46 (config #2: netG_input_nc = 35 + 3)
    "GlobalGenerator_config2": ("Pix2Pix_NET.GlobalGenerator", [38, 3, 64, 4, 9, "instance", "reflect", False]),
    # Pix2Pix_NET.py:8-47 (config #4: 35 labels + edge + 3)
    "LocalEnhancer_config4": ("Pix2Pix_NET.LocalEnhancer", [39, 3, 32, 4, 9, 1, 3, "instance", "reflect"]),
    # scripts/train_mask2image_city.sh flag set (:47-50): input_nc 35, ngf 64, 3 down, 9 blocks, skip, ctx_label, gate
    "GlobalTwoStreamGenerator_shipped": ("Pix2Pix_NET.GlobalTwoStreamGenerator",
                                         [35, 3, 64, 3, 9, "instance", "reflect", True, "ctx_label", True, "early_add"]),
    # pix2pixHD_condImg_model.py:73-76
    "MultiscaleDiscriminator_config2": ("Discriminator_NET.MultiscaleDiscriminator", [41, 64, 3, "instance", False, 3, True]),
    "MultiscaleDiscriminator_no_ganFeat": ("Discriminator_NET.MultiscaleDiscriminator", [41, 64, 3, "instance", False, 3, False]),
}


def build(spec):
    import importlib
    mod, cls = spec[0].split(".")
    return getattr(importlib.import_module(mod), cls)(*spec[1])


def main():
    sys.path.insert(0, os.path.join(REF, "models"))
    sys.path.insert(0, REF)
    out = {}
    for name, spec in CASES.items():
        net = build(spec)
        out[name] = dict(cls=spec[0], args=spec[1], keys={k: list(v.shape) for k, v in net.state_dict().items()})
        print(name, len(out[name]["keys"]), "entries")
    with open(OUT, "w") as fh:
        json.dump(out, fh, indent=0, sort_keys=True)


if __name__ == "__main__":
    main()

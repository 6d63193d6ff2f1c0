"""oracle/ -- TEST INFRASTRUCTURE ONLY.  CPU (torch fp32 / fp64) restatement of the reference's mask2image
training hot path, used as the checker for the CUDA implementation.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import
this package; nothing under `neurips18_hierchical_image_manipulation_b200/` does (tests/test_abi.py greps for it).

Pinning status: the reference ships no tests, fixtures or golden vectors for this path, so
parity is UNPINNED BY THE REFERENCE'S OWN TESTS.  Instead the restatement is pinned against the reference's own
nn.Module classes (GlobalGenerator, LocalEnhancer, ResnetBlock, MultiscaleDiscriminator, GANLoss, weights_init,
sn_utils.max_singular_value) imported from /root/reference and run here on CPU with fixed seeds:
`oracle/make_golden.py` (+ `make_golden_sn.py`, `make_golden_twostream.py`) wrote tests/golden/*.npz from those classes,
`oracle/make_golden_model.py` wrote model_*.npz from the reference's own MODEL-LEVEL forward
(Pix2PixHDModel_condImg.__init__ / forward run unmodified on CPU with run-time stubs for `.cuda()` / `util.util` / the
VGG download), and tests/test_oracle_golden.py checks every function below against them.  /root/reference is not
needed (and not read) at test or bench time.

VGG19: the reference downloads ImageNet weights (models/layer_util.py:384); there is no network here, so both the
oracle and the product use a seeded random-initialised VGG19 of the same topology.  Loss VALUES therefore differ
from a real run; kernel parity is unaffected.
"""

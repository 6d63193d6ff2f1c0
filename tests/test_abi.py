"""CPU tests of the drop-in boundary: the C-ABI library builds for sm_100a, loads, and exports exactly the symbols
include/hm_b200.h declares (no compute calls -- there is no GPU here); the product never touches oracle/."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "neurips18_hierchical_image_manipulation_b200")


def _header_functions():
    src = open(os.path.join(ROOT, "include", "hm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(hm_\w+)\s*\(([^;{]*?)\)\s*;", src):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("void", "") else len(args.split(","))
    return out


def test_library_builds_and_exports_every_declared_symbol():
    from neurips18_hierchical_image_manipulation_b200 import build, _lib
    lib_path = build.build()
    assert os.path.exists(lib_path)
    lib = _lib.load()
    decl = _header_functions()
    assert len(decl) >= 25
    nm = subprocess.check_output(["nm", "-D", "--defined-only", lib_path], text=True)
    exported = set(re.findall(r"\bT (hm_\w+)", nm))
    for name, nargs in decl.items():
        assert name in exported, "header declares %s but the .so does not export it" % name
        assert name in _lib.PROTOTYPES, "no ctypes prototype for %s" % name
        assert len(_lib.PROTOTYPES[name][1]) == nargs, "arity mismatch for %s" % name
        assert getattr(lib, name) is not None
    assert set(_lib.PROTOTYPES) == set(decl), "ctypes table and header disagree"
    assert b"sm_100a" in lib.hm_version()


def test_library_is_sm100a_tcgen05_tma():
    """The SASS carries the Blackwell-native mnemonics (B200_PROFILING.md): UTC*MMA (tcgen05.mma), LDTM (tcgen05.ld),
    UTMALDG (TMA)."""
    from neurips18_hierchical_image_manipulation_b200 import build
    sass = subprocess.check_output(["cuobjdump", "-sass", build.build()], text=True)
    assert "UTCHMMA" in sass and "LDTM" in sass and "UTMALDG" in sass
    assert "sm_100a" in sass
    assert "HMMA.16816" not in sass  # no legacy mma.sync path


def test_pure_host_abi_helpers():
    from neurips18_hierchical_image_manipulation_b200 import _lib
    lib = _lib.load()
    assert [lib.hm_pick_bn(c) for c in (1, 3, 16, 17, 64, 65, 128, 129, 1024)] == [16, 16, 16, 32, 64, 128, 128, 256, 256]
    assert lib.hm_rows_pad(1024) == 1024 and lib.hm_rows_pad(3) == 16 and lib.hm_rows_pad(192) == 256
    assert lib.hm_k_pad(38) == 64 and lib.hm_k_pad(64) == 64 and lib.hm_k_pad(65) == 128
    assert lib.hm_wgrad_ws_bytes(3, 3, 1024, 1024) == 9 * 1024 * 1024 * 4
    assert lib.hm_wgrad_ws_bytes(7, 7, 38, 64) == 49 * 64 * 64 * 4
    assert lib.hm_in_ws_bytes(4, 2048, 1024) > 0


def test_invalid_arguments_are_rejected_without_a_gpu():
    import ctypes as C
    from neurips18_hierchical_image_manipulation_b200 import _lib
    lib = _lib.load()
    assert lib.hm_pack_weight(None, 1, 1, 1, 1, 1, 1, None, None, None) == -1
    assert lib.hm_conv_fprop(None, None, None, 64, 64, None, 3, 3, 1, 1, 8, 8, 64, 0, 0.0, None, None, None, None) == -1
    op = _lib.Operand(None, None, 1, 8, 8, 64, 64)
    assert lib.hm_conv_fprop(C.byref(op), None, None, 64, 64, None, 3, 3, 3, 1, 8, 8, 64, 0, 0.0, None, None, None, None) == -1
    assert lib.hm_adam_step(None, None, None, None, 10, 1e-3, 0.5, 0.999, 1e-8, 1, 1.0, None) == -1
    assert lib.hm_in_stats(None, 1, 1, 4, 1e-5, None, None, None, None) == -1
    # entry points added for the thin-side lowering, the two-stream generator, K13 and CUDA-graph replays
    assert lib.hm_pack_weight_ex(None, 1, 1, 1, 0, 1, 1, 1, 0, 1, 1, None, None, None, None) == -1
    assert lib.hm_wgrad_unpack_cols(None, 7, 7, 64, 3, None, 0, None) == -1
    assert lib.hm_tap_unroll(None, None, 1, 8, 8, 3, 8, 1, 7, 0, 0, 1, -1, None, None, 8, 14, 24, None) == -1
    assert lib.hm_tap_combine(None, 1, 8, 14, 24, 1, 7, 3, 0, 0, 1, 1, None, 0, 0.0, None, 8, 8, 3, None) == -1
    assert lib.hm_mask_maxpool(None, 1, 8, 8, 2, None, None) == -1
    assert lib.hm_mask_blend(None, None, None, 1, 4, 4, 8, None, None, None, 8, 0, None) == -1
    assert lib.hm_mask_blend_bwd(None, None, 16, 8, None, None, None) == -1
    assert lib.hm_concat_operands(None, None, 8, 8, None, None, 8, 8, None, None, 16, 16, None) == -1
    assert lib.hm_cond_image_operand(None, None, 1, 8, 8, None, None, 8, 3, None) == -1
    assert lib.hm_sn_power_iteration(None, 1, 8, 8, 1, None) == -1
    assert lib.hm_sn_weight_grad(None, 1, 8, 8, None) == -1
    assert lib.hm_adam_step_dev(None, None, None, None, 10, 1e-3, 0.5, 0.999, 1e-8, None, 1.0, None) == -1
    assert lib.hm_sn_stash_floats(512, 4096) == 2 * 512 + 4096 + 8


def test_options_mirror_the_reference_flag_names_and_defaults():
    """options/mask2image_base_options.py:15-74, options/mask2image_train_options.py:9-46."""
    from neurips18_hierchical_image_manipulation_b200.models import Options
    o = Options()
    assert (o.netG, o.ngf, o.n_downsample_global, o.n_blocks_global, o.norm) == ("global", 64, 4, 9, "instance")
    assert (o.num_D, o.n_layers_D, o.ndf, o.lambda_feat, o.lr, o.beta1) == (2, 3, 64, 10.0, 0.0002, 0.5)
    assert (o.which_encoder, o.feat_fusion, o.use_skip, o.use_output_gate) == ("ctx", "early_add", False, False)
    assert not (o.no_imgCond or o.mask_gan_input or o.use_soft_mask or o.no_lsgan or o.no_vgg_loss or o.no_ganFeat_loss)
    assert (o.precision, o.cuda_graph, o.sn_D) == ("bf16x3", True, False)


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), os.path.join(dirpath, f)
                assert "/root/reference" not in txt, os.path.join(dirpath, f)


def test_model_fails_loudly_without_cuda():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from neurips18_hierchical_image_manipulation_b200.models import Options, create_model
    with pytest.raises(RuntimeError):
        create_model(Options(ngf=8, n_downsample_global=1, n_blocks_global=1))
    with pytest.raises(RuntimeError):      # box2mask (N3) exists now and, like the rest, refuses to run without CUDA
        create_model(Options(model="AE_maskgen_twostream"))
    with pytest.raises(NotImplementedError):
        create_model(Options(model="pix2pixHD_condImgColor"))


def test_synthetic_batch_contract_matches_oracle_generator():
    import torch
    from neurips18_hierchical_image_manipulation_b200.synthetic import synthetic_batch
    from oracle.model import synthetic_batch as oracle_batch
    a, b = synthetic_batch(2, 32, 64, 35, seed=5), oracle_batch(2, 32, 64, 35, seed=5)
    for k in a:
        assert torch.equal(a[k], b[k]), k
    assert a["label"].shape == (2, 1, 32, 64) and a["image"].shape == (2, 3, 32, 64)
    assert float(a["label"].max()) <= 34 and float(a["image"].abs().max()) <= 1.0
    assert set(a["mask_in"].unique().tolist()) <= {0.0, 1.0}
    assert bool((a["mask_out"] >= a["mask_in"]).all())


def test_visual_conversions_match_the_reference_util(golden_dir):
    """get_current_visuals (pix2pixHD_condImg_model.py:293-299): tensor2im and the label colour maps against golden
    vectors produced by the reference's own util/util.py (oracle/make_golden_visuals.py)."""
    import numpy as np
    import torch
    from neurips18_hierchical_image_manipulation_b200.models import colorize_labels, tensor2im
    z = np.load(os.path.join(golden_dir, "visuals.npz"))
    for n in (35, 20, 6):
        assert np.array_equal(colorize_labels(z["label_%d" % n], n), z["color_%d" % n]), n
    assert np.array_equal(tensor2im(torch.from_numpy(z["img"])), z["img_u8"])


def test_vgg19_weight_loading_remaps_torchvision_keys(tmp_path, capsys):
    """VGGLoss must be able to use the real ImageNet VGG19 (layer_util.py:384).  A torchvision-style state dict
    ('features.N.*', classifier entries, deeper convs) is remapped to the reference Vgg19 module tree 'slice{K}.N.*';
    the seeded random stand-in is only silent when asked for explicitly."""
    import pytest
    import torch
    from neurips18_hierchical_image_manipulation_b200 import models as M
    from neurips18_hierchical_image_manipulation_b200.networks import VGG19_CONVS, VGG19_SLICE_OF
    g = torch.Generator().manual_seed(0)
    tv = {}
    for idx, cin, cout in VGG19_CONVS + [(30, 512, 512), (32, 512, 512)]:        # features[30:] exist in torchvision's file
        tv["features.%d.weight" % idx] = torch.randn(cout, cin, 3, 3, generator=g)
        tv["features.%d.bias" % idx] = torch.randn(cout, generator=g)
    tv["classifier.0.weight"] = torch.zeros(8, 8)
    sd = M.remap_torchvision_vgg19(tv)
    assert list(sd) == [("slice%d.%d.%s" % (VGG19_SLICE_OF[i], i, leaf)) for i, _, _ in VGG19_CONVS for leaf in ("weight", "bias")]
    assert torch.equal(sd["slice5.28.weight"], tv["features.28.weight"]) and torch.equal(sd["slice1.0.bias"], tv["features.0.bias"])
    path = str(tmp_path / "vgg19.pth")
    torch.save(tv, path)
    got = M.load_vgg19_state_dict(M.Options(vgg_weights=path))
    assert torch.equal(got["slice3.7.weight"], tv["features.7.weight"])
    with pytest.raises(FileNotFoundError):
        M.load_vgg19_state_dict(M.Options(vgg_weights=str(tmp_path / "missing.pth")))
    bad = dict(tv); del bad["features.19.bias"]
    with pytest.raises(KeyError):
        M.remap_torchvision_vgg19(bad)
    capsys.readouterr()
    M.load_vgg19_state_dict(M.Options(vgg_weights="random"))
    assert "WARNING" not in capsys.readouterr().err
    if not os.path.isfile(os.path.join(torch.hub.get_dir(), "checkpoints", "vgg19-dcbb9e9d.pth")) and not os.environ.get("HM_VGG19_WEIGHTS"):
        M.load_vgg19_state_dict(M.Options())                                     # implicit fallback: loud
        assert "SEEDED RANDOM VGG19" in capsys.readouterr().err

"""CPU-side checks of the C-ABI boundary: the library builds, loads, and exports every symbol of the header;
the Python binding lists exactly those symbols; host-side helpers behave.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(REPO, "include", "stylemesh_b200.h")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(smb_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib_path():
    from stylemesh_b200 import build
    return build.build(verbose=False)


def test_library_exports_every_header_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    syms = header_symbols()
    assert len(syms) >= 24
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/stylemesh_b200.h but not exported"


def test_binding_matches_header(lib_path):
    from stylemesh_b200 import _abi
    bound = sorted(name for name, _, _ in _abi.PROTOTYPES)
    assert bound == header_symbols()
    lib = _abi.load()
    assert lib.smb_abi_version() == 1
    assert _abi.last_error() == ""


def test_argument_errors_are_reported_not_crashing(lib_path):
    from stylemesh_b200 import _abi
    lib = _abi.load()
    rc = lib.smb_ctx_set_impl(None, 0, 0)
    assert rc < 0 and "null context" in _abi.last_error()
    rc = lib.smb_level_begin(None, 64, 64)
    assert rc < 0


def test_product_refuses_to_run_without_cuda():
    """no CPU fallback: the engine must fail loudly when there is no CUDA device."""
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from stylemesh_b200 import _abi, engine
    with pytest.raises(_abi.StyleMeshB200Error):
        engine.VGGEngine({})
    with pytest.raises(_abi.StyleMeshB200Error):
        engine.uv_sample_fwd([torch.zeros(3, 4, 4)], torch.zeros(2, 2, 2))


def test_product_never_imports_oracle():
    pkg = os.path.join(REPO, "stylemesh_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(root, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, os.path.join(root, f)


def test_layer_geometry_and_plan_host_logic():
    from stylemesh_b200.model.losses.content_and_style_losses import build_loss_plan, layer_hw
    assert layer_hw(0, 480, 640) == (480, 640)
    assert layer_hw(2, 480, 640) == (240, 320)
    assert layer_hw(12, 480, 640) == (30, 40)
    assert layer_hw(12, 256, 341) == (16, 21)           # 341 -> 170 -> 85 -> 42 -> 21 (floor mode)
    H, W = 32, 48
    m0 = torch.zeros(1, 1, H, W); m0[..., :, :24] = 1
    m1 = torch.zeros(1, 1, 2 * H, 2 * W); m1[..., :, :48] = 1
    ang = torch.full((1, 1, H, W), 10.0); ang[..., :, 12:] = 80.0
    plan = build_loss_plan([(H, W), (2 * H, 2 * W)], [m0, m1], ang, 30.0, ["r11", "r42"], True)
    r = plan.levels[0]["layers"]["r11"]
    assert r["n"] == H * 24 and r["n_pass"] + r["n_fail"] == r["n"]
    f0 = plan.levels[0]["layers"]["r11"]["f"]; f1 = plan.levels[1]["layers"]["r11"]["f"]
    assert abs(f0 + f1 - 1.0) < 1e-6 and abs(f0 - 0.5) < 1e-6
    assert plan.levels[0]["layers"]["r42"]["hw"] == (4, 6)


def test_cli_parser_accepts_reference_script_flags():
    from stylemesh_b200.model.optimize import build_parser
    argv = ("--gpus 1 --root_path x --dataset scannet --resize_size 256 --texture_size 4096,4096 --min_images 1 "
            "--max_images 1000 --scene s --hierarchical --hierarchical_layers 4 --loss_weight content=7e1 "
            "--loss_weight style=1e-4 --style_weights=1000,1000,10,10,1000 --loss_weight tex_reg=5e3 "
            "--vgg_gatys_model_path p --renderer_mipmap r --learning_rate 1 --decay_step_size 3 --log_images_nth 5000 "
            "--batch_size 1 --max_epochs 7 --train_split 0.99 --val_split 0.01 --sampler_mode repeat --index_repeat 20 "
            "--save_texture --split_mode sequential --num_workers 4 --style_image_path s.jpg --style_pyramid_mode multi "
            "--gram_mode current --angle_threshold 30 --pyramid_levels 4 --min_pyramid_depth 0.25 "
            "--min_pyramid_height 256").split()
    a = build_parser().parse_args(argv)
    assert a.texture_size == [4096, 4096] and a.style_weights == [1000.0, 1000.0, 10.0, 10.0, 1000.0]
    assert a.loss_weights == [["content", "7e1"], ["style", "1e-4"], ["tex_reg", "5e3"]]
    assert a.max_epochs == 7 and a.gpus == 1 and a.style_pyramid_mode == "multi"


def test_reference_entry_point_shim_resolves_to_b200_modules():
    """`python -m model.optimize` / `from model.model import ...` (what the reference scripts call) must resolve to
    the B200 implementation."""
    import importlib
    import sys
    for name in [m for m in sys.modules if m == "model" or m.startswith("model.")]:
        del sys.modules[name]
    m = importlib.import_module("model.model")
    o = importlib.import_module("model.optimize")
    c = importlib.import_module("model.losses.content_and_style_losses")
    assert m.TextureOptimizationStyleTransferPipeline.__module__ == "stylemesh_b200.model.model"
    assert c.ContentAndStyleLoss.__module__ == "stylemesh_b200.model.losses.content_and_style_losses"
    assert callable(o.main)

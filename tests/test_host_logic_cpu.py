"""CPU: host-side logic that needs no GPU - the synthetic workload helpers, the rule that the product
(and bench.py's own arm) never touches oracle/, option presets, and the SRModel mode table."""
import ast
import os

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_synthetic_batch_shapes_and_determinism():
    from deepsee_b200.options.configurations import make_opt
    from deepsee_b200.util.synthetic import synthetic_batch
    o = make_opt("8x_independent_256x256")
    a, b = synthetic_batch(o, 3, seed=7), synthetic_batch(o, 3, seed=7)
    assert set(a) == {"image", "label"}
    assert tuple(a["image"].shape) == (3, 3, 256, 256) and a["image"].dtype == torch.float32
    assert tuple(a["label"].shape) == (3, 1, 256, 256) and a["label"].dtype == torch.int64
    assert torch.equal(a["image"], b["image"]) and torch.equal(a["label"], b["label"])
    assert float(a["image"].min()) >= -1 and float(a["image"].max()) <= 1
    assert int(a["label"].min()) >= 0 and int(a["label"].max()) < o.label_nc
    # blocky maps are constant on the 16 x 16 grid cells
    cell = a["label"][:, :, :16, :16]
    assert (cell == cell[:, :, :1, :1]).all()
    iid = synthetic_batch(o, 1, seed=7, blocky=False)["label"]
    assert iid[0, 0, :16, :16].unique().numel() > 4
    g = synthetic_batch(make_opt("32x_guided_512x512"), 1)
    assert set(g) == {"image", "label", "guiding_image", "guiding_label"}
    assert tuple(g["guiding_label"].shape) == (1, 1, 512, 512)


def test_settle_spectral_norm_converges_to_the_largest_singular_value():
    from deepsee_b200.util.synthetic import settle_spectral_norm
    torch.manual_seed(3)
    conv = torch.nn.utils.spectral_norm(torch.nn.Conv2d(6, 10, 3))
    net = torch.nn.Sequential(conv, torch.nn.Conv2d(10, 4, 1))
    assert settle_spectral_norm(net, iters=200) == 1
    w = conv.weight_orig.detach().flatten(1)
    sigma = torch.dot(conv.weight_u, torch.mv(w, conv.weight_v))
    assert abs(float(sigma) - float(torch.linalg.matrix_norm(w, 2))) < 1e-4 * float(sigma)


def _oracle_imports(path):
    """(lineno, enclosing function or None) of every import of the oracle package in a source file."""
    tree = ast.parse(open(path).read())
    found = []

    def visit(node, fn):
        for ch in ast.iter_child_nodes(node):
            f = ch.name if isinstance(ch, (ast.FunctionDef, ast.AsyncFunctionDef)) else fn
            if isinstance(ch, ast.Import) and any(a.name.split(".")[0] == "oracle" for a in ch.names):
                found.append((ch.lineno, fn))
            if isinstance(ch, ast.ImportFrom) and (ch.module or "").split(".")[0] == "oracle":
                found.append((ch.lineno, fn))
            visit(ch, f)
    visit(tree, None)
    return found


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "deepsee_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                assert _oracle_imports(os.path.join(dirpath, f)) == [], f
    # bench.py: only the CPU-baseline leg (which also serves --impl reference) and the parity check that
    # runs next to it (the oracle as the checker of the benched precision mode) may use it
    where = _oracle_imports(os.path.join(ROOT, "bench.py"))
    assert where and all(fn in ("cpu_train_iteration_timer", "parity_leg") for _, fn in where), where


def test_option_presets_match_the_reference_names():
    from deepsee_b200.options.configurations import make_opt
    o = make_opt("8x_independent_256x256")
    assert (o.start_size, o.crop_size, o.add_noise, o.netE, o.guiding_style_image) == (32, 256, True, "combinedstyle", False)
    o = make_opt("32x_guided_512x512")
    assert (o.start_size, o.crop_size, o.add_noise, o.netE, o.guiding_style_image) == (16, 512, False, "fullstyle", True)
    o = make_opt("8x_guided_128x128", ngf=8)
    assert (o.start_size, o.crop_size, o.dataset, o.ngf) == (16, 128, "celeba", 8)
    with pytest.raises(ValueError):
        make_opt("4x_something")


def test_srmodel_mode_table_covers_the_reference_modes():
    """Every mode string of the reference's SRModel.forward (sr_model.py:76-444) is dispatched."""
    src = open(os.path.join(ROOT, "deepsee_b200", "deepsee_models", "sr_model.py")).read()
    for mode in ("generator", "discriminator", "inference", "encode_only", "demo", "baseline",
                 "inference_noise", "inference_multi_modal", "inference_replace_semantics",
                 "inference_reference_semantics", "inference_interpolation", "inference_interpolation_style",
                 "inference_particular_combined", "inference_particular_full", "inference_reference",
                 "inference_reference_interpolation"):
        assert repr(mode) in src or '"%s"' % mode in src, mode


def test_style_sweep_layout_with_stub_networks():
    """SRModel._style_sweep's host logic (which style matrix goes with which sample, how variants are
    laid out) on CPU, with stub networks: the 'generator' paints every pixel with the mean of the
    style rows it was given plus 10 x the sample's LR mean, so the output identifies (sample, variant)."""
    import numpy as np
    from deepsee_b200.deepsee_models.sr_model import SRModel
    from deepsee_b200.options.configurations import make_opt

    B, S = 2, 8
    opt = make_opt("8x_independent_256x256", n_interpolation=3, noise_delta=0.5, region_idx=[1, 2],
                   dont_merge_fake=False, batchSize=B)
    m = SRModel.__new__(SRModel)
    torch.nn.Module.__init__(m)
    m.opt = opt
    m.model_variant = "independent"
    style = torch.linspace(-0.4, 0.4, B * 19 * 4).view(B, 19, 4)

    def encode_style(**kw):
        return style.clone(), None

    def generate_fake(input_semantics, image_downsized, encoded_style=None, **kw):
        n = image_downsized.shape[0]
        val = encoded_style[:, [1, 2]].mean(dim=(1, 2)) + 10 * image_downsized.mean(dim=(1, 2, 3))
        return val.view(n, 1, 1, 1).expand(n, 3, S, S).clone(), None, encoded_style

    m.encode_style, m.generate_fake = encode_style, generate_fake
    data = {"input_semantics": torch.zeros(B, 19, S, S), "image_hr": torch.zeros(B, 3, S, S),
            "image_lr": torch.stack([torch.full((3, 2, 2), float(b + 1)) for b in range(B)])}

    out = m._style_sweep("inference_interpolation", data)
    assert tuple(out["fake_image"].shape) == (B, 3, S, 3 * S)
    for b in range(B):
        for i, step in enumerate(np.linspace(-0.5, 0.5, num=3)):
            want = (style[b, [1, 2]] + float(step)).clamp(-1, 1).mean() + 10 * (b + 1)
            got = out["fake_image"][b, 0, 0, i * S]
            assert abs(float(got) - float(want)) < 1e-5, (b, i)

    opt.dont_merge_fake = True
    out = m._style_sweep("inference_reference", data)
    assert tuple(out["fake_image"].shape) == (B, B, 3, S, S)
    # sample 0 rendered with sample 1's regions 1, 2
    want = style[1, [1, 2]].mean() + 10 * 1
    assert abs(float(out["fake_image"][0, 1, 0, 0, 0]) - float(want)) < 1e-5

    d2 = dict(data, style_from=style, style_to=style.flip(0))
    out = m._style_sweep("inference_interpolation_style", d2)
    mid = 0.5 * style[0, [1, 2]].mean() + 0.5 * style[1, [1, 2]].mean() + 10 * 1
    assert abs(float(out["fake_image"][0, 1, 0, 0, 0]) - float(mid)) < 1e-5
    assert len(out["style"]) == B and tuple(out["style"][0].shape) == (3, 19, 4)


def test_committed_bench_lines_carry_the_contract_keys():
    """The bench lines committed under profiles/ are what DESIGN.md quotes: each must be one JSON object with
    the keys of the bench contract (metric / value / e2e / roofline / cpu_baseline / clocks / gpu_launches),
    a parity figure inside the bound, and - for the multi-GPU lines - the sharded-vs-whole-batch check."""
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for n in (1, 2, 8):
        path = os.path.join(root, "profiles", "r2_bench_c2_n%d_final.json" % n)
        d = json.load(open(path))
        assert d["n_gpus"] == n and d["unit"] == "images/sec" and d["higher_is_better"] is True
        assert d["value"] > 0 and d["ms_per_step"] > 0 and d["gpu_launches"] > 0
        assert d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
        r = d["roofline"]
        assert r["bound"] in ("hbm", "tensor") and 0 < r["frac"] <= 1.05 and r["peak"] > 0
        assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-6
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        assert "workload" in d["config"] and "model" not in d["config"]
        if n == 1:
            assert d["parity"]["within_tolerance"] and d["parity"]["max_abs"] < d["parity"]["tolerance"] == 1e-3
            assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] > 0
            assert d["value"] > 100 * d["cpu_baseline"]["value"]
        else:
            assert d["ddp_parity"]["grads_only_where_single_process_has_them"]
            assert d["ddp_parity"]["loss_rel_diff"] < 1e-5

"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol
include/manipose_sm100.h declares; the host modules expose the reference's names / state_dict; the product path fails
loudly without a GPU (no CPU fallback) and never imports the oracle."""
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "manipose_sm100.h")


@pytest.fixture(scope="module")
def lib():
    from manipose_b200 import _build, _lib
    _build.build(verbose=False)
    return _lib.load()


def _declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mp_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from manipose_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/manipose_sm100.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes SIGNATURES and the header disagree"


def test_no_compute_without_gpu_fails_loudly(lib):
    from manipose_b200 import _lib
    assert lib.mp_abi_version() == 1
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert lib.mp_device_check() == _lib.MP_EDEVICE
    assert "no CPU fallback" in _lib.last_error()
    import manipose_b200 as mb
    model = mb.RMCLManifoldMixSTE(mb.h36m17_skeleton(), num_frame=9, n_hyp=2, depth_rot=1, depth_seg=1)
    with pytest.raises(_lib.ManiposeLibraryError):
        model(torch.zeros(1, 9, 17, 2))
    with pytest.raises(_lib.ManiposeLibraryError):
        mb.metrics.wta_l2_loss_and_activate_head(torch.zeros(1, 2, 9, 17, 3), torch.zeros(1, 9, 17, 3))


def test_skeleton_validation(lib):
    import ctypes
    from manipose_b200 import _lib
    from manipose_b200.data import h36m17_skeleton, skeleton_tables
    par, ops_rows = skeleton_tables(h36m17_skeleton())
    flat = [v for r in ops_rows for v in r]
    assert lib.mp_set_skeleton(17, (ctypes.c_int32 * 17)(*par), (ctypes.c_float * 51)(*flat)) == 0
    bad = list(par)
    bad[11] = 9
    assert lib.mp_set_skeleton(17, (ctypes.c_int32 * 17)(*bad), (ctypes.c_float * 51)(*flat)) == _lib.MP_EUNSUPPORTED
    assert lib.mp_set_skeleton(15, (ctypes.c_int32 * 17)(*par), (ctypes.c_float * 51)(*flat)) == _lib.MP_EUNSUPPORTED


def test_state_dict_matches_reference_layout():
    """290 tensors with the reference's names, shapes and order (frozen in tests/golden/forward.pt)."""
    import manipose_b200 as mb
    g = torch.load(os.path.join(ROOT, "tests", "golden", "forward.pt"), weights_only=False)
    m = mb.RMCLManifoldMixSTE(mb.h36m17_skeleton(), num_frame=27, n_hyp=5, drop_path_rate=0.1)
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == [tuple(x) for x in g["t27k5_init"]["keys"]]
    full = mb.RMCLManifoldMixSTE(mb.h36m17_skeleton())
    assert len(full.state_dict()) == 290 and sum(p.numel() for p in full.parameters()) == 34_440_062
    for attr in ("n_hyp", "num_joints", "rotations_module", "segments_module", "decoder", "aggregate", "concat_hyp_and_scores",
                 "poses_from_hyp_idx"):
        assert hasattr(full, attr)


def test_product_never_imports_the_oracle():
    code = "import sys; import manipose_b200, manipose_b200.ops, manipose_b200.metrics; " \
           "bad=[m for m in sys.modules if m == 'oracle' or m.startswith('oracle.')]; print(bad); sys.exit(1 if bad else 0)"
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    for dirpath, _, files in os.walk(os.path.join(ROOT, "manipose_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f"{f} imports the oracle"


def test_install_rebinds_reference_names():
    from oracle.ref_loader import reference_available, load_reference
    if not reference_available():
        pytest.skip("/root/reference not present")
    load_reference()
    import importlib
    import manipose_b200 as mb
    replaced = mb.install()
    try:
        arch = importlib.import_module("mh_so3_hpe.architectures")
        met = importlib.import_module("mh_so3_hpe.metrics")
        assert arch.RMCLManifoldMixSTE is mb.RMCLManifoldMixSTE
        assert met.wta_with_scoring_loss is mb.metrics.wta_with_scoring_loss
        model = arch.RMCLManifoldMixSTE(mb.h36m17_skeleton(), num_frame=9, n_hyp=2, depth_rot=1, depth_seg=1)
        assert isinstance(model, importlib.import_module("mh_so3_hpe.architectures.rmcl_manifold_mix_ste").RMCLManifoldMixSTE)
    finally:
        for qual, obj in replaced.items():
            modname, attr = qual.rsplit(".", 1)
            setattr(importlib.import_module(modname), attr, obj)


def test_fused_adam_checkpoints_interoperate_with_torch_adam():
    """optimizer state in torch.optim.Adam's checkpoint layout both ways, and reference-style model checkpoints
    ("model_pos" wrapper, DataParallel "module." prefix).  Host logic only (the step kernel needs a GPU)."""
    import torch
    import manipose_b200 as mb
    from manipose_b200.optim import FusedAdam
    torch.manual_seed(0)
    kw = dict(num_frame=9, n_hyp=2, depth_rot=1, depth_seg=1)
    ref = mb.RMCLManifoldMixSTE(mb.h36m17_skeleton(), **kw)
    opt_ref = torch.optim.Adam(ref.parameters(), lr=4e-5, weight_decay=1e-6)
    for _ in range(3):                       # synthetic gradients: the reference optimizer's own CPU step
        for p in ref.parameters():
            p.grad = torch.randn_like(p)
        opt_ref.step()
    ckpt = {"model_pos": {"module." + k: v.clone() for k, v in ref.state_dict().items()}, "optimizer": opt_ref.state_dict()}

    m = mb.RMCLManifoldMixSTE(mb.h36m17_skeleton(), **kw)
    mb.load_checkpoint(m, ckpt)
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), ref.state_dict().values()))
    opt = FusedAdam(m, lr=1e-3)
    opt.load_state_dict(ckpt["optimizer"])
    assert opt.param_groups[0]["lr"] == 4e-5 and opt.param_groups[0]["weight_decay"] == 1e-6
    assert int(opt.step_dev) == 3
    back = opt.state_dict()
    want = opt_ref.state_dict()
    assert back["param_groups"][0]["params"] == want["param_groups"][0]["params"]
    for i, st in want["state"].items():
        assert torch.equal(back["state"][i]["exp_avg"], st["exp_avg"]) and torch.equal(back["state"][i]["exp_avg_sq"], st["exp_avg_sq"])
        assert float(back["state"][i]["step"]) == float(st["step"])
    opt_ref2 = torch.optim.Adam(ref.parameters(), lr=1.0)
    opt_ref2.load_state_dict(back)          # and the reference optimizer accepts what we write
    assert opt_ref2.param_groups[0]["lr"] == 4e-5


def test_droppath_scales_follow_timm_semantics_on_cpu():
    """Host logic of the training forward (no kernel involved): the batched DropPath factors are 0 or 1 / keep per SAMPLE of each block's
    batch (timm DropPath under mix_ste.py:334-336): a (clip, frame) in spatial blocks, a (clip, token) track in temporal blocks; block 0
    (Identity drop_path) gets none; eval mode gets none."""
    import manipose_b200 as mb
    from manipose_b200 import _lib as L
    torch.manual_seed(0)
    m = mb.RMCLManifoldMixSTE(mb.h36m17_skeleton(), num_frame=9, n_hyp=2, drop_path_rate=0.5).rotations_module
    blocks = m._block_list()
    n_clips, T_, J_ = 64, 9, 17
    m.train()
    scales = m._droppath_scales(blocks, n_clips, torch.device("cpu"))
    assert len(scales) == len(blocks) and scales[0] == [None, None] and scales[1] == [None, None]     # dpr[0] = 0 for the first STE / TTE block
    seen = 0
    for (blk, _, mode, _), (s1, s2) in zip(blocks, scales):
        if s1 is None:
            continue
        keep = 1.0 - blk.drop_path.drop_prob
        for s in (s1, s2):
            seen += 1
            assert s.shape == (n_clips * T_ * J_,) and s.is_contiguous()
            v = s.view(n_clips, T_, J_)
            assert bool(((v == 0) | ((v - 1.0 / keep).abs() < 1e-6)).all())
            if mode == L.MP_ATTN_SPATIAL:
                assert torch.equal(v, v[:, :, :1].expand_as(v))          # one coin per (clip, frame)
            else:
                assert torch.equal(v, v[:, :1, :].expand_as(v))          # one coin per (clip, token) track
        assert not torch.equal(s1, s2)                                    # the two residual branches draw independently
    assert seen == 2 * (len(blocks) - 2)
    last = scales[-1][0].view(n_clips, T_, J_)
    assert 0.3 <= float((last > 0).float().mean()) <= 0.7                 # keep = 0.5 for the deepest block
    m.eval()
    assert all(s == [None, None] for s in m._droppath_scales(blocks, n_clips, torch.device("cpu")))


def test_stacked_gradient_views_on_cpu():
    """train_ops.stacked_grads / accumulate_stacked (gradient accumulation of the K hypothesis heads): one strided view over K
    same-shaped gradients that sit at a constant stride of the flat gradient buffer, per-parameter adds otherwise."""
    from manipose_b200 import train_ops as T
    from manipose_b200.optim import FlatParameters

    class Head(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.norm = torch.nn.LayerNorm(8)
            self.lin = torch.nn.Linear(8, 3)

    heads = torch.nn.ModuleList([Head() for _ in range(4)])
    flat = FlatParameters(heads)
    ws = [h.lin.weight for h in heads]
    view = T.stacked_grads(ws)
    assert view is not None and view.shape == (4, 3, 8)
    add = torch.arange(4 * 3 * 8, dtype=torch.float32).view(4, 3, 8)
    T.accumulate_stacked(ws, add)
    T.accumulate_stacked(ws, add)
    for k, w in enumerate(ws):
        assert torch.equal(w.grad, 2 * add[k]) and w.grad.data_ptr() == flat.flat_grad.data_ptr() + 4 * flat.offsets[id(w)][0]
    assert float(flat.flat_grad.sum()) == float(2 * add.sum())            # nothing was written outside the four weights
    # scalars / biases and the fallback (gradients that do not form a strided stack)
    bs = [h.lin.bias for h in heads]
    T.accumulate_stacked(bs, torch.ones(4, 3))
    assert all(torch.equal(b.grad, torch.ones(3)) for b in bs)
    loose = [torch.nn.Parameter(torch.zeros(5)) for _ in range(3)]
    assert T.stacked_grads(loose) is None
    T.accumulate_stacked(loose, torch.ones(3, 5))
    assert all(torch.equal(p.grad, torch.ones(5)) for p in loose)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` runs without a GPU (the CPU oracle on the host cores) and prints ONE JSON line with the contract's keys."""
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "lifted_frames_per_sec_T243_H36M" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # ranks other than 0 exit 0 without work
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r1 = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                        capture_output=True, text=True, timeout=600, cwd=root, env=env)
    assert r1.returncode == 0 and not [ln for ln in r1.stdout.splitlines() if ln.startswith("{")]


def _reference_names(path):
    """(names imported from mh_so3_hpe.* , attributes read off `model` / `model_pos`) in one unmodified reference source file."""
    import ast
    tree = ast.parse(open(path).read())
    imported, attrs = {}, set()
    for node in ast.walk(tree):
        if isinstance(node, ast.ImportFrom) and node.module and node.module.startswith("mh_so3_hpe"):
            for a in node.names:
                imported[a.name] = node.module
        if isinstance(node, ast.Attribute) and isinstance(node.value, ast.Name) and node.value.id in ("model", "model_pos"):
            attrs.add(node.attr)
    return imported, attrs


def test_every_name_the_reference_callers_use_resolves():
    """The unmodified hot-path callers (hpe/eval_utils.py:16-203, the metric imports of hpe/main_h36m_lifting.py:33-37 and
    hpe/main_3dhp.py:30-32) only touch names that manipose_b200 provides: metrics / augmentations / architectures exports and the
    attributes they read off the model.  Nothing of the reference is executed here (AST scan)."""
    from oracle.ref_loader import reference_available, REFERENCE_ROOT
    if not reference_available():
        pytest.skip("/root/reference not present")
    import manipose_b200 as mb
    from manipose_b200 import architectures, augmentations, data, metrics
    provided = {"mh_so3_hpe.metrics": metrics, "mh_so3_hpe.architectures": architectures, "mh_so3_hpe.augmentations": augmentations,
                "mh_so3_hpe.data": data}
    model = mb.RMCLManifoldMixSTE(mb.h36m17_skeleton(), num_frame=9, n_hyp=2, depth_rot=1, depth_seg=1)
    imported, attrs = _reference_names(os.path.join(REFERENCE_ROOT, "hpe", "eval_utils.py"))
    assert {"mpjpe_error", "pose_flip", "RMCLManifoldMixSTE"} <= set(imported)
    for name, module in imported.items():
        assert hasattr(provided[module], name), f"eval_utils.py imports {module}.{name}: not provided"
    for a in attrs:
        assert hasattr(model, a), f"eval_utils.py reads model.{a}: not provided"
    # the drivers' metric imports: every metric of the hot path and of the §8f analytics is ours; what is left to the reference's
    # own torch code is listed here explicitly (viz-only stretch statistics)
    left_to_reference = {"segments_max_strech_per_bone", "segments_max_diff_strech_per_bone"}
    for driver in ("main_h36m_lifting.py", "main_3dhp.py"):
        imported, _ = _reference_names(os.path.join(REFERENCE_ROOT, "hpe", driver))
        for name, module in imported.items():
            if module == "mh_so3_hpe.metrics" and name not in left_to_reference:
                assert hasattr(metrics, name), f"{driver} imports {module}.{name}: not provided"

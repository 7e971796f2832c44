"""Loader for the UNMODIFIED reference (cedricrommel/manipose) from /root/reference.

TEST INFRASTRUCTURE ONLY.  Used by ``scripts/make_goldens.py`` and by the CPU tests
that pin ``oracle/manipose_oracle.py`` against the real reference code.  Nothing in
the product package (``manipose_b200``) may import this module, and it is never used
on the GPU box (``/root/reference`` does not exist there).

The reference needs three non-invasive accommodations to run on CPU:

1. ``sys.path`` gets ``/root/reference/hpe`` (``mh_so3_hpe`` is a namespace package:
   its ``__init_.py`` is mis-named).
2. ``timm.models.layers.DropPath`` and ``mup.MuReadout`` are absent from this image and
   are injected as stub modules (timm 0.9.16 / mup 1.0.0 semantics, see
   ``hpe/mh_so3_hpe/architectures/mix_ste.py:8-9,334-336``).
3. ``rotation_tools.normalize_vector`` hard-codes ``.cuda()``
   (``hpe/mh_so3_hpe/architectures/utils/rotation_tools.py:9-14``); it is rebound to the
   same arithmetic with the 1e-8 tensor created on ``v.device``.
"""
import os
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("MANIPOSE_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "hpe", "mh_so3_hpe"))


class _DropPath(nn.Module):
    """timm 0.9.16 ``DropPath``: per-sample (dim 0) Bernoulli(keep)/keep in training."""

    def __init__(self, drop_prob: float = 0.0, scale_by_keep: bool = True):
        super().__init__()
        self.drop_prob = drop_prob
        self.scale_by_keep = scale_by_keep

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        shape = (x.shape[0],) + (1,) * (x.ndim - 1)
        mask = x.new_empty(shape).bernoulli_(keep)
        if keep > 0.0 and self.scale_by_keep:
            mask.div_(keep)
        return x * mask


class _MuReadout(nn.Linear):
    """mup 1.0.0 ``MuReadout`` stand-in; only instantiated when ``mup=True`` (never here)."""

    def __init__(self, *args, readout_zero_init=False, output_mult=1.0, **kwargs):
        super().__init__(*args, **kwargs)


def _install_stubs():
    if "timm" not in sys.modules:
        timm = types.ModuleType("timm")
        models = types.ModuleType("timm.models")
        layers = types.ModuleType("timm.models.layers")
        layers.DropPath = _DropPath
        timm.models = models
        models.layers = layers
        sys.modules["timm"] = timm
        sys.modules["timm.models"] = models
        sys.modules["timm.models.layers"] = layers
    if "mup" not in sys.modules:
        mup = types.ModuleType("mup")
        mup.MuReadout = _MuReadout
        sys.modules["mup"] = mup


_LOADED = None


def load_reference():
    """Returns a namespace with the reference's architectures, metrics and Skeleton."""
    global _LOADED
    if _LOADED is not None:
        return _LOADED
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    _install_stubs()
    hpe = os.path.join(REFERENCE_ROOT, "hpe")
    if hpe not in sys.path:
        sys.path.insert(0, hpe)

    from mh_so3_hpe.architectures.utils import rotation_tools

    def normalize_vector(v):
        # same arithmetic as rotation_tools.py:6-17, device-agnostic epsilon
        batch = v.shape[0]
        v_mag = torch.sqrt(v.pow(2).sum(1))
        v_mag = torch.max(v_mag, torch.tensor([1e-8], dtype=v.dtype, device=v.device))
        v_mag = v_mag.view(batch, 1).expand(batch, v.shape[1])
        return v / v_mag

    rotation_tools.normalize_vector = normalize_vector

    from mh_so3_hpe import architectures, metrics
    from mh_so3_hpe.architectures.pose_decoder import PoseDecoder
    from mh_so3_hpe.architectures.utils.forward_kinematics import forward_kinematics
    from mh_so3_hpe.data.skeleton import Skeleton

    ns = types.SimpleNamespace(
        architectures=architectures,
        metrics=metrics,
        PoseDecoder=PoseDecoder,
        forward_kinematics=forward_kinematics,
        rotation_tools=rotation_tools,
        Skeleton=Skeleton,
    )
    ns.make_skeleton = lambda: make_reference_skeleton(ns)
    _LOADED = ns
    return ns


# hpe/mh_so3_hpe/data/h36m_lifting.py:40-57 values restated (dataset module needs files we lack)
_T_POSE_OPS = {
    1: (1, 0, 0), 2: (0, -1, 0), 3: (0, -1, 0), 4: (-1, 0, 0), 5: (0, -1, 0), 6: (0, -1, 0),
    7: (0, 1, 0), 8: (0, 1, 0), 9: (0, 1, 0), 10: (0, 1, 0), 11: (-1, 0, 0), 12: (-1, 0, 0),
    13: (-1, 0, 0), 14: (1, 0, 0), 15: (1, 0, 0), 16: (1, 0, 0),
}


def make_reference_skeleton(ns):
    """17-joint skeleton built with the reference's own ``Skeleton`` class, as
    ``hpe/mh_so3_hpe/data/dataset_3dhp.py:132-138`` does."""
    ops = {j: torch.tensor(v, dtype=torch.float) for j, v in _T_POSE_OPS.items()}
    return ns.Skeleton(
        parents=[-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 9, 8, 11, 12, 8, 14, 15],
        joints_left=[4, 5, 6, 11, 12, 13],
        joints_right=[1, 2, 3, 14, 15, 16],
        t_pose_operators=ops,
    )

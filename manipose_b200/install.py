"""Drop-in: rebind the hot-path names inside the reference's ``mh_so3_hpe`` package so its unmodified drivers
(hpe/main_h36m_lifting.py, hpe/main_3dhp.py, hpe/eval_utils.py, hpe/viz.py) construct and call the sm_100a
implementations and every ``isinstance(model, RMCLManifoldMixSTE)`` check they make passes.

    import sys; sys.path.insert(0, "<reference>/hpe")
    import manipose_b200; manipose_b200.install()          # BEFORE importing the driver module
    import runpy; runpy.run_path("<reference>/hpe/main_h36m_lifting.py", run_name="__main__")
"""
import importlib
import sys

_ARCH = ("MixSTE", "ManifoldMixSTE", "RMCLManifoldMixSTE")
_METRICS = ("wta_l2_loss_and_activate_head", "wta_with_scoring_loss", "weighted_mpjpe_loss", "weighted_mse_loss",
            "mean_velocity_error", "smoothness_regularization", "mpjpe_error", "STANDARD_H36M_WEIGHTS",
            "measure_bones_length", "segments_time_consistency", "segments_time_consistency_per_bone", "sagittal_symmetry",
            "sagittal_symmetry_per_bone", "p_mpjpe", "keypoint_3d_pck", "keypoint_3d_auc", "mse_error", "jointwise_error",
            "jointwise_mse", "coordwise_error", "segments_len_err")
_CONSISTENCY = ("segments_time_consistency", "segments_time_consistency_per_bone", "sagittal_symmetry", "sagittal_symmetry_per_bone")


def install(package: str = "mh_so3_hpe") -> dict:
    """Rebinds the names in ``<package>.architectures`` (and its defining submodules) and ``<package>.metrics``.
    Returns {qualified name: replaced object} so a caller can undo it."""
    from . import architectures as A
    from . import metrics as M
    replaced = {}

    def rebind(modname, names, src):
        try:
            mod = importlib.import_module(modname)
        except ImportError:
            return
        for n in names:
            if hasattr(mod, n):
                replaced[f"{modname}.{n}"] = getattr(mod, n)
                setattr(mod, n, getattr(src, n))

    rebind(f"{package}.architectures", _ARCH, A)
    rebind(f"{package}.architectures.mix_ste", ("MixSTE",), A)
    rebind(f"{package}.architectures.manifold_mix_ste", ("MixSTE", "ManifoldMixSTE", "PoseDecoder"), A)
    rebind(f"{package}.architectures.rmcl_manifold_mix_ste", ("MixSTE", "ManifoldMixSTE", "RMCLManifoldMixSTE"), A)
    rebind(f"{package}.architectures.pose_decoder", ("PoseDecoder",), A)
    rebind(f"{package}.metrics", _METRICS, M)
    rebind(f"{package}.metrics.losses", _METRICS, M)
    rebind(f"{package}.metrics.regularizations", ("smoothness_regularization", "measure_bones_length") + _CONSISTENCY, M)
    rebind(f"{package}.metrics.utils", ("measure_bones_length",), M)
    rebind(f"{package}.metrics.mean_joint_errors", ("mpjpe_error", "p_mpjpe", "mse_error", "jointwise_error", "jointwise_mse",
                                                    "coordwise_error", "segments_len_err"), M)
    rebind(f"{package}.metrics.pck", ("keypoint_3d_pck", "keypoint_3d_auc"), M)
    return replaced


def load_checkpoint(model, checkpoint, strict: bool = True):
    """Loads a reference checkpoint into a manipose_b200 model: unwraps the ``"model_pos"`` entry the drivers write
    (hpe/main_h36m_lifting.py:755-761) and strips the ``module.`` prefix of weights saved from ``nn.DataParallel``."""
    if isinstance(checkpoint, dict) and "model_pos" in checkpoint:
        checkpoint = checkpoint["model_pos"]
    sd = {(k[len("module."):] if k.startswith("module.") else k): v for k, v in checkpoint.items()}
    return model.load_state_dict(sd, strict=strict)

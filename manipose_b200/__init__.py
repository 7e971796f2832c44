"""manipose_b200 — B200-native (sm_100a) implementation of the ManiPose lifting hot path.

Same nn.Module / loss API as cedricrommel/manipose's ``hpe/mh_so3_hpe`` for the path SURVEY.md §8 scopes:
``RMCLManifoldMixSTE`` (MixSTE backbone -> K hypothesis heads -> manifold decoder) and the WTA / scoring losses and
hypothesis metrics.  All arithmetic runs in hand-written CUDA (libmanipose_sm100.so, C ABI in include/manipose_sm100.h);
there is no CPU or PyTorch fallback.
"""
from . import _lib
from .architectures import MixSTE, ManifoldMixSTE, RMCLManifoldMixSTE, PoseDecoder
from .data import Skeleton, h36m17_skeleton
from . import metrics
from .install import install, load_checkpoint

__all__ = ["MixSTE", "ManifoldMixSTE", "RMCLManifoldMixSTE", "PoseDecoder", "Skeleton", "h36m17_skeleton", "metrics", "install", "load_checkpoint"]

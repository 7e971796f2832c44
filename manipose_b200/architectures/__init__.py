from .mix_ste import MixSTE
from .manifold_mix_ste import ManifoldMixSTE, BonesMixSTE
from .rmcl_manifold_mix_ste import RMCLManifoldMixSTE, RMCLRotMixSTE, MCLHead
from .pose_decoder import PoseDecoder

__all__ = ["MixSTE", "ManifoldMixSTE", "RMCLManifoldMixSTE", "BonesMixSTE", "RMCLRotMixSTE", "MCLHead", "PoseDecoder"]

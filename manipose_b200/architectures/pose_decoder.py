"""Manifold decoder behind the reference's ``PoseDecoder`` surface (hpe/mh_so3_hpe/architectures/pose_decoder.py:10-120).

6-D -> SO(3) Gram-Schmidt, T-pose from bone lengths and forward kinematics run as ONE sm_100a kernel
(libmanipose_sm100: mp_decoder_fwd / mp_decoder_bwd); there is no CPU path.
"""
import torch
from torch import nn

from .. import ops
from ..data.skeleton import skeleton_tables


class PoseDecoder(nn.Module):
    def __init__(self, skeleton, rot_rep_dim: int = 6):
        super().__init__()
        self.skeleton = skeleton
        self.rot_rep_dim = rot_rep_dim
        assert rot_rep_dim in [4, 6], f"Unsupported rotations representation dimension: {self.rot_rep_dim}"
        self._tables = skeleton_tables(skeleton)
        # exact = one correctly-rounded IEEE op per reference op (bit-identical to oracle.pose_decoder_ieee); False = rsqrt + FMA
        self.exact = True

    def forward(self, rotations_repr: torch.Tensor, bones_lengths_repr: torch.Tensor, root_positions: torch.Tensor) -> torch.Tensor:
        """rotations_repr [(B H L), J, D]; bones_lengths_repr [B, S, 1]; root_positions [(B H L), 3] -> [(B H L), J, 3]."""
        assert rotations_repr.shape[-1] == self.rot_rep_dim
        ops.set_skeleton(*self._tables)
        n = rotations_repr.shape[0]
        b = bones_lengths_repr.shape[0]
        assert n % b == 0                                   # pose_decoder.py:94
        root = root_positions   # None == zeros (what the reference models pass): the kernel then skips the read
        return ops.decode(rotations_repr, bones_lengths_repr.reshape(b, -1), root, b, 1, n // b, self.rot_rep_dim, self.exact)

    def build_t_pose_from_bone_lengths(self, bones_length: torch.Tensor) -> torch.Tensor:
        """pose_decoder.py:98-120 == decoding identity rotations ([1,0,0,0,1,0]) with a zero root."""
        n = bones_length.shape[0]
        assert bones_length.shape[1] == self.skeleton.num_bones
        ops.set_skeleton(*self._tables)
        ident = torch.tensor([1.0, 0, 0, 0, 1.0, 0], device=bones_length.device).expand(n, self.skeleton.num_joints, 6).contiguous()
        poses, _ = ops.decoder_fwd(ident, bones_length.reshape(n, -1), None, None, n, 1, 1, 6, True)
        return poses

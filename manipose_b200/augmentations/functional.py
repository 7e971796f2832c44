"""Horizontal flip + left/right joint swap (hpe/mh_so3_hpe/augmentations/functional.py:7-28).  Index permutation on
whatever device the tensors live on; kept because ``evaluate`` calls it around the hot path."""
from typing import Tuple

import torch


def pose_flip(poses_tuple: Tuple[torch.Tensor], skeleton) -> Tuple[torch.Tensor]:
    assert isinstance(poses_tuple, tuple)
    out = []
    for pose in poses_tuple:
        assert pose.shape[-1] in [2, 3]
        assert pose.shape[-2] == skeleton.num_joints
        pose[..., 0] *= -1  # in place, like the reference
        pose[..., skeleton.joints_left + skeleton.joints_right, :] = pose[..., skeleton.joints_right + skeleton.joints_left, :]
        out.append(pose)
    return tuple(out)

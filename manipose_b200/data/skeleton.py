"""Kinematic-tree metadata (mirror of hpe/mh_so3_hpe/data/skeleton.py:7-172 for the 17-joint tree the hot path uses).

Only what the hot path and its callers read is kept: parents, has_children, children, joints_left / joints_right,
bones, bones_left / bones_right and t_pose_operators.  The kernels are specialised at compile time for this tree
(SURVEY.md §A.1); any other tree is rejected by ``manipose_b200.ops.set_skeleton``.
"""
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

# hpe/mh_so3_hpe/data/h36m_lifting.py:40-57 after the 17-joint reduction (:649-660) == dataset_3dhp.py:132-138
H36M17_PARENTS = [-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 9, 8, 11, 12, 8, 14, 15]
H36M17_JOINTS_LEFT = [4, 5, 6, 11, 12, 13]
H36M17_JOINTS_RIGHT = [1, 2, 3, 14, 15, 16]
H36M17_NAMES = ["Hip", "RHip", "RKnee", "RFoot", "LHip", "LKnee", "LFoot", "Spine", "Thorax", "Neck/Nose", "Head",
                "LShoulder", "LElbow", "LWrist", "RShoulder", "RElbow", "RWrist"]
_OPS = [(0, 0, 0), (1, 0, 0), (0, -1, 0), (0, -1, 0), (-1, 0, 0), (0, -1, 0), (0, -1, 0), (0, 1, 0), (0, 1, 0), (0, 1, 0),
        (0, 1, 0), (-1, 0, 0), (-1, 0, 0), (-1, 0, 0), (1, 0, 0), (1, 0, 0), (1, 0, 0)]


class Skeleton:
    """Same constructor arguments and read-only properties as the reference's ``Skeleton`` (skeleton.py:8-32)."""

    def __init__(self, parents: Sequence[int], joints_left: List[int], joints_right: List[int],
                 t_pose_operators: Optional[Dict[int, torch.Tensor]] = None, joints_group=None,
                 joints_names: Optional[List[str]] = None):
        assert len(joints_left) == len(joints_right)
        self.t_pose_operators = t_pose_operators
        self._parents = np.array(parents)
        self._joints_left = list(joints_left)
        self._joints_right = list(joints_right)
        self._joints_group = joints_group
        self._joints_names = joints_names if joints_names is not None else [""] * len(self._parents)
        self._compute_metadata()

    def _compute_metadata(self):
        """skeleton.py:87-120."""
        n = len(self._parents)
        self._has_children = np.zeros(n).astype(bool)
        for i, p in enumerate(self._parents):
            if p != -1:
                self._has_children[p] = True
        self._children = [[] for _ in range(n)]
        for i, p in enumerate(self._parents):
            if p != -1:
                self._children[p].append(i)
        # (joint, joint_parent) tuples, like the reference (hpe/mh_so3_hpe/data/skeleton.py:100-103)
        self._bones = tuple((j, int(p)) for j, p in enumerate(self._parents) if p >= 0)
        self._bones_names = tuple(f"{self._joints_names[p]}->{self._joints_names[j]}" for j, p in self._bones)
        bone_index = {b: i for i, b in enumerate(self._bones)}
        bone_parent = dict(self._bones)
        self._bones_left = tuple(bone_index[(j, bone_parent[j])] for j in self._joints_left if j >= 0)       # skeleton.py:110-120
        self._bones_right = tuple(bone_index[(j, bone_parent[j])] for j in self._joints_right if j >= 0)

    num_joints = property(lambda self: len(self._parents))
    num_bones = property(lambda self: len([p for p in self._parents if p >= 0]))
    parents = property(lambda self: self._parents)
    has_children = property(lambda self: self._has_children)
    children = property(lambda self: self._children)
    joints_left = property(lambda self: self._joints_left)
    joints_right = property(lambda self: self._joints_right)
    joints_group = property(lambda self: self._joints_group)
    joints_names = property(lambda self: self._joints_names)
    bones = property(lambda self: self._bones)
    bones_left = property(lambda self: self._bones_left)
    bones_right = property(lambda self: self._bones_right)
    bones_names = property(lambda self: self._bones_names)


def h36m17_skeleton() -> Skeleton:
    """The 17-joint skeleton both drivers end up with (h36m_lifting.py:649-660, dataset_3dhp.py:132-138)."""
    ops = {j: torch.tensor(_OPS[j], dtype=torch.float) for j in range(1, 17)}
    return Skeleton(H36M17_PARENTS, H36M17_JOINTS_LEFT, H36M17_JOINTS_RIGHT, ops, joints_names=H36M17_NAMES)


def skeleton_tables(skeleton):
    """(parents list, [J][3] operator rows) from our Skeleton or the reference's (operators are a dict keyed by joint)."""
    parents = [int(p) for p in skeleton.parents]
    ops = skeleton.t_pose_operators
    rows = []
    for j in range(len(parents)):
        if j == 0 or ops is None:
            rows.append([0.0, 0.0, 0.0] if j == 0 else list(map(float, _OPS[j])))
        else:
            rows.append([float(v) for v in ops[j]])
    return parents, rows

"""Clip windowing on the device: ``PoseSequenceGenerator`` (hpe/mh_so3_hpe/data/generators.py:45-219) without a
Python DataLoader in the way of a path that lifts ~450 k frames/s.

All sequences are uploaded once, back to back; a batch is one gather kernel (``mp_gather_windows``) driven by the same
index -> (sequence, start frame) table the reference builds (generators.py:83-104), with the reference's replicate padding of the last,
shorter window when ``drop_last`` is False.  The training-time randomness of the generator — random start frames (:121-127), the
occlusion masks of every ``miss_type`` (:157-205) and the "noisy" input perturbation (:206-210) — is sampled ON THE HOST with the very RNG
calls the reference makes, item by item in batch order, so a seeded run reproduces a seeded single-worker reference loader bit for bit;
only the sampled parameters (start frame, flip flag, mask, noise) travel to the device, where the gather applies them.  The generator's
``transform`` hook is carried over for the one transform the drivers install — ``PoseFlip(skeleton, probability)``
(augmentations/transforms.py:8-31, main_h36m_lifting.py:583-595) — as ``flip_probability``; arbitrary callables are not.
(The reference's in-place ``pose_flip`` / noise write through to the dataset arrays when those are float32; the device feed leaves the
uploaded sequences untouched.)"""
import math
from typing import List, Sequence, Tuple

import numpy as np
import torch

from .. import _lib as L
from .. import ops


class DeviceSequenceWindows:
    possible_miss_types_rates = {            # generators.py:50-57
        "no_miss": 0.2,
        "random": 0.2,
        "random_left_arm_right_leg": 0.4,
        "structured_joint": 0.4,
        "structured_frame": 0.2,
    }

    def __init__(self, poses_3d: List[np.ndarray], poses_2d: List[np.ndarray], seq_len: int = 243, drop_last: bool = True,
                 device: str = "cuda", random_start: bool = False, miss_type: str = "no_miss", miss_rate: float = 0.2,
                 noise_sigma: float = 5, flip_probability=None, joints_left=(4, 5, 6, 11, 12, 13), joints_right=(1, 2, 3, 14, 15, 16)):
        assert poses_3d is not None and len(poses_3d) == len(poses_2d)
        self.seq_len = int(seq_len)
        self.drop_last = drop_last
        self.random_start = random_start
        self.miss_type, self.miss_rate, self.noise_sigma = miss_type, miss_rate, noise_sigma
        self.flip_probability = flip_probability
        lengths = [int(p.shape[0]) for p in poses_3d]
        offsets = np.concatenate([[0], np.cumsum(lengths)[:-1]]).astype(np.int64) if lengths else np.zeros(0, np.int64)
        rows = []
        for i, n in enumerate(lengths):                       # generators.py:93-104
            size = n // self.seq_len
            if not drop_last and n % self.seq_len > 0:
                size += 1
            rows += [(int(offsets[i]), n, k * self.seq_len) for k in range(size)]
        self.table = torch.tensor(rows, dtype=torch.int64).reshape(-1, 3)
        self.n_joints, self.in_chans = int(poses_2d[0].shape[1]), int(poses_2d[0].shape[2])
        dev = torch.device(device)
        self.frames_3d = torch.from_numpy(np.concatenate(poses_3d)).float().contiguous().to(dev)
        self.frames_2d = torch.from_numpy(np.concatenate(poses_2d)).float().contiguous().to(dev)
        perm = list(range(self.n_joints))
        for a, b in zip(joints_left, joints_right):
            perm[a], perm[b] = b, a
        self.joint_perm = torch.tensor(perm, dtype=torch.int32, device=dev)

    def __len__(self) -> int:
        return self.table.shape[0]

    _LEFT_ARM_RIGHT_LEG = (1, 2, 3, 11, 12, 13)      # generators.py:184
    _RIGHT_LEG = (1, 2, 3)                           # generators.py:196

    def _sample_mask(self):
        """One item's occlusion pattern (generators.py:157-210), drawn from numpy's global RNG with the reference's calls in the reference's
        order: -> (keep [L, J] float64 or None when nothing is occluded, noise [L, J, C] float64 or None)."""
        n_frames, n_joints = self.seq_len, self.n_joints
        kind, rate = self.miss_type, self.miss_rate
        if kind == "all":                                                  # one pattern per item, with that pattern's own rate
            kind = np.random.choice(list(self.possible_miss_types_rates.keys()))
            rate = self.possible_miss_types_rates[kind]
        if kind == "no_miss":
            return None, None
        if kind == "noisy":
            return None, np.random.normal(0, self.noise_sigma, size=(n_frames, n_joints, self.in_chans))
        keep = np.ones((n_frames, n_joints))
        if kind == "random":                                               # independent (frame, joint) drop-outs
            keep = (np.random.uniform(0.0, 1.0, size=(n_frames, n_joints)) > rate).astype(np.float64)
        elif kind == "random_left_arm_right_leg":                          # two limbs vanish in a random subset of the frames
            frames = np.random.choice(n_frames, size=math.floor(rate * n_frames), replace=False)
            keep[np.ix_(frames, self._LEFT_ARM_RIGHT_LEG)] = 0.0
        elif kind in ("structured_joint", "structured_frame"):             # one contiguous span: the right leg only, or every joint
            span = int(n_frames * rate)
            first = int(np.random.choice(n_frames - span, size=1, replace=False)[0])
            cols = list(self._RIGHT_LEG) if kind == "structured_joint" else slice(None)
            keep[first:first + span, cols] = 0.0
        else:
            raise ValueError(f"Unexpected miss_type: {self.miss_type}")
        return keep, None

    def batch(self, indices: Sequence[int]) -> Tuple[torch.Tensor, torch.Tensor]:
        """-> (pose_2d [B, L, J, in_chans], pose_3d [B, L, J, 3]) on the device, equal to stacking the reference generator's items
        ``[gen[i] for i in indices]`` (with the same torch / numpy seeds when the generator is randomised)."""
        ops._need_cuda(self.frames_2d)
        dev = self.frames_2d.device
        idx = list(indices)
        rows = self.table[idx].clone()
        masks, noises, flips = [], [], []
        for r in range(rows.shape[0]):                                    # the reference's per-item order: start frame, transform, mask
            if self.random_start:
                rows[r, 2] = torch.randint(low=0, high=int(rows[r, 1]) - self.seq_len, size=(1,)).item()   # generators.py:121-127
            if self.flip_probability is not None:
                flips.append(1 if torch.rand(1).item() <= self.flip_probability else 0)                    # transforms.py:26
            m, nz = self._sample_mask()
            masks.append(m)
            noises.append(nz)
        b, t, j = rows.shape[0], self.seq_len, self.n_joints
        mask = noise = None
        if any(m is not None for m in masks):
            mask = torch.from_numpy(np.stack([np.ones((t, j)) if m is None else m for m in masks])).float().to(dev)
        if any(nz is not None for nz in noises):
            noise = torch.from_numpy(np.stack([np.zeros((t, j, self.in_chans)) if nz is None else nz for nz in noises])).double().to(dev)
        flip = torch.tensor(flips, dtype=torch.uint8).to(dev) if flips and any(flips) else None
        rows = rows.contiguous().to(dev)
        out2d = torch.empty((b, t, j, self.in_chans), dtype=torch.float32, device=dev)
        out3d = torch.empty((b, t, j, 3), dtype=torch.float32, device=dev)
        rc = L.load().mp_gather_windows(L.ptr(self.frames_2d), L.ptr(self.frames_3d), L.ptr(rows), L.ptr(mask), L.ptr(noise), L.ptr(flip),
                                        L.ptr(self.joint_perm), L.ptr(out2d), L.ptr(out3d), b, t, j, self.in_chans, L.stream_ptr())
        L.check(rc, "mp_gather_windows")
        ops._count()
        return out2d, out3d

    def batches(self, batch_size: int):
        for s in range(0, len(self), batch_size):
            yield self.batch(range(s, min(len(self), s + batch_size)))

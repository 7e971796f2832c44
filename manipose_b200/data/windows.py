"""Clip windowing on the device: ``PoseSequenceGenerator`` (hpe/mh_so3_hpe/data/generators.py:45-219) without a
Python DataLoader in the way of a path that lifts ~450 k frames/s.

All sequences are uploaded once, back to back; a batch is one gather kernel (``mp_gather_windows``) driven by the same
index -> (sequence, start frame) table the reference builds (generators.py:83-104), with the reference's replicate padding of the last,
shorter window when ``drop_last`` is False.  Random starts, occlusion masks and noise (training-time augmentations of the generator)
are not built: this is the deterministic evaluation / inference feed."""
from typing import List, Sequence, Tuple

import numpy as np
import torch

from .. import _lib as L
from .. import ops


class DeviceSequenceWindows:
    def __init__(self, poses_3d: List[np.ndarray], poses_2d: List[np.ndarray], seq_len: int = 243, drop_last: bool = True,
                 device: str = "cuda"):
        assert poses_3d is not None and len(poses_3d) == len(poses_2d)
        self.seq_len = int(seq_len)
        self.drop_last = drop_last
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
        self.table_dev = self.table.to(dev)

    def __len__(self) -> int:
        return self.table.shape[0]

    def batch(self, indices: Sequence[int]) -> Tuple[torch.Tensor, torch.Tensor]:
        """-> (pose_2d [B, L, J, in_chans], pose_3d [B, L, J, 3]) on the device, equal to stacking the reference generator's items."""
        ops._need_cuda(self.frames_2d)
        idx = torch.as_tensor(list(indices), dtype=torch.int64, device=self.table_dev.device)
        rows = self.table_dev.index_select(0, idx).contiguous()
        b, t, j = rows.shape[0], self.seq_len, self.n_joints
        out2d = torch.empty((b, t, j, self.in_chans), dtype=torch.float32, device=rows.device)
        out3d = torch.empty((b, t, j, 3), dtype=torch.float32, device=rows.device)
        rc = L.load().mp_gather_windows(L.ptr(self.frames_2d), L.ptr(self.frames_3d), L.ptr(rows), L.ptr(out2d), L.ptr(out3d), b, t, j,
                                        self.in_chans, L.stream_ptr())
        L.check(rc, "mp_gather_windows")
        ops._count()
        return out2d, out3d

    def batches(self, batch_size: int):
        for s in range(0, len(self), batch_size):
            yield self.batch(range(s, min(len(self), s + batch_size)))

"""ManiPose: K rotation hypotheses + scores, shared bone lengths, manifold decoder
(hpe/mh_so3_hpe/architectures/rmcl_manifold_mix_ste.py:15-298) on sm_100a kernels.

``forward(x[B,L,17,2]) -> (poses[B,K,L,17,3], scores[B,K,L,1])`` — same names, kwargs and return types as the
reference so it drops in under hpe/main_h36m_lifting.py, hpe/main_3dhp.py and hpe/eval_utils.py.
"""
import os
from typing import Optional

import torch
import torch.nn as nn

from .. import _lib as L
from .. import ops
from .. import train_ops as T
from .manifold_mix_ste import ManifoldMixSTE
from .mix_ste import MixSTE, _version_key


class MCLHead(nn.Module):
    """Parameter holder for rmcl_manifold_mix_ste.py:267-298: LN(eps 1e-5) -> Linear(C -> D+1) -> (rot, Linear(J -> 1))."""

    def __init__(self, embed_dim: int, out_dim: int, num_joints: int, mup: bool = False):
        super().__init__()
        if mup:
            raise NotImplementedError("mup=True is off in every BASELINE config")
        self.norm = nn.LayerNorm(embed_dim)
        self.prediction_head = nn.Linear(embed_dim, out_dim + 1)
        self.score_head = nn.Linear(num_joints, 1)

    def forward(self, x: torch.Tensor):
        """rmcl_manifold_mix_ste.py:290-298 on its own: x [B, L, J, C] (already through Temporal_norm) -> (rotations [B, L, J, D],
        score logit [B, L, 1]).  The models run all K heads in one mp_heads_fwd launch; this is the same kernel with K = 1 and the
        shared post-norm switched off.  Inference only, fp32 like the reference."""
        ops._need_cuda(x)
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            raise NotImplementedError("MCLHead.forward on its own is inference-only (call it under torch.no_grad()); gradients flow "
                                      "through RMCLManifoldMixSTE")
        if x.dim() != 4 or x.shape[2] != self.score_head.in_features:
            raise ValueError(f"MCLHead expects [B, L, {self.score_head.in_features}, C], got {tuple(x.shape)}")
        b, l, j, _ = x.shape
        d = self.prediction_head.out_features - 1
        x = ops._f32(x)
        rot = torch.empty((b, 1, l, j, d), dtype=torch.float32, device=x.device)
        logits = torch.empty((b, 1, l), dtype=torch.float32, device=x.device)
        ops.heads_fwd(x.reshape(-1, x.shape[-1]), None, None, 0.0, self.norm.weight.detach().unsqueeze(0).contiguous(),
                      self.norm.bias.detach().unsqueeze(0).contiguous(), self.prediction_head.weight.detach().unsqueeze(0).contiguous(),
                      self.prediction_head.bias.detach().unsqueeze(0).contiguous(), self.score_head.weight.detach().contiguous(),
                      self.score_head.bias.detach().contiguous(), rot, logits, b, l, 1, d, True)
        return rot[:, 0], logits[:, 0].unsqueeze(-1)


class RMCLRotMixSTE(MixSTE):
    """rmcl_manifold_mix_ste.py:188-264."""

    def __init__(self, num_frame=243, num_joints=17, in_chans=2, out_dim=6, embed_dim=512, depth=8, num_heads=8, mlp_ratio=2.0,
                 qkv_bias=True, qk_scale=None, drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.2, norm_layer=None, n_hyp=5,
                 mup=False):
        # like the reference (:208-223) the parent is built without mup
        super().__init__(num_frame, num_joints, in_chans, out_dim, embed_dim, depth, num_heads, mlp_ratio, qkv_bias, qk_scale,
                         drop_rate, attn_drop_rate, drop_path_rate, norm_layer)
        self.n_hyp = n_hyp
        self.head = nn.ModuleList([MCLHead(embed_dim=embed_dim, out_dim=out_dim, num_joints=num_joints, mup=mup)
                                   for _ in range(self.n_hyp)])
        self._head_key = None
        self._head_stack = None
        self._fold_key = None

    def _stacked_heads(self):
        """K heads as stacked fp32 tensors ([K,C], [K,C], [K,D+1,C], [K,D+1], [K,J], [K]), refreshed on parameter change."""
        params = [p for h in self.head for p in h.parameters()]
        key = _version_key(params)
        if key != self._head_key:
            with torch.no_grad():
                self._head_stack = (
                    torch.stack([h.norm.weight for h in self.head]).contiguous(),
                    torch.stack([h.norm.bias for h in self.head]).contiguous(),
                    torch.stack([h.prediction_head.weight for h in self.head]).contiguous(),
                    torch.stack([h.prediction_head.bias for h in self.head]).contiguous(),
                    torch.stack([h.score_head.weight[0] for h in self.head]).contiguous(),
                    torch.stack([h.score_head.bias[0] for h in self.head]).contiguous(),
                )
            self._head_key = key
        return self._head_stack

    # K heads on the tensor cores (C = 512, fused trunk): False keeps the fp32 CUDA-core projection (mp_heads_fwd)
    heads_on_tensor_cores = True

    def _folded_heads(self):
        """LN_k(y) = yhat * gamma_k + beta_k shares yhat between the heads, so the K heads are ONE Linear with folded parameters
        W_k * gamma_k (16-bit, zero-padded to 128 rows) and W_k beta_k + b_k (fp32); refreshed when a head parameter changes."""
        params = [p for h in self.head for p in h.parameters()]
        key = (self.compute_dtype,) + _version_key(params)
        if key != getattr(self, "_fold_key", None):
            hg, hb, hw, hbias, sw, sb = self._stacked_heads()
            with torch.no_grad():
                k, d1, c = hw.shape
                n_pad = (k * d1 + 127) // 128 * 128
                wf = torch.zeros((n_pad, c), dtype=torch.float32, device=hw.device)
                wf[:k * d1] = (hw * hg[:, None, :]).reshape(k * d1, c)
                bf = torch.zeros(n_pad, dtype=torch.float32, device=hw.device)
                bf[:k * d1] = ((hw * hb[:, None, :]).sum(-1) + hbias).reshape(k * d1)
                self._fold = (ops.cast16(wf, ops.DTYPE_CODE[self.compute_dtype]), bf, sw, sb)
            self._fold_key = key
        return self._fold

    def hypotheses_into(self, x2d: torch.Tensor, n_clips: int, rot: torch.Tensor, logits: torch.Tensor) -> None:
        """One micro-batch: rot fp32 [n_clips, K, L, J, D], logits fp32 [n_clips, K, L]."""
        eps = self.head[0].norm.eps
        if (self.heads_on_tensor_cores and self.embed_dim == 512 and self.fuse_layernorm and self.num_tokens == ops.J
                and all(h.norm.eps == eps for h in self.head)):
            wf16, bf, sw, sb = self._folded_heads()
            with ops.nvtx("manipose.rotations.trunk"):
                xhat = self.trunk(x2d, n_clips, head_norm_eps=eps)
            n_tokens = n_clips * self.num_frame * self.num_tokens
            ws = getattr(self, "_heads_ws", None)
            if ws is None or ws.numel() < n_tokens * wf16.shape[0] or ws.device != xhat.device:
                ws = torch.empty(n_tokens * wf16.shape[0], dtype=torch.float32, device=xhat.device)
                self._heads_ws = ws
            with ops.nvtx("manipose.rotations.heads"):
                ops.heads_fwd16(xhat, wf16, bf, sw, sb, rot, logits, ws, n_clips, self.num_frame, self.n_hyp, self.out_dim, True)
            return
        with ops.nvtx("manipose.rotations.trunk"):
            feat = self.trunk(x2d, n_clips)
        hg, hb, hw, hbias, sw, sb = self._stacked_heads()
        with ops.nvtx("manipose.rotations.heads"):
            ops.heads_fwd(feat, self.Temporal_norm.weight, self.Temporal_norm.bias, self.Temporal_norm.eps, hg, hb, hw, hbias, sw, sb,
                          rot, logits, n_clips, self.num_frame, self.n_hyp, self.out_dim, True)

    def hypotheses_with_grad(self, x: torch.Tensor):
        """Differentiable K heads over the whole batch -> (rot [B,K,L,J,D], logits [B,K,L]).

        LN_k(y) = yhat * gamma_k + beta_k shares yhat between the heads, so all K heads are one Linear with folded parameters
        W_k * gamma_k, W_k beta_k + b_k, zero-padded to 128 outputs for the tensor-core Linear kernel (fp32 output); the score
        Linear(J -> 1) is a J-term dot product per frame and head (train_ops.FoldedHeadsFn: forward and hand-unfolded backward)."""
        b, l, j, _ = x.shape
        c = self.embed_dim
        feat = self.trunk_autograd(x, b)
        y = T.layer_norm(feat, self.Temporal_norm.weight, self.Temporal_norm.bias, self.Temporal_norm.eps)
        one, zero = torch.ones(c, dtype=torch.float32, device=x.device), torch.zeros(c, dtype=torch.float32, device=x.device)
        yhat = T.layer_norm(y, one, zero, self.head[0].norm.eps, out16=ops.DTYPE_CODE[self.compute_dtype])
        rot, logits = T.FoldedHeadsFn.apply(yhat, list(self.head), b, l, j, self.out_dim, self.head[0].norm.weight)
        return rot, logits

    def forward(self, x: torch.Tensor):
        """-> (hypothesis [B,K,L,J,D], scores [B,K,L,1])  (rmcl_manifold_mix_ste.py:239-264)."""
        ops._need_cuda(x)
        b, l, j, _ = self._check_input(x)
        x = ops._f32(x)
        if self._grad_mode():
            rot, logits = self.hypotheses_with_grad(x)
            return rot, ops.softmax_hyp(logits.contiguous()).unsqueeze(-1)
        rot = torch.empty((b, self.n_hyp, l, j, self.out_dim), dtype=torch.float32, device=x.device)
        logits = torch.empty((b, self.n_hyp, l), dtype=torch.float32, device=x.device)
        mb = self.clips_per_micro_batch()
        for s in range(0, b, mb):
            n = min(mb, b - s)
            self.hypotheses_into(x[s:s + n], n, rot[s:s + n], logits[s:s + n])
        return rot, ops.softmax_hyp(logits).unsqueeze(-1)


_BRANCH_STREAMS = {}


def _branch_stream(dev, priority: int = 0):
    key = torch.device(dev).index if torch.device(dev).index is not None else torch.cuda.current_device()
    st = _BRANCH_STREAMS.get((key, priority))
    if st is None:
        st = _BRANCH_STREAMS[(key, priority)] = torch.cuda.Stream(device=key, priority=priority)
    return st


class RMCLManifoldMixSTE(ManifoldMixSTE):
    """rmcl_manifold_mix_ste.py:15-185."""

    # training: the bone-length backbone runs on a second stream, concurrently with the rotations backbone
    overlap_branches = os.environ.get("MANIPOSE_BRANCH_STREAM", "1") != "0"
    # inference: the rotations backbone on a HIGH-priority stream of its own, so that its persistent kernels (one CTA per SM, statically
    # assigned tiles: a CTA that starts late delays the whole launch) are placed before the pending CTAs of the bone-length stream
    trunk_priority = os.environ.get("MANIPOSE_TRUNK_PRIORITY", "0") != "0"

    def __init__(self, skeleton, num_frame: int = 243, num_joints: int = 17, num_bones: int = 16, in_chans: int = 2,
                 rot_rep_dim: int = 6, embed_dim_rot: int = 512, depth_rot: int = 8, num_heads_rot: int = 8, embed_dim_seg: int = 128,
                 depth_seg: int = 2, num_heads_seg: int = 8, mlp_ratio: float = 2.0, qkv_bias: bool = True, qk_scale: float = None,
                 drop_rate: float = 0.0, attn_drop_rate: float = 0.0, drop_path_rate: float = 0.2, norm_layer: nn.Module = None,
                 n_hyp: int = 5, mup: bool = False):
        super().__init__(skeleton=skeleton, num_frame=num_frame, num_joints=num_joints, num_bones=num_bones, in_chans=in_chans,
                         rot_rep_dim=rot_rep_dim, embed_dim_rot=embed_dim_rot, depth_rot=depth_rot, num_heads_rot=num_heads_rot,
                         embed_dim_seg=embed_dim_seg, depth_seg=depth_seg, num_heads_seg=num_heads_seg, mlp_ratio=mlp_ratio,
                         qkv_bias=qkv_bias, qk_scale=qk_scale, drop_rate=drop_rate, attn_drop_rate=attn_drop_rate,
                         drop_path_rate=drop_path_rate, norm_layer=norm_layer, mup=mup)
        self.n_hyp = n_hyp
        self.rotations_module = RMCLRotMixSTE(num_frame=num_frame, num_joints=num_joints, in_chans=in_chans, out_dim=rot_rep_dim,
                                              embed_dim=embed_dim_rot, depth=depth_rot, num_heads=num_heads_rot, mlp_ratio=mlp_ratio,
                                              qkv_bias=qkv_bias, qk_scale=qk_scale, drop_rate=drop_rate,
                                              attn_drop_rate=attn_drop_rate, drop_path_rate=drop_path_rate, norm_layer=norm_layer,
                                              n_hyp=n_hyp, mup=mup)
        self._scratch = None

    def forward(self, x: torch.Tensor):
        """Per micro-batch of clips: rotations trunk -> K heads, segments trunk -> bone lengths, then the fused decoder
        (Gram-Schmidt + FK + softmax over K) writes straight into the [B,K,L,J,3] / [B,K,L,1] outputs."""
        ops._need_cuda(x)
        rm, sm = self.rotations_module, self.segments_module
        b, l, j, _ = rm._check_input(x)
        ops.set_skeleton(*self.decoder._tables)
        x = ops._f32(x)
        k, d = self.n_hyp, rm.out_dim
        dev = x.device
        if rm._grad_mode() or sm._grad_mode():
            # differentiable path (training): whole batch at once, activations kept for the backward sweep
            if self.overlap_branches:
                # the bone-length backbone (C = 128: ~150 short kernels) on its own stream, under the rotations backbone; autograd runs
                # its backward on that stream too.  MANIPOSE_BRANCH_STREAM=0: one after the other.
                main, side = torch.cuda.current_stream(dev), _branch_stream(dev)
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    bones = sm.bone_lengths_with_grad(x)
                rot, logits = rm.hypotheses_with_grad(x)
                main.wait_stream(side)
                bones.record_stream(main)
            else:
                rot, logits = rm.hypotheses_with_grad(x)
                bones = sm.bone_lengths_with_grad(x)
            poses = ops.decode(rot.reshape(b * k * l, j, d), bones, None, b, k, l, d, self.decoder.exact).view(b, k, l, j, 3)
            return poses, ops.softmax_hyp(logits.contiguous()).unsqueeze(-1)
        poses = torch.empty((b, k, l, j, 3), dtype=torch.float32, device=dev)
        scores = torch.empty((b, k, l, 1), dtype=torch.float32, device=dev)
        mb = min(rm.clips_per_micro_batch(), max(b, 1))
        if self._scratch is None or self._scratch[0].shape[0] < mb or self._scratch[0].device != dev:
            self._scratch = (torch.empty((mb, k, l, j, d), dtype=torch.float32, device=dev),
                             torch.empty((mb, k, l), dtype=torch.float32, device=dev),
                             torch.empty((mb, sm.num_bones), dtype=torch.float32, device=dev))
        rot, logits, bones = self._scratch
        caller = torch.cuda.current_stream(dev)
        side = _branch_stream(dev) if self.overlap_branches else None
        if side is not None and self.trunk_priority:
            main = _branch_stream(dev, -1)
            main.wait_stream(caller)
            with torch.cuda.stream(main):
                self._lift_micro_batches(x, b, mb, main, side, rot, logits, bones, poses, scores)
            caller.wait_stream(main)
        else:
            self._lift_micro_batches(x, b, mb, caller, side, rot, logits, bones, poses, scores)
        return poses, scores

    def _lift_micro_batches(self, x, b, mb, main, side, rot, logits, bones, poses, scores) -> None:
        rm, sm = self.rotations_module, self.segments_module
        k, l, d = self.n_hyp, rm.num_frame, rm.out_dim
        lib = L.load()
        for s in range(0, b, mb):
            n = min(mb, b - s)
            xs = x[s:s + n]
            if side is not None:
                # the bone-length backbone of this micro-batch on the second stream (after the previous micro-batch's decoder, which
                # reads `bones`), under the rotations backbone; joined before the decoder
                side.wait_stream(main)
                with torch.cuda.stream(side), ops.nvtx("manipose.segments"):
                    sm.bone_lengths_into(xs, n, bones[:n])
                rm.hypotheses_into(xs, n, rot[:n], logits[:n])
                main.wait_stream(side)
            else:
                rm.hypotheses_into(xs, n, rot[:n], logits[:n])
                with ops.nvtx("manipose.segments"):
                    sm.bone_lengths_into(xs, n, bones[:n])
            with ops.nvtx("manipose.decoder"):
                rc = lib.mp_decoder_fwd(L.ptr(rot), L.ptr(bones), None, L.ptr(logits), L.ptr(poses[s:s + n]), L.ptr(scores[s:s + n]),
                                        n, k, l, d, L.MP_DEC_EXACT if self.decoder.exact else L.MP_DEC_FAST, L.stream_ptr())
                L.check(rc, "mp_decoder_fwd")

    def concat_hyp_and_scores(self, hypothesis: torch.Tensor, scores: torch.Tensor) -> torch.Tensor:
        """rmcl_manifold_mix_ste.py:108-119 -> [B,K,L,J,4]."""
        return torch.cat((hypothesis, scores.unsqueeze(3).expand(-1, -1, -1, self.num_joints, -1)), dim=-1)

    def poses_from_hyp_idx(self, hypothesis: torch.Tensor, hyp_indices: torch.Tensor) -> torch.Tensor:
        """rmcl_manifold_mix_ste.py:121-139: gather hypothesis hyp_indices[b,l] -> [B,L,J,3] (device-side gather)."""
        ops._need_cuda(hypothesis)
        b, k, l, j, _ = hypothesis.shape
        idx = hyp_indices.to(hypothesis.device)[:, None, :, None, None].expand(b, 1, l, j, 3)
        return hypothesis.gather(1, idx)[:, 0]

    def aggregate(self, hypothesis: torch.Tensor, scores: torch.Tensor = None, mode: str = "weighted_ave",
                  ground_truth: Optional[torch.Tensor] = None) -> torch.Tensor:
        """rmcl_manifold_mix_ste.py:141-185 (oracle mode returns a tuple, like the reference)."""
        if mode == "best_score":
            assert scores is not None, "Scores required to compute hypothesis with best confidence."
            pose, _, _ = ops.aggregate(hypothesis, scores.reshape(scores.shape[:3]), None, L.MP_AGG_BEST_SCORE)
            return pose
        elif mode == "weighted_ave":
            assert scores is not None, "Scores required to compute weighted hypothesis average."
            pose, _, _ = ops.aggregate(hypothesis, scores.reshape(scores.shape[:3]), None, L.MP_AGG_WEIGHTED_AVE)
            return pose
        elif mode == "oracle":
            assert ground_truth is not None, "Ground-truth required to compute best hypothesis."
            pose, val, _ = ops.aggregate(hypothesis, None, ground_truth, L.MP_AGG_ORACLE)
            return val, pose
        else:
            raise ValueError(f"Only best_score and weighted_ave modes are implemented.Got {mode}.")

"""Single-hypothesis constrained model and the bone-length backbone
(hpe/mh_so3_hpe/architectures/manifold_mix_ste.py:10-154)."""
import torch
import torch.nn as nn

from .. import ops
from .. import train_ops as T
from .mix_ste import MixSTE
from .pose_decoder import PoseDecoder


class BonesMixSTE(MixSTE):
    """manifold_mix_ste.py:91-154: Linear(J*in -> S*C) per frame, MixSTE over S segment tokens, LN + Linear(C -> 1), mean over time."""

    def __init__(self, num_frame=243, num_joints=17, num_bones=16, in_chans=2, out_dim=1, embed_dim=128, depth=2, num_heads=8,
                 mlp_ratio=2, qkv_bias=True, qk_scale=None, drop_rate=0, attn_drop_rate=0, drop_path_rate=0.2, norm_layer=None,
                 mup=False):
        super().__init__(num_frame=num_frame, num_joints=num_bones, in_chans=in_chans, out_dim=out_dim, embed_dim=embed_dim,
                         depth=depth, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale,
                         drop_rate=drop_rate, attn_drop_rate=attn_drop_rate, drop_path_rate=drop_path_rate, norm_layer=norm_layer,
                         mup=mup)
        self.num_joints = num_joints
        self.num_bones = num_bones
        self.embed_dim = embed_dim
        self.Spatial_patch_to_embedding = nn.Identity()
        self.joints_to_segments_proj = nn.Linear(in_features=num_joints * in_chans, out_features=num_bones * embed_dim)
        self._head_ws = None

    def _embed(self, x2d, n_clips, n_frames, x, h):
        blk0 = self.STEblocks[0]
        ops.embed_segments(x2d, self.joints_to_segments_proj.weight, self.joints_to_segments_proj.bias, self.Spatial_pos_embed,
                           blk0.norm1.weight, blk0.norm1.bias, blk0.norm1.eps, x, h, n_clips * n_frames,
                           self.joints_to_segments_proj.in_features, self.num_bones, self.embed_dim,
                           ops.DTYPE_CODE[self.compute_dtype])

    def bone_lengths_into(self, x2d: torch.Tensor, n_clips: int, out: torch.Tensor) -> None:
        """One micro-batch: x2d fp32 [n_clips, L, J, in] -> out fp32 [n_clips, S] (signed, no activation)."""
        feat = self.trunk(x2d, n_clips)
        n_tokens = feat.shape[0]
        if self._head_ws is None or self._head_ws.numel() < n_tokens or self._head_ws.device != feat.device:
            self._head_ws = torch.empty(n_tokens, dtype=torch.float32, device=feat.device)
        norm, lin = self.head[0], self.head[1]
        ops.bones_head(feat, self.Temporal_norm.weight, self.Temporal_norm.bias, self.Temporal_norm.eps, norm.weight, norm.bias,
                       lin.weight, lin.bias, out, n_clips, self.num_frame, self.num_bones, self.embed_dim, self._head_ws)

    def _embed_backward(self, x2d, dx0, n_clips):
        """Gradients of joints_to_segments_proj and Spatial_pos_embed (manifold_mix_ste.py:139-148) from dx0 [frames*S, C]."""
        lin = self.joints_to_segments_proj
        n_rows = n_clips * self.num_frame
        T.small_wgrad(dx0.view(n_rows, self.num_bones * self.embed_dim), x2d.reshape(n_rows, lin.in_features), T.grad_of(lin.weight),
                      T.grad_of(lin.bias))
        T.group_rowsum(dx0, T.grad_of(self.Spatial_pos_embed), 1, self.num_bones)

    def bone_lengths_with_grad(self, x: torch.Tensor) -> torch.Tensor:
        """Differentiable bone lengths [B, S] of the whole batch (Temporal_norm -> head LayerNorm(1e-5) -> Linear(C -> 1) -> mean over
        time); the head Linear is zero-padded to 128 outputs so that it runs on the tensor-core Linear kernel with fp32 output."""
        b, l = x.shape[:2]
        feat = self.trunk_autograd(x, b)
        norm, lin = self.head[0], self.head[1]
        y = T.layer_norm(feat, self.Temporal_norm.weight, self.Temporal_norm.bias, self.Temporal_norm.eps)
        z = T.layer_norm(y, norm.weight, norm.bias, norm.eps, out16=ops.DTYPE_CODE[self.compute_dtype])
        w = torch.cat([lin.weight, lin.weight.new_zeros(127, self.embed_dim)])
        bias = torch.cat([lin.bias, lin.bias.new_zeros(127)])
        return T.linear_f32(z, w, bias)[:, 0].reshape(b, l, self.num_bones).mean(dim=1)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        ops._need_cuda(x)
        b, l, j, _ = self._check_input(x)
        x = ops._f32(x)
        if self._grad_mode():
            return self.bone_lengths_with_grad(x).unsqueeze(-1)
        out = torch.empty((b, self.num_bones), dtype=torch.float32, device=x.device)
        mb = self.clips_per_micro_batch()
        for s in range(0, b, mb):
            n = min(mb, b - s)
            self.bone_lengths_into(x[s:s + n], n, out[s:s + n])
        return out.unsqueeze(-1)


class ManifoldMixSTE(nn.Module):
    """manifold_mix_ste.py:10-88: rotations backbone + bone-length backbone + manifold decoder (one hypothesis)."""

    def __init__(self, skeleton, num_frame=243, num_joints=17, num_bones=16, in_chans=2, rot_rep_dim=6, embed_dim_rot=512,
                 depth_rot=8, num_heads_rot=8, embed_dim_seg=128, depth_seg=2, num_heads_seg=8, mlp_ratio=2.0, qkv_bias=True,
                 qk_scale=None, drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.2, norm_layer=None, mup=False):
        super().__init__()
        self.num_joints = num_joints
        self.rotations_module = MixSTE(num_frame=num_frame, num_joints=num_joints, in_chans=in_chans, out_dim=rot_rep_dim,
                                       embed_dim=embed_dim_rot, depth=depth_rot, num_heads=num_heads_rot, mlp_ratio=mlp_ratio,
                                       qkv_bias=qkv_bias, qk_scale=qk_scale, drop_rate=drop_rate, attn_drop_rate=attn_drop_rate,
                                       drop_path_rate=drop_path_rate, norm_layer=norm_layer, mup=mup)
        self.segments_module = BonesMixSTE(num_frame=num_frame, num_joints=num_joints, num_bones=num_bones, in_chans=in_chans,
                                           out_dim=1, embed_dim=embed_dim_seg, depth=depth_seg, num_heads=num_heads_seg,
                                           mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale, drop_rate=drop_rate,
                                           attn_drop_rate=attn_drop_rate, drop_path_rate=drop_path_rate, norm_layer=norm_layer,
                                           mup=mup)
        self.decoder = PoseDecoder(skeleton=skeleton, rot_rep_dim=rot_rep_dim)

    def set_compute_dtype(self, name: str) -> "ManifoldMixSTE":
        """"bf16" (BASELINE config 3) or "fp16" (same speed, 3 more mantissa bits) for both backbones."""
        if name not in ("bf16", "fp16"):
            raise ValueError(f"compute dtype must be 'bf16' or 'fp16', got {name!r}")
        self.rotations_module.compute_dtype = name
        self.segments_module.compute_dtype = name
        return self

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        b, l, _, _ = x.shape
        rotations = self.rotations_module(x)            # (B, L, J, D)
        bones_lengths = self.segments_module(x)         # (B, S, 1)
        poses = self.decoder(rotations_repr=rotations.reshape(b * l, self.num_joints, -1), bones_lengths_repr=bones_lengths,
                             root_positions=None)
        return poses.reshape(b, l, self.num_joints, 3)

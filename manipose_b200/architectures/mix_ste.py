"""MixSTE backbone behind the reference's nn.Module surface (hpe/mh_so3_hpe/architectures/mix_ste.py:12-368).

The module tree (``STEblocks.{i}.attn.qkv`` ...) exists so that ``state_dict()`` / ``load_state_dict()`` /
``Adam(model.parameters())`` interoperate with the reference (290 tensors).  The arithmetic does not
go through ``nn.Linear``: ``MixSTE.trunk`` drives the sm_100a kernels of libmanipose_sm100.so over ONE activation
layout, [clip, frame, token, C] — fp32 residual stream, 16-bit (``compute_dtype``: "bf16" | "fp16") tensor-core operands:

    embed (+spatial pos-embed, +norm1)                                  mp_embed_joints / mp_embed_segments
    per block:  qkv GEMM -> attention (spatial | temporal) -> proj GEMM + residual
                -> norm2 -> fc1 GEMM + GELU -> fc2 GEMM + residual      mp_linear (tcgen05/TMEM/TMA), mp_attention
                -> shared post-norm (+temporal pos-embed) fused with the next block's norm1      mp_layernorm

The reference's "(B L) J C <-> (B J) L C" rearranges (mix_ste.py:131,144,167,171,184) are strided reads inside the
temporal attention kernel; nothing is transposed in memory.
"""
from functools import partial
from math import sqrt
from typing import Dict, List, Optional, Tuple

import os

import torch
import torch.nn as nn

from .. import _lib as L
from .. import ops
from .. import train_ops as T


class DropPath(nn.Module):
    """Stochastic depth holder (timm 0.9.16 semantics, used at mix_ste.py:334-336).  Identity in eval mode; the kernels
    take the per-row keep mask as an explicit input in training (SURVEY.md §7 hard parts)."""

    def __init__(self, drop_prob: float = 0.0, scale_by_keep: bool = True):
        super().__init__()
        self.drop_prob = drop_prob
        self.scale_by_keep = scale_by_keep

    def extra_repr(self):
        return f"drop_prob={round(self.drop_prob, 3):0.3f}"


# ---- standalone entry points of the block pieces (mix_ste.py:194-368).  The models run these fused inside MixSTE.trunk; the
# reference also lets a caller invoke a Block / Attention / Mlp on its own, so these do too: inference only (no tape is kept),
# fp32 in and out like the reference, the same kernels as the trunk (16-bit tensor-core operands, fp32 accumulation).
def _standalone_input(module, name: str, x: torch.Tensor, c: int):
    ops._need_cuda(x)
    if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in module.parameters())):
        raise NotImplementedError(
            f"{name}.forward on its own is inference-only in manipose_b200 (call it under torch.no_grad()); gradients flow through "
            "the enclosing MixSTE / ManifoldMixSTE / RMCLManifoldMixSTE, whose backward sweep is hand-written")
    if x.shape[-1] != c:
        raise RuntimeError(f"{name}: expected {c} channels, got input of shape {tuple(x.shape)}")
    return ops._f32(x).reshape(-1, c)


def _w16(w: torch.Tensor, code: int) -> torch.Tensor:
    return ops.cast16(w.detach(), code)


def _bias(lin: nn.Linear) -> torch.Tensor:
    return lin.bias.detach() if lin.bias is not None else torch.zeros(lin.out_features, dtype=torch.float32, device=lin.weight.device)


def _mlp_into(mlp, h16: torch.Tensor, x_acc: torch.Tensor, code: int) -> None:
    """x_acc (fp32 [M, C]) += fc2(gelu(fc1(h16)))."""
    hid = torch.empty((h16.shape[0], mlp.fc1.out_features), dtype=h16.dtype, device=h16.device)
    ops.linear(h16, _w16(mlp.fc1.weight, code), _bias(mlp.fc1), hid, L.MP_EPI_GELU)
    ops.linear(hid, _w16(mlp.fc2.weight, code), _bias(mlp.fc2), x_acc, L.MP_EPI_RESIDUAL, resid=x_acc)


def _attention_into(attn, h16: torch.Tensor, x_acc: torch.Tensor, n_seq: int, seq_len: int, code: int) -> None:
    """x_acc (fp32 [n_seq * seq_len, C]) += proj(softmax(q k^T / sqrt(hd)) v), sequences of seq_len consecutive rows."""
    c = attn.qkv.in_features
    if seq_len > 256:
        raise NotImplementedError(f"Attention: sequences longer than 256 tokens are not built (got {seq_len})")
    qkv = torch.empty((h16.shape[0], 3 * c), dtype=h16.dtype, device=h16.device)
    ops.linear(h16, _w16(attn.qkv.weight, code), _bias(attn.qkv), qkv, L.MP_EPI_BIAS)
    o = torch.empty_like(h16)
    ops.attention(qkv, o, n_seq, seq_len, 1, c, attn.num_heads, L.MP_ATTN_TEMPORAL)   # one "track" per sequence
    ops.linear(o, _w16(attn.proj.weight, code), _bias(attn.proj), x_acc, L.MP_EPI_RESIDUAL, resid=x_acc)


class Mlp(nn.Module):
    """mix_ste.py:194-222 (fc1 -> exact GELU -> fc2)."""
    compute_dtype = "bf16"

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0, changedim=False,
                 currentdim=0, depth=0):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)

    def forward(self, x):
        c = self.fc1.in_features
        x2d = _standalone_input(self, "Mlp", x, c)
        code = ops.DTYPE_CODE[self.compute_dtype]
        out = torch.zeros((x2d.shape[0], self.fc2.out_features), dtype=torch.float32, device=x.device)
        _mlp_into(self, ops.cast16(x2d, code), out, code)
        return out.reshape(*x.shape[:-1], self.fc2.out_features)


class Attention(nn.Module):
    """mix_ste.py:225-282: x [B, N, C] -> [B, N, C]."""
    compute_dtype = "bf16"

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0.0, proj_drop=0.0, comb=False, vis=False,
                 mup=False):
        super().__init__()
        self.num_heads = num_heads
        head_dim = dim // num_heads
        default_scale = 1 / head_dim if mup else head_dim ** -0.5
        self.scale = qk_scale or default_scale
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        self.comb = comb
        self.vis = vis

    def forward(self, x, vis=False):
        c = self.qkv.in_features
        x2d = _standalone_input(self, "Attention", x, c)
        if x.dim() != 3:
            raise ValueError(f"Attention expects [B, N, C], got {tuple(x.shape)}")
        code = ops.DTYPE_CODE[self.compute_dtype]
        out = torch.zeros_like(x2d)
        _attention_into(self, ops.cast16(x2d, code), out, x.shape[0], x.shape[1], code)
        return out.reshape(x.shape)


class Block(nn.Module):
    """mix_ste.py:285-368 (pre-LN attention + MLP, residual_scale = 1 without muP): x [B, N, C] -> [B, N, C]."""
    compute_dtype = "bf16"

    def __init__(self, dim, num_heads, mlp_ratio=4.0, attention=Attention, qkv_bias=False, qk_scale=None, drop=0.0, attn_drop=0.0,
                 drop_path=0.0, act_layer=nn.GELU, norm_layer=nn.LayerNorm, comb=False, changedim=False, currentdim=0, depth=0,
                 vis=False, mup=False):
        super().__init__()
        if mup:
            raise NotImplementedError("mup=True (muP residual scaling / readouts) is off in every BASELINE config (config.yaml:52)")
        if changedim or comb:
            raise NotImplementedError("changedim / comb blocks are never instantiated by the reference models")
        self.changedim, self.currentdim, self.depth = changedim, currentdim, depth
        self.norm1 = norm_layer(dim)
        self.attn = attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop, proj_drop=drop,
                              comb=comb, vis=vis, mup=mup)
        self.residual_scale = 1.0
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)
        self.vis = vis

    def forward(self, x, vis=False):
        c = self.norm1.normalized_shape[0]
        x2d = _standalone_input(self, "Block", x, c)
        if x.dim() != 3:
            raise ValueError(f"Block expects [B, N, C], got {tuple(x.shape)}")
        if self.training and isinstance(self.drop_path, DropPath) and self.drop_path.drop_prob > 0.0:
            raise NotImplementedError("Block.forward on its own has no stochastic depth: call .eval() (the training trunk applies DropPath)")
        code = ops.DTYPE_CODE[self.compute_dtype]
        acc = x2d.clone()                                     # the fp32 residual stream of this block
        h = torch.empty(acc.shape, dtype=ops.TORCH_DTYPE[code], device=acc.device)
        ops.layernorm(acc, None, h, ln=(self.norm1.weight, self.norm1.bias), ln_eps=self.norm1.eps, dtype=code)
        _attention_into(self.attn, h, acc, x.shape[0], x.shape[1], code)
        ops.layernorm(acc, None, h, ln=(self.norm2.weight, self.norm2.bias), ln_eps=self.norm2.eps, dtype=code)
        _mlp_into(self.mlp, h, acc, code)
        return acc.reshape(x.shape)


def _version_key(params) -> Tuple:
    return tuple((p.data_ptr(), p._version) for p in params)


class _TrunkFn(torch.autograd.Function):
    """Autograd node of a whole MixSTE trunk.  Parameter gradients are accumulated into ``p.grad`` by the backward sweep itself
    (like fused weight-gradient accumulation in Megatron-style trainers), so the only tensor input is an ``anchor`` parameter
    that makes the output require grad; its returned gradient is None."""

    @staticmethod
    def forward(ctx, module, x2d, n_clips, anchor):
        feat, tape = module._train_forward(x2d, n_clips)
        ctx.module, ctx.tape = module, tape
        return feat

    @staticmethod
    def backward(ctx, dfeat):
        ctx.module._train_backward(ctx.tape, ops._f32(dfeat))
        ctx.tape = None
        return None, None, None, None


_WGRAD_STREAMS: Dict[int, "torch.cuda.Stream"] = {}   # one second stream per device for the weight-gradient GEMMs (MixSTE._train_backward)


class MixSTE(nn.Module):
    """mix_ste.py:12-191.  ``forward(x[B,L,J,in_chans]) -> [B,L,J,out_dim]``."""

    # tokens per micro-batch of the trunk (128 clips of 243 x 17); activations of one micro-batch are 12*C bytes per token (3.2 GB).
    # Larger micro-batches amortise launches and tile-wave tails (measured at 1024 clips: 339k / 365k / 383k frames/s at 8 / 16 / 32
    # clips on the first GEMM kernels; 428k / 435k / 439k at 32 / 64 / 128 clips now)
    micro_batch_tokens = 530000
    # 16-bit format of everything that feeds a tensor-core contraction ("bf16": BASELINE config 3; "fp16": same speed and
    # bytes, 3 more mantissa bits — needed for the 0.05 mm end-to-end MPJPE gate)
    compute_dtype = "bf16"
    # fuse the residual Linear with the LayerNorms that follow it (C = 512 only); False keeps the separate kernels
    fuse_layernorm = True
    # fc1 -> GELU -> fc2 + residual + LayerNorms in ONE launch with the hidden activation kept on chip (mp_mlp_ln; C = 512, hidden =
    # 1024 only).  Bit-identical to the two launches it replaces, but SLOWER on B200 (1784 vs 1367 us per 528,768 tokens): the fc2
    # accumulator and two fc1 chunk accumulators fill TMEM, so the LayerNorm epilogue cannot overlap the next tile's fc2 (DESIGN.md
    # §5, negative results).  Off by default; kept for A/B measurements (MANIPOSE_FUSE_MLP=1 turns it on).
    fuse_mlp = os.environ.get("MANIPOSE_FUSE_MLP", "0") == "1"
    # training: weight-gradient GEMMs on a second stream, under the data-gradient kernels that follow them (MANIPOSE_WGRAD_STREAM=0: in line)
    overlap_wgrad = os.environ.get("MANIPOSE_WGRAD_STREAM", "1") != "0"

    @staticmethod
    def _side_stream(dev):
        key = torch.device(dev).index if torch.device(dev).index is not None else torch.cuda.current_device()
        st = _WGRAD_STREAMS.get(key)
        if st is None:
            st = _WGRAD_STREAMS[key] = torch.cuda.Stream(device=key)
        return st

    def __init__(self, num_frame=243, num_joints=17, in_chans=2, out_dim=3, embed_dim=512, depth=8, num_heads=8, mlp_ratio=2.0,
                 qkv_bias=True, qk_scale=None, drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.2, norm_layer=None, mup=False):
        super().__init__()
        if mup:
            raise NotImplementedError("mup=True is off in every BASELINE config (hpe/conf/config.yaml:52)")
        if drop_rate != 0.0 or attn_drop_rate != 0.0:
            raise NotImplementedError("drop_rate / attn_drop_rate are never set by the reference drivers (SURVEY.md §A.2)")
        if qk_scale is not None:
            raise NotImplementedError("qk_scale overrides are not built; the kernels use head_dim ** -0.5 (mix_ste.py:240-244)")
        norm_layer = norm_layer or partial(nn.LayerNorm, eps=1e-6)
        self.embed_dim = embed_dim
        self.num_frame = num_frame
        self.num_tokens = num_joints
        self.num_heads = num_heads
        self.in_chans = in_chans
        self.out_dim = out_dim
        self.Spatial_patch_to_embedding = nn.Linear(in_chans, embed_dim)
        self.Spatial_pos_embed = nn.Parameter(torch.zeros(1, num_joints, embed_dim))
        self.Temporal_pos_embed = nn.Parameter(torch.zeros(1, num_frame, embed_dim))
        self.pos_drop = nn.Dropout(p=drop_rate)
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, depth)]
        self.block_depth = depth
        self.STEblocks = nn.ModuleList([
            Block(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale, drop=drop_rate,
                  attn_drop=attn_drop_rate, drop_path=dpr[i], norm_layer=norm_layer, depth=0, mup=mup) for i in range(depth)])
        self.TTEblocks = nn.ModuleList([
            Block(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale, drop=drop_rate,
                  attn_drop=attn_drop_rate, drop_path=dpr[i], norm_layer=norm_layer, comb=False, changedim=False,
                  currentdim=i + 1, depth=depth, mup=mup) for i in range(depth)])
        self.Spatial_norm = norm_layer(embed_dim)
        self.Temporal_norm = norm_layer(embed_dim)
        self.head = nn.Sequential(nn.LayerNorm(embed_dim), nn.Linear(embed_dim, out_dim))
        self._shadow: Dict[str, torch.Tensor] = {}
        self._shadow_key = None
        self._ws = None

    # ------------------------------------------------------------------------------------------ weight shadows
    def _gemm_params(self):
        out = []
        for blk in list(self.STEblocks) + list(self.TTEblocks):
            out += [blk.attn.qkv.weight, blk.attn.proj.weight, blk.mlp.fc1.weight, blk.mlp.fc2.weight]
        return out

    def _shadow_weights(self) -> List[torch.Tensor]:
        """16-bit shadows of the GEMM weights, refreshed (ONE launch for all of them, mp_refresh_shadows) when a parameter changed
        (optimizer.step / load_state_dict).  The buffers and the device-side table of pointers are allocated once."""
        params = self._gemm_params()
        dt = ops.DTYPE_CODE[self.compute_dtype]
        key = (dt,) + _version_key(params)
        if key != self._shadow_key:
            self._refresh_shadows(params, dt)
            self._shadow_key = key
        return self._shadow_list

    def _refresh_shadows(self, params, dt) -> None:
        want_t = getattr(self, "_shadow_want_t", False)
        alloc_key = (dt, want_t) + tuple(p.data_ptr() for p in params)
        if getattr(self, "_shadow_alloc_key", None) != alloc_key:
            td, dev = ops.TORCH_DTYPE[dt], params[0].device
            self._shadow_list = [torch.empty(p.shape, dtype=td, device=dev) for p in params]
            self._shadow_t_list = [torch.empty((p.shape[1], p.shape[0]), dtype=td, device=dev) for p in params] if want_t else None
            rows = [[p.data_ptr(), s.data_ptr(), self._shadow_t_list[i].data_ptr() if want_t else 0, p.shape[0], p.shape[1]]
                    for i, (p, s) in enumerate(zip(params, self._shadow_list))]
            self._shadow_table = torch.tensor(rows, dtype=torch.int64).to(dev)
            self._shadow_max_tiles = max(p.shape[0] * p.shape[1] // 4096 for p in params)
            self._shadow_alloc_key = alloc_key
        rc = L.load().mp_refresh_shadows(L.ptr(self._shadow_table), len(params), self._shadow_max_tiles, dt, L.stream_ptr())
        L.check(rc, "mp_refresh_shadows")
        ops._count()

    def _workspace(self, n_tokens: int, device) -> Dict[str, torch.Tensor]:
        c = self.embed_dim
        hidden = self.STEblocks[0].mlp.fc1.out_features
        wide = max(3 * c, hidden)
        td = ops.TORCH_DTYPE[ops.DTYPE_CODE[self.compute_dtype]]
        if self._ws is None or self._ws["x"].shape[0] < n_tokens or self._ws["x"].device != device or self._ws["h"].dtype != td:
            self._ws = {
                "x": torch.empty((n_tokens, c), dtype=torch.float32, device=device),   # residual stream (fp32)
                "h": torch.empty((n_tokens, c), dtype=td, device=device),              # normalised / attention out
                "wide": torch.empty((n_tokens, wide), dtype=td, device=device),        # qkv, then the MLP hidden
            }
        return self._ws

    # ------------------------------------------------------------------------------------------ fused trunk
    def _embed(self, x2d: torch.Tensor, n_clips: int, n_frames: int, x, h):
        blk0 = self.STEblocks[0]
        ops.embed_joints(x2d, self.Spatial_patch_to_embedding.weight, self.Spatial_patch_to_embedding.bias, self.Spatial_pos_embed,
                         blk0.norm1.weight, blk0.norm1.bias, blk0.norm1.eps, x, h, n_clips * n_frames * self.num_tokens,
                         self.num_tokens, self.embed_dim, ops.DTYPE_CODE[self.compute_dtype])

    def trunk(self, x2d: torch.Tensor, n_clips: int, head_norm_eps: Optional[float] = None) -> torch.Tensor:
        """STE_forward + TTE_foward + ST_foward (mix_ste.py:128-173) on one micro-batch.

        x2d: fp32 [n_clips, L, J, in_chans] (contiguous).  Returns the fp32 [n_clips*L*tokens, C] output of the last temporal
        block BEFORE ``Temporal_norm`` (the head kernels apply it, fused with their own LayerNorm).

        ``head_norm_eps`` (fused C = 512 trunk only): the last block's epilogue also applies ``Temporal_norm`` and the affine-free
        LayerNorm(eps) that every head of the model shares, and the 16-bit normalised activations [tokens, C] are returned instead
        (the A operand of the folded head projection, mp_heads_fwd16); the fp32 stream is not written out."""
        if self.training and any(isinstance(b.drop_path, DropPath) and b.drop_path.drop_prob > 0 for b in self.STEblocks):
            raise NotImplementedError("the fused inference trunk has no stochastic depth: call model.eval(), or run the forward "
                                      "with gradients enabled (the training trunk applies DropPath)")
        n_frames = self.num_frame
        n_tok, c, heads = self.num_tokens, self.embed_dim, self.num_heads
        n_tokens = n_clips * n_frames * n_tok
        ws = self._workspace(n_tokens, x2d.device)
        x, h = ws["x"][:n_tokens], ws["h"][:n_tokens]
        hidden_dim = self.STEblocks[0].mlp.fc1.out_features
        flat = ws["wide"].view(-1)
        qkv = flat[:n_tokens * 3 * c].view(n_tokens, 3 * c)
        hid = flat[:n_tokens * hidden_dim].view(n_tokens, hidden_dim)   # aliases qkv: never live at the same time
        w = self._shadow_weights()
        dt = ops.DTYPE_CODE[self.compute_dtype]
        self._embed(x2d, n_clips, n_frames, x, h)
        depth = self.block_depth
        blocks = []
        for i in range(depth):
            blocks.append((self.STEblocks[i], 4 * i, L.MP_ATTN_SPATIAL, self.Spatial_norm))
            blocks.append((self.TTEblocks[i], 4 * (depth + i), L.MP_ATTN_TEMPORAL, self.Temporal_norm))
        fused = c == 512 and self.fuse_layernorm     # residual GEMM + LayerNorms in one kernel (CTA pairs, N = 512)
        for bi, (blk, wi, mode, post) in enumerate(blocks):
            last = bi + 1 == len(blocks)
            nxt = None if last else blocks[bi + 1][0]
            pos = self.Temporal_pos_embed if bi == 0 else None   # TTE_foward adds it once, after the first STE block
            ops.linear(h, w[wi + 0], blk.attn.qkv.bias, qkv, L.MP_EPI_BIAS)
            ops.attention(qkv, h, n_clips, n_frames, n_tok, c, heads, mode)
            if fused:
                ops.linear_ln(h, w[wi + 1], blk.attn.proj.bias, x, x, h, ln=(blk.norm2.weight, blk.norm2.bias), ln_eps=blk.norm2.eps)
            else:
                ops.linear(h, w[wi + 1], blk.attn.proj.bias, x, L.MP_EPI_RESIDUAL, resid=x)
                ops.layernorm(x, None, h, ln=(blk.norm2.weight, blk.norm2.bias), ln_eps=blk.norm2.eps, dtype=dt)
            if fused and self.fuse_mlp and hidden_dim == 1024:
                # h is read (a tile's 64 rows are resident in shared memory before anything of the tile is written) and rewritten in place
                if last and head_norm_eps is not None:
                    one, zero = self._unit_affine(x.device)
                    ops.mlp_ln(h, w[wi + 2], blk.mlp.fc1.bias, w[wi + 3], blk.mlp.fc2.bias, x, None, h, post=(post.weight, post.bias),
                               post_eps=post.eps, ln=(one, zero), ln_eps=head_norm_eps)
                    return h
                if last:
                    ops.mlp_ln(h, w[wi + 2], blk.mlp.fc1.bias, w[wi + 3], blk.mlp.fc2.bias, x, x, None)
                else:
                    ops.mlp_ln(h, w[wi + 2], blk.mlp.fc1.bias, w[wi + 3], blk.mlp.fc2.bias, x, x, h, post=(post.weight, post.bias),
                               post_eps=post.eps, pos=pos, pos_div=n_tok, pos_mod=n_frames, ln=(nxt.norm1.weight, nxt.norm1.bias),
                               ln_eps=nxt.norm1.eps)
                continue
            ops.linear(h, w[wi + 2], blk.mlp.fc1.bias, hid, L.MP_EPI_GELU)
            if fused:
                if last and head_norm_eps is not None:
                    one, zero = self._unit_affine(x.device)
                    ops.linear_ln(hid, w[wi + 3], blk.mlp.fc2.bias, x, None, h, post=(post.weight, post.bias), post_eps=post.eps,
                                  ln=(one, zero), ln_eps=head_norm_eps)
                    return h
                if last:
                    ops.linear_ln(hid, w[wi + 3], blk.mlp.fc2.bias, x, x, None)
                else:
                    ops.linear_ln(hid, w[wi + 3], blk.mlp.fc2.bias, x, x, h, post=(post.weight, post.bias), post_eps=post.eps, pos=pos,
                                  pos_div=n_tok, pos_mod=n_frames, ln=(nxt.norm1.weight, nxt.norm1.bias), ln_eps=nxt.norm1.eps)
            else:
                ops.linear(hid, w[wi + 3], blk.mlp.fc2.bias, x, L.MP_EPI_RESIDUAL, resid=x)
                if not last:
                    ops.layernorm(x, x, h, post=(post.weight, post.bias), post_eps=post.eps, pos=pos, pos_div=n_tok, pos_mod=n_frames,
                                  ln=(nxt.norm1.weight, nxt.norm1.bias), ln_eps=nxt.norm1.eps, dtype=dt)
        return x

    def _unit_affine(self, device):
        ua = getattr(self, "_unit_affine_cache", None)
        if ua is None or ua[0].device != device:
            ua = (torch.ones(self.embed_dim, dtype=torch.float32, device=device), torch.zeros(self.embed_dim, dtype=torch.float32, device=device))
            self._unit_affine_cache = ua
        return ua

    # ------------------------------------------------------------------------------------------ training path
    def _grad_mode(self) -> bool:
        """True when a forward must record what the backward sweep needs (any parameter wants a gradient)."""
        return torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())

    def _shadow_weights_t(self) -> List[torch.Tensor]:
        """Transposed 16-bit shadows [K, N] of the GEMM weights (operand of dgrad = dY W), written by the same refresh launch."""
        if not getattr(self, "_shadow_want_t", False):
            self._shadow_want_t = True
            self._shadow_key = None          # re-run the refresh with the transposed outputs in the table
        self._shadow_weights()
        return self._shadow_t_list

    def _block_list(self):
        depth = self.block_depth
        blocks = []
        for i in range(depth):
            blocks.append((self.STEblocks[i], 4 * i, L.MP_ATTN_SPATIAL, self.Spatial_norm))
            blocks.append((self.TTEblocks[i], 4 * (depth + i), L.MP_ATTN_TEMPORAL, self.Temporal_norm))
        return blocks

    def _droppath_scales(self, blocks, n_clips: int, device):
        """Per-token branch scales of timm's DropPath (mix_ste.py:334-336) for BOTH residual branches of every block of one forward:
        Bernoulli(keep) / keep per sample of the block's batch — a (clip, frame) in spatial blocks, a (clip, token) track in temporal
        blocks.  Returns [(s1, s2)] per block, entries None where the block keeps everything.  All masks of a forward come out of two
        ``torch.rand`` calls (one per block kind): drawing them block by block cost ~110 tiny launches per training step."""
        n_frames, n_tok = self.num_frame, self.num_tokens
        out = [[None, None] for _ in blocks]
        want = {L.MP_ATTN_SPATIAL: [], L.MP_ATTN_TEMPORAL: []}
        for bi, (blk, _, mode, _) in enumerate(blocks):
            dp = blk.drop_path
            if self.training and isinstance(dp, DropPath) and dp.drop_prob > 0.0:
                keep = 1.0 - dp.drop_prob
                inv = 1.0 / keep if (keep > 0.0 and dp.scale_by_keep) else 1.0
                want[mode] += [(bi, 0, keep, inv), (bi, 1, keep, inv)]
        cache = self.__dict__.setdefault("_dp_consts", {})
        for mode, items in want.items():
            if not items:
                continue
            key = (mode, str(device), tuple((k, i) for _, _, k, i in items))
            if key not in cache:                 # built eagerly during warm-up, so a CUDA-graph capture of the step only sees device tensors
                if len(cache) > 8:
                    cache.clear()
                cache[key] = (torch.tensor([k for _, _, k, _ in items], dtype=torch.float32, device=device).view(-1, 1, 1, 1),
                              torch.tensor([i for _, _, _, i in items], dtype=torch.float32, device=device).view(-1, 1, 1, 1))
            keep_t, inv_t = cache[key]
            shape = (len(items), n_clips, n_frames, 1) if mode == L.MP_ATTN_SPATIAL else (len(items), n_clips, 1, n_tok)
            mask = (torch.rand(shape, dtype=torch.float32, device=device) < keep_t) * inv_t
            full = mask.expand(len(items), n_clips, n_frames, n_tok).reshape(len(items), -1).contiguous()
            for r, (bi, slot, _, _) in enumerate(items):
                out[bi][slot] = full[r]
        return out

    def _train_forward(self, x2d: torch.Tensor, n_clips: int):
        """The trunk with every tensor the backward sweep needs kept on a tape (no aliasing, no fused residual+LayerNorm):
        per block x0 (input), h1 = norm1(x0), qkv, o (attention out), x1, h2 = norm2(x1), u (fc1 pre-activation), a = gelu(u), x2."""
        n_frames, n_tok, c, heads = self.num_frame, self.num_tokens, self.embed_dim, self.num_heads
        n_tokens = n_clips * n_frames * n_tok
        dev = x2d.device
        dt = ops.DTYPE_CODE[self.compute_dtype]
        td = ops.TORCH_DTYPE[dt]
        hidden = self.STEblocks[0].mlp.fc1.out_features
        w = self._shadow_weights()
        f32 = lambda: torch.empty((n_tokens, c), dtype=torch.float32, device=dev)
        b16 = lambda cols: torch.empty((n_tokens, cols), dtype=td, device=dev)
        x, h = f32(), b16(c)
        self._embed(x2d, n_clips, n_frames, x, h)
        blocks = self._block_list()
        tape = {"x2d": x2d, "n_clips": n_clips, "blocks": []}
        scales = self._droppath_scales(blocks, n_clips, dev)
        for bi, (blk, wi, mode, post) in enumerate(blocks):
            last = bi + 1 == len(blocks)
            s1, s2 = scales[bi]
            qkv, o = b16(3 * c), b16(c)
            ops.linear(h, w[wi + 0], blk.attn.qkv.bias, qkv, L.MP_EPI_BIAS)
            ops.attention(qkv, o, n_clips, n_frames, n_tok, c, heads, mode)
            x1, h2 = f32(), b16(c)
            if c == 512:
                # proj + DropPath-scaled residual add + norm2 in one kernel (x0 stays intact for the tape: resid != x_out)
                ops.linear_ln(o, w[wi + 1], blk.attn.proj.bias, x, x1, h2, ln=(blk.norm2.weight, blk.norm2.bias), ln_eps=blk.norm2.eps,
                              row_scale=s1)
            else:
                if s1 is None:
                    ops.linear(o, w[wi + 1], blk.attn.proj.bias, x1, L.MP_EPI_RESIDUAL, resid=x)
                else:
                    T.residual_rowscale(x, ops.linear(o, w[wi + 1], blk.attn.proj.bias, b16(c), L.MP_EPI_BIAS), s1, x1)
                ops.layernorm(x1, None, h2, ln=(blk.norm2.weight, blk.norm2.bias), ln_eps=blk.norm2.eps, dtype=dt)
            u, a = b16(hidden), b16(hidden)
            if hidden % 256 == 0:
                ops.linear_gelu2(h2, w[wi + 2], blk.mlp.fc1.bias, u, a)      # pre-activation (for the backward) and GELU in one launch
            else:
                ops.linear(h2, w[wi + 2], blk.mlp.fc1.bias, u, L.MP_EPI_BIAS)
                T.gelu_fwd(u, a)
            x2 = f32()
            fused_tail = c == 512 and not last      # fc2 + DropPath-scaled residual + post-norm (+ pos-embed) + next norm1 in one launch
            if fused_tail:
                nxt = blocks[bi + 1][0]
                x3, h_next = f32(), b16(c)
                pos = self.Temporal_pos_embed if bi == 0 else None   # TTE_foward adds it once, after the first STE block
                ops.linear_ln(a, w[wi + 3], blk.mlp.fc2.bias, x1, x3, h_next, post=(post.weight, post.bias), post_eps=post.eps, pos=pos,
                              pos_div=n_tok, pos_mod=n_frames, ln=(nxt.norm1.weight, nxt.norm1.bias), ln_eps=nxt.norm1.eps, row_scale=s2,
                              x_pre=x2)                              # x2 (before the post-norm) stays on the tape for its backward
            elif s2 is None:
                ops.linear(a, w[wi + 3], blk.mlp.fc2.bias, x2, L.MP_EPI_RESIDUAL, resid=x1)
            elif c == 512:
                ops.linear_ln(a, w[wi + 3], blk.mlp.fc2.bias, x1, x2, None, row_scale=s2)   # fc2 + DropPath-scaled residual add
            else:
                T.residual_rowscale(x1, ops.linear(a, w[wi + 3], blk.mlp.fc2.bias, b16(c), L.MP_EPI_BIAS), s2, x2)
            rec = {"x0": x, "h1": h, "qkv": qkv, "o": o, "x1": x1, "h2": h2, "u": u, "a": a, "x2": x2, "s1": s1, "s2": s2,
                   "pos": bi == 0}
            tape["blocks"].append(rec)
            if last:
                x, h = x2, None
            elif fused_tail:
                x, h = x3, h_next
            else:
                nxt = blocks[bi + 1][0]
                x3, h = f32(), b16(c)
                pos = self.Temporal_pos_embed if bi == 0 else None   # TTE_foward adds it once, after the first STE block
                ops.layernorm(x2, x3, h, post=(post.weight, post.bias), post_eps=post.eps, pos=pos, pos_div=n_tok, pos_mod=n_frames,
                              ln=(nxt.norm1.weight, nxt.norm1.bias), ln_eps=nxt.norm1.eps, dtype=dt)
                x = x3
        return x, tape

    def _embed_backward(self, x2d: torch.Tensor, dx0: torch.Tensor, n_clips: int) -> None:
        """Gradients of Spatial_patch_to_embedding and Spatial_pos_embed (mix_ste.py:128-138) from dx0 [tokens, C]."""
        lin = self.Spatial_patch_to_embedding
        T.small_wgrad(dx0, x2d.reshape(-1, self.in_chans), T.grad_of(lin.weight), T.grad_of(lin.bias))
        T.group_rowsum(dx0, T.grad_of(self.Spatial_pos_embed), 1, self.num_tokens)

    def _train_backward(self, tape, dfeat: torch.Tensor) -> None:
        """Reverse sweep over the tape: accumulates every parameter gradient of the trunk into ``p.grad`` (fp32).  dfeat is the
        gradient w.r.t. the trunk output (the last block's x2, before Temporal_norm).  ``_on_block_grads(block)`` is called when
        a block's parameter gradients are complete (the data-parallel reducer hooks in there)."""
        n_clips = tape["n_clips"]
        n_frames, n_tok, c, heads = self.num_frame, self.num_tokens, self.embed_dim, self.num_heads
        hidden = self.STEblocks[0].mlp.fc1.out_features
        dt = ops.DTYPE_CODE[self.compute_dtype]
        td = ops.TORCH_DTYPE[dt]
        w_t = self._shadow_weights_t()
        blocks = self._block_list()
        n_tokens = dfeat.shape[0]
        dev = dfeat.device
        g = T.grad_of
        b16 = lambda cols: torch.empty((n_tokens, cols), dtype=td, device=dev)
        dx = dfeat.contiguous().clone()            # fp32 gradient of the residual stream, updated in place below
        dy16, wide16 = b16(c), b16(max(3 * c, hidden))
        flat = wide16.view(-1)
        # Weight gradients on a second stream.  dgrad and wgrad of a layer read the same output gradient and are independent, and at a few
        # thousand rows a GEMM leaves SMs idle (13,770 rows = 108 tiles on 74 CTA pairs: two rounds where 1.46 would do): the weight-
        # gradient kernel of a layer runs under the data-gradient kernels that follow it and fills those SMs.  `wg(...)` returns the event
        # the main stream waits for before it overwrites a buffer that weight gradient still reads; a block's last one is joined before
        # its tape entry is released and the reducer hook runs.  Captured in the step's CUDA graph like any fork / join.
        main = torch.cuda.current_stream(dev)
        side = self._side_stream(dev) if self.overlap_wgrad else None

        def wg(dy, act, gw, gb):
            if side is None:
                T.wgrad(dy, act, gw, gb)
                return None
            side.wait_stream(main)
            with torch.cuda.stream(side):
                T.wgrad(dy, act, gw, gb)
                ev = torch.cuda.Event()
                ev.record(side)
            return ev

        def after(ev):
            if ev is not None:
                main.wait_event(ev)

        for bi in range(len(blocks) - 1, -1, -1):
            blk, wi, mode, post = blocks[bi]
            rec = tape["blocks"][bi]
            last = bi + 1 == len(blocks)
            if not last:
                # x3 = post-norm(x2) (+ Temporal_pos_embed after the first block); dx currently is d/dx3
                if rec["pos"]:
                    T.group_rowsum(dx, g(self.Temporal_pos_embed), n_tok, n_frames)
                # (the bias gradients are column sums of the 16-bit output gradients: each is taken by the kernel that writes that matrix)
                T.layernorm_bwd(rec["x2"], post.weight, post.eps, dx, None, dx, g(post.weight), g(post.bias), dt, dx16=dy16,
                                rowscale=rec["s2"], dx16_colsum=g(blk.mlp.fc2.bias))
                fc2_db = None
            else:
                T.cast_rowscale(dx, rec["s2"], dy16)
                fc2_db = g(blk.mlp.fc2.bias)
            # ---- MLP branch: x2 = x1 + s2 * (fc2(gelu(fc1(norm2(x1))))); dy16 = 16-bit(s2 * dx)
            e_fc2 = wg(dy16, rec["a"], g(blk.mlp.fc2.weight), fc2_db)
            da = flat[:n_tokens * hidden].view(n_tokens, hidden)
            T.dgrad(dy16, w_t[wi + 3], da)
            T.gelu_bwd(rec["u"], da, da, colsum=g(blk.mlp.fc1.bias))
            e_fc1 = wg(da, rec["h2"], g(blk.mlp.fc1.weight), None)
            after(e_fc2)                               # dy16 is rewritten next
            T.dgrad(da, w_t[wi + 2], dy16)
            # norm2 backward reads dy16 (= d h2) and rewrites it with 16-bit(s1 * dx1), the operand of the attention branch
            T.layernorm_bwd(rec["x1"], blk.norm2.weight, blk.norm2.eps, dy16, dx, dx, g(blk.norm2.weight), g(blk.norm2.bias), dt, dx16=dy16,
                            rowscale=rec["s1"], dx16_colsum=g(blk.attn.proj.bias))
            # ---- attention branch: x1 = x0 + s1 * proj(attention(qkv(norm1(x0))))
            e_proj = wg(dy16, rec["o"], g(blk.attn.proj.weight), None)
            do = b16(c)
            T.dgrad(dy16, w_t[wi + 1], do)
            dqkv = flat[:n_tokens * 3 * c].view(n_tokens, 3 * c)
            qkv_db = g(blk.attn.qkv.bias) if blk.attn.qkv.bias is not None else None
            after(e_fc1)                               # dqkv shares its buffer with da
            T.attention_bwd(rec["qkv"], rec["o"], do, dqkv, n_clips, n_frames, n_tok, c, heads, mode, colsum=qkv_db)
            e_qkv = wg(dqkv, rec["h1"], g(blk.attn.qkv.weight), None)
            after(e_proj)                              # dy16 is rewritten next
            T.dgrad(dqkv, w_t[wi + 0], dy16)
            T.layernorm_bwd(rec["x0"], blk.norm1.weight, blk.norm1.eps, dy16, dx, dx, g(blk.norm1.weight), g(blk.norm1.bias), dt)
            after(e_qkv)                               # the block's gradients are complete; its activations may go
            rec.clear()
            hook = getattr(self, "_on_block_grads", None)
            if hook is not None:
                hook(blk)
        self._embed_backward(tape["x2d"], dx, n_clips)

    def trunk_autograd(self, x2d: torch.Tensor, n_clips: int) -> torch.Tensor:
        """Differentiable trunk: same result as ``trunk`` (separate kernels instead of the fused residual+LayerNorm ones); its
        backward runs ``_train_backward`` and writes parameter gradients straight into ``.grad``."""
        anchor = next(p for p in self.parameters() if p.requires_grad)
        return _TrunkFn.apply(self, x2d, n_clips, anchor)

    def _check_input(self, x: torch.Tensor):
        if x.dim() != 4:
            raise ValueError(f"expected x of shape [B, L, J, C], got {tuple(x.shape)}")
        b, l, j, cin = x.shape
        if l != self.num_frame:
            # the reference fails here too: Temporal_pos_embed is [1, num_frame, C] (mix_ste.py:63-65,149)
            raise RuntimeError(f"The size of tensor a ({l}) must match the size of tensor b ({self.num_frame}) at non-singleton dimension 1")
        return b, l, j, cin

    def clips_per_micro_batch(self) -> int:
        """Clips per micro-batch; a multiple of 4 keeps every per-clip output slice 16-byte aligned (bulk-copy paths)."""
        n = max(1, self.micro_batch_tokens // (self.num_frame * self.num_tokens))
        return n - n % 4 if n >= 4 else n

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """mix_ste.py:175-191 with the plain head (LayerNorm eps 1e-5 + Linear): [B,L,J,in] -> [B,L,J,out_dim]."""
        ops._need_cuda(x)
        b, l, j, _ = self._check_input(x)
        x = ops._f32(x)
        if self._grad_mode():
            return self._forward_with_grad(x)
        out = torch.empty((b, 1, l, j, self.out_dim), dtype=torch.float32, device=x.device)
        mb = self.clips_per_micro_batch()
        norm, lin = self.head[0], self.head[1]
        for s in range(0, b, mb):
            n = min(mb, b - s)
            feat = self.trunk(x[s:s + n], n)
            ops.heads_fwd(feat, self.Temporal_norm.weight, self.Temporal_norm.bias, self.Temporal_norm.eps, norm.weight, norm.bias,
                          lin.weight, lin.bias, None, None, out[s:s + n], None, n, l, 1, self.out_dim, False)
        return out[:, 0]

    def _forward_with_grad(self, x: torch.Tensor) -> torch.Tensor:
        """Differentiable forward of the plain head (Temporal_norm -> LayerNorm(1e-5) -> Linear), whole batch at once."""
        b, l, j, _ = x.shape
        feat = self.trunk_autograd(x, b)
        norm, lin = self.head[0], self.head[1]
        code = ops.DTYPE_CODE[self.compute_dtype]
        y = T.layer_norm(feat, self.Temporal_norm.weight, self.Temporal_norm.bias, self.Temporal_norm.eps)
        z = T.layer_norm(y, norm.weight, norm.bias, norm.eps, out16=code)
        n_pad = (self.out_dim + 127) // 128 * 128
        w = torch.cat([lin.weight, lin.weight.new_zeros(n_pad - self.out_dim, self.embed_dim)])
        bias = torch.cat([lin.bias, lin.bias.new_zeros(n_pad - self.out_dim)])
        out = T.linear_f32(z, w, bias)[:, :self.out_dim]
        return out.reshape(b, l, j, self.out_dim)

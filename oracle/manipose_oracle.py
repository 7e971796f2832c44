"""CPU oracle for the ManiPose lifting hot path (TEST INFRASTRUCTURE — NOT PRODUCT CODE).

A plain-PyTorch fp32 *functional* restatement of the reference's algorithm for the hot path
SURVEY.md §8(a) lists (MixSTE backbone -> K hypothesis heads -> manifold decoder -> WTA loss /
hypothesis metrics).  Every function cites the reference ``file:line`` it follows (paths are
relative to ``/root/reference/``).  It works from a ``state_dict`` with the reference's
parameter names, so it can check the product on identical weights.

Who may import this: ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs — only as the checker / the timed CPU baseline,
never as a product code path.  ``manipose_b200`` must never import it.

Parity pin: the reference ships NO tests, golden vectors or known-answer fixtures for this
path, so this oracle is pinned against *outputs of the reference itself run in
the build container* (``oracle/ref_loader.py`` imports the unmodified reference with the three
shims of SURVEY.md §8c): ``tests/test_oracle_vs_reference.py`` compares every function here with
the reference on seeded inputs whenever ``/root/reference`` exists, and
``scripts/make_goldens.py`` froze reference outputs into ``tests/golden/*.pt`` which
``tests/test_oracle_golden.py`` checks everywhere (including the GPU box).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# Skeleton constants (hpe/mh_so3_hpe/data/h36m_lifting.py:40-57,649-660 after the 17-joint
# reduction == hpe/mh_so3_hpe/data/dataset_3dhp.py:132-138); SURVEY.md §A.1.
# --------------------------------------------------------------------------------------
H36M17_PARENTS = [-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 9, 8, 11, 12, 8, 14, 15]
H36M17_T_POSE_OPERATORS = [
    (0, 0, 0),
    (1, 0, 0), (0, -1, 0), (0, -1, 0),
    (-1, 0, 0), (0, -1, 0), (0, -1, 0),
    (0, 1, 0), (0, 1, 0), (0, 1, 0), (0, 1, 0),
    (-1, 0, 0), (-1, 0, 0), (-1, 0, 0),
    (1, 0, 0), (1, 0, 0), (1, 0, 0),
]
H36M17_JOINTS_LEFT = [4, 5, 6, 11, 12, 13]
H36M17_JOINTS_RIGHT = [1, 2, 3, 14, 15, 16]
# hpe/mh_so3_hpe/metrics/losses.py:6-8
STANDARD_H36M_WEIGHTS = torch.tensor(
    [1, 1, 2.5, 2.5, 1, 2.5, 2.5, 1, 1, 1, 1.5, 1.5, 4, 4, 1.5, 4, 4], dtype=torch.float32
)


def has_children(parents: List[int]) -> List[bool]:
    """hpe/mh_so3_hpe/data/skeleton.py:85-89."""
    out = [False] * len(parents)
    for p in parents:
        if p != -1:
            out[p] = True
    return out


# --------------------------------------------------------------------------------------
# D1-D3: 6-D -> SO(3)   (hpe/mh_so3_hpe/architectures/utils/rotation_tools.py)
# --------------------------------------------------------------------------------------
def normalize_vector(v: torch.Tensor) -> torch.Tensor:
    """rotation_tools.py:6-17 — v / max(sqrt(sum v^2), 1e-8) (epsilon on v.device)."""
    mag = torch.sqrt(v.pow(2).sum(1))
    mag = torch.max(mag, torch.tensor([1e-8], dtype=v.dtype, device=v.device))
    return v / mag[:, None]


def cross_product(u: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """rotation_tools.py:21-32."""
    i = u[:, 1] * v[:, 2] - u[:, 2] * v[:, 1]
    j = u[:, 2] * v[:, 0] - u[:, 0] * v[:, 2]
    k = u[:, 0] * v[:, 1] - u[:, 1] * v[:, 0]
    return torch.stack((i, j, k), dim=1)


def rotation_matrix_from_ortho6d(r6: torch.Tensor) -> torch.Tensor:
    """rotation_tools.py:35-57 — x = n(a), z = n(x × b), y = z × x; COLUMNS are [x y z]."""
    x = normalize_vector(r6[:, 0:3])
    z = normalize_vector(cross_product(x, r6[:, 3:6]))
    y = cross_product(z, x)
    return torch.stack((x, y, z), dim=2)


def rotation_matrix_from_ortho4d(r4: torch.Tensor) -> torch.Tensor:
    """rotation_tools.py:60-116 — R_theta · R_phi from two normalised 2-vectors."""
    n = r4.shape[0]
    cs_t = normalize_vector(r4[:, 0:2])
    cs_p = normalize_vector(r4[:, 2:4])
    zeros = torch.zeros((n, 1), dtype=r4.dtype)
    theta_y = torch.cat([cs_t, zeros], dim=1)
    theta_z = torch.tensor([0.0, 0.0, 1.0], dtype=r4.dtype).expand(n, -1)
    theta_x = cross_product(theta_y, theta_z)
    phi_y = torch.cat([zeros, cs_p], dim=1)
    phi_x = torch.tensor([1.0, 0.0, 0.0], dtype=r4.dtype).expand(n, -1)
    phi_z = cross_product(phi_x, phi_y)
    r_theta = torch.stack((theta_x, theta_y, theta_z), dim=2)
    r_phi = torch.stack((phi_x, phi_y, phi_z), dim=2)
    return r_theta.bmm(r_phi)


# --------------------------------------------------------------------------------------
# D4-D7: decoder   (hpe/mh_so3_hpe/architectures/pose_decoder.py, utils/forward_kinematics.py)
# --------------------------------------------------------------------------------------
def build_t_pose(bones_length: torch.Tensor, parents=H36M17_PARENTS,
                 operators=H36M17_T_POSE_OPERATORS) -> torch.Tensor:
    """pose_decoder.py:98-120 — t[b+1] = t[parent[b+1]] + op[b+1] * len[b]; [N,16,1] -> [N,17,3]."""
    n, n_parts, _ = bones_length.shape
    assert n_parts == len(parents) - 1
    t_pose = torch.zeros((n, len(parents), 3), dtype=torch.float32)
    for b in range(n_parts):
        op = torch.tensor(operators[b + 1], dtype=torch.float32)
        t_pose[:, b + 1, :] = t_pose[:, parents[b + 1], :] + op * bones_length[:, b]
    return t_pose


def forward_kinematics(t_pose: torch.Tensor, rotations: torch.Tensor, root_positions: torch.Tensor,
                       parents=H36M17_PARENTS) -> torch.Tensor:
    """forward_kinematics.py:6-48 — Rw[j] = Rw[par] R[j]; p[j] = Rw[j](t[j]-t[par]) + p[par]."""
    kids = has_children(parents)
    pos: List[torch.Tensor] = []
    rot: List[Optional[torch.Tensor]] = []
    for j in range(rotations.shape[1]):
        if parents[j] == -1:
            pos.append(root_positions)
            rot.append(rotations[:, 0])
        else:
            p = parents[j]
            offset = (t_pose[:, j, :] - t_pose[:, p, :]).view(-1, 3, 1)
            rw = rot[p].matmul(rotations[:, j])
            pos.append(rw.matmul(offset).view(-1, 3) + pos[p])
            rot.append(rw if kids[j] else None)
    return torch.stack(pos, dim=2).permute(0, 2, 1)


def pose_decoder(rotations_repr: torch.Tensor, bones_lengths_repr: torch.Tensor,
                 root_positions: torch.Tensor, rot_rep_dim: int = 6,
                 parents=H36M17_PARENTS, operators=H36M17_T_POSE_OPERATORS) -> torch.Tensor:
    """pose_decoder.py:32-96 — rotations_repr [N,J,D], bones [B,16,1] (row n uses clip n // (N/B))."""
    assert rot_rep_dim in (4, 6), f"Unsupported rotations representation dimension: {rot_rep_dim}"
    assert rotations_repr.shape[-1] == rot_rep_dim
    n, j, _ = rotations_repr.shape
    b = bones_lengths_repr.shape[0]
    assert n % b == 0
    reps = n // b
    bones = torch.stack([bones_lengths_repr] * reps, dim=1).reshape(n, -1, 1)  # :85-96
    flat = rotations_repr.reshape(-1, rot_rep_dim)
    mats = rotation_matrix_from_ortho6d(flat) if rot_rep_dim == 6 else rotation_matrix_from_ortho4d(flat)
    mats = mats.reshape(n, j, 3, 3)
    t_pose = build_t_pose(bones, parents, operators)
    return forward_kinematics(t_pose, mats, root_positions, parents)


def pose_decoder_ieee(rotations_repr: torch.Tensor, bones_lengths_repr: torch.Tensor, root_positions: torch.Tensor,
                      parents=H36M17_PARENTS, operators=H36M17_T_POSE_OPERATORS) -> torch.Tensor:
    """The same algorithm as ``pose_decoder`` (6-D, or 4-D when the last dim is 4: rotation_tools.py:60-116, whose products with
    the exact 0 / 1 entries of R_theta and R_phi are exact) restated in numpy float32, where every +, -, *, / and sqrt is
    one correctly-rounded IEEE-754 operation and nothing is fused or reassociated: sum of squares as (x^2 + y^2) + z^2
    (rotation_tools.py:6-17), cross products as two products and a subtraction (:21-32), 3x3 products as
    ((a0 b0 + a1 b1) + a2 b2) (forward_kinematics.py:31-40), T-pose by cumulative adds with the offset recovered by
    subtraction (pose_decoder.py:115-119, forward_kinematics.py:31-33).

    Why it exists: torch's CPU ``sqrt`` on this build (MKL/AVX-512 path) is NOT correctly rounded (0.7 % of inputs are 1 ulp
    off, measured in this container), so ``pose_decoder`` itself is only reproducible to ~1e-7 across platforms.  This
    restatement is the bit-exact target for the CUDA decoder's EXACT mode; tests pin it to ``pose_decoder`` / the reference
    fixtures within 1e-6 relative."""
    import numpy as np
    r = rotations_repr.detach().cpu().numpy().astype(np.float32)
    n, nj, _ = r.shape
    b = bones_lengths_repr.shape[0]
    assert n % b == 0
    lens = np.repeat(bones_lengths_repr.detach().cpu().numpy().astype(np.float32).reshape(b, -1), n // b, axis=0)   # [N,16]
    root = root_positions.detach().cpu().numpy().astype(np.float32)
    eps = np.float32(1e-8)

    def normalize(x, y, z):
        m = np.maximum(np.sqrt((x * x + y * y) + z * z), eps)
        return x / m, y / m, z / m

    def cross(u, v):
        return (u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0])

    kids = has_children(parents)
    rw = [None] * nj     # world rotations as [row][col] lists of [N] arrays
    pos = [None] * nj
    tp = [None] * nj     # T-pose x / y
    rd = r.shape[2]
    assert rd in (4, 6), f"Unsupported rotations representation dimension: {rd}"
    zero = np.zeros(n, np.float32)

    def normalize2(x, y):
        m = np.maximum(np.sqrt(x * x + y * y), eps)
        return x / m, y / m

    for j in range(nj):
        a = [r[:, j, i] for i in range(rd)]
        if rd == 6:
            x = normalize(a[0], a[1], a[2])
            z = normalize(*cross(x, (a[3], a[4], a[5])))
            y = cross(z, x)
            rl = [[x[i], y[i], z[i]] for i in range(3)]      # rows of the local rotation, columns [x y z]
        else:
            ct, st = normalize2(a[0], a[1])
            cp, sp = normalize2(a[2], a[3])
            rl = [[st, ct * cp, -(ct * sp)], [-ct, st * cp, -(st * sp)], [zero, sp, cp]]
        if parents[j] == -1:
            rw[j] = rl
            pos[j] = [root[:, 0], root[:, 1], root[:, 2]]
            tp[j] = [np.zeros(n, np.float32), np.zeros(n, np.float32)]
            continue
        p = parents[j]
        op = operators[j]
        ax = 0 if op[0] != 0 else 1
        step = lens[:, j - 1] if op[ax] > 0 else -lens[:, j - 1]
        tp[j] = list(tp[p])
        tp[j][ax] = tp[p][ax] + step
        off = tp[j][ax] - tp[p][ax]
        rp = rw[p]
        w = [[(rp[row][0] * rl[0][col] + rp[row][1] * rl[1][col]) + rp[row][2] * rl[2][col] for col in range(3)] for row in range(3)]
        pos[j] = [w[row][ax] * off + pos[p][row] for row in range(3)]
        rw[j] = w if kids[j] else None
    out = np.stack([np.stack(pj, axis=-1) for pj in pos], axis=1)
    return torch.from_numpy(out)


# --------------------------------------------------------------------------------------
# B1-B5: MixSTE backbone   (hpe/mh_so3_hpe/architectures/mix_ste.py)
# --------------------------------------------------------------------------------------
def _ln(x, sd, prefix, eps):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], eps)


def attention(x: torch.Tensor, sd: Dict[str, torch.Tensor], prefix: str, num_heads: int) -> torch.Tensor:
    """mix_ste.py:255-282 (comb=False) — qkv rows ordered [q|k|v] x heads; scale = head_dim**-0.5."""
    bsz, n, c = x.shape
    hd = c // num_heads
    qkv = F.linear(x, sd[prefix + ".qkv.weight"], sd.get(prefix + ".qkv.bias"))
    qkv = qkv.reshape(bsz, n, 3, num_heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    attn = (q @ k.transpose(-2, -1)) * (hd ** -0.5)
    attn = attn.softmax(dim=-1)
    out = (attn @ v).transpose(1, 2).reshape(bsz, n, c)
    return F.linear(out, sd[prefix + ".proj.weight"], sd[prefix + ".proj.bias"])


def block(x: torch.Tensor, sd: Dict[str, torch.Tensor], prefix: str, num_heads: int) -> torch.Tensor:
    """mix_ste.py:352-358 in eval mode (DropPath = identity, residual_scale = 1), LN eps 1e-6 (:49)."""
    x = x + attention(_ln(x, sd, prefix + ".norm1", 1e-6), sd, prefix + ".attn", num_heads)
    h = _ln(x, sd, prefix + ".norm2", 1e-6)
    h = F.linear(h, sd[prefix + ".mlp.fc1.weight"], sd[prefix + ".mlp.fc1.bias"])
    h = F.gelu(h)  # exact erf GELU (nn.GELU default, mix_ste.py:200)
    h = F.linear(h, sd[prefix + ".mlp.fc2.weight"], sd[prefix + ".mlp.fc2.bias"])
    return x + h


def mixste_trunk(x_emb: torch.Tensor, sd: Dict[str, torch.Tensor], prefix: str, depth: int,
                 num_heads: int) -> torch.Tensor:
    """mix_ste.py:128-173 after the patch embedding: x_emb is [B, L, J, C] (embedding done)."""
    b, l, j, c = x_emb.shape
    x = x_emb.reshape(b * l, j, c) + sd[prefix + "Spatial_pos_embed"]                # :137
    x = block(x, sd, prefix + "STEblocks.0", num_heads)                               # :140-141
    x = _ln(x, sd, prefix + "Spatial_norm", 1e-6)                                     # :143
    x = x.reshape(b, l, j, c).permute(0, 2, 1, 3).reshape(b * j, l, c)                # :144
    x = x + sd[prefix + "Temporal_pos_embed"]                                         # :149
    x = block(x, sd, prefix + "TTEblocks.0", num_heads)                               # :151-152
    x = _ln(x, sd, prefix + "Temporal_norm", 1e-6)                                    # :154
    x = x.reshape(b, j, l, c).permute(0, 2, 1, 3)                                     # rmcl:249 / :181
    for i in range(1, depth):                                                         # :160-171
        x = x.reshape(b * l, j, c)
        x = block(x, sd, prefix + f"STEblocks.{i}", num_heads)
        x = _ln(x, sd, prefix + "Spatial_norm", 1e-6)
        x = x.reshape(b, l, j, c).permute(0, 2, 1, 3).reshape(b * j, l, c)
        x = block(x, sd, prefix + f"TTEblocks.{i}", num_heads)
        x = _ln(x, sd, prefix + "Temporal_norm", 1e-6)
        x = x.reshape(b, j, l, c).permute(0, 2, 1, 3)
    return x.contiguous()                                                             # [B, L, J, C]


def _depth_of(sd: Dict[str, torch.Tensor], prefix: str) -> int:
    d = 0
    while f"{prefix}STEblocks.{d}.norm1.weight" in sd:
        d += 1
    return d


def rotations_module(x: torch.Tensor, sd: Dict[str, torch.Tensor], num_heads: int = 8,
                     prefix: str = "rotations_module.") -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """RMCLRotMixSTE.forward, rmcl_manifold_mix_ste.py:239-264 (+ MCLHead :291-298).

    Returns (rot6d [B,K,L,J,D], scores [B,K,L,1], logits [B,K,L,1])."""
    depth = _depth_of(sd, prefix)
    emb = F.linear(x, sd[prefix + "Spatial_patch_to_embedding.weight"],
                   sd[prefix + "Spatial_patch_to_embedding.bias"])                   # mix_ste.py:134
    feat = mixste_trunk(emb, sd, prefix, depth, num_heads)
    preds, logits = [], []
    k = 0
    while f"{prefix}head.{k}.norm.weight" in sd:
        hp = f"{prefix}head.{k}"
        h = _ln(feat, sd, hp + ".norm", 1e-5)                                         # rmcl:277,292
        pe = F.linear(h, sd[hp + ".prediction_head.weight"], sd[hp + ".prediction_head.bias"])
        preds.append(pe[..., :-1])                                                    # rmcl:294
        logits.append(F.linear(pe[..., -1], sd[hp + ".score_head.weight"], sd[hp + ".score_head.bias"]))
        k += 1
    hyp = torch.stack(preds, dim=1)
    lg = torch.stack(logits, dim=1)
    return hyp, lg.softmax(dim=1), lg


def single_rotations_module(x: torch.Tensor, sd: Dict[str, torch.Tensor], num_heads: int = 8,
                            prefix: str = "rotations_module.") -> torch.Tensor:
    """MixSTE.forward, mix_ste.py:175-191 with the plain head (LN eps 1e-5 + Linear) -> [B,L,J,D]."""
    depth = _depth_of(sd, prefix)
    emb = F.linear(x, sd[prefix + "Spatial_patch_to_embedding.weight"],
                   sd[prefix + "Spatial_patch_to_embedding.bias"])
    feat = mixste_trunk(emb, sd, prefix, depth, num_heads)
    h = _ln(feat, sd, prefix + "head.0", 1e-5)
    return F.linear(h, sd[prefix + "head.1.weight"], sd[prefix + "head.1.bias"])


def segments_module(x: torch.Tensor, sd: Dict[str, torch.Tensor], num_heads: int = 8,
                    prefix: str = "segments_module.") -> torch.Tensor:
    """BonesMixSTE.forward, manifold_mix_ste.py:139-154 -> bone lengths [B, S, 1] (signed)."""
    b, l = x.shape[:2]
    depth = _depth_of(sd, prefix)
    w = sd[prefix + "joints_to_segments_proj.weight"]
    c = sd[prefix + "Spatial_pos_embed"].shape[-1]
    s = w.shape[0] // c
    emb = F.linear(x.reshape(b * l, -1), w, sd[prefix + "joints_to_segments_proj.bias"]).reshape(b, l, s, c)
    feat = mixste_trunk(emb, sd, prefix, depth, num_heads)
    h = _ln(feat, sd, prefix + "head.0", 1e-5)                                        # mix_ste.py:123-126
    out = F.linear(h, sd[prefix + "head.1.weight"], sd[prefix + "head.1.bias"])       # [B, L, S, 1]
    return out.mean(dim=1)                                                            # :153


def rmcl_forward(x: torch.Tensor, sd: Dict[str, torch.Tensor], num_heads_rot: int = 8,
                 num_heads_seg: int = 8) -> Tuple[torch.Tensor, torch.Tensor]:
    """RMCLManifoldMixSTE.forward, rmcl_manifold_mix_ste.py:83-106 -> (poses [B,K,L,J,3], scores [B,K,L,1])."""
    b, l = x.shape[:2]
    rot, scores, _ = rotations_module(x, sd, num_heads_rot)
    k, j, d = rot.shape[1], rot.shape[3], rot.shape[4]
    bones = segments_module(x, sd, num_heads_seg)
    root = torch.zeros(b * l * k, 3)
    poses = pose_decoder(rot.reshape(b * k * l, j, d), bones, root, rot_rep_dim=d)
    return poses.reshape(b, k, l, j, 3), scores


def manifold_forward(x: torch.Tensor, sd: Dict[str, torch.Tensor], num_heads_rot: int = 8,
                     num_heads_seg: int = 8) -> torch.Tensor:
    """ManifoldMixSTE.forward (single hypothesis), manifold_mix_ste.py:74-88 -> [B,L,J,3]."""
    b, l = x.shape[:2]
    rot = single_rotations_module(x, sd, num_heads_rot)
    j, d = rot.shape[2], rot.shape[3]
    bones = segments_module(x, sd, num_heads_seg)
    poses = pose_decoder(rot.reshape(b * l, j, d), bones, torch.zeros(b * l, 3), rot_rep_dim=d)
    return poses.reshape(b, l, j, 3)


# --------------------------------------------------------------------------------------
# L1-L6: losses   (hpe/mh_so3_hpe/metrics/losses.py, regularizations.py)
# --------------------------------------------------------------------------------------
def l2_loss_per_hyp(hyp: torch.Tensor, y: torch.Tensor, weights: Optional[torch.Tensor] = None,
                    squared: bool = False) -> torch.Tensor:
    """losses.py:104-123 (+ :14-72) — [B,K,L,J,3],[B,L,J,3] -> [B,K,L]."""
    tgt = y[:, None].expand_as(hyp)
    if squared:
        if weights is None:  # F.mse_loss over everything: a scalar (reference quirk, :57-58)
            return F.mse_loss(hyp, tgt)
        return (weights[None, None, :, None] * (hyp - tgt) ** 2).mean(dim=4).mean(dim=3)
    if weights is None:
        weights = torch.ones(y.shape[-2])
    return (weights[None, None, :] * torch.norm(hyp - tgt, p=2, dim=4)).mean(dim=3)


def wta_l2_loss_and_activate_head(hyp, y, weights=None, squared=False):
    """losses.py:126-138 — torch.min over the hypothesis dim -> (values [B,L], int64 indices [B,L])."""
    return torch.min(l2_loss_per_hyp(hyp, y, weights, squared), dim=1)


def wta_with_scoring_loss(hyp, scores, y, beta, weights=None, squared=False):
    """losses.py:141-170 — WTA mean + beta * BCE(scores, one_hot(winner)); bare scalar if beta == 0."""
    wta, idx = wta_l2_loss_and_activate_head(hyp, y, weights, squared)
    if beta == 0:
        return wta.mean()
    b, k, l = hyp.shape[:3]
    gt = F.one_hot(idx, k).permute(0, 2, 1).to(torch.float32)                         # :158-163
    bce = F.binary_cross_entropy(scores.view(b, k, l), gt)
    return wta.mean() + beta * bce, beta * bce


def mean_velocity_error(pred: torch.Tensor, target: torch.Tensor, axis: int = 1, squared: bool = False):
    """losses.py:75-101."""
    if pred.dim() > target.dim():
        target = target.unsqueeze(1).expand_as(pred)
    dv = torch.diff(pred, dim=axis) - torch.diff(target, dim=axis)
    if squared:
        return torch.mean(dv ** 2)
    return torch.mean(torch.norm(dv, dim=target.dim() - 1))


def smoothness_regularization(pred: torch.Tensor, weights: Optional[torch.Tensor] = None, axis: int = 1):
    """regularizations.py:160-174."""
    v = torch.diff(pred, dim=axis)
    if weights is None:
        # reference quirk kept: ones_like(v[0, 0, :, 0]) has J entries only for 4-D input (:165-166)
        weights = torch.ones_like(v[0, 0, :, 0])
    assert weights.shape[0] == v.shape[-2]
    return torch.mean(weights[None, None, :, None] * v ** 2)


def training_loss(poses, scores, y, beta=0.1, vel_w=2.0, smooth_w=0.5, weights=STANDARD_H36M_WEIGHTS,
                  squared=False):
    """make_loss + compute_and_acc_loss for the RMCL model, hpe/main_h36m_lifting.py:101-209 with the
    hpe/conf/config.yaml:33-37 defaults.  Returns (total, dict of terms)."""
    wl = wta_l2_loss_and_activate_head(poses, y, weights, squared)[0].mean()
    terms = {"wloss": wl}
    total = wl
    if beta != 0:
        terms["score_reg"] = wta_with_scoring_loss(poses, scores, y, beta, weights, squared)[1]
        total = total + terms["score_reg"]
    if vel_w > 0:
        terms["vloss"] = vel_w * mean_velocity_error(poses, y, axis=2, squared=squared)
        total = total + terms["vloss"]
    if smooth_w > 0:
        terms["sreg"] = smooth_w * smoothness_regularization(poses, weights, axis=2)
        total = total + terms["sreg"]
    return total, terms


# --------------------------------------------------------------------------------------
# M1-M3: hypothesis aggregation / metrics
# --------------------------------------------------------------------------------------
def poses_from_hyp_idx(hyp: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """rmcl_manifold_mix_ste.py:121-139 — gather hypothesis idx[b,l] -> [B,L,J,3]."""
    b, k, l, j, _ = hyp.shape
    gather = idx[:, None, :, None, None].expand(b, 1, l, j, 3)
    return hyp.gather(1, gather)[:, 0]


def aggregate(hyp, scores=None, mode="weighted_ave", ground_truth=None):
    """rmcl_manifold_mix_ste.py:141-185."""
    if mode == "best_score":
        assert scores is not None
        return poses_from_hyp_idx(hyp, torch.argmax(scores, dim=1)[..., 0])
    if mode == "weighted_ave":
        assert scores is not None
        return torch.sum(hyp * scores.unsqueeze(-1), dim=1)
    if mode == "oracle":
        assert ground_truth is not None
        val, idx = wta_l2_loss_and_activate_head(hyp, ground_truth, None, False)
        return val, poses_from_hyp_idx(hyp, idx)
    raise ValueError(f"Only best_score and weighted_ave modes are implemented.Got {mode}.")


def concat_hyp_and_scores(hyp: torch.Tensor, scores: torch.Tensor) -> torch.Tensor:
    """rmcl_manifold_mix_ste.py:108-119 -> [B,K,L,J,4]."""
    return torch.cat((hyp, scores.unsqueeze(3).expand(-1, -1, -1, hyp.shape[3], -1)), dim=-1)


def mpjpe_error(pred: torch.Tensor, gt: torch.Tensor, mode: str):
    """mean_joint_errors.py:8-36."""
    d = torch.norm(gt.reshape(-1, 3) - pred.reshape(-1, 3), 2, 1)
    if mode == "average":
        return d.mean()
    if mode == "sum":
        return d.sum()
    if mode == "no_agg":
        return d
    raise ValueError(f"Unexpected value for 'mode' encoutered: {mode}.")


# --------------------------------------------------------------------------------------
# P-MPJPE ("Protocol #2")
# --------------------------------------------------------------------------------------
def p_mpjpe(predicted: torch.Tensor, target: torch.Tensor) -> float:
    """hpe/mh_so3_hpe/metrics/mean_joint_errors.py:144-189: MPJPE after the optimal similarity alignment of every frame
    (orthogonal Procrustes through numpy's batched SVD of X0^T Y0, reflections excluded)."""
    import numpy as np
    assert predicted.shape == target.shape and predicted.shape[-1] == 3
    n_joints = predicted.shape[-2]
    pr = predicted.contiguous().view(-1, n_joints, 3).detach().cpu().numpy()
    tg = target.contiguous().view(-1, n_joints, 3).detach().cpu().numpy()
    mu_x = np.mean(tg, axis=1, keepdims=True)
    mu_y = np.mean(pr, axis=1, keepdims=True)
    x0, y0 = tg - mu_x, pr - mu_y
    norm_x = np.sqrt(np.sum(x0 ** 2, axis=(1, 2), keepdims=True))
    norm_y = np.sqrt(np.sum(y0 ** 2, axis=(1, 2), keepdims=True))
    x0 = x0 / norm_x
    y0 = y0 / norm_y
    h = np.matmul(x0.transpose(0, 2, 1), y0)
    u, sv, vt = np.linalg.svd(h)
    v = vt.transpose(0, 2, 1)
    r = np.matmul(v, u.transpose(0, 2, 1))
    sign_det = np.sign(np.expand_dims(np.linalg.det(r), axis=1))
    v[:, :, -1] *= sign_det
    sv[:, -1] *= sign_det.flatten()
    r = np.matmul(v, u.transpose(0, 2, 1))
    tr = np.expand_dims(np.sum(sv, axis=1, keepdims=True), axis=2)
    a = tr * norm_x / norm_y
    t = mu_x - a * np.matmul(mu_y, r)
    aligned = a * np.matmul(pr, r) + t
    return float(np.mean(np.linalg.norm(aligned - tg, axis=2)))


MISS_TYPE_RATES = {"no_miss": 0.2, "random": 0.2, "random_left_arm_right_leg": 0.4, "structured_joint": 0.4, "structured_frame": 0.2}


def occlusion_mask(seq_len: int, n_joints: int, in_chans: int, miss_type: str, miss_rate: float, noise_sigma: float):
    """hpe/mh_so3_hpe/data/generators.py:157-210 for one item, drawing from numpy's global RNG in the reference's order:
    -> (mask [L, J] float64, noise [L, J, C] float64 or None)."""
    import math
    import numpy as np
    shape = (seq_len, n_joints)
    if miss_type == "all":
        miss_type = np.random.choice(list(MISS_TYPE_RATES.keys()))
        miss_rate = MISS_TYPE_RATES[miss_type]
    mask, noise = np.ones(shape), None
    if miss_type == "no_miss":
        pass
    elif miss_type == "random":
        mask = (np.random.uniform(0.0, 1.0, size=shape) > miss_rate).astype(np.float64)
    elif miss_type == "random_left_arm_right_leg":
        frames = np.random.choice(seq_len, size=math.floor(miss_rate * seq_len), replace=False).tolist()
        for joint in (1, 2, 3, 11, 12, 13):
            mask[frames, joint] = 0.0
    elif miss_type in ("structured_joint", "structured_frame"):
        occl = int(seq_len * miss_rate)
        first = np.random.choice(seq_len - occl, size=1, replace=False)[0]
        if miss_type == "structured_joint":
            mask[first:first + occl, [1, 2, 3]] = 0.0
        else:
            mask[first:first + occl] = 0.0
    elif miss_type == "noisy":
        noise = np.random.normal(0, noise_sigma, size=shape + (in_chans,))
    else:
        raise ValueError(f"Unexpected miss_type: {miss_type}")
    return mask, noise


def sequence_windows(poses_3d, poses_2d, seq_len: int, drop_last: bool = True, random_start: bool = False, miss_type: str = "no_miss",
                     miss_rate: float = 0.2, noise_sigma: float = 5, indices=None, flip_probability=None):
    """hpe/mh_so3_hpe/data/generators.py:83-219 (PoseSequenceGenerator): the items (pose_2d [L,J,C] * mask, pose_3d [L,J,3]) for ``indices``
    (default: the whole dataset in order); the last, shorter window of a sequence is replicate-padded when drop_last is False; random
    starts come from torch's global RNG, the PoseFlip transform's coin (augmentations/transforms.py:21-31, ``flip_probability``) from
    torch's, masks / noise from numpy's, per item in that order like the reference."""
    table = []
    for s, p3 in enumerate(poses_3d):
        n = p3.shape[0]
        size = n // seq_len + (1 if (not drop_last and n % seq_len > 0) else 0)
        table += [(s, k * seq_len) for k in range(size)]
    items = []
    for i in (range(len(table)) if indices is None else indices):
        s, start = table[i]
        t3, t2 = torch.from_numpy(poses_3d[s]).float(), torch.from_numpy(poses_2d[s]).float()
        n = t3.shape[0]
        if random_start:
            start = torch.randint(low=0, high=n - seq_len, size=(1,)).item()
        idx = torch.clamp(torch.arange(start, start + seq_len), max=n - 1)          # replicate padding = clamp to the last frame
        p2, p3 = t2[idx], t3[idx]
        if flip_probability is not None and torch.rand(1).item() <= flip_probability:
            p2, p3 = pose_flip(p2), pose_flip(p3)
        mask, noise = occlusion_mask(seq_len, p2.shape[1], p2.shape[2], miss_type, miss_rate, noise_sigma)
        if noise is not None:
            p2 = p2.double() + torch.from_numpy(noise)   # `pose_2d += noise` with a float64 ndarray re-binds pose_2d to the float64 sum; callers `.float()` it
        items.append((p2 * torch.from_numpy(mask[..., None]).float(), p3))
    return items


def keypoint_3d_pck(pred: torch.Tensor, gt: torch.Tensor, threshold: float = 150.0) -> float:
    """hpe/mh_so3_hpe/metrics/pck.py:77-141 with alignment='none', mask=None: pred / gt [N, K, 3]."""
    import numpy as np
    p, g = pred.detach().cpu().numpy(), gt.detach().cpu().numpy()
    mask = np.ones(g.shape[:2]).astype(bool)
    error = np.linalg.norm(p - g, ord=2, axis=-1)
    return float((error < threshold).astype(np.float32)[mask].mean() * 100)


def keypoint_3d_auc(pred: torch.Tensor, gt: torch.Tensor) -> float:
    """pck.py:144-198 with alignment='none', mask=None."""
    import numpy as np
    p, g = pred.detach().cpu().numpy(), gt.detach().cpu().numpy()
    mask = np.ones(g.shape[:2]).astype(bool)
    error = np.linalg.norm(p - g, ord=2, axis=-1)
    thresholds = np.linspace(0., 150, 31)
    pck_values = np.zeros(len(thresholds))
    for i in range(len(thresholds)):
        pck_values[i] = (error < thresholds[i]).astype(np.float32)[mask].mean()
    return float(pck_values.mean() * 100)


# --------------------------------------------------------------------------------------
# evaluation epilogue with flip test-time augmentation
# --------------------------------------------------------------------------------------
def tta_prediction(x: torch.Tensor, sd: Dict[str, torch.Tensor], mode: str = "weighted_ave") -> torch.Tensor:
    """hpe/eval_utils.py:51-56,83-142 for the RMCL model: aggregate(model(x)), aggregate(model(pose_flip(x))) flipped back, averaged.
    ``pose_flip`` (augmentations/functional.py:7-28) works in place on its argument; a copy is flipped here."""
    poses, scores = rmcl_forward(x, sd)
    pred = aggregate(poses, scores, mode)
    poses_f, scores_f = rmcl_forward(pose_flip(x.clone()), sd)
    pred_f = pose_flip(aggregate(poses_f, scores_f, mode).clone())
    return (pred + pred_f) / 2


def evaluate(batches, sd: Dict[str, torch.Tensor], tta: bool, return_hyps: bool = False, compute_oracle: bool = True, forward=None):
    """hpe/eval_utils.py:16-203 (``evaluate`` + ``evaluation_metrics``) for the RMCL model, bookkeeping included: per batch the
    weighted-average prediction, the oracle figure (non-TTA: sum of the per-frame joint-MEAN errors divided by J once more, :57-64 — the
    reference's normalisation quirk, kept), the best-score ("per-sample oracle") figure, their flip-TTA variants (:95-135), MPJPE
    average / sum in mm; totals normalised as :186-197.  ``batches`` = iterable of (input_2d [B,L,J,2], target_3d [B,L,J,3]).
    ``forward`` (x -> poses, scores) replaces the oracle forward when only the bookkeeping is under test."""
    forward = forward or (lambda inp: rmcl_forward(inp, sd))
    mpjpe_total, m_p3d, n, batch_no = 0.0, 0.0, 0, 0
    oracle_total, psoracle_total = 0, 0
    all_pred, all_target, all_oracle = [], [], []
    for batch_no, (x, y) in enumerate(batches, start=1):
        b, l, j, _ = y.shape
        y = y.float()
        poses, scores = forward(x.float())
        hyp = concat_hyp_and_scores(poses, scores)
        pred = aggregate(poses, scores, "weighted_ave")
        if compute_oracle:
            val, oracle_preds = aggregate(hyp[..., :-1], mode="oracle", ground_truth=y)
            oracle_mpjpe = val.sum() / j
            ps_preds = aggregate(hyp[..., :-1], scores=hyp[..., -1], mode="best_score")
            ps_mpjpe = mpjpe_error(ps_preds, y, "sum") / j
        if tta:
            poses_f, scores_f = forward(pose_flip(x.float()))
            pred_f = aggregate(poses_f, scores_f, "weighted_ave")
            if compute_oracle:
                hyp_f = pose_flip(poses_f)
                _, oracle_f = aggregate(hyp_f, mode="oracle", ground_truth=y)
                oracle_preds = (oracle_preds + oracle_f) / 2
                oracle_mpjpe = mpjpe_error(oracle_preds, y, "sum") / j
                ps_f = aggregate(hyp_f, scores=scores_f.expand(-1, -1, -1, j), mode="best_score")
                ps_mpjpe = mpjpe_error((ps_preds + ps_f) / 2, y, "sum") / j
            pred = (pred + pose_flip(pred_f)) / 2
        n += b
        if return_hyps:
            hyp = hyp.clone()
            hyp[..., :-1] *= 1000
            all_pred.append(hyp)
        else:
            all_pred.append(pred * 1000)
        mpjpe_total += (mpjpe_error(pred, y, "average") * 1000).item()
        m_p3d += (mpjpe_error(pred, y, "sum") * 1000).numpy()
        if compute_oracle:
            oracle_total = oracle_total + oracle_mpjpe
            psoracle_total = psoracle_total + ps_mpjpe
            all_oracle.append(oracle_preds * 1000)
        all_target.append(y)
    performance = m_p3d / (n * l * j)
    if not compute_oracle:
        return all_pred, all_target, performance
    return all_pred, all_target, performance, oracle_total / (n * l) * 1000, psoracle_total / (n * l) * 1000, all_oracle


def _agg(mode: str):
    """mean_joint_errors.py:8-28: 'average' -> torch.mean, 'sum' -> torch.sum, 'no_agg' -> identity."""
    if mode == "average":
        return torch.mean
    if mode == "sum":
        return torch.sum
    if mode == "no_agg":
        return lambda x, dim=None: x
    raise ValueError(f"Unexpected value for 'mode' encoutered: {mode}.")


def mse_error(pred: torch.Tensor, gt: torch.Tensor, mode: str):
    """mean_joint_errors.py:39-44: squared distance of every 3-D point."""
    return _agg(mode)(torch.sum((gt.reshape(-1, 3) - pred.reshape(-1, 3)) ** 2, dim=1))


def jointwise_error(pred: torch.Tensor, gt: torch.Tensor, mode: str, squared: bool = False):
    """mean_joint_errors.py:47-80 (jointwise_error; squared=True: jointwise_mse): per-joint error aggregated over dim 0 -> [J]."""
    j = gt.shape[-2]
    d = gt.contiguous().view(-1, j, 3) - pred.contiguous().view(-1, j, 3)
    e = torch.sum(d ** 2, dim=2) if squared else torch.norm(d, 2, 2)
    return _agg(mode)(e, dim=0)


def coordwise_error(pred: torch.Tensor, gt: torch.Tensor, mode: str):
    """mean_joint_errors.py:132-141: |gt - pred| per coordinate aggregated over every point -> [3]."""
    return _agg(mode)(torch.abs(gt.contiguous().view(-1, 3) - pred.contiguous().view(-1, 3)), dim=0)


def segments_len_err(pred_jc: torch.Tensor, gt_jc: torch.Tensor, mode: str, signed: bool = True, bones=None):
    """mean_joint_errors.py:83-129: gt - predicted bone lengths, inputs [B,3,J,L] -> scalar, or [B*L, num_bones] for 'no_agg'."""
    bones = H36M17_BONES if bones is None else bones
    b, _, _, l = pred_jc.shape
    pl = measure_bones_length(pred_jc, bones).permute(0, 2, 1).reshape(b * l, -1)
    gl = measure_bones_length(gt_jc, bones).permute(0, 2, 1).reshape(b * l, -1)
    diff = gl - pl
    if not signed:
        diff = torch.abs(diff)
    return _agg(mode)(diff)


# --------------------------------------------------------------------------------------
# pose-consistency metrics (bone lengths, MPSCE, MPSSE)
# --------------------------------------------------------------------------------------
H36M17_BONES = tuple((j, p) for j, p in enumerate(H36M17_PARENTS) if p >= 0)      # skeleton.py:100-103: (joint, parent)
H36M17_BONES_LEFT = (3, 4, 5, 10, 11, 12)                                          # skeleton.py:110-120 on joints_left / joints_right
H36M17_BONES_RIGHT = (0, 1, 2, 13, 14, 15)


def measure_bones_length(joints_coords: torch.Tensor, bones=H36M17_BONES) -> torch.Tensor:
    """hpe/mh_so3_hpe/metrics/utils.py:4-20: joints_coords [B,3,J,L] -> [B,num_bones,L]."""
    b, three, n_joints, length = joints_coords.shape
    assert three == 3 and n_joints == len(bones) + 1
    out = torch.empty((b, len(bones), length), dtype=joints_coords.dtype)
    for i, (j, p) in enumerate(bones):
        out[:, i, :] = torch.sum((joints_coords[:, :, j, :] - joints_coords[:, :, p, :]) ** 2, axis=1).sqrt()
    return out


def segments_time_consistency(joints_coords: torch.Tensor, mode: str, per_bone: bool = False, bones=H36M17_BONES):
    """regularizations.py:8-61: var (std for mode 'std') over time of the bone lengths, aggregated over (batch, bone) or over batch."""
    lengths = measure_bones_length(joints_coords, bones)
    stat = torch.std if mode == "std" else torch.var
    agg = {"average": torch.mean, "std": torch.mean, "sum": torch.sum, "min": torch.min, "max": torch.max}[mode]
    v = stat(lengths, dim=2)
    return agg(v, dim=0) if per_bone else agg(v)


def sagittal_symmetry(joints_coords: torch.Tensor, mode: str, squared: bool = True, per_bone: bool = False, bones=H36M17_BONES,
                      left=H36M17_BONES_LEFT, right=H36M17_BONES_RIGHT):
    """regularizations.py:103-157: |len[left] - len[right]| (squared by default), mean or sum."""
    lengths = measure_bones_length(joints_coords, bones)
    agg = {"average": torch.mean, "sum": torch.sum}[mode]
    diff = (lengths[:, list(left), :] - lengths[:, list(right), :]).abs()
    if squared:
        diff = diff ** 2.0
    return agg(diff.permute(0, 2, 1).reshape(-1, len(left)), dim=0) if per_bone else agg(diff)


def pose_flip(x: torch.Tensor, left=H36M17_JOINTS_LEFT, right=H36M17_JOINTS_RIGHT) -> torch.Tensor:
    """hpe/mh_so3_hpe/augmentations/functional.py:7-28 — negate x coordinate, swap L/R joints (copy)."""
    out = x.clone()
    out[..., 0] *= -1
    out[..., left + right, :] = out[..., right + left, :]
    return out


# --------------------------------------------------------------------------------------
# Synthetic weights (no reference needed): same names/shapes as SURVEY.md §A.3
# --------------------------------------------------------------------------------------
def make_state_dict(num_frame=243, n_hyp=5, num_joints=17, num_bones=16, in_chans=2, rot_rep_dim=6,
                    embed_dim_rot=512, depth_rot=8, embed_dim_seg=128, depth_seg=2, mlp_ratio=2.0,
                    seed=0, std=0.02, single_head=False) -> Dict[str, torch.Tensor]:
    """Seeded synthetic state_dict with the reference's key names and shapes.  Linear weights ~
    N(0, 1/fan_in) (close to nn.Linear's default scale), LN affine and pos-embeds perturbed by ``std``
    so zero-init pos-embeds and unit LN affine are exercised."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def lin(name, out_f, in_f):
        sd[name + ".weight"] = torch.randn(out_f, in_f, generator=g) / math.sqrt(in_f)
        sd[name + ".bias"] = torch.randn(out_f, generator=g) * std

    def ln(name, c):
        sd[name + ".weight"] = 1.0 + torch.randn(c, generator=g) * std
        sd[name + ".bias"] = torch.randn(c, generator=g) * std

    def trunk(prefix, c, depth, n_tok):
        sd[prefix + "Spatial_pos_embed"] = torch.randn(1, n_tok, c, generator=g) * std
        sd[prefix + "Temporal_pos_embed"] = torch.randn(1, num_frame, c, generator=g) * std
        hidden = int(c * mlp_ratio)
        for kind in ("STEblocks", "TTEblocks"):
            for i in range(depth):
                p = f"{prefix}{kind}.{i}"
                ln(p + ".norm1", c)
                lin(p + ".attn.qkv", 3 * c, c)
                lin(p + ".attn.proj", c, c)
                ln(p + ".norm2", c)
                lin(p + ".mlp.fc1", hidden, c)
                lin(p + ".mlp.fc2", c, hidden)
        ln(prefix + "Spatial_norm", c)
        ln(prefix + "Temporal_norm", c)

    rp = "rotations_module."
    lin(rp + "Spatial_patch_to_embedding", embed_dim_rot, in_chans)
    trunk(rp, embed_dim_rot, depth_rot, num_joints)
    if single_head:
        ln(rp + "head.0", embed_dim_rot)
        lin(rp + "head.1", rot_rep_dim, embed_dim_rot)
    else:
        for k in range(n_hyp):
            ln(f"{rp}head.{k}.norm", embed_dim_rot)
            lin(f"{rp}head.{k}.prediction_head", rot_rep_dim + 1, embed_dim_rot)
            lin(f"{rp}head.{k}.score_head", 1, num_joints)
    sp = "segments_module."
    lin(sp + "joints_to_segments_proj", num_bones * embed_dim_seg, num_joints * in_chans)
    trunk(sp, embed_dim_seg, depth_seg, num_bones)
    ln(sp + "head.0", embed_dim_seg)
    lin(sp + "head.1", 1, embed_dim_seg)
    return sd

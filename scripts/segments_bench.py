"""Development aid: the CUDA-core kernels of the bone-length backbone (C = 128) alone, checked against torch fp32 and timed with an
L2 flush between launches.  Usage: python scripts/segments_bench.py [clips]"""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from manipose_b200 import _lib as L, ops  # noqa: E402

dev = torch.device("cuda")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    clips = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    t, s, c = 243, 16, 128
    frames = clips * t
    n = frames * s
    g = torch.Generator(device=dev).manual_seed(0)
    r = lambda *sh: torch.randn(*sh, generator=g, device=dev)
    out = {"clips": clips, "tokens": n}
    hbm = 6436.1

    # ---- segment embedding: Linear(34 -> 16 * 128) + spatial position embedding + norm1
    x2d = 0.3 * r(frames, 34)
    w, b, spos = r(s * c, 34) / 6.0, 0.1 * r(s * c), 0.02 * r(s, c)
    lg, lb = 1.0 + 0.1 * r(c), 0.1 * r(c)
    x = torch.empty(n, c, device=dev)
    h = torch.empty(n, c, dtype=torch.bfloat16, device=dev)
    run = lambda: ops.embed_segments(x2d, w, b, spos.reshape(-1), lg, lb, 1e-6, x, h, frames, 34, s, c, L.MP_DTYPE_BF16)
    us = timeit(run)
    want = (F.linear(x2d.double(), w.double(), b.double()).view(frames, s, c) + spos.double()).view(n, c)
    err_x = float((x.double() - want).abs().max())
    err_h = float((h.double() - F.layer_norm(want, (c,), lg.double(), lb.double(), 1e-6)).abs().max())
    out["embed_segments"] = {"us": us, "gbs": n * c * 6 / us / 1e3, "frac": n * c * 6 / us / 1e3 / hbm, "max_err_x": err_x, "max_err_h": err_h}

    # ---- LayerNorm, both forms the trunk uses
    xin = r(n, c)
    us = timeit(lambda: ops.layernorm(xin, None, h, ln=(lg, lb), ln_eps=1e-6, dtype=L.MP_DTYPE_BF16))
    err = float((h.double() - F.layer_norm(xin.double(), (c,), lg.double(), lb.double(), 1e-6)).abs().max())
    out["layernorm"] = {"us": us, "gbs": n * c * 6 / us / 1e3, "frac": n * c * 6 / us / 1e3 / hbm, "max_err_h": err}
    pg, pb, pos = 1.0 + 0.1 * r(c), 0.1 * r(c), 0.02 * r(t, c)
    xo = torch.empty_like(xin)
    us = timeit(lambda: ops.layernorm(xin, xo, h, post=(pg, pb), post_eps=1e-6, pos=pos, pos_div=s, pos_mod=t, ln=(lg, lb), ln_eps=1e-6,
                                      dtype=L.MP_DTYPE_BF16))
    wx = F.layer_norm(xin.double(), (c,), pg.double(), pb.double(), 1e-6).view(clips, t, s, c) + pos.double()[None, :, None, :]
    wx = wx.view(n, c)
    err_x = float((xo.double() - wx).abs().max())
    err_h = float((h.double() - F.layer_norm(wx, (c,), lg.double(), lb.double(), 1e-6)).abs().max())
    out["layernorm_post_pos"] = {"us": us, "gbs": n * c * 10 / us / 1e3, "frac": n * c * 10 / us / 1e3 / hbm, "max_err_x": err_x, "max_err_h": err_h}

    # ---- bone-length head: Temporal_norm -> LN(1e-5) -> Linear(128 -> 1) -> mean over frames
    hg, hb_, hw, hbias = 1.0 + 0.1 * r(c), 0.1 * r(c), r(c) / 11.0, 0.1 * r(1)
    bone = torch.empty(clips, s, device=dev)
    ws = torch.empty(n, device=dev)
    us = timeit(lambda: ops.bones_head(xin, pg, pb, 1e-6, hg, hb_, hw, hbias, bone, clips, t, s, c, ws))
    y = F.layer_norm(F.layer_norm(xin.double(), (c,), pg.double(), pb.double(), 1e-6), (c,), hg.double(), hb_.double(), 1e-5)
    wb = (y @ hw.double() + hbias.double()).view(clips, t, s).mean(1)
    out["bones_head"] = {"us": us, "gbs": n * c * 4 / us / 1e3, "frac": n * c * 4 / us / 1e3 / hbm, "max_err": float((bone.double() - wb).abs().max())}
    # ---- joint embedding of the rotations backbone (C = 512): Linear(2 -> 512) + spatial position embedding + norm1
    del x, h, xin, xo
    nj = clips * t * 17
    c5 = 512
    x2 = 0.3 * r(nj, 2)
    w5, b5, sp5 = r(c5, 2), 0.1 * r(c5), 0.02 * r(17, c5)
    g5, bt5 = 1.0 + 0.1 * r(c5), 0.1 * r(c5)
    x5 = torch.empty(nj, c5, device=dev)
    h5 = torch.empty(nj, c5, dtype=torch.bfloat16, device=dev)
    us = timeit(lambda: ops.embed_joints(x2, w5, b5, sp5, g5, bt5, 1e-6, x5, h5, nj, 17, c5, L.MP_DTYPE_BF16))
    want = (F.linear(x2.double(), w5.double(), b5.double()).view(-1, 17, c5) + sp5.double()).view(nj, c5)
    err_x = float((x5[:100000].double() - want[:100000]).abs().max())
    err_h = float((h5[:100000].double() - F.layer_norm(want[:100000], (c5,), g5.double(), bt5.double(), 1e-6)).abs().max())
    out["embed_joints"] = {"us": us, "gbs": nj * c5 * 6 / us / 1e3, "frac": nj * c5 * 6 / us / 1e3 / hbm, "max_err_x": err_x, "max_err_h": err_h}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

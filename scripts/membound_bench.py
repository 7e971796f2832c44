"""The memory-bound rows alone (decoder forward / backward, objective forward / backward, aggregation, MPJPE) through the C ABI,
CUDA events, L2 flushed between launches: achieved GB/s of the ALGORITHMIC bytes against MEASURED_PEAKS.json.
Development aid; `python scripts/membound_bench.py [--once]` (--once: one launch per row, for ncu)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from manipose_b200 import ops, _lib as L  # noqa: E402

dev = torch.device("cuda")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
ONCE = "--once" in sys.argv


def timeit(fn, iters=20):
    if ONCE:
        fn()
        torch.cuda.synchronize()
        return 0.0, 0.0
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
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    lib = L.load()
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = peaks.get("hbm_gbs") or 6436.1
    out = {}
    g = torch.Generator(device=dev).manual_seed(0)

    def row(name, fn, nbytes, units=None):
        med, best = timeit(fn)
        r = {"us": med * 1e3, "best_us": best * 1e3}
        if med > 0:
            r["gbs"] = nbytes / med / 1e6
            r["frac"] = r["gbs"] / hbm
            if units:
                r["gunits_s"] = units / med / 1e6
        out[name] = r
        print(name, {k: round(v, 3) for k, v in r.items()}, flush=True)

    # ---- decoder, BASELINE config 2: 824 clips x 5 x 243
    nc, k, t = 824, 5, 243
    n = nc * k * t
    rot = torch.randn(n, 17, 6, generator=g, device=dev)
    bones = 0.1 + 0.4 * torch.rand(nc, 16, generator=g, device=dev)
    logits = torch.randn(nc, k, t, generator=g, device=dev)
    poses = torch.empty(n, 17, 3, device=dev)
    scores = torch.empty(nc, k, t, device=dev)
    for exact in (True, False):
        fl = 0 if exact else L.MP_DEC_FAST
        row(f"decoder_fwd_{'exact' if exact else 'fast'}",
            lambda: L.check(lib.mp_decoder_fwd(L.ptr(rot), L.ptr(bones), None, L.ptr(logits), L.ptr(poses), L.ptr(scores), nc, k, t, 6, fl,
                                               L.stream_ptr()), "dec"), 620.0 * n, n)
    gp = torch.randn(n, 17, 3, generator=g, device=dev)
    grot = torch.empty_like(rot)
    gbone = torch.zeros(nc, 16, device=dev)
    dwsb = L.load().mp_decoder_bwd_workspace_bytes(nc, k, t)
    dws = torch.empty(dwsb, dtype=torch.uint8, device=dev)
    row("decoder_bwd", lambda: L.check(lib.mp_decoder_bwd(L.ptr(rot), L.ptr(bones), L.ptr(gp), L.ptr(grot), L.ptr(gbone), None, nc, k, t, 6,
                                                           L.ptr(dws), dwsb, L.stream_ptr()), "bwd"), 1020.0 * n, n)
    del rot, poses, gp, grot

    # ---- objective, 1024 clips x 5 x 243
    B, K, T = 1024, 5, 243
    y = 0.3 * torch.randn(B, T, 17, 3, generator=g, device=dev)
    hyp = y[:, None] + 0.1 * torch.randn(B, K, T, 17, 3, generator=g, device=dev)
    sc = torch.softmax(torch.randn(B, K, T, generator=g, device=dev), 1)
    w = torch.tensor([1, 1, 2.5, 2.5, 1, 2.5, 2.5, 1, 1, 1, 1.5, 1.5, 4, 4, 1.5, 4, 4.0], device=dev)
    nfr = B * T
    terms = torch.empty(8, device=dev)
    val = torch.empty(B, T, device=dev)
    idx = torch.empty(B, T, dtype=torch.int64, device=dev)
    wsb = lib.mp_loss_workspace_bytes(B, K, T)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    row("loss_fwd", lambda: L.check(lib.mp_loss_fwd(L.ptr(hyp), L.ptr(sc), L.ptr(y), L.ptr(w), 0, 0.1, 2.0, 0.5, L.ptr(terms), L.ptr(val), L.ptr(idx),
                                                    B, K, T, L.ptr(ws), wsb, L.stream_ptr()), "loss"), (204.0 * 6 + 20 + 12) * nfr, nfr)
    gt = torch.zeros(8, device=dev)
    gt[4] = 1.0
    gh = torch.empty_like(hyp)
    gs = torch.empty_like(sc)
    row("loss_bwd", lambda: L.check(lib.mp_loss_bwd(L.ptr(hyp), L.ptr(sc), L.ptr(y), L.ptr(w), L.ptr(idx), 0, 0.1, 2.0, 0.5, L.ptr(gt), None, L.ptr(gh),
                                                    L.ptr(gs), B, K, T, L.stream_ptr()), "lossb"), (204.0 * 11 + 40 + 8) * nfr, nfr)
    del gh
    row("wta_fwd", lambda: ops.wta_fwd(hyp, y, w, False), (204.0 * 6 + 12) * nfr, nfr)
    for mode, nm, byt in ((0, "weighted_ave", 204.0 * 6 + 20), (1, "best_score", 204.0 * 2 + 28), (2, "oracle", 204.0 * 8 + 12)):
        row(f"aggregate_{nm}", lambda: ops.aggregate(hyp, sc, y, mode), byt * nfr, nfr)
    pred = ops.aggregate(hyp, sc, None, 0)[0]
    row("mpjpe", lambda: ops.mpjpe(pred, y), 408.0 * nfr, nfr)
    print(json.dumps({"hbm_gbs": hbm, "rows": out}))


if __name__ == "__main__":
    main()

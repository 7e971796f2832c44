"""Per-kernel timings on one B200 (CUDA events, L2 flushed between iterations): achieved TFLOP/s / GB/s against
MEASURED_PEAKS.json.  Development aid; the judged numbers come from bench.py."""
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from manipose_b200 import ops, _lib as L  # noqa: E402

dev = torch.device("cuda")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10, flush_l2=True):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        if flush_l2:
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
    clips = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    T, J = 243, 17
    m = clips * T * J
    out = {}
    g = torch.Generator(device=dev).manual_seed(0)
    for name, n, k, epi in (("qkv", 1536, 512, 0), ("proj", 512, 512, 2), ("fc1", 1024, 512, 1), ("fc2", 512, 1024, 2)):
        a = torch.randn(m, k, generator=g, device=dev).bfloat16()
        w = (torch.randn(n, k, generator=g, device=dev) / math.sqrt(k)).bfloat16()
        b = torch.randn(n, generator=g, device=dev)
        y = torch.randn(m, n, generator=g, device=dev).bfloat16()
        yres = torch.randn(m, n, generator=g, device=dev)
        for fl in (True, False):
            med, best = timeit(lambda: ops.linear(a, w, b, (yres if epi == 2 else y), epi, resid=yres if epi == 2 else None), flush_l2=fl)
            fl_ops = 2.0 * m * n * k
            byt = (m * k + n * k) * 2 + m * n * (8 if epi == 2 else 2)
            out[f"gemm_{name}_M{m}" + ("" if fl else "_warmL2")] = {"ms": med, "best_ms": best, "tflops": fl_ops / med / 1e9, "gbs": byt / med / 1e6}
        t_ref, _ = timeit(lambda: torch.matmul(a, w.t()), flush_l2=True)
        out[f"cublas_{name}_M{m}"] = {"ms": t_ref, "tflops": 2.0 * m * n * k / t_ref / 1e9}
    for name, k in (("proj+ln2", 512), ("fc2+post+ln1", 1024)):
        a = torch.randn(m, k, generator=g, device=dev).bfloat16()
        w = (torch.randn(512, k, generator=g, device=dev) / math.sqrt(k)).bfloat16()
        b = torch.randn(512, generator=g, device=dev)
        xx = torch.randn(m, 512, generator=g, device=dev)
        hh = torch.empty(m, 512, dtype=torch.bfloat16, device=dev)
        pp = [torch.randn(512, generator=g, device=dev) for _ in range(4)]
        post = (pp[0], pp[1]) if k == 1024 else None
        med, best = timeit(lambda: ops.linear_ln(a, w, b, xx, xx, hh, post=post, ln=(pp[2], pp[3])))
        byt = (m * k + 512 * k) * 2 + m * 512 * (8 + 2)
        out[f"fused_{name}_M{m}"] = {"ms": med, "best_ms": best, "tflops": 2.0 * m * 512 * k / med / 1e9, "gbs": byt / med / 1e6}
    qkv = torch.randn(m, 1536, generator=g, device=dev).bfloat16()
    o = torch.empty(m, 512, dtype=torch.bfloat16, device=dev)
    for mode, nm in ((1, "temporal"), (0, "spatial")):
        med, best = timeit(lambda: ops.attention(qkv, o, clips, T, J, 512, 8, mode))
        seq = T if mode == 1 else J
        fl_ops = 4.0 * m * seq * 512
        out[f"attn_{nm}"] = {"ms": med, "best_ms": best, "tflops": fl_ops / med / 1e9, "gbs": m * 2048 * 2 / med / 1e6}
    # K = 5 hypothesis heads: fp32 CUDA-core kernel (mp_heads_fwd) vs the folded 16-bit tensor-core projection (mp_heads_fwd16)
    kh, d1 = 5, 7
    xf = torch.randn(m, 512, generator=g, device=dev)
    xh = torch.randn(m, 512, generator=g, device=dev).bfloat16()
    hp = [torch.randn(512, generator=g, device=dev) for _ in range(2)]
    hg, hb = torch.randn(kh, 512, generator=g, device=dev), torch.randn(kh, 512, generator=g, device=dev)
    hw, hbias = torch.randn(kh, d1, 512, generator=g, device=dev) * 0.02, torch.randn(kh, d1, generator=g, device=dev)
    sw, sb = torch.randn(kh, 17, generator=g, device=dev), torch.randn(kh, generator=g, device=dev)
    rot = torch.empty(clips, kh, T, J, 6, device=dev)
    lg = torch.empty(clips, kh, T, device=dev)
    med, _ = timeit(lambda: ops.heads_fwd(xf, hp[0], hp[1], 1e-6, hg, hb, hw, hbias, sw, sb, rot, lg, clips, T, kh, 6, True))
    out["heads_fp32"] = {"ms": med, "gbs": (m * 2048 + rot.numel() * 4) / med / 1e6}
    wf16 = torch.zeros(128, 512, device=dev).bfloat16()
    bfold = torch.zeros(128, device=dev)
    ws = torch.empty(m * 128, device=dev)
    med, _ = timeit(lambda: ops.heads_fwd16(xh, wf16, bfold, sw, sb, rot, lg, ws, clips, T, kh, 6, True))
    out["heads_tensor_core"] = {"ms": med, "gbs": (m * 1024 + rot.numel() * 4) / med / 1e6,
                                "note": "algorithmic bytes: 16-bit normalised input + rot/logits out; the fp32 [tokens, 128] intermediate adds 2 x 512 B per token"}
    del xf, xh, ws
    x = torch.randn(m, 512, generator=g, device=dev)
    h = torch.empty(m, 512, dtype=torch.bfloat16, device=dev)
    p = [torch.randn(512, generator=g, device=dev) for _ in range(4)]
    med, _ = timeit(lambda: ops.layernorm(x, x, h, post=(p[0], p[1]), ln=(p[2], p[3])))
    out["layernorm_post+pre"] = {"ms": med, "gbs": m * 512 * 10 / med / 1e6}
    med, _ = timeit(lambda: ops.layernorm(x, None, h, ln=(p[2], p[3])))
    out["layernorm_pre"] = {"ms": med, "gbs": m * 512 * 6 / med / 1e6}
    # decoder, BASELINE config 2
    nc, k, t = 824, 5, 243
    n = nc * k * t
    rot = torch.randn(n, 17, 6, generator=g, device=dev)
    bones = 0.1 + 0.4 * torch.rand(nc, 16, generator=g, device=dev)
    logits = torch.randn(nc, k, t, generator=g, device=dev)
    for exact in (True, False):
        med, best = timeit(lambda: ops.decoder_fwd(rot, bones, None, logits, nc, k, t, 6, exact))
        out[f"decoder_fwd_{'exact' if exact else 'fast'}_1M"] = {"ms": med, "best_ms": best, "gbs": 620.0 * n / med / 1e6, "gposes_s": n / med / 1e6}
    # decoder backward: reads rot6d 408 + grad_pose 204, writes grad_rot6d 408 (+ per-clip bone-length grads)
    gp = torch.randn(n, 17, 3, generator=g, device=dev)
    grot = torch.empty_like(rot)
    gbone = torch.zeros(nc, 16, device=dev)
    dwsb = L.load().mp_decoder_bwd_workspace_bytes(nc, k, t)
    dws = torch.empty(dwsb, dtype=torch.uint8, device=dev)
    def dec_bwd():
        gbone.zero_()
        L.check(L.load().mp_decoder_bwd(L.ptr(rot), L.ptr(bones), L.ptr(gp), L.ptr(grot), L.ptr(gbone), None, nc, k, t, 6, L.ptr(dws), dwsb, L.stream_ptr()), "bwd")
    med, best = timeit(dec_bwd)
    out["decoder_bwd_1M"] = {"ms": med, "best_ms": best, "gbs": 1020.0 * n / med / 1e6, "gposes_s": n / med / 1e6}
    y = 0.3 * torch.randn(1024, 243, 17, 3, generator=g, device=dev)
    hyp = y[:, None] + 0.1 * torch.randn(1024, 5, 243, 17, 3, generator=g, device=dev)
    sc = torch.softmax(torch.randn(1024, 5, 243, generator=g, device=dev), 1)
    w = torch.tensor([1, 1, 2.5, 2.5, 1, 2.5, 2.5, 1, 1, 1, 1.5, 1.5, 4, 4, 1.5, 4, 4.0], device=dev)
    nfr = 1024 * 243
    med, best = timeit(lambda: ops.loss_terms(hyp, sc, y, w, False, 0.1, 2.0, 0.5))
    out["loss_fwd_B1024"] = {"ms": med, "gbs": 1252.0 * nfr / med / 1e6, "gframes_s": nfr / med / 1e6}
    hg = hyp.clone().requires_grad_()
    sg = sc.clone().requires_grad_()
    terms = ops.loss_terms(hg, sg, y, w, False, 0.1, 2.0, 0.5)[0]
    gt = torch.zeros(8, device=dev)
    gt[4] = 1.0
    med, best = timeit(lambda: torch.autograd.grad(terms, (hg, sg), gt, retain_graph=True))
    out["loss_bwd_B1024"] = {"ms": med, "gbs": 2292.0 * nfr / med / 1e6, "gframes_s": nfr / med / 1e6}
    med, best = timeit(lambda: ops.wta_fwd(hyp, y, w, False))
    out["wta_fwd_B1024"] = {"ms": med, "gbs": (204.0 * 6 + 12) * nfr / med / 1e6}
    for mode, nm, byt in ((0, "weighted_ave", 204.0 * 6 + 20), (1, "best_score", 204.0 * 2 + 28), (2, "oracle", 204.0 * 8 + 12)):
        med, best = timeit(lambda: ops.aggregate(hyp, sc, y, mode))
        out[f"aggregate_{nm}_B1024"] = {"ms": med, "gbs": byt * nfr / med / 1e6}
    pred = ops.aggregate(hyp, sc, None, 0)[0]
    med, best = timeit(lambda: ops.mpjpe(pred, y))
    out["mpjpe_B1024"] = {"ms": med, "gbs": 408.0 * nfr / med / 1e6}
    # the same rows on the host cores with the CPU oracle (reference algorithm), bounded samples
    if "--cpu" in sys.argv:
        import time
        from oracle import manipose_oracle as O
        torch.set_num_threads(os.cpu_count() or 1)
        cpu = {}
        nb = 103   # 103 clips x 5 x 243 = 125,145 poses (1/8 of config 2)
        rc, bc = rot[:nb * k * t].cpu(), bones[:nb].cpu().unsqueeze(-1)
        t0 = time.perf_counter(); O.pose_decoder(rc, bc, torch.zeros(nb * k * t, 3)); dt = time.perf_counter() - t0
        cpu["decoder_fwd"] = {"s": dt, "mposes_s": nb * k * t / dt / 1e6, "sample": f"{nb * k * t} poses"}
        hc, scc, yc = hyp[:128].cpu(), sc[:128].cpu().unsqueeze(-1), y[:128].cpu()
        t0 = time.perf_counter(); O.training_loss(hc, scc, yc); dt = time.perf_counter() - t0
        cpu["loss_fwd"] = {"s": dt, "mframes_s": 128 * 243 / dt / 1e6, "sample": "128 clips x 243 frames"}
        hcg = hc.clone().requires_grad_()
        t0 = time.perf_counter(); O.training_loss(hcg, scc, yc)[0].backward(); dt = time.perf_counter() - t0
        cpu["loss_fwd_bwd"] = {"s": dt, "mframes_s": 128 * 243 / dt / 1e6, "sample": "128 clips x 243 frames"}
        out["cpu_oracle"] = {"threads": torch.get_num_threads(), **cpu}
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    print(json.dumps({"clips": clips, "tokens": m, "peaks": {k: peaks.get(k) for k in ("hbm_gbs", "bf16_tflops", "bf16_tflops_sustained")}, "kernels": out}, indent=1))


if __name__ == "__main__":
    main()

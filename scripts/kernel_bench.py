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
    y = 0.3 * torch.randn(1024, 243, 17, 3, generator=g, device=dev)
    hyp = y[:, None] + 0.1 * torch.randn(1024, 5, 243, 17, 3, generator=g, device=dev)
    sc = torch.softmax(torch.randn(1024, 5, 243, generator=g, device=dev), 1)
    w = torch.tensor([1, 1, 2.5, 2.5, 1, 2.5, 2.5, 1, 1, 1, 1.5, 1.5, 4, 4, 1.5, 4, 4.0], device=dev)
    med, best = timeit(lambda: ops.loss_terms(hyp, sc, y, w, False, 0.1, 2.0, 0.5))
    out["loss_fwd_B1024"] = {"ms": med, "gbs": 1252.0 * 1024 * 243 / med / 1e6}
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    print(json.dumps({"clips": clips, "tokens": m, "peaks": {k: peaks.get(k) for k in ("hbm_gbs", "bf16_tflops", "bf16_tflops_sustained")}, "kernels": out}, indent=1))


if __name__ == "__main__":
    main()

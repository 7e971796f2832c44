"""Development aid: time the non-residual Linear kernels of the C = 512 backbone (qkv, fc1 + GELU) alone at a micro-batch size.
Usage: python scripts/gemm_ab.py [clips] [dtype]      The library reads MANIPOSE_PAIR_AS once per process (0: streaming A, default:
A-stationary), so an A/B comparison is two processes."""
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from manipose_b200 import ops  # noqa: E402

dev = torch.device("cuda")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, iters=8):
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
    clips = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    td = torch.float16 if (len(sys.argv) > 2 and sys.argv[2] == "fp16") else torch.bfloat16
    m = clips * 243 * 17
    out = {"pair_as": os.environ.get("MANIPOSE_PAIR_AS", "default"), "tokens": m}
    g = torch.Generator(device=dev).manual_seed(0)
    for name, n, k, epi in (("qkv", 1536, 512, 0), ("fc1_gelu", 1024, 512, 1)):
        a = torch.randn(m, k, generator=g, device=dev).to(td)
        w = (torch.randn(n, k, generator=g, device=dev) / math.sqrt(k)).to(td)
        b = torch.randn(n, generator=g, device=dev)
        y = torch.empty(m, n, dtype=td, device=dev)
        med, best = timeit(lambda: ops.linear(a, w, b, y, epi))
        ref = a[:4096].float() @ w.float().t() + b
        if epi == 1:
            ref = torch.nn.functional.gelu(ref)
        err = float((y[:4096].float() - ref).abs().max())
        tail = a[-300:].float() @ w.float().t() + b
        if epi == 1:
            tail = torch.nn.functional.gelu(tail)
        err_tail = float((y[-300:].float() - tail).abs().max())
        t_ref, _ = timeit(lambda: torch.matmul(a, w.t()))
        out[name] = {"us": med * 1e3, "best_us": best * 1e3, "tflops": 2.0 * m * n * k / med / 1e9, "max_err_head": err, "max_err_tail": err_tail,
                     "cublas_us": t_ref * 1e3, "cublas_tflops": 2.0 * m * n * k / t_ref / 1e9}
        del a, y
    print(json.dumps(out))


if __name__ == "__main__":
    main()

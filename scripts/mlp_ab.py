"""Development aid: the MLP branch of a C = 512 block at a micro-batch size — fused (mp_mlp_ln) against the two launches it replaces
(mp_linear with GELU + mp_linear_ln).  Usage: python scripts/mlp_ab.py [clips]"""
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
    return ts[len(ts) // 2]


def main():
    clips = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    m, c, hid = clips * 243 * 17, 512, 1024
    g = torch.Generator(device=dev).manual_seed(0)
    td = torch.bfloat16
    h = torch.randn(m, c, generator=g, device=dev).to(td)
    w1 = (torch.randn(hid, c, generator=g, device=dev) / math.sqrt(c)).to(td)
    w2 = (torch.randn(c, hid, generator=g, device=dev) / math.sqrt(hid)).to(td)
    b1, b2 = torch.randn(hid, generator=g, device=dev), torch.randn(c, generator=g, device=dev)
    x = torch.randn(m, c, generator=g, device=dev)
    pp = [torch.randn(c, generator=g, device=dev) for _ in range(4)]
    hidden = torch.empty(m, hid, dtype=td, device=dev)
    ho = torch.empty(m, c, dtype=td, device=dev)
    kw = dict(post=(pp[0], pp[1]), ln=(pp[2], pp[3]))

    def two():
        ops.linear(h, w1, b1, hidden, 1)
        ops.linear_ln(hidden, w2, b2, x, x, ho, **kw)

    def fused():
        ops.mlp_ln(h, w1, b1, w2, b2, x, x, ho, **kw)

    t2, t1 = timeit(two), timeit(fused)
    flops = 4.0 * m * hid * c
    out = {"tokens": m, "two_launches_us": t2 * 1e3, "fused_us": t1 * 1e3, "fused_tflops": flops / t1 / 1e9, "two_tflops": flops / t2 / 1e9,
           "fused_hbm_gbs": (m * c * (2 + 4 + 4 + 2)) / t1 / 1e6}
    print(json.dumps(out))


if __name__ == "__main__":
    main()

"""Development aid: time the fused residual + LayerNorm GEMM (mp_linear_ln) alone.  Usage: python scripts/ln_experiment.py [clips ...]
The library reads MANIPOSE_LN_CFG once per process (1: split accumulation, 2: single accumulation, unset: by K)."""
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
    out = {"cfg": os.environ.get("MANIPOSE_LN_CFG", "default")}
    g = torch.Generator(device=dev).manual_seed(0)
    for clips in [int(a) for a in sys.argv[1:]] or [32]:
        m = clips * 243 * 17
        for name, k in (("proj+ln2", 512), ("fc2+post+ln1", 1024)):
            a = torch.randn(m, k, generator=g, device=dev).bfloat16()
            w = (torch.randn(512, k, generator=g, device=dev) / math.sqrt(k)).bfloat16()
            b = torch.randn(512, generator=g, device=dev)
            xx = torch.randn(m, 512, generator=g, device=dev)
            hh = torch.empty(m, 512, dtype=torch.bfloat16, device=dev)
            pp = [torch.randn(512, generator=g, device=dev) for _ in range(4)]
            post = (pp[0], pp[1]) if k == 1024 else None
            med = timeit(lambda: ops.linear_ln(a, w, b, xx, xx, hh, post=post, ln=(pp[2], pp[3])))
            byt = (m * k + 512 * k) * 2 + m * 512 * (8 + 2)
            out[f"{name}_M{m}"] = {"us": med * 1e3, "tflops": 2.0 * m * 512 * k / med / 1e9, "gbs": byt / med / 1e6}
            del a, xx, hh
    print(json.dumps(out))


if __name__ == "__main__":
    main()

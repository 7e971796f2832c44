"""Development aid: time temporal / spatial attention alone at a micro-batch size.  Usage: python scripts/attn_ab.py [clips] [frames]
The library reads MANIPOSE_ATTN_TC2 once per process (set: the shared-memory-P kernel; unset: P in TMEM), so an A/B is two processes."""
import json
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
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 243
    J = 17
    m = clips * T * J
    out = {"tc2": os.environ.get("MANIPOSE_ATTN_TC2") is not None, "tokens": m, "frames": T}
    g = torch.Generator(device=dev).manual_seed(0)
    for td, name in ((torch.bfloat16, "bf16"), (torch.float16, "fp16")):
        qkv = torch.randn(m, 1536, generator=g, device=dev).to(td)
        o = torch.empty(m, 512, dtype=td, device=dev)
        for mode, nm in ((1, "temporal"), (0, "spatial")):
            med = timeit(lambda: ops.attention(qkv, o, clips, T, J, 512, 8, mode))
            out[f"{nm}_{name}"] = {"us": med * 1e3, "gbs": m * 2048 * 2 / med / 1e6, "tflops": 4.0 * m * (T if mode == 1 else J) * 512 / med / 1e9}
    print(json.dumps(out))


if __name__ == "__main__":
    main()

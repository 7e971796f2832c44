"""Development probe: time the qkv / fc1 GEMM shapes (cold L2) — run with MANIPOSE_SINGLE_CTA / MANIPOSE_DBG variants."""
import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from manipose_b200 import ops
dev = torch.device("cuda")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
m = 8 * 243 * 17
g = torch.Generator(device=dev).manual_seed(0)
res = []
for name, n, k, epi in (("qkv", 1536, 512, 0), ("fc1", 1024, 512, 1)):
    a = torch.randn(m, k, generator=g, device=dev).bfloat16()
    w = (torch.randn(n, k, generator=g, device=dev) / math.sqrt(k)).bfloat16()
    b = torch.randn(n, generator=g, device=dev)
    y = torch.empty(m, n, dtype=torch.bfloat16, device=dev)
    for _ in range(3):
        ops.linear(a, w, b, y, epi)
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.linear(a, w, b, y, epi); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    res.append(f"{name}={ts[5]*1000:.1f}us")
print(os.environ.get("MANIPOSE_SINGLE_CTA", "pair"), "dbg", os.environ.get("MANIPOSE_DBG", "0"), " ".join(res))

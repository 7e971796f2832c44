"""Development probe: fused residual GEMM + LayerNorm kernels at a 32-clip micro-batch."""
import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from manipose_b200 import ops
dev = torch.device("cuda")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
clips = int(sys.argv[1]) if len(sys.argv) > 1 else 32
m = clips * 243 * 17
g = torch.Generator(device=dev).manual_seed(0)
for name, k, post in (("proj+ln2", 512, False), ("fc2+post+ln1", 1024, True)):
    a = torch.randn(m, k, generator=g, device=dev).bfloat16()
    w = (torch.randn(512, k, generator=g, device=dev) / math.sqrt(k)).bfloat16()
    b = torch.randn(512, generator=g, device=dev)
    x = torch.randn(m, 512, generator=g, device=dev)
    h = torch.empty(m, 512, dtype=torch.bfloat16, device=dev)
    p = [torch.randn(512, generator=g, device=dev) for _ in range(4)]
    fn = lambda: ops.linear_ln(a, w, b, x, x, h, post=(p[0], p[1]) if post else None, ln=(p[2], p[3]))
    for _ in range(2):
        fn()
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[2]
    byt = (m * k + 512 * k) * 2 + m * 512 * 10
    print(name, f"M={m} {t*1000:.1f}us  {byt/t/1e6:.0f} GB/s  {2.0*m*512*k/t/1e9:.0f} TF/s")

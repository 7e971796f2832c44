"""Development probe: temporal attention cold vs warm L2, and vs clip count."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from manipose_b200 import ops
dev = torch.device("cuda")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
T, J = 243, 17
for clips in (2, 8, 32):
    m = clips * T * J
    g = torch.Generator(device=dev).manual_seed(0)
    qkv = torch.randn(m, 1536, generator=g, device=dev).bfloat16()
    o = torch.empty(m, 512, dtype=torch.bfloat16, device=dev)
    for warm in (False, True):
        for _ in range(2):
            ops.attention(qkv, o, clips, T, J, 512, 8, 1)
        ts = []
        for _ in range(5):
            if not warm:
                flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.attention(qkv, o, clips, T, J, 512, 8, 1); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = sorted(ts)[2]
        print(f"clips={clips} {'warm' if warm else 'cold'} L2: {t*1000:.1f}us  per head-SM {t*1000*148/(clips*J*8):.2f}us  {m*2048*2/t/1e6:.0f} GB/s")

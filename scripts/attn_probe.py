"""Development probe: temporal / spatial attention at the 8-clip micro-batch shape (cold L2)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from manipose_b200 import ops
dev = torch.device("cuda")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
clips, T, J = 8, 243, 17
m = clips * T * J
g = torch.Generator(device=dev).manual_seed(0)
qkv = torch.randn(m, 1536, generator=g, device=dev).bfloat16()
o = torch.empty(m, 512, dtype=torch.bfloat16, device=dev)
for mode, name in ((1, "temporal"), (0, "spatial")):
    for _ in range(2):
        ops.attention(qkv, o, clips, T, J, 512, 8, mode)
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.attention(qkv, o, clips, T, J, 512, 8, mode); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(name, f"{sorted(ts)[2]*1000:.1f}us")

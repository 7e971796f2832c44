"""Development probe: temporal attention (head_dim 64) at ~530 k tokens for a given clip length.  Usage: python scripts/attn_probe.py T [T ...]
MANIPOSE_ATTN_TRACKS_OFF=1 forces the T = 243 kernel for short clips (A/B)."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from manipose_b200 import ops
dev = torch.device("cuda")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
out = {"tracks_off": os.environ.get("MANIPOSE_ATTN_TRACKS_OFF", "0")}
for T in [int(a) for a in sys.argv[1:]] or [27]:
    clips = max(1, 530000 // (T * 17))
    m = clips * T * 17
    qkv = torch.randn(m, 1536, device=dev).bfloat16()
    o = torch.empty(m, 512, dtype=torch.bfloat16, device=dev)
    ts = []
    for it in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.attention(qkv, o, clips, T, 17, 512, 8, 1)
        e1.record()
        torch.cuda.synchronize()
        if it >= 2:
            ts.append(e0.elapsed_time(e1))
    ts.sort()
    out[f"T{T}"] = {"tokens": m, "us": 1e3 * ts[len(ts) // 2], "gbs": m * 2048 * 2 / ts[len(ts) // 2] / 1e6}
print(json.dumps(out))

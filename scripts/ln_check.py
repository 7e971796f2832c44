"""Development aid: one mp_linear_ln call at a given M / K / mode against fp32 torch (one process per case: a device fault is sticky)."""
import math
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from manipose_b200 import ops  # noqa: E402

m, k, mode = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
gen = torch.Generator(device="cuda").manual_seed(m + k)
a = torch.randn(m, k, generator=gen, device="cuda").bfloat16()
w = (torch.randn(512, k, generator=gen, device="cuda") / math.sqrt(k)).bfloat16()
bias = torch.randn(512, generator=gen, device="cuda")
resid = torch.randn(m, 512, generator=gen, device="cuda") * 1.5 + 0.3
pg, pb, lg, lb = (torch.randn(512, generator=gen, device="cuda") for _ in range(4))
x_ref = resid + a.float() @ w.float().t() + bias
post = (pg, pb) if mode.startswith("post") else None
ln = (lg, lb) if mode != "plain" else None
if post is not None:
    x_ref = F.layer_norm(x_ref, (512,), pg, pb, 1e-6)
h_ref = F.layer_norm(x_ref, (512,), lg, lb, 1e-6) if ln is not None else None
x = resid.clone()
h = torch.full((m, 512), float("nan"), dtype=torch.bfloat16, device="cuda") if ln is not None else None
try:
    ops.linear_ln(a, w, bias, x, x, h, post=post, post_eps=1e-6, ln=ln, ln_eps=1e-6)
    torch.cuda.synchronize()
    ex = float((x - x_ref).abs().max())
    eh = float((h.float() - h_ref).abs().max()) if ln is not None else 0.0
    print(f"M={m} K={k} {mode}: max|dx|={ex:.2e} max|dh|={eh:.2e}", "OK" if ex < 1e-3 and eh < 0.1 else "MISMATCH")
except Exception as e:  # noqa: BLE001
    print(f"M={m} K={k} {mode}: FAULT {str(e).splitlines()[0][:80]}")

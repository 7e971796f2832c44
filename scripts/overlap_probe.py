"""Development aid: do two micro-batches in flight on HALF the SMs each beat one micro-batch at a time on all of them?

The forward alternates tensor-bound launches (qkv, fc1 + GELU) and HBM-bound ones (proj / fc2 + residual + LayerNorms, attention).
With `mp_set_sm_limit(74)` every persistent grid takes 37 CTA pairs, so two streams can run side by side, a tensor-bound launch
of one lane next to an HBM-bound launch of the other.  This script runs the kernels of one transformer block pair (spatial +
temporal) for two micro-batches (a) one after the other at full width and (b) on two streams at half width, the second lane
started half a block later, for a few seconds each (board power settles), A B A B.

Usage: python scripts/overlap_probe.py [clips] [seconds] [gemm_only]"""
import json
import math
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from manipose_b200 import _lib as L, ops  # noqa: E402

dev = torch.device("cuda")


class Lane:
    def __init__(self, clips, seed):
        g = torch.Generator(device=dev).manual_seed(seed)
        self.clips = clips
        m = self.m = clips * 243 * 17
        r = lambda *s: torch.randn(*s, generator=g, device=dev)
        self.x = r(m, 512)
        self.h = r(m, 512).bfloat16()
        self.wide = torch.empty(m * 1536, dtype=torch.bfloat16, device=dev)
        self.qkv = self.wide.view(m, 1536)
        self.hid = self.wide[:m * 1024].view(m, 1024)
        self.w_qkv = (r(1536, 512) / math.sqrt(512)).bfloat16()
        self.w_proj = (r(512, 512) / math.sqrt(512)).bfloat16()
        self.w_fc1 = (r(1024, 512) / math.sqrt(512)).bfloat16()
        self.w_fc2 = (r(512, 1024) / math.sqrt(1024)).bfloat16()
        self.b_qkv, self.b_proj, self.b_fc1, self.b_fc2 = r(1536), r(512), r(1024), r(512)
        self.g = [1.0 + 0.1 * r(512) for _ in range(3)]
        self.b = [0.1 * r(512) for _ in range(3)]

    def kernels(self, gemm_only):
        ks = []
        for mode in (L.MP_ATTN_SPATIAL, L.MP_ATTN_TEMPORAL):
            ks.append(lambda: ops.linear(self.h, self.w_qkv, self.b_qkv, self.qkv, L.MP_EPI_BIAS))
            if not gemm_only:
                ks.append(lambda mode=mode: ops.attention(self.qkv, self.h, self.clips, 243, 17, 512, 8, mode))
            ks.append(lambda: ops.linear_ln(self.h, self.w_proj, self.b_proj, self.x, self.x, self.h, ln=(self.g[0], self.b[0])))
            ks.append(lambda: ops.linear(self.h, self.w_fc1, self.b_fc1, self.hid, L.MP_EPI_GELU))
            ks.append(lambda: ops.linear_ln(self.hid, self.w_fc2, self.b_fc2, self.x, self.x, self.h, post=(self.g[1], self.b[1]),
                                            ln=(self.g[2], self.b[2])))
        return ks


def main():
    clips = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 3.0
    gemm_only = len(sys.argv) > 3 and sys.argv[3] == "gemm_only"
    lib = L.load()
    a, b = Lane(clips, 1), Lane(clips, 2)
    ka, kb = a.kernels(gemm_only), b.kernels(gemm_only)
    n = len(ka)
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()

    def sequential(reps):
        lib.mp_set_sm_limit(0)
        for _ in range(reps):
            for k in ka:
                k()
            for k in kb:
                k()

    def dual(reps, limit, shift):
        lib.mp_set_sm_limit(limit)
        main_s = torch.cuda.current_stream()
        sa.wait_stream(main_s)
        sb.wait_stream(main_s)
        for _ in range(reps):
            for i in range(n):
                with torch.cuda.stream(sa):
                    ka[i]()
                with torch.cuda.stream(sb):
                    kb[(i + shift) % n]()
        main_s.wait_stream(sa)
        main_s.wait_stream(sb)
        lib.mp_set_sm_limit(0)

    def timed(fn, reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(reps)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    out = {"clips": clips, "gemm_only": gemm_only, "launches_per_lane_rep": n}
    t1 = timed(sequential, 3)
    reps = max(3, int(seconds * 1e3 / t1))
    variants = [("sequential_full", sequential)]
    for limit, shift in ((74, n // 4), (74, 1), (74, 0), (148, n // 4), (100, n // 4)):
        variants.append((f"dual_limit{limit}_shift{shift}", lambda r, limit=limit, shift=shift: dual(r, limit, shift)))
    for rnd in range(2):
        for name, fn in variants:
            t0 = time.time()
            ms = timed(fn, reps)
            out.setdefault(name, []).append(round(ms, 3))
            print(name, round(ms, 3), "ms per rep (two micro-batch block pairs)", round(time.time() - t0, 1), "s", flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()

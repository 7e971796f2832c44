"""BASELINE config 5: K in {1, 5, 10} hypotheses x T in {27, 81, 243} frames, inference, ~250k frames per point (1 B200).
Prints one JSON object: frames/s and the fraction of the measured sustained bf16 peak at F(T, K) flops per frame."""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import manipose_b200 as mb
from bench import flops_per_frame, measured_peaks

dtype = sys.argv[1] if len(sys.argv) > 1 else "bf16"
dev = torch.device("cuda")
peaks = measured_peaks()
out = {"dtype": dtype, "points": []}
for T in (27, 81, 243):
    for K in (1, 5, 10):
        B = max(1, 250_000 // T)
        torch.manual_seed(42)
        m = mb.RMCLManifoldMixSTE(mb.h36m17_skeleton(), num_frame=T, n_hyp=K, drop_path_rate=0.1).to(dev).eval().set_compute_dtype(dtype)
        x = 0.3 * torch.randn(B, T, 17, 2, generator=torch.Generator().manual_seed(1234)).to(dev)
        with torch.no_grad():
            for _ in range(3):
                m(x)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                m(x)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        fps = B * T / (ms / 1000.0)
        tf = flops_per_frame(T, K) * fps / 1e12
        out["points"].append({"T": T, "K": K, "clips": B, "frames_per_s": fps, "ms": ms, "gflop_per_frame": flops_per_frame(T, K) / 1e9,
                              "tflops": tf, "frac_of_sustained_bf16_peak": tf / peaks["bf16_sustained"]})
        del m
        torch.cuda.empty_cache()
print(json.dumps(out, indent=1))

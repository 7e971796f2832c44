"""Headline benchmark: T=243 H36M-shape lifted frames/sec (BASELINE.json) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--clips B] [--impl reference]

A step is one pass of the lifting hot path (RMCLManifoldMixSTE.forward: MixSTE backbone -> K=5 hypothesis heads ->
manifold decoder + hypothesis softmax) over one batch of B synthetic clips of T=243 frames, random-init weights
(hpe/conf/config.yaml defaults, seed 42).  BASELINE config 3 (B=1024 clips IN TOTAL, bf16 backbone, fp32 decoder).  Multi-GPU:
one process per GPU (torchrun), the 1024 clips sharded 1024/N per GPU ("scaling": "strong", as BASELINE config 3 /
SURVEY.md §8d state it), NO collective on the data path; the weak-scaled figure (1024 clips per GPU) rides along as `weak`.

The one JSON line carries: value (device-resident inputs), e2e (pinned-host inputs, H2D + forward + D2H of poses/scores
inside the timed region), roofline for the dominant kernel family (sampled with CUDA events inside the timed region),
cpu_baseline (the reference algorithm — oracle/manipose_oracle.py, a PyTorch-CPU restatement pinned to the reference —
timed on the host cores on a bounded sample), and, measured in the same run:
  parity   |dMPJPE| (mm) and score-arg-max agreement of the TIMED dtype against the fp32 oracle on BASELINE config 1 inputs
  fp16     the same workload with fp16 tensor-core operands (value, e2e, parity): the dtype that meets the 0.05 mm gate
  train    BASELINE config 4 (T=27 training step, captured graph, data parallel with the NCCL gradient all-reduce)
  decoder  BASELINE config 2 (manifold decoder alone on 1,001,160 poses; N = 1 only)
--impl reference times the CPU path alone.  --config 4 / --config 2 print the training / decoder line alone.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

T, J, K = 243, 17, 5
METRIC = "lifted_frames_per_sec_T243_H36M"
UNIT = "frames/s"


def flops_per_frame(t=T, k=K):
    """Closed form of SURVEY.md §8(d) / BASELINE.md §3 (matmul flops of one forward, per lifted frame)."""
    return (17 * 8 * 2 * (16 * 512 ** 2) + 17 * 8 * (4 * 17 * 512 + 4 * t * 512)
            + 16 * 2 * 2 * (16 * 128 ** 2) + 16 * 2 * (4 * 16 * 128 + 4 * t * 128)
            + (17 * 2 * 2 * 512 + 2 * 34 * 2048) + k * (17 * 2 * 512 * 7 + 2 * 17) + 16 * 2 * 128)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_burst": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"], "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "src": "fallback"}


# dram__bytes_read.sum + dram__bytes_write.sum per launch of the residual Linear + LayerNorm kernel (pair_linear_ln64_kernel, M = 528,768 rows
# = the default 128-clip micro-batch) from the committed `ncu --set full` capture profiles/r2_ln64_tc3_full.csv, averaged over its two uses
# per block like `achieved` (proj + norm2: 3.191 GB against 3.249 GB algorithmic; fc2 + post-norm + norm1: 3.734 GB against 3.791 GB)
PROFILED_TRAFFIC_LN = {"bytes_per_launch": 3462.5e6, "algorithmic_bytes": 3520.2e6,
                       "launches": {"proj + norm2 (K=512)": {"dram_bytes": 3191.0e6, "algorithmic_bytes": 3249.3e6, "us_under_ncu": 537.5},
                                    "fc2 + post-norm + norm1 (K=1024)": {"dram_bytes": 3734.2e6, "algorithmic_bytes": 3790.7e6, "us_under_ncu": 746.7}},
                       "launch": "M=528768 (128 clips)", "source": "profiles/r2_ln64_tc3_full.csv, profiles/r2_launch_and_kernel_summary.md"}


def roofline_object(dom, rl, peaks, step_tflops, traffic):
    """Roofline of the dominant kernel family (by measured share of the step) + the other family and the whole step for context.
    linear_ln (residual GEMM with fused LayerNorms, 5-7 B per output element) is HBM-bound; linear (qkv, fc1) is tensor-bound."""
    if dom is None:
        return None
    out = {}
    if dom == "linear_ln":
        a = rl[dom]["gbs"]
        out = {"bound": "hbm", "kernel": "pair_linear_ln64_kernel (tcgen05 cta_group::2 residual GEMM + fused LayerNorm epilogue on 64-row tiles, "
                                        "two TMEM accumulators: proj / fc2)",
               "achieved": a, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": a / peaks["hbm_gbs"], "traffic": traffic.get(dom),
               "peak_source": f"{peaks['src']} (STREAM-style copy)"}
    else:
        a = rl[dom]["tflops"]
        out = {"bound": "tensor", "kernel": "pair_linear_kernel (tcgen05 cta_group::2: qkv, fc1 + GELU)", "achieved": a,
               "peak": peaks["bf16_sustained"], "unit": "TFLOP/s", "frac": a / peaks["bf16_sustained"], "traffic": traffic.get(dom),
               "peak_source": f"{peaks['src']} (sustained cuBLAS bf16; burst {peaks['bf16_burst']})"}
    out["share_of_step"] = rl[dom]["share_of_step"]
    out["avg_launch_us"] = rl[dom]["avg_us"]
    out["families"] = rl
    out["whole_step"] = {"achieved": step_tflops, "unit": "TFLOP/s", "frac": step_tflops / peaks["bf16_sustained"],
                         "flops": "algorithmic matmul flops of the whole forward / step time, vs the sustained bf16 peak"}
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = max(mx, float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------- CPU reference arm
def _oracle_setup(seed=42):
    import torch
    from oracle import manipose_oracle as O
    sd = O.make_state_dict(num_frame=T, n_hyp=K, seed=seed)
    return O, sd


def cpu_forward_rate(clips, reps, warm=True, budget_s=25.0):
    """Reference algorithm on the host cores: frames/s of rmcl_forward on `clips` synthetic clips (fp32, all threads)."""
    import torch
    O, sd = _oracle_setup()
    torch.set_num_threads(os.cpu_count() or 1)
    x = 0.3 * torch.randn(clips, T, J, 2, generator=torch.Generator().manual_seed(1234))
    with torch.no_grad():
        if warm:
            O.rmcl_forward(x[:1], sd)
        best, t_start, done = None, time.perf_counter(), 0
        for _ in range(reps):
            t0 = time.perf_counter()
            O.rmcl_forward(x, sd)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
            done += 1
            if time.perf_counter() - t_start > budget_s:
                break
    return clips * T / best, torch.get_num_threads(), done


def cpu_train_rate(t, clips=2, budget_s=25.0):
    """Reference algorithm on the host cores for the training step: forward + default objective + backward (torch autograd through
    the fp32 oracle; no optimizer step) on a bounded sample, frames/s."""
    import torch
    from oracle import manipose_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    sd = {k: v.requires_grad_() for k, v in O.make_state_dict(num_frame=t, n_hyp=K, seed=42).items()}
    g = torch.Generator().manual_seed(1234)
    x = 0.3 * torch.randn(clips, t, J, 2, generator=g)
    y = 0.3 * torch.randn(clips, t, J, 3, generator=g)
    best, t_start, done = None, time.perf_counter(), 0
    for _ in range(4):
        t0 = time.perf_counter()
        poses, scores = O.rmcl_forward(x, sd)
        loss, _ = O.training_loss(poses, scores, y)
        loss.backward()
        dt = time.perf_counter() - t0
        done += 1
        if done > 1:                      # first pass warms the allocator
            best = dt if best is None else min(best, dt)
        if time.perf_counter() - t_start > budget_s and best is not None:
            break
    return {"value": clips * t / best, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"best of {done - 1} passes of forward + objective + backward (no optimizer) over {clips} clips x {t} frames, fp32, torch CPU autograd"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    clips = 4
    import torch
    O, sd = _oracle_setup()
    torch.set_num_threads(os.cpu_count() or 1)
    x = 0.3 * torch.randn(clips, T, J, 2, generator=torch.Generator().manual_seed(1234))
    with torch.no_grad():
        for _ in range(args.warmup):
            O.rmcl_forward(x[:1], sd)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            O.rmcl_forward(x, sd)
        dt = time.perf_counter() - t0
    value = clips * T * args.steps / dt
    sample = f"{clips} clips x {T} frames per step (of the {args.clips}-clip workload), fp32, torch CPU, {torch.get_num_threads()} threads"
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    strong = args.scaling == "strong"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1000.0, "higher_is_better": True, "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"ManiPose H36M lifting forward (RMCLManifoldMixSTE, config.yaml defaults), T={T}, K={K}, J={J}, "
                                   f"{args.clips} clips " + (f"in total, sharded {max(1, args.clips // max(world, 1))} per GPU" if strong else "per GPU")
                                   + f" = BASELINE config 3; this CPU arm times a {clips}-clip SAMPLE of that workload per step on the host cores "
                                   f"of rank 0 (frames/s of the reference algorithm does not depend on the batch size)",
                       "reference_arm": "reference algorithm restated in oracle/manipose_oracle.py (pinned to the unmodified reference; "
                                        "the reference itself is Python and needs timm/mup/.cuda() shims, SURVEY.md §8c)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- GPU arm
def _dist_env():
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    return world, rank, local_rank, dev


def _make_timers(world, dev):
    import torch
    import torch.distributed as dist

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """K calls bracketed by barrier + synchronize on both sides, CUDA events on the launching stream, max over ranks."""
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    return sync_all, timed


def parity_config1(model, dtype, ref_cache):
    """BASELINE config 1 inputs (4 clips x T=243, x = 0.3 randn seed 1234) through the benchmarked model at `dtype` and through the
    fp32 CPU oracle on the same weights: |d MPJPE| in mm of the weighted-average aggregate (worst of three synthetic targets, the
    construction of tests/test_gpu_backbone.py::test_end_to_end_mpjpe_gate_config1) and agreement of the score arg-max."""
    import torch
    from oracle import manipose_oracle as O
    x = 0.3 * torch.randn(4, T, J, 2, generator=torch.Generator().manual_seed(1234))
    if "ref" not in ref_cache:
        sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
        torch.set_num_threads(os.cpu_count() or 1)
        with torch.no_grad():
            pr, sr = O.rmcl_forward(x, sd)
        ref_cache["ref"] = (O.aggregate(pr, sr, "weighted_ave"), sr)
    agg_ref, scores_ref = ref_cache["ref"]
    was = model.rotations_module.compute_dtype
    model.set_compute_dtype(dtype)
    with torch.no_grad():
        poses, scores = model(x.to(next(model.parameters()).device))
        agg = model.aggregate(poses, scores, "weighted_ave").cpu()
    model.set_compute_dtype(was)
    mm = lambda p, y: float((p - y).norm(dim=-1).mean() * 1000.0)
    worst = 0.0
    for seed in (5, 6, 7):
        y = 0.3 * torch.randn(4, T, J, 3, generator=torch.Generator().manual_seed(seed))
        worst = max(worst, abs(mm(agg, y) - mm(agg_ref, y)))
    gate = {"fp16": 0.05, "bf16": 0.25}[dtype]
    return {"dtype": dtype, "d_mpjpe_mm": worst, "gate_mm": gate, "gate_met": worst <= gate,
            "gate_source": "north_star 0.05 mm end-to-end gate" if dtype == "fp16" else "bf16 backbone tolerance, stated separately (DESIGN.md §3)",
            "score_argmax_agreement": float((scores.cpu().argmax(1) == scores_ref.argmax(1)).float().mean()),
            "against": "fp32 CPU oracle (oracle/manipose_oracle.py) on the same weights, BASELINE config 1 inputs (4 clips x 243 frames)"}


def measure_inference(model, B, dtype, args, world, rank, local_rank, dev, sample_kernels=True):
    """`B` clips per GPU through RMCLManifoldMixSTE.forward at `dtype`: device-resident and end-to-end timing (max over ranks)."""
    import torch
    from manipose_b200 import ops
    sync_all, timed = _make_timers(world, dev)
    model.set_compute_dtype(dtype)
    gen = torch.Generator().manual_seed(1234 + rank)
    x_host = (0.3 * torch.randn(B, T, J, 2, generator=gen)).pin_memory()
    x_dev = x_host.to(dev)
    out_host = (torch.empty((B, K, T, J, 3), dtype=torch.float32).pin_memory(), torch.empty((B, K, T, 1), dtype=torch.float32).pin_memory())

    def step_resident():
        with torch.no_grad():
            torch.cuda.nvtx.range_push("manipose.forward")
            out = model(x_dev)
            torch.cuda.nvtx.range_pop()
            return out

    def step_e2e():
        with torch.no_grad():
            xd = x_host.to(dev, non_blocking=True)
            poses, scores = model(xd)
            out_host[0].copy_(poses, non_blocking=True)
            out_host[1].copy_(scores, non_blocking=True)

    for _ in range(max(args.warmup, 3)):
        step_resident()
    # ---- device-resident throughput, with the GEMMs of one micro-batch in `sample_every` bracketed by CUDA events
    sampled = []
    orig_trunk = type(model.rotations_module).trunk
    counter = {"n": 0}
    sample_every = 4

    def trunk_sampled(self, x2d, n_clips, **kw):
        counter["n"] += 1
        ops.GEMM_TIMING = sampled if counter["n"] % sample_every == 1 % sample_every else None
        try:
            return orig_trunk(self, x2d, n_clips, **kw)
        finally:
            ops.GEMM_TIMING = None

    if sample_kernels:
        model.rotations_module.trunk = trunk_sampled.__get__(model.rotations_module)
    launches0 = ops.LAUNCHES
    steps = args.steps
    with ClockSampler(local_rank) as clocks:
        ms = timed(step_resident, steps)
        if ms < 1500.0:                 # short regions (strong scaling at N = 8): repeat so that nvidia-smi sees the load
            extra = int(1500.0 / max(ms / steps, 1e-3)) + 1
            ms += timed(step_resident, extra)
            steps += extra
    launches = ops.LAUNCHES - launches0
    if sample_kernels:
        del model.rotations_module.trunk
    frames = B * T * world
    res = {"value": frames * steps / (ms / 1000.0), "ms_per_step": ms / steps, "steps": steps, "frames_per_step": frames,
           "gpu_launches": launches, "clocks": clocks.summary(), "clips_per_gpu": B}
    fam = {}
    for a, b, fl, by, tag in sampled:
        f = fam.setdefault(tag, {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "n": 0})
        f["ms"] += a.elapsed_time(b)
        f["flops"] += fl
        f["bytes"] += by
        f["n"] += 1
    n_mb = max(1, counter["n"])
    sampled_mb = max(1, (n_mb + sample_every - 1) // sample_every)
    rl = {}
    for tag, f in fam.items():
        t_s = f["ms"] / 1000.0
        rl[tag] = {"launches_sampled": f["n"], "avg_us": f["ms"] * 1000.0 / f["n"], "tflops": f["flops"] / t_s / 1e12,
                   "gbs": f["bytes"] / t_s / 1e9, "share_of_step": f["ms"] * (n_mb / sampled_mb) / ms}
    res["families"] = rl
    res["step_tflops"] = flops_per_frame() * B * T * steps / (ms / 1000.0) / 1e12          # per GPU (max-over-ranks time)
    # ---- end to end through the public API with host buffers
    for _ in range(2):
        step_e2e()
    e2e_steps = max(args.steps, 3)
    ms_e2e = timed(step_e2e, e2e_steps)
    res["e2e"] = {"value": frames * e2e_steps / (ms_e2e / 1000.0), "unit": UNIT, "h2d_bytes_per_step": x_host.numel() * 4,
                  "d2h_bytes_per_step": (out_host[0].numel() + out_host[1].numel()) * 4, "ms_per_step": ms_e2e / e2e_steps}
    return res


def run_gpu(args):
    import torch
    import torch.distributed as dist

    world, rank, local_rank, dev = _dist_env()
    from manipose_b200 import _build
    if rank == 0 and not os.path.exists(os.path.join(ROOT, "manipose_b200", "libmanipose_sm100.so")):
        _build.build(verbose=False)
    if world > 1:
        dist.barrier()
    import manipose_b200 as mb

    torch.manual_seed(42)
    model = mb.RMCLManifoldMixSTE(mb.h36m17_skeleton(), num_frame=T, n_hyp=K, drop_path_rate=0.1)
    model = model.to(dev).eval().set_compute_dtype(args.dtype)
    if args.micro_batch_clips:
        model.rotations_module.micro_batch_tokens = args.micro_batch_clips * T * J
        model.segments_module.micro_batch_tokens = args.micro_batch_clips * T * J
    total = args.clips
    strong = args.scaling == "strong"
    B = max(1, total // world) if strong else total
    peaks = measured_peaks()
    main = measure_inference(model, B, args.dtype, args, world, rank, local_rank, dev)
    extras = {}
    if not args.headline_only:
        other = "fp16" if args.dtype == "bf16" else "bf16"
        extras["other"] = measure_inference(model, B, other, args, world, rank, local_rank, dev, sample_kernels=False)
        if world > 1 and strong:
            extras["weak"] = measure_inference(model, total, args.dtype, args, world, rank, local_rank, dev, sample_kernels=False)
    parity = {}
    if rank == 0 and not args.headline_only:
        cache = {}
        parity[args.dtype] = parity_config1(model, args.dtype, cache)
        other = "fp16" if args.dtype == "bf16" else "bf16"
        parity[other] = parity_config1(model, other, cache)
    model.set_compute_dtype(args.dtype)
    micro = model.rotations_module.clips_per_micro_batch()
    del model
    torch.cuda.empty_cache()
    train = None
    if not args.headline_only:
        train = measure_train(args, world, rank, local_rank, dev, with_cpu=False)
    decoder = None
    if rank == 0 and world == 1 and not args.headline_only:
        decoder = measure_decoder(args, with_cpu=False)

    if rank == 0:
        rl = main["families"]
        dom = max(rl, key=lambda k: rl[k]["share_of_step"]) if rl else None
        traffic = {"linear_ln": PROFILED_TRAFFIC_LN, "linear": None}
        # the CPU oracle is timed beside the GPU number at N = 1 only (the other ranks would sit in the barrier meanwhile)
        skip_cpu = args.skip_cpu_baseline or world > 1
        cpu_value, cpu_threads, cpu_reps = (0.0, 0, 0) if skip_cpu else cpu_forward_rate(clips=4, reps=3)
        frames = main["frames_per_step"]
        line = {
            "metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": world, "steps": main["steps"], "warmup": max(args.warmup, 3),
            "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": f"ManiPose H36M lifting forward (RMCLManifoldMixSTE, config.yaml defaults), T={T}, K={K}, J={J}, "
                                   f"{total} clips " + (f"in total, sharded {B} per GPU" if strong else "per GPU") + " = BASELINE config 3; "
                                   f"{args.dtype} tensor-core operands (fp32 accumulate, fp32 residual stream), fp32 decoder; the CPU arm "
                                   f"(cpu_baseline / --impl reference) is timed on a 4-clip SAMPLE of this workload",
                       "clips_per_gpu": B, "frames_per_step": frames, "parallelism": f"clip-sharded x{world}, no collective",
                       "micro_batch_clips": micro,
                       "l2": "inputs+activations per step (>1 GB) exceed the 126 MB L2; no explicit flush",
                       "gflop_per_frame": flops_per_frame() / 1e9},
            "e2e": main["e2e"],
            "gpu_launches": main["gpu_launches"],
            "clocks": main["clocks"],
            "roofline": roofline_object(dom, rl, peaks, main["step_tflops"], traffic),
            "cpu_baseline": None if skip_cpu else {
                "value": cpu_value, "unit": UNIT, "cores": cpu_threads, "kind": "port",
                "sample": f"best of {cpu_reps} passes over 4 clips x {T} frames (972 frames) of the same workload, fp32, torch CPU"},
        }
        if parity:
            line["parity"] = parity[args.dtype]
        if "other" in extras:
            o = extras["other"]
            other = "fp16" if args.dtype == "bf16" else "bf16"
            line[other] = {"value": o["value"], "unit": UNIT, "ms_per_step": o["ms_per_step"], "e2e": o["e2e"], "clocks": o["clocks"],
                           "whole_step": {"achieved": o["step_tflops"], "unit": "TFLOP/s", "frac": o["step_tflops"] / peaks["bf16_sustained"]},
                           "parity": parity.get(other)}
        if "weak" in extras:
            wk = extras["weak"]
            line["weak"] = {"value": wk["value"], "unit": UNIT, "clips_per_gpu": wk["clips_per_gpu"], "ms_per_step": wk["ms_per_step"],
                            "e2e": wk["e2e"], "clocks": wk["clocks"], "scaling": "weak"}
        if train is not None:
            line["train"] = train
        if decoder is not None:
            line["decoder"] = decoder
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------- training step (BASELINE config 4)
def measure_train(args, world, rank, local_rank, dev, with_cpu=True):
    """One data-parallel training step at the 3DHP shape (T=27, 30 clips per GPU, K=5): forward + default
    objective + backward + bucketed gradient all-reduce + Adam, weak-scaled (30 clips per GPU, as nn.DataParallel's per-GPU batch
    would be).  The step is ONE captured CUDA graph (--eager: kernel by kernel).  Returns the JSON object (on every rank)."""
    import torch
    import torch.distributed as dist
    import manipose_b200 as mb
    from manipose_b200 import metrics, ops
    from manipose_b200.optim import FusedAdam, GradientReducer

    t4, b4 = args.frames, args.train_clips
    torch.manual_seed(42)                      # identical replicas (FusedAdam broadcasts rank 0's state anyway)
    model = mb.RMCLManifoldMixSTE(mb.h36m17_skeleton(), num_frame=t4, n_hyp=K, drop_path_rate=args.drop_path)
    model = model.to(dev).train().set_compute_dtype(args.dtype)
    opt = FusedAdam(model, lr=4e-5, weight_decay=1e-6)
    gen = torch.Generator().manual_seed(1234 + rank)
    x = (0.3 * torch.randn(b4, t4, J, 2, generator=gen)).to(dev)
    y = 0.3 * torch.randn(b4, t4, J, 3, generator=gen)
    y[:, :, 0] = 0
    y = y.to(dev)
    loss_box = [None]
    loss_fn = lambda out, yy: metrics.losses.training_loss(out[0], out[1], yy)[0]

    def step():
        opt.zero_grad()
        poses, scores = model(x)
        loss, _ = metrics.losses.training_loss(poses, scores, y)
        loss.backward()
        opt.step()
        loss_box[0] = loss.detach()

    sync_all, timed = _make_timers(world, dev)
    for _ in range(max(args.warmup, 3)):
        step()
    launches0 = ops.LAUNCHES
    step()
    launches = ops.LAUNCHES - launches0
    run = step
    graph = None
    if args.cuda_graph:
        # the step is shape-static: capture forward + loss + backward + Adam once, replay it (removes ~500 launches of host work)
        from manipose_b200.optim import CapturedTrainStep
        sync_all()
        graph = CapturedTrainStep(model, opt, loss_fn, x, y, warmup=1)

        def run():
            loss_box[0] = graph(x, y)
        for _ in range(2):
            run()
    steps = max(args.steps, 20)
    with ClockSampler(local_rank) as clocks:
        ms = timed(run, steps)
        if ms < 2000.0:                        # ~10 ms steps: repeat until the 100 ms clock sampler has seen ~2 s of load
            extra = int(2000.0 / max(ms / steps, 1e-3)) + 1
            ms += timed(run, extra)
            steps += extra
    # the same step without the exchange (the reducer sees a world of 1): the difference is the exposed all-reduce time
    ms_local, local_steps = None, 100
    if world > 1:
        orig = GradientReducer.world_size
        GradientReducer.world_size = property(lambda self: 1)
        try:
            run_local, graph_local = step, None
            if args.cuda_graph:
                graph_local = CapturedTrainStep(model, opt, loss_fn, x, y, warmup=1)
                run_local = lambda: graph_local(x, y)
            for _ in range(2):
                run_local()
            ms_local = timed(run_local, local_steps)
            if graph_local is not None:
                graph_local.close()
        finally:
            GradientReducer.world_size = orig
        opt.broadcast_state()                  # the replicas drifted apart during the exchange-free steps
    frames = b4 * t4 * world
    value = frames * steps / (ms / 1000.0)
    peaks = measured_peaks()
    tflops = 3.0 * flops_per_frame(t4, K) * b4 * t4 * steps / (ms / 1000.0) / 1e12
    n_params = sum(p.numel() for p in model.parameters())
    cpu = None
    if rank == 0 and with_cpu and not args.skip_cpu_baseline and world == 1:
        cpu = cpu_train_rate(t4)
    line = {"metric": f"train_frames_per_sec_T{t4}" + ("_3DHP" if t4 == 27 else ""), "value": value, "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": f"ManiPose training step (forward + wta/bce/velocity/smoothness objective + backward + gradient all-reduce + Adam), "
                                   f"T={t4}, K={K}, J={J}, {b4} clips/GPU = BASELINE config 4; drop_path_rate {args.drop_path}",
                       "clips_per_gpu": b4, "frames_per_step": frames, "parallelism": f"data-parallel x{world}, bucketed NCCL all-reduce of "
                       f"{n_params} fp32 gradients ({n_params * 4 / 1e6:.1f} MB) overlapped with the backward sweep",
                       "cuda_graph": bool(args.cuda_graph), "gflop_per_frame_fwd_bwd": 3.0 * flops_per_frame(t4, K) / 1e9},
            "gpu_launches": launches * steps, "launches_per_step": launches, "clocks": clocks.summary(), "loss": float(loss_box[0]),
            "roofline": {"bound": "tensor", "achieved": tflops, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                         "frac": tflops / peaks["bf16_sustained"], "traffic": None,
                         "note": "whole step: 3 x forward matmul flops / step time; the step is launch / latency bound at 810 frames per GPU"},
            "allreduce": None if ms_local is None else {
                "payload_mb": n_params * 4 / 1e6, "exposed_ms_per_step": ms / steps - ms_local / local_steps,
                "ms_per_step_without_exchange": ms_local / local_steps,
                "how": "same step re-captured with the reducer seeing a world of 1, timed for 100 steps; exposed = difference"},
            "cpu_baseline": cpu}
    if graph is not None:
        graph.close()                          # before the process group goes: NCCL waits for graphs holding its collectives
    del model, opt
    torch.cuda.empty_cache()
    return line


def run_train(args):
    import torch.distributed as dist
    world, rank, local_rank, dev = _dist_env()
    line = measure_train(args, world, rank, local_rank, dev)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------- decoder alone (BASELINE config 2)
def measure_decoder(args, with_cpu=True):
    """BASELINE config 2: 6D -> SO(3) + forward kinematics + hypothesis softmax on N = 824 x 5 x 243 = 1,001,160 synthetic
    poses, fp32, one B200.  HBM roofline: 620 algorithmic bytes per pose (408 rot6d + 4 logit in, 204 pose + 4 score out)."""
    import torch
    import manipose_b200  # noqa: F401
    from manipose_b200 import ops
    dev = torch.device("cuda", torch.cuda.current_device())
    nc, k, t = 824, K, T
    n = nc * k * t
    gen = torch.Generator().manual_seed(1234)
    rot_host = torch.randn(n, J, 6, generator=gen).pin_memory()
    bones_host = (0.1 + 0.4 * torch.rand(nc, 16, generator=gen)).pin_memory()
    logits_host = torch.randn(nc, k, t, generator=gen).pin_memory()
    rot, bones, logits = rot_host.to(dev), bones_host.to(dev), logits_host.to(dev)
    poses_host, scores_host = torch.empty((n, J, 3)).pin_memory(), torch.empty((nc, k, t)).pin_memory()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)       # > the 126 MB L2, written between timed launches
    exact = not args.fast_decoder

    def step_resident():
        return ops.decoder_fwd(rot, bones, None, logits, nc, k, t, 6, exact)

    def step_e2e():
        p, sc = ops.decoder_fwd(rot_host.to(dev, non_blocking=True), bones_host.to(dev, non_blocking=True), None,
                                logits_host.to(dev, non_blocking=True), nc, k, t, 6, exact)
        poses_host.copy_(p, non_blocking=True)
        scores_host.copy_(sc, non_blocking=True)
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step_resident()
    steps = max(args.steps, 2000)               # ~0.5 s of launches so that the clock sampler sees the region
    launches0 = ops.LAUNCHES
    ms_kernel = 0.0
    with ClockSampler(dev.index or 0) as clocks:
        for _ in range(steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step_resident()
            e1.record()
            torch.cuda.synchronize()
            ms_kernel += e0.elapsed_time(e1)
    launches = ops.LAUNCHES - launches0
    for _ in range(2):
        step_e2e()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        step_e2e()
    ms_e2e = (time.perf_counter() - t0) * 1000.0 / 5
    ms = ms_kernel / steps
    peaks = measured_peaks()
    gbs = 620.0 * n / (ms / 1000.0) / 1e9
    cpu = None
    if with_cpu and not args.skip_cpu_baseline:
        from oracle import manipose_oracle as O
        torch.set_num_threads(os.cpu_count() or 1)
        nb = 103                                  # 103 clips x 5 x 243 = 125,145 poses (1/8 of the workload)
        rc, bc, zero = rot_host[:nb * k * t].clone(), bones_host[:nb].clone().unsqueeze(-1), torch.zeros(nb * k * t, 3)
        best = float("inf")
        for _ in range(3):
            t0 = time.perf_counter()
            O.pose_decoder(rc, bc, zero)
            best = min(best, time.perf_counter() - t0)
        cpu = {"value": nb * k * t / best, "unit": "poses/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"best of 3 passes over {nb * k * t} poses (1/8 of the workload), fp32, torch CPU (oracle.pose_decoder)"}
    line = {"metric": "decoded_poses_per_sec_1M", "value": n / (ms / 1000.0), "unit": "poses/s", "n_gpus": 1, "steps": steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"manifold decoder alone (6D -> SO(3) Gram-Schmidt, T-pose from bone lengths, forward kinematics, softmax over K) "
                                   f"on {n} poses = {nc} clips x {k} x {t} = BASELINE config 2",
                       "mode": "exact (bit-identical to the correctly rounded oracle)" if exact else "fast (one-MUFU reciprocals / rsqrt)",
                       "l2": "256 MB written between timed launches (L2 flush)"},
            "e2e": {"value": n / (ms_e2e / 1000.0), "unit": "poses/s", "h2d_bytes_per_step": (rot_host.numel() + bones_host.numel() + logits_host.numel()) * 4,
                    "d2h_bytes_per_step": (poses_host.numel() + scores_host.numel()) * 4, "ms_per_step": ms_e2e},
            "gpu_launches": launches, "clocks": clocks.summary(),
            "roofline": {"bound": "hbm", "kernel": "decoder_fwd_kernel", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": gbs / peaks["hbm_gbs"], "traffic": None, "algorithmic_bytes_per_pose": 620,
                         "peak_source": "measured (STREAM-style copy)" if peaks["src"] == "measured" else "fallback"},
            "cpu_baseline": cpu}
    return line


def run_decoder(args):
    import torch
    torch.cuda.set_device(0)
    print(json.dumps(measure_decoder(args)), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--clips", type=int, default=1024, help="clips per step: in total (strong scaling, the default: BASELINE config 3 = 1024 "
                    "clips sharded over the GPUs) or per GPU (--scaling weak)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="N > 1: strong = --clips in total, 1024/N per GPU as BASELINE config 3 states; weak = --clips per GPU")
    ap.add_argument("--train-clips", type=int, default=30, help="config 4: clips per GPU per training step")
    ap.add_argument("--headline-only", action="store_true",
                    help="profiling runs: only the config-3 line at --dtype (no fp16 / weak / parity / train / decoder sub-objects)")
    ap.add_argument("--micro-batch-clips", type=int, default=0)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp16"], help="16-bit operand format of the backbone")
    ap.add_argument("--skip-cpu-baseline", action="store_true", help="profiling runs only: do not time the CPU oracle")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=[2, 3, 4],
                    help="3: T=243 inference (headline); 4: T=27 training step; 2: manifold decoder alone on 1M poses")
    ap.add_argument("--fast-decoder", action="store_true", help="config 2: the FAST decoder mode instead of the bit-exact one")
    ap.add_argument("--drop-path", type=float, default=0.1, help="config 4: stochastic depth rate (drivers use 0.1)")
    ap.add_argument("--frames", type=int, default=27, help="config 4: frames per clip (27 = 3DHP shape of BASELINE config 4; 243 = H36M)")
    ap.add_argument("--cuda-graph", action="store_true", help="config 4: capture the whole training step (NCCL all-reduces included) in a CUDA graph (default)")
    ap.add_argument("--eager", action="store_true", help="config 4: launch the step eagerly instead (also reports the exposed all-reduce time)")
    args = ap.parse_args()
    args.cuda_graph = not args.eager
    if args.impl == "reference":
        run_reference(args)
    elif args.config == 4:
        run_train(args)
    elif args.config == 2:
        run_decoder(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()

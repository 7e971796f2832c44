"""Run under torchrun on >= 2 GPUs: the CUDA-graph replay of the data-parallel training step (NCCL all-reduces captured) performs the
same updates as the eager data-parallel step.  Prints one JSON line on rank 0 and exits non-zero on disagreement.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/check_graph_ddp.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import manipose_b200 as mb
    from manipose_b200 import metrics
    from manipose_b200.optim import FusedAdam, CapturedTrainStep
    gen = torch.Generator().manual_seed(3 + rank)                      # different data per rank
    xs = [(0.3 * torch.randn(4, 27, 17, 2, generator=gen)).cuda() for _ in range(3)]
    ys = [(0.3 * torch.randn(4, 27, 17, 3, generator=gen)).cuda() for _ in range(3)]
    loss_fn = lambda out, y: metrics.losses.training_loss(out[0], out[1], y)[0]

    def make():
        torch.manual_seed(0)
        m = mb.RMCLManifoldMixSTE(mb.h36m17_skeleton(), num_frame=27, n_hyp=5, drop_path_rate=0.0).cuda().train()
        return m, FusedAdam(m, lr=1e-4, weight_decay=1e-6)

    rel = lambda a, b: float((a - b).norm() / b.norm())
    m1, o1 = make()
    sd0 = {k: v.clone() for k, v in m1.state_dict().items()}
    eager = []
    for x, y in zip(xs, ys):
        o1.zero_grad()
        loss = loss_fn(m1(x), y)
        loss.backward()
        o1.step()
        eager.append(float(loss.detach()))
    m2, o2 = make()
    step = CapturedTrainStep(m2, o2, loss_fn, xs[0], ys[0], warmup=2)
    m2.load_state_dict(sd0)
    o2.exp_avg.zero_()
    o2.exp_avg_sq.zero_()
    o2.step_dev.zero_()
    o2._invalidate_shadows()
    graphed = [float(step(x, y)) for x, y in zip(xs, ys)]
    torch.cuda.synchronize()
    r1, r2 = rel(o2.exp_avg, o1.exp_avg), rel(o2.exp_avg_sq, o1.exp_avg_sq)
    # replicas must stay identical across ranks (same reduced gradients everywhere)
    p = o2.flat.flat_param.clone()
    lo, hi = p.clone(), p.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    spread = float((hi - lo).abs().max())
    ok = (abs(eager[0] - graphed[0]) <= 1e-5 * abs(eager[0]) and all(abs(a - b) <= 2e-3 * abs(a) for a, b in zip(eager, graphed))
          and r1 <= 1.5e-1 and r2 <= 5e-2 and spread == 0.0 and int(o2.step_dev) == 3)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"check": "graph_ddp", "world": world, "ok": bool(int(flag)), "eager_loss": eager, "graph_loss": graphed,
                          "exp_avg_rel": r1, "exp_avg_sq_rel": r2, "replica_spread": spread}), flush=True)
    step.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag) else 1)


if __name__ == "__main__":
    main()

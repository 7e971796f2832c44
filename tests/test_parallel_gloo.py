"""N > 1 host logic on CPU: world_size-2 gloo process group.  The data path has no collective; what is
checked is that the clip shards partition the batch, that per-rank partial metric sums reduce to the global sums, and that
gathered shards reassemble the batch in order."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp_


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_clips, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from manipose_b200 import parallel as P
    g = torch.Generator().manual_seed(0)
    x = torch.randn(n_clips, 9, 17, 2, generator=g)                     # same batch on every rank
    y = torch.randn(n_clips, 9, 17, 3, generator=g)
    pred = y + 0.01 * torch.randn(n_clips, 9, 17, 3, generator=g)      # stands in for the lifted poses of each clip
    lo, hi = P.shard_range(n_clips)
    xs = P.shard_clips(x)
    assert xs.shape[0] == hi - lo and torch.equal(xs, x[lo:hi])
    err = (pred[lo:hi] - y[lo:hi]).norm(dim=-1)                         # per-joint errors of this rank's clips only
    sums = torch.tensor([float(err.sum()), float(err.numel())], dtype=torch.float64)
    P.reduce_metric_sums(sums)
    full = (pred - y).norm(dim=-1)
    assert abs(float(sums[0]) - float(full.sum())) < 1e-6 * float(full.sum())
    assert int(sums[1]) == full.numel()
    gathered = P.gather_clips(pred[lo:hi].contiguous(), n_clips)
    if rank == 0:
        assert torch.equal(gathered, pred)
    else:
        assert gathered is None
    torch.save({"lo": lo, "hi": hi}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_clips", [8, 7, 1])
def test_clip_sharding_world2_gloo(tmp_path, n_clips):
    world = 2
    mp_.spawn(_worker, args=(world, _free_port(), n_clips, str(tmp_path)), nprocs=world, join=True)
    spans = [torch.load(tmp_path / f"r{r}.pt") for r in range(world)]
    assert spans[0]["lo"] == 0 and spans[-1]["hi"] == n_clips
    assert all(a["hi"] == b["lo"] for a, b in zip(spans, spans[1:]))    # contiguous, disjoint, complete


def test_shard_range_is_balanced():
    from manipose_b200.parallel import shard_range
    for n in (0, 1, 5, 1024, 1031):
        for w in (1, 2, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            sizes = [b - a for a, b in spans]
            assert sum(sizes) == n and max(sizes) - min(sizes) <= 1
            assert spans[0][0] == 0 and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _grad_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import manipose_b200 as mb
    from manipose_b200.optim import FlatParameters, GradientReducer, block_buckets
    torch.manual_seed(0)                                               # same replica on every rank
    m = mb.RMCLManifoldMixSTE(mb.h36m17_skeleton(), num_frame=9, n_hyp=2, depth_rot=2, depth_seg=1)
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    flat = FlatParameters(m)
    assert all(torch.equal(v, sd0[k]) for k, v in m.state_dict().items())          # re-homing keeps the values
    buckets, index = block_buckets(flat, m)
    assert len(index) == 4 and buckets[0][0] == 0 and buckets[-1][1] == flat.numel
    red = GradientReducer(flat.flat_grad, buckets)
    g = torch.Generator().manual_seed(100 + rank)                      # rank-specific "gradients" written through p.grad
    local = {}
    for name, p in m.named_parameters():
        local[name] = torch.randn(p.shape, generator=g)
        p.grad.copy_(local[name])
    # blocks report in backward order (as MixSTE._train_backward does), the rest is swept up by finish()
    rot = m.rotations_module
    for blk in reversed([b for pair in zip(rot.STEblocks, rot.TTEblocks) for b in pair]):
        red.bucket_ready(index[id(blk)])
    scale = red.finish()
    assert scale == 1.0 / world
    torch.save({n: p.grad.clone() for n, p in m.named_parameters()}, os.path.join(out_dir, f"g{rank}.pt"))
    torch.save(local, os.path.join(out_dir, f"l{rank}.pt"))
    flat.zero_grad()
    assert float(flat.flat_grad.abs().sum()) == 0.0 and all(p.grad.data_ptr() >= flat.flat_grad.data_ptr() for p in m.parameters())
    dist.barrier()
    dist.destroy_process_group()


def test_bucketed_gradient_allreduce_world2_gloo(tmp_path):
    """Training exchange step: per-block buckets of the flat gradient buffer are sum-all-reduced; every rank ends
    with the same sums, equal to the sum of the per-rank gradients."""
    world = 2
    mp_.spawn(_grad_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    reduced = [torch.load(tmp_path / f"g{r}.pt") for r in range(world)]
    local = [torch.load(tmp_path / f"l{r}.pt") for r in range(world)]
    for name in reduced[0]:
        want = local[0][name] + local[1][name]
        assert torch.allclose(reduced[0][name], want, rtol=1e-6, atol=1e-6), name
        assert torch.equal(reduced[0][name], reduced[1][name]), name

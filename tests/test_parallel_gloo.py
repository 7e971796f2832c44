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

"""CPU, world_size 2, gloo: the host-side sharding logic of latentfusion_b200/dist.py (the collectives
that NCCL runs on the GPU box)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, result_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from latentfusion_b200 import dist as lfdist
    torch.manual_seed(0)
    V = 5
    views = torch.randn(V, 3, 4, 4, 4)                 # every rank builds the same full set
    lo, hi = lfdist.shard_range(V, rank, world)
    local = views[lo:hi]
    ok = True
    ok &= torch.allclose(lfdist.fuse_views_sharded(local, 'mean', V, rank, world), views.mean(0, keepdim=True), atol=1e-6)
    ok &= torch.equal(lfdist.fuse_views_sharded(local, 'max', V, rank, world), views.max(0, keepdim=True)[0])
    ok &= torch.equal(lfdist.fuse_views_sharded(local, 'gather', V, rank, world), views)
    # hypothesis sharding: global ranking from per-rank losses
    losses = torch.tensor([3.0, 0.5, 2.0]) if rank == 0 else torch.tensor([1.0, 4.0])
    params = torch.arange(losses.numel() * 2, dtype=torch.float32).view(-1, 2) + 100 * rank
    top_l, top_p, owner = lfdist.merge_rankings(losses, params, ranking_size=3)
    ok &= top_l.tolist() == [0.5, 1.0, 2.0] and owner.tolist() == [0, 1, 0]
    ok &= top_p.tolist() == [[2.0, 3.0], [100.0, 101.0], [4.0, 5.0]]
    open(os.path.join(result_dir, f'ok{rank}'), 'w').write(str(bool(ok)))
    dist.destroy_process_group()


def test_sharding_collectives_world2(tmp_path):
    port = 29500 + (os.getpid() % 500)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert [open(tmp_path / f'ok{r}').read() for r in range(2)] == ['True', 'True']


def test_shard_range_covers_everything():
    sys.path.insert(0, ROOT)
    from latentfusion_b200.dist import shard_range
    for n in (1, 7, 8, 16, 17):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1

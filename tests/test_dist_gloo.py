"""world_size-2 gloo test (CPU) of the data-parallel host logic: flat-parameter broadcast,
bucketed gradient all-reduce and the 1/world fold-in, exactly as the Trainer wires them."""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    b2 = ge.load_package()
    from b2pose import parallel as P
    import torch.distributed as dist
    globals()["P"] = P
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, _, w = P.init_from_env("gloo")
    assert (r, w) == (rank, world)

    class Flat:
        pass
    f = Flat()
    f.n = 1000
    torch.manual_seed(rank)
    f.w = torch.randn(f.n)
    f.w16 = None
    f.g = torch.full((f.n,), float(rank + 1))
    P.broadcast_flat(f, dist.group.WORLD)
    buckets = P.GradBuckets(f, dist.group.WORLD, bucket_mb=256 * 4 / (1 << 20))
    assert len(buckets.bounds) == 4 and buckets.bounds[0][1] == f.n      # reverse order
    buckets.allreduce()
    # two-stage exchange (overlapped all-reduce of the deep ranges, then the rest): every element reduced exactly once
    f2 = Flat()
    f2.n = 1000
    f2.g = torch.arange(f2.n, dtype=torch.float32) * (rank + 1)
    staged = P.GradBuckets(f2, dist.group.WORLD, bucket_mb=128 * 4 / (1 << 20), deep=[(192, 448), (448, 512), (896, 1000)])
    deep_elems = sorted(i for lo_, hi_ in staged.deep_bounds for i in range(lo_, hi_))
    shallow_elems = sorted(i for lo_, hi_ in staged.shallow_bounds for i in range(lo_, hi_))
    assert deep_elems == list(range(192, 512)) + list(range(896, 1000))
    assert sorted(deep_elems + shallow_elems) == list(range(1000))
    assert max(hi_ - lo_ for lo_, hi_ in staged.deep_bounds + staged.shallow_bounds) <= 128
    staged.start_deep()
    mid = f2.g.clone()
    staged.finish()
    lo, hi = P.shard_range(10, rank, world)
    torch.save(dict(w=f.w, g=f.g, shard=(lo, hi), g2=f2.g, mid=mid), out % rank)
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_allreduce(tmp_path):
    port = _free_port()
    out = str(tmp_path / "r%d.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    a, b = torch.load(out % 0), torch.load(out % 1)
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    ge.load_package()
    from b2pose import parallel as P
    assert torch.equal(a["w"], b["w"])                                    # broadcast from rank 0
    torch.manual_seed(0)
    assert torch.equal(a["w"], torch.randn(1000))
    assert torch.equal(a["g"], torch.full((1000,), 3.0)) and torch.equal(b["g"], a["g"])   # 1 + 2
    assert a["shard"] == (0, 5) and b["shard"] == (5, 10)
    want = torch.arange(1000, dtype=torch.float32) * 3
    assert torch.equal(a["g2"], want) and torch.equal(b["g2"], want)
    assert P.merge_ranges([(5, 9), (0, 3), (3, 5), (20, 30)]) == [(0, 9), (20, 30)]
    assert P.split_ranges(40, [(20, 30), (0, 9)]) == ([(0, 9), (20, 30)], [(9, 20), (30, 40)])

"""N>1 host logic on CPU: two gloo ranks shard one ragged batch with
starflate_b200.sharding.partition_streams, each handles its contiguous range (here with the
oracle standing in for a GPU — this is a test of the partition / gather arithmetic, not of the
kernels), and the gathered per-stream results equal the unsharded run.  No data-path collective:
the only traffic is the gather of results and the max-over-ranks of a timing."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle import bindings
    from starflate_b200 import sharding
    from tests import deflate_tools as T
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(11)  # same batch on every rank
        streams, caps = [], []
        for i in range(90):
            kind = ["dynamic", "fixed", "stored", "repetitive"][i % 4]
            plain, comp = T.make_stream(kind, int(rng.integers(1, 20000)), 600 + i)
            streams.append(comp if i % 10 else comp[: len(comp) // 2])
            caps.append(len(plain))
        b = T.Batch(streams, caps)
        parts = sharding.partition_streams(b.src_len, b.dst_cap, world)
        assert sum(c for _, c in parts) == b.n and parts[0][0] == 0
        first, count = parts[rank]
        orc = bindings.load_oracle()
        dst = b.new_dst()
        sl = slice(first, first + count)
        st, wr, _ = orc.decompress_batch(b.src, b.src_off[sl].copy(), b.src_len[sl].copy(), dst,
                                         b.dst_off[sl].copy(), b.dst_cap[sl].copy())
        # gather the per-stream results on every rank (variable counts: pad to n)
        mine = torch.full((b.n, 2), -1, dtype=torch.int64)
        mine[sl, 0] = torch.from_numpy(np.asarray(st, dtype=np.int64))
        mine[sl, 1] = torch.from_numpy(np.asarray(wr, dtype=np.int64))
        dist.all_reduce(mine, op=dist.ReduceOp.MAX)
        slowest = sharding.max_over_ranks(float(rank + 1))
        assert slowest == float(world)
        if rank == 0:
            full = b.new_dst()
            fst, fwr, _ = orc.decompress_batch(b.src, b.src_off, b.src_len, full, b.dst_off, b.dst_cap)
            assert (mine[:, 0].numpy() == np.asarray(fst)).all()
            assert (mine[:, 1].numpy() == np.asarray(fwr, dtype=np.int64)).all()
            cost = (b.src_len + b.dst_cap).astype(np.float64)
            loads = [cost[f:f + c].sum() for f, c in parts]
            assert max(loads) <= 1.0 * cost.sum() / world + cost.max()  # as even as a contiguous cut allows
            open(os.path.join(out_dir, "ok"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_over_gloo(tmp_path):
    import torch.multiprocessing as mp
    from oracle import bindings
    bindings.build(with_reference=None)
    port = 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()


def test_partition_properties():
    from starflate_b200 import sharding
    rng = np.random.default_rng(1)
    for world in (1, 2, 4, 8):
        for n in (0, 1, 3, 1000):
            sl = rng.integers(0, 5000, n).astype(np.uint64)
            dc = rng.integers(0, 70000, n).astype(np.uint64)
            parts = sharding.partition_streams(sl, dc, world)
            assert len(parts) == world and parts[0][0] == 0
            assert all(parts[r][0] + parts[r][1] == parts[r + 1][0] for r in range(world - 1))
            assert parts[-1][0] + parts[-1][1] == n

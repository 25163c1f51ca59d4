"""world_size-2 gloo test of the multi-GPU plumbing (proof sharding, max-over-ranks timing, gather)."""
import os
import sys

import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from shielded_pool_pinocchio_solana_b200 import dist as gd
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = gd.shard_range(9, rank, world)
    mx = gd.max_over_ranks(10.0 + rank)
    proofs = [bytes([rank * 16 + i]) * 388 for i in range(2)]
    allp = gd.gather_proofs(proofs, 388)
    q.put((rank, lo, hi, mx, [p[0] for p in allp], [len(p) for p in allp]))
    dist.destroy_process_group()


def test_shard_max_gather_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [(0, 5), (5, 9)]       # 9 proofs over 2 ranks
    assert all(r[3] == 11.0 for r in res)                         # max over ranks
    assert all(r[4] == [0, 1, 16, 17] and r[5] == [388] * 4 for r in res)


def test_shard_range_partitions():
    sys.path.insert(0, ROOT)
    from shielded_pool_pinocchio_solana_b200.dist import shard_range
    for n in (0, 1, 7, 4096):
        for world in (1, 2, 4, 8):
            parts = [shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1

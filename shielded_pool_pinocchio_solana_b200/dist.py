"""Multi-GPU plumbing for the batch path: one process per GPU, proofs sharded
one-per-GPU-slot with NO data-path collective; `torch.distributed` only carries the timing barrier,
the max-over-ranks reduction and (for result collection) a gather of the proof bytes."""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous, balanced slice [lo, hi) of n_items for `rank` (first n_items % world ranks get one more)."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def max_over_ranks(value, device="cpu"):
    """Device/host scalar -> max over all ranks (the multi-GPU timing rule)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_proofs(local_proofs, proof_len, device="cpu"):
    """All ranks contribute their proofs (bytes, equal count per rank); rank 0 gets them in rank order."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return list(local_proofs)
    flat = torch.frombuffer(bytearray(b"".join(local_proofs)), dtype=torch.uint8).to(device)
    out = [torch.empty_like(flat) for _ in range(dist.get_world_size())]
    dist.all_gather(out, flat)
    res = []
    for t in out:
        b = bytes(t.cpu().numpy().tobytes())
        res += [b[i:i + proof_len] for i in range(0, len(b), proof_len)]
    return res


# ---- one large MSM split across GPUs ------------------------------------
def combine_partials(ctx, partials, group="g1"):
    """Sum of the per-rank partial results (gnark raw point bytes) -- itself a tiny MSM with unit
    scalars on the device, so no host curve arithmetic exists anywhere."""
    one = (1).to_bytes(32, "big")
    bases = ctx.load_bases(b"".join(partials), group, window=4)
    try:
        return bases.msm(one * len(partials), batch=1)
    finally:
        bases.free()


def msm_sharded(ctx, points_be, scalars_be, group="g1", bases_cache=None):
    """out = sum_i s_i P_i with the point set split into contiguous ranges, one per rank; the partial
    sums (64 B / 128 B each) are exchanged with ONE all-gather over NCCL and added on every rank.
    `bases_cache`: dict reused across calls so a rank's slice stays resident (a proving key is static)."""
    size = 64 if group == "g1" else 128
    n = len(points_be) // size
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    lo, hi = shard_range(n, rank, world)
    key = (group, lo, hi)
    bases = bases_cache.get(key) if bases_cache is not None else None
    if bases is None and hi > lo:
        bases = ctx.load_bases(points_be[lo * size:hi * size], group)
        if bases_cache is not None:
            bases_cache[key] = bases
    inf = b"\x00" * size
    partial = bases.msm(scalars_be[lo * 32:hi * 32], batch=1) if hi > lo else inf
    if world == 1:
        return partial
    t = torch.frombuffer(bytearray(partial), dtype=torch.uint8).cuda()
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return combine_partials(ctx, [bytes(x.cpu().numpy().tobytes()) for x in out], group)

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

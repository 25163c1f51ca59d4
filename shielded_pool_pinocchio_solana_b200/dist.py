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


# ---- one large PROOF split across GPUs (BASELINE.json configs[3]) -------------------
def share_unique_id(rank, world):
    """Rank 0 draws the NCCL id of the library's communicator; torch.distributed only carries the 128 bytes."""
    from . import api
    box = [api.comm_unique_id() if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(box, src=0)
    return box[0]


def bench_single_proof(_ctx, torch_mod, stream, rank, world, logn=22, reps=3, seed=0x5EED22):
    """One proof of a synthetic 2^logn-constraint circuit with every MSM point set split across `world`
    ranks (contiguous index ranges resident per GPU); SpMV and the quotient run replicated, the partial
    G1 / G2 sums meet in one NCCL all-gather inside g16_prove_wires_dev.  Returns the bench sub-record
    (strong scaling: the work is fixed as N grows); `proof_sha256` must be identical for every N."""
    import hashlib
    import time as _time
    from . import api
    local = torch_mod.cuda.current_device()
    ctx = api.Context(local)
    ctx.set_stream(stream.cuda_stream)
    ctx.comm_init(share_unique_id(rank, world), rank, world)
    t0 = _time.time()
    ccs = api.synth_ccs(1 << logn, 2, 64, seed)
    pk, _vk = ctx.setup(ccs, b"single-proof-%d" % logn)
    circ = ctx.load_circuit(ccs, pk)
    del pk
    setup_s = _time.time() - t0
    nw = circ.info["nb_wires"]
    gen = torch_mod.Generator(device="cuda").manual_seed(seed)          # same wires on every rank
    wires = torch_mod.randint(-2**31, 2**31 - 1, (nw, 8), dtype=torch_mod.int32, device="cuda", generator=gen)
    wires[:, 7] &= 0x0FFFFFFF                                           # any value < r is a valid representative
    out = torch_mod.empty((1, 80), dtype=torch_mod.int32, device="cuda")
    rnd = bytes(range(1, 97))
    circ.prove_wires_dev(wires.data_ptr(), 1, out.data_ptr(), rnd)      # warm-up (scratch allocation)
    torch_mod.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch_mod.cuda.Event(enable_timing=True), torch_mod.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        circ.prove_wires_dev(wires.data_ptr(), 1, out.data_ptr(), rnd)
    e1.record(stream)
    torch_mod.cuda.synchronize()
    ms = max_over_ranks(e0.elapsed_time(e1) / reps, device="cuda")
    digest = hashlib.sha256(out.cpu().numpy().tobytes()).hexdigest()
    digests = [digest]
    if world > 1:
        digests = [None] * world
        dist.all_gather_object(digests, digest)
    info = circ.info
    circ.free()
    ctx.close()
    return {"workload": "synthetic R1CS, 2^%d constraints, random wires, one proof; MSM point sets split across ranks "
                        "(BASELINE.json configs[3])" % logn,
            "n_gpus": world, "scaling": "strong", "ms_per_proof": ms, "proofs_per_s": 1e3 / ms, "reps": reps,
            "msm_points_total": {k: info[k] for k in ("n_a", "n_b", "n_k", "n_z")},
            "collective": "2 x ncclAllGather per proof (4 G1 + 1 G2 partial sums per rank)" if world > 1 else "none",
            "proof_sha256": digest, "identical_on_all_ranks": len(set(digests)) == 1,
            "setup_and_load_s": setup_s}

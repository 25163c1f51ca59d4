"""CUDA MSM (through the C ABI) vs the big-integer oracle -- bit-exact.

Reference behaviour: gnark-crypto G1Jac.MultiExp / G2Jac.MultiExp as used by `sunspot prove`
(/root/reference/client/proof.helper.ts:64).
"""
import json
import os
import random

import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "msm_small.json")


@pytest.mark.parametrize("group", ["g1", "g2"])
@pytest.mark.parametrize("window", [0, 4, 9, 13, 16])
def test_msm_golden(ctx, group, window):
    case = json.load(open(GOLDEN))[group]
    bases = ctx.load_bases(bytes.fromhex(case["points"]), group, window=window, batch_hint=case["batch"])
    got = bases.msm(bytes.fromhex(case["scalars"]), batch=case["batch"])
    assert got.hex() == case["results"]
    bases.free()


def test_generated_points_match_oracle(ctx):
    import bn254 as B
    import serialize as S
    M = (1 << 64) - 1

    def splitmix64(x):
        x = (x + 0x9e3779b97f4a7c15) & M
        x = ((x ^ (x >> 30)) * 0xbf58476d1ce4e5b9) & M
        x = ((x ^ (x >> 27)) * 0x94d049bb133111eb) & M
        return x ^ (x >> 31)

    def k_of(seed, i):
        k = 0
        for w in range(4):
            k |= splitmix64((seed + 4 * i + w) & M) << (64 * w)
        return k & ((1 << 253) - 1)

    seed = 0xB200
    pts = ctx.generate_points(8, seed, "g1")
    for i in range(8):
        assert pts[64 * i:64 * i + 64] == S.g1_to_bytes(B.g1_mul(B.G1_GEN, k_of(seed, i)))
    pts2 = ctx.generate_points(3, seed, "g2")
    for i in range(3):
        assert pts2[128 * i:128 * i + 128] == S.g2_to_bytes(B.g2_mul(B.G2_GEN, k_of(seed, i)))


@pytest.mark.parametrize("n,batch", [(1 << 12, 4), (5000, 1), (1 << 14, 2)])
def test_msm_linearity_large(ctx, n, batch):
    """Size-independent property at sizes the big-int oracle cannot reach quickly:
    MSM(P, s) + MSM(P, t) == MSM(P, s + t)  and  MSM(P, e_j) == P_j."""
    import bn254 as B
    import serialize as S
    rng = random.Random(n)
    pts = ctx.generate_points(n, 0x51, "g1")
    bases = ctx.load_bases(pts, "g1", batch_hint=3 * batch)
    s = [[rng.randrange(B.R) for _ in range(n)] for _ in range(batch)]
    t = [[rng.randrange(B.R) for _ in range(n)] for _ in range(batch)]
    u = [[(a + b) % B.R for a, b in zip(sr, tr)] for sr, tr in zip(s, t)]
    enc = lambda rows: b"".join(S.fr_to_bytes(x) for row in rows for x in row)
    out = bases.msm(enc(s + t + u), batch=3 * batch)
    P = [S.g1_from_bytes(out[64 * i:64 * i + 64]) for i in range(3 * batch)]
    for b in range(batch):
        assert B.g1_add(P[b], P[batch + b]) == P[2 * batch + b]
    # unit vectors pick out single bases
    j = rng.randrange(n)
    e = [[1 if i == j else 0 for i in range(n)]]
    assert bases.msm(enc(e), batch=1) == pts[64 * j:64 * j + 64]
    bases.free()


@pytest.mark.parametrize("group", ["g1", "g2"])
def test_point_range_split_and_combine(ctx, group):
    """Multi-GPU single-MSM path (point ranges per rank + all-gather of partial sums), emulated on
    one GPU: three unequal shards, partials combined on the device, equals the unsplit golden result."""
    from shielded_pool_pinocchio_solana_b200.dist import combine_partials, shard_range
    case = json.load(open(GOLDEN))[group]
    size = 64 if group == "g1" else 128
    pts, sc, n = bytes.fromhex(case["points"]), bytes.fromhex(case["scalars"]), case["n"]
    partials = []
    for r in range(3):
        lo, hi = shard_range(n, r, 3)
        b = ctx.load_bases(pts[lo * size:hi * size], group)
        partials.append(b.msm(sc[lo * 32:hi * 32], batch=1))
        b.free()
    assert combine_partials(ctx, partials, group).hex() == case["results"][:2 * size]
    # an all-zero shard contributes the point at infinity
    inf = b"\x00" * size
    assert combine_partials(ctx, partials + [inf], group).hex() == case["results"][:2 * size]


@pytest.mark.parametrize("group", ["g1", "g2"])
def test_sliced_single_msm_matches_golden(ctx, group, monkeypatch):
    """One large standalone MSM runs as 4 point slices pipelined on 4 streams (g16_msm_dev, >= 2^20 points): forced
    here on the golden vectors (300 / 96 points incl. infinities, repeated and all-equal points), same bytes."""
    import numpy as np
    import torch
    case = json.load(open(GOLDEN))[group]
    n = len(bytes.fromhex(case["points"])) // (64 if group == "g1" else 128)
    bases = ctx.load_bases(bytes.fromhex(case["points"]), group, window=0, batch_hint=1)
    sc_be = bytes.fromhex(case["scalars"])[:32 * n]                       # first batch element
    want = bases.msm(sc_be, batch=1)
    limbs = np.frombuffer(sc_be, dtype=">u4").reshape(n, 8)[:, ::-1].astype(np.uint32)
    d_sc = torch.from_numpy(np.ascontiguousarray(limbs).view(np.int32)).cuda()
    out = torch.zeros(64 if group == "g2" else 32, dtype=torch.int32, device="cuda")
    monkeypatch.setenv("G16_MSM_SLICE_MIN", "16")
    torch.cuda.synchronize()
    bases.msm_dev(d_sc.data_ptr(), 1, out.data_ptr(), montgomery=False)
    ctx.sync()
    torch.cuda.synchronize()
    got_mont = out.cpu().numpy().view(np.uint32)
    monkeypatch.setenv("G16_MSM_SLICE_MIN", "0")
    out2 = torch.zeros_like(out)
    bases.msm_dev(d_sc.data_ptr(), 1, out2.data_ptr(), montgomery=False)
    ctx.sync()
    torch.cuda.synchronize()
    assert (got_mont == out2.cpu().numpy().view(np.uint32)).all()          # sliced == unsliced (Montgomery limbs)
    assert want == bytes.fromhex(case["results"])[:len(want)]
    bases.free()

"""CUDA NTT / quotient kernels (through the C ABI) vs the big-integer oracle -- bit-exact.

Reference behaviour: gnark-crypto fr/fft (DIF/DIT, OnCoset) and gnark computeH as run inside
`sunspot prove` (/root/reference/client/proof.helper.ts:64).
"""
import random

import pytest

import bn254 as B
import serialize as S

pytestmark = pytest.mark.gpu


def enc(v):
    return b"".join(S.fr_to_bytes(x) for x in v)


def dec(b):
    return [int.from_bytes(b[i:i + 32], "big") for i in range(0, len(b), 32)]


@pytest.mark.parametrize("logn", [1, 2, 3, 6, 7, 8, 10, 11, 12, 13])
@pytest.mark.parametrize("coset", [False, True])
def test_ntt_matches_oracle(ctx, logn, coset):
    n = 1 << logn
    rng = random.Random(logn * 2 + coset)
    batch = 3 if logn <= 10 else 1
    vecs = [[rng.randrange(B.R) for _ in range(n)] for _ in range(batch)]
    vecs[0][0] = 0
    vecs[0][n - 1] = B.R - 1
    got = dec(ctx.ntt(b"".join(enc(v) for v in vecs), logn, batch, inverse=False, coset=coset))
    for k, v in enumerate(vecs):
        want = B.bitrev_permute(B.ntt_natural(v, coset=B.COSET_GEN if coset else None))
        assert got[k * n:(k + 1) * n] == want
    # inverse takes bit-reversed input back to the natural-order original
    back = dec(ctx.ntt(enc(got), logn, batch, inverse=True, coset=coset))
    assert back == [x for v in vecs for x in v]


@pytest.mark.parametrize("logn", [14, 15, 18, 20])
def test_ntt_roundtrip_large(ctx, logn):
    """Size-independent properties at sizes the oracle is too slow for: inverse(forward(x)) == x,
    a delta transforms to all-ones, and linearity."""
    n = 1 << logn
    rng = random.Random(logn)
    x = [rng.randrange(B.R) for _ in range(n)]
    fx = ctx.ntt(enc(x), logn)
    assert dec(ctx.ntt(fx, logn, inverse=True)) == x
    delta = [1] + [0] * (n - 1)
    assert dec(ctx.ntt(enc(delta), logn)) == [1] * n
    if logn <= 15:
        y = [rng.randrange(B.R) for _ in range(n)]
        fy = dec(ctx.ntt(enc(y), logn))
        fxy = dec(ctx.ntt(enc([(a + b) % B.R for a, b in zip(x, y)]), logn))
        assert fxy == [(a + b) % B.R for a, b in zip(dec(fx), fy)]
        # spot-check 3 output points against the definition (Horner on the oracle side)
        w = B.fr_root(n)
        for k in (1, n // 3, n - 1):
            pt = pow(w, k, B.R)
            acc = 0
            for coef in reversed(x):
                acc = (acc * pt + coef) % B.R
            assert dec(fx)[B.bitrev(k, logn)] == acc


@pytest.mark.parametrize("logn,nproofs", [(3, 2), (8, 3), (11, 2), (12, 1)])
def test_compute_h_matches_oracle(ctx, logn, nproofs):
    n = 1 << logn
    rng = random.Random(100 + logn)
    blob, want = b"", []
    for _ in range(nproofs):
        # satisfiable rows: c = a*b on the first rows, zero padding after (as gnark pads to n)
        used = n - rng.randrange(0, n // 4 + 1)
        a = [rng.randrange(B.R) for _ in range(used)] + [0] * (n - used)
        b = [rng.randrange(B.R) for _ in range(used)] + [0] * (n - used)
        c = [x * y % B.R for x, y in zip(a, b)]
        blob += enc(a) + enc(b) + enc(c)
        h = B.quotient_h(a, b, c)
        assert h[n - 1] == 0
        want += B.bitrev_permute(h)
    assert dec(ctx.compute_h(blob, logn, nproofs)) == want

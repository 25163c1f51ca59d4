"""CPU checks that pin the big-integer oracle to definitions and to the reference's artifacts."""
import random

import bn254 as B


def test_constants():
    assert pow(B.ROOT_2_28, 1 << 28, B.R) == 1 and pow(B.ROOT_2_28, 1 << 27, B.R) == B.R - 1
    assert B.g1_on_curve(B.G1_GEN) and B.g2_on_curve(B.G2_GEN)
    assert B.g1_mul(B.G1_GEN, B.R - 1) == B.g1_neg(B.G1_GEN)
    assert B.g2_mul(B.G2_GEN, B.R - 1) == B.g2_neg(B.G2_GEN)


def test_ntt_is_the_definition():
    rng = random.Random(1)
    n = 16
    a = [rng.randrange(B.R) for _ in range(n)]
    w = B.fr_root(n)
    fa = B.ntt_natural(a)
    for k in range(n):
        assert fa[k] == sum(a[j] * pow(w, k * j, B.R) for j in range(n)) % B.R
    ca = B.ntt_natural(a, coset=5)
    for k in range(n):
        assert ca[k] == sum(a[j] * pow(5 * pow(w, k, B.R), j, B.R) for j in range(n)) % B.R
    assert B.ntt_natural(fa, inverse=True) == a
    assert B.ntt_natural(ca, inverse=True, coset=5) == a


def test_quotient_h_is_polynomial_division():
    """quotient_h against schoolbook multiplication and exact division by X^n - 1."""
    rng = random.Random(2)
    n = 8
    a = [rng.randrange(B.R) for _ in range(n)]
    b = [rng.randrange(B.R) for _ in range(n)]
    c = [x * y % B.R for x, y in zip(a, b)]
    A, Bc, C = (B.ntt_natural(v, inverse=True) for v in (a, b, c))
    prod = [0] * (2 * n - 1)
    for i, x in enumerate(A):
        for j, y in enumerate(Bc):
            prod[i + j] = (prod[i + j] + x * y) % B.R
    for i, z in enumerate(C):
        prod[i] = (prod[i] - z) % B.R
    # divide by X^n - 1: q[i] = prod[i + n] (+ q[i + n]); remainder must vanish
    q = [0] * (n - 1)
    rem = prod[:]
    for i in range(2 * n - 2, n - 1, -1):
        q[i - n] = rem[i]
        rem[i - n] = (rem[i - n] + rem[i]) % B.R
        rem[i] = 0
    assert all(v == 0 for v in rem)
    h = B.quotient_h(a, b, c)
    assert h[:n - 1] == q and h[n - 1] == 0


def test_msm_and_fixed_base():
    rng = random.Random(3)
    fb = B.g1_fixed_base()
    ks = [rng.randrange(B.R) for _ in range(5)]
    pts = [fb.mul(k) for k in ks]
    assert all(p == B.g1_mul(B.G1_GEN, k) for p, k in zip(pts, ks))
    ss = [rng.randrange(B.R) for _ in range(5)]
    want = B.g1_mul(B.G1_GEN, sum(k * s for k, s in zip(ks, ss)) % B.R)
    assert B.g1_msm(pts, ss) == want

"""The __host__ __device__ field / curve routines of csrc/ff.cuh and csrc/ec.cuh, compiled for the host
(libg16_hostshim.so, test tooling) and checked against Python big integers: the same source the kernels
run (gnark-crypto ecc/bn254/fp, fr: Montgomery R = 2^256, little-endian limbs)."""
import ctypes
import os
import random

import bn254 as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = ctypes.CDLL(os.path.join(ROOT, "shielded_pool_pinocchio_solana_b200", "libg16_hostshim.so"))
U8 = ctypes.c_uint32 * 8
U16 = ctypes.c_uint32 * 16
MONT = 1 << 256


def limbs(x, n=8):
    return (ctypes.c_uint32 * n)(*[(x >> (32 * i)) & 0xFFFFFFFF for i in range(n)])


def val(a):
    return sum(int(w) << (32 * i) for i, w in enumerate(a))


def fop(fn, op, a, b=0):
    out = U8()
    fn(op, limbs(a), limbs(b), out)
    return val(out)


EDGE = [0, 1, 2, 3, (1 << 253), (1 << 128) - 1, (1 << 128), 0xFFFFFFFF, 1 << 32]


def test_field_ops_match_bigint():
    rng = random.Random(11)
    for fn, m in ((SHIM.shim_fp_op, B.P), (SHIM.shim_fr_op, B.R)):
        xs = [e % m for e in EDGE] + [m - 1, m - 2, (m - 1) // 2, (m + 1) // 2] + [rng.randrange(m) for _ in range(200)]
        rinv = B.inv(MONT, m)
        for a in xs:
            b = xs[rng.randrange(len(xs))]
            assert fop(fn, 0, a, b) == a * b * rinv % m            # Montgomery product
            assert fop(fn, 7, a) == a * a * rinv % m               # squaring
            assert fop(fn, 1, a, b) == (a + b) % m
            assert fop(fn, 2, a, b) == (a - b) % m
            assert fop(fn, 3, a) == (-a) % m
            assert fop(fn, 9, a) == a * B.inv(2, m) % m            # halve
            assert fop(fn, 5, a) == a * MONT % m
            assert fop(fn, 6, a) == a * rinv % m


def test_binary_gcd_inverse_equals_fermat_and_bigint():
    """inverse() (fixed-sequence binary GCD on 64-bit approximations) = inverse_euclid() (shift/subtract extended
    Euclid) = inverse_fermat() = the big-integer inverse: for X = aR the result is a^-1 R = R^2 / X."""
    rng = random.Random(12)
    for fn, m in ((SHIM.shim_fp_op, B.P), (SHIM.shim_fr_op, B.R)):
        xs = [e % m for e in EDGE] + [m - 1, m - 2, (m - 1) // 2] + [rng.randrange(1, m) for _ in range(2000)]
        xs += [1 << k for k in range(0, 254, 7)] + [(m - (1 << k)) % m for k in range(0, 254, 11)]
        for a in xs:
            want = 0 if a == 0 else MONT * MONT * B.inv(a, m) % m
            assert fop(fn, 4, a) == want
        for a in xs[:400]:
            want = 0 if a == 0 else MONT * MONT * B.inv(a, m) % m
            assert fop(fn, 8, a) == want
            assert fop(fn, 10, a) == want


def test_fp2_ops():
    rng = random.Random(13)
    rinv = B.inv(MONT, B.P)
    for _ in range(50):
        a = (rng.randrange(B.P), rng.randrange(B.P))
        b = (rng.randrange(B.P), rng.randrange(B.P))
        def call(op, x, y):
            out = U16()
            SHIM.shim_fp2_op(op, limbs(x[0] | (x[1] << 256), 16), limbs(y[0] | (y[1] << 256), 16), out)
            v = val(out)
            return (v & ((1 << 256) - 1), v >> 256)
        # Montgomery: (aR)(bR)/R = abR
        want = B.f2_mul(a, b)
        assert call(0, a, b) == (want[0] * rinv % B.P, want[1] * rinv % B.P)
        sq = B.f2_sqr(a)
        assert call(1, a, a) == (sq[0] * rinv % B.P, sq[1] * rinv % B.P)
        iv = B.f2_inv(a)
        assert call(2, a, a) == (iv[0] * MONT * MONT % B.P, iv[1] * MONT * MONT % B.P)

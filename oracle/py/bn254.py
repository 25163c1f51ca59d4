"""BN254 field / curve arithmetic on Python big integers -- TEST ORACLE ONLY.

This file is test infrastructure (oracle/): only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline leg may import it.  It is never on the product path.

PARITY UNPINNED: the arithmetic being restated lives in gnark-crypto
(github.com/consensys/gnark-crypto, `ecc/bn254/{g1,g2,multiexp}.go`,
`ecc/bn254/fr/fft`), a dependency of gnark v0.14.0 (pinned by the CBOR body of
/root/reference/noir_circuit/target/shielded_pool_verifier.ccs, `GnarkVersion`)
that is NOT vendored in /root/reference.  What is restated here is the published
mathematics (short Weierstrass group law, radix-2 NTT); the constants are pinned
against the reference's committed artifacts in tests/test_oracle_kat.py
(ScalarField in the .ccs, Montgomery coefficient table, every point of both .vk
files on-curve, Fp limbs in client/merkle.ts:48).

Everything is written for obviousness, not speed.
"""

# --- moduli (`.ccs` ScalarField; client/merkle.ts:48) -------------
P = 21888242871839275222246405745257275088696311157297823662689037894645226208583
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
MONT_R = 1 << 256          # Montgomery radix used by gnark-crypto for both fields
B1 = 3                     # G1: y^2 = x^3 + 3
TWO_ADICITY = 28
# gnark-crypto fr root of unity of order 2^28
ROOT_2_28 = 19103219067921713944291392827692070036145651957329286315305642004821462161904
COSET_GEN = 5              # fr.MultiplicativeGen


def inv(a, m):
    return pow(a % m, -1, m)


# --- Fp2 = Fp[u]/(u^2+1), elements are (a0, a1) = a0 + a1*u ----------------------
def f2(a0, a1=0):
    return (a0 % P, a1 % P)

F2_ZERO = (0, 0)
F2_ONE = (1, 0)

def f2_add(a, b): return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)
def f2_sub(a, b): return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)
def f2_neg(a): return ((-a[0]) % P, (-a[1]) % P)
def f2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)
def f2_sqr(a): return f2_mul(a, a)
def f2_muls(a, s): return ((a[0] * s) % P, (a[1] * s) % P)
def f2_inv(a):
    d = inv(a[0] * a[0] + a[1] * a[1], P)
    return ((a[0] * d) % P, (-a[1] * d) % P)

# twist: y^2 = x^3 + 3/(9+u)   (verified against both .vk files)
B2 = f2_mul((3, 0), f2_inv((9, 1)))

G1_GEN = (1, 2)
G2_GEN = (
    (10857046999023057135944570762232829481370756359578518086990519993285655852781,
     11559732032986387107991004021392285783925812861821192530917403151452391805634),
    (8495653923123431417604973247489272438418190587263600148770280649306958101930,
     4082367875863433681332203403145435568316851327593401208105741076214120093531),
)


class _Fp:
    """Field ops table so the same group law serves G1 (Fp) and G2 (Fp2)."""
    zero = 0
    one = 1
    b = B1
    @staticmethod
    def add(a, b): return (a + b) % P
    @staticmethod
    def sub(a, b): return (a - b) % P
    @staticmethod
    def mul(a, b): return (a * b) % P
    @staticmethod
    def neg(a): return (-a) % P
    @staticmethod
    def inv(a): return inv(a, P)
    @staticmethod
    def is_zero(a): return a % P == 0


class _Fp2:
    zero = F2_ZERO
    one = F2_ONE
    b = B2
    add = staticmethod(f2_add)
    sub = staticmethod(f2_sub)
    mul = staticmethod(f2_mul)
    neg = staticmethod(f2_neg)
    inv = staticmethod(f2_inv)
    @staticmethod
    def is_zero(a): return a[0] % P == 0 and a[1] % P == 0


# --- affine group law; None is the point at infinity ------------------------------
def _on_curve(F, pt):
    if pt is None:
        return True
    x, y = pt
    return F.sub(F.mul(y, y), F.add(F.mul(F.mul(x, x), x), F.b)) == F.zero

def _neg(F, pt):
    return None if pt is None else (pt[0], F.neg(pt[1]))

def _add(F, p, q):
    if p is None: return q
    if q is None: return p
    x1, y1 = p
    x2, y2 = q
    if x1 == x2:
        if F.is_zero(F.add(y1, y2)):
            return None
        three_x2 = F.mul(F.add(F.add(F.one, F.one), F.one), F.mul(x1, x1))
        lam = F.mul(three_x2, F.inv(F.add(y1, y1)))
    else:
        lam = F.mul(F.sub(y2, y1), F.inv(F.sub(x2, x1)))
    x3 = F.sub(F.sub(F.mul(lam, lam), x1), x2)
    y3 = F.sub(F.mul(lam, F.sub(x1, x3)), y1)
    return (x3, y3)

# Jacobian (X, Y, Z): x = X/Z^2, y = Y/Z^3.  Z == zero <=> infinity.
def _jac_from_affine(F, p):
    return (F.one, F.one, F.zero) if p is None else (p[0], p[1], F.one)

def _jac_to_affine(F, p):
    X, Y, Z = p
    if F.is_zero(Z):
        return None
    zi = F.inv(Z)
    zi2 = F.mul(zi, zi)
    return (F.mul(X, zi2), F.mul(Y, F.mul(zi2, zi)))

def _jac_double(F, p):
    X, Y, Z = p
    if F.is_zero(Z) or F.is_zero(Y):
        return (F.one, F.one, F.zero)
    A = F.mul(X, X); Bq = F.mul(Y, Y); C = F.mul(Bq, Bq)
    t = F.add(X, Bq)
    D = F.sub(F.sub(F.mul(t, t), A), C); D = F.add(D, D)
    E = F.add(F.add(A, A), A)
    Fq = F.mul(E, E)
    X3 = F.sub(Fq, F.add(D, D))
    C8 = F.add(C, C); C8 = F.add(C8, C8); C8 = F.add(C8, C8)
    Y3 = F.sub(F.mul(E, F.sub(D, X3)), C8)
    Z3 = F.mul(F.add(Y, Y), Z)
    return (X3, Y3, Z3)

def _jac_add(F, p, q):
    X1, Y1, Z1 = p
    X2, Y2, Z2 = q
    if F.is_zero(Z1): return q
    if F.is_zero(Z2): return p
    Z1Z1 = F.mul(Z1, Z1); Z2Z2 = F.mul(Z2, Z2)
    U1 = F.mul(X1, Z2Z2); U2 = F.mul(X2, Z1Z1)
    S1 = F.mul(Y1, F.mul(Z2, Z2Z2)); S2 = F.mul(Y2, F.mul(Z1, Z1Z1))
    if U1 == U2:
        if S1 == S2:
            return _jac_double(F, p)
        return (F.one, F.one, F.zero)
    H = F.sub(U2, U1); Rr = F.sub(S2, S1)
    HH = F.mul(H, H); HHH = F.mul(H, HH); V = F.mul(U1, HH)
    X3 = F.sub(F.sub(F.mul(Rr, Rr), HHH), F.add(V, V))
    Y3 = F.sub(F.mul(Rr, F.sub(V, X3)), F.mul(S1, HHH))
    Z3 = F.mul(F.mul(Z1, Z2), H)
    return (X3, Y3, Z3)

def _mul(F, pt, k, order=R):
    """[k]pt by double-and-add on Jacobian coordinates; k is reduced mod `order`."""
    k %= order
    acc = (F.one, F.one, F.zero)
    if pt is None or k == 0:
        return None
    base = _jac_from_affine(F, pt)
    for bit in bin(k)[2:]:
        acc = _jac_double(F, acc)
        if bit == '1':
            acc = _jac_add(F, acc, base)
    return _jac_to_affine(F, acc)

def _msm(F, points, scalars):
    """Naive sum_i [s_i]P_i -- the definition, not an algorithm."""
    acc = (F.one, F.one, F.zero)
    for pt, s in zip(points, scalars):
        s %= R
        if pt is None or s == 0:
            continue
        t = _mul(F, pt, s)
        acc = _jac_add(F, acc, _jac_from_affine(F, t))
    return _jac_to_affine(F, acc)


class FixedBase:
    """Windowed fixed-base multiplier (8-bit windows) -- used by the oracle setup."""
    def __init__(self, F, base, bits=254, w=8):
        self.F, self.w = F, w
        self.nwin = (bits + w - 1) // w
        self.table = []
        cur = base
        for _ in range(self.nwin):
            row = [None]
            acc = None
            for _ in range((1 << w) - 1):
                acc = _add(F, acc, cur)
                row.append(acc)
            self.table.append(row)
            cur = _add(F, acc, cur)      # 2^w * cur
    def mul(self, k):
        F = self.F
        k %= R
        acc = (F.one, F.one, F.zero)
        i = 0
        while k:
            d = k & ((1 << self.w) - 1)
            if d:
                acc = _jac_add(F, acc, _jac_from_affine(F, self.table[i][d]))
            k >>= self.w
            i += 1
        return _jac_to_affine(F, acc)


# public names
def g1_on_curve(p): return _on_curve(_Fp, p)
def g2_on_curve(p): return _on_curve(_Fp2, p)
def g1_add(p, q): return _add(_Fp, p, q)
def g2_add(p, q): return _add(_Fp2, p, q)
def g1_neg(p): return _neg(_Fp, p)
def g2_neg(p): return _neg(_Fp2, p)
def g1_mul(p, k): return _mul(_Fp, p, k)
def g2_mul(p, k): return _mul(_Fp2, p, k)
def g1_msm(points, scalars): return _msm(_Fp, points, scalars)
def g2_msm(points, scalars): return _msm(_Fp2, points, scalars)
def g1_fixed_base(base=G1_GEN): return FixedBase(_Fp, base)
def g2_fixed_base(base=G2_GEN): return FixedBase(_Fp2, base)


# --- Fr NTT (gnark-crypto fr/fft semantics) -----------------
def fr_root(n):
    """Generator of the size-n domain: ROOT_2_28^(2^28/n)."""
    assert n & (n - 1) == 0 and n <= (1 << TWO_ADICITY)
    return pow(ROOT_2_28, (1 << TWO_ADICITY) // n, R)

def bitrev(i, logn):
    r = 0
    for _ in range(logn):
        r = (r << 1) | (i & 1)
        i >>= 1
    return r

def bitrev_permute(a):
    n = len(a)
    logn = n.bit_length() - 1
    return [a[bitrev(i, logn)] for i in range(n)]

def ntt_natural(a, inverse=False, coset=None):
    """Evaluate (or interpolate) in NATURAL order, O(n log n), no in-place tricks.

    forward: out[k] = sum_j a[j] * (g*w^k)^j   (g = coset shift, default 1)
    inverse: the exact inverse map.
    """
    n = len(a)
    w = fr_root(n)
    if inverse:
        w = inv(w, R)
    a = list(a)
    if coset is not None and not inverse:
        g, acc = coset % R, 1
        for j in range(n):
            a[j] = a[j] * acc % R
            acc = acc * g % R
    out = _ntt_rec(a, w)
    if inverse:
        ninv = inv(n, R)
        out = [x * ninv % R for x in out]
        if coset is not None:
            gi, acc = inv(coset, R), 1
            for j in range(n):
                out[j] = out[j] * acc % R
                acc = acc * gi % R
    return out

def _ntt_rec(a, w):
    n = len(a)
    if n == 1:
        return a
    ev = _ntt_rec(a[0::2], w * w % R)
    od = _ntt_rec(a[1::2], w * w % R)
    out = [0] * n
    t = 1
    h = n // 2
    for k in range(h):
        x = t * od[k] % R
        out[k] = (ev[k] + x) % R
        out[k + h] = (ev[k] - x) % R
        t = t * w % R
    return out


def to_mont(x, m): return (x * MONT_R) % m
def from_mont(x, m): return (x * inv(MONT_R, m)) % m


# --- Groth16 quotient (gnark backend/groth16/bn254/prove.go `computeH`) ----
def quotient_h(a, b, c):
    """H = (A*B - C) / (X^n - 1) where A,B,C interpolate the rows a,b,c on the size-n domain.

    Returns the n coefficients of H in NATURAL order (coefficient n-1 is always zero).
    Restated from the definition: evaluate on the coset g*H, divide by the constant
    Z(g w^k) = g^n - 1, interpolate back.
    """
    n = len(a)
    ea = ntt_natural(ntt_natural(a, inverse=True), coset=COSET_GEN)
    eb = ntt_natural(ntt_natural(b, inverse=True), coset=COSET_GEN)
    ec = ntt_natural(ntt_natural(c, inverse=True), coset=COSET_GEN)
    den = inv(pow(COSET_GEN, n, R) - 1, R)
    eh = [(x * y - z) * den % R for x, y, z in zip(ea, eb, ec)]
    return ntt_natural(eh, inverse=True, coset=COSET_GEN)

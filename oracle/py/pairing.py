"""Optimal-ate pairing on BN254, written for obviousness -- TEST ORACLE ONLY (see bn254.py).

Stands in for the pairing check `sunspot verify` / the on-chain `sol_alt_bn128_group_op`
verifier perform (/root/reference/noir_circuit/prove_linux.sh:87,
/root/reference/audit_circuit/target/audit_verifier.so).  Fp12 is the plain
polynomial ring Fp[w]/(w^12 - 18 w^6 + 82) (so that Fp2's u = w^6 - 9), arithmetic is
schoolbook; a pairing takes about a second.
"""
from bn254 import P, R, inv

ATE_LOOP_COUNT = 29793968203157093288          # 6x + 2, x = 4965661367192848881
LOG_ATE_LOOP_COUNT = 63
_MOD_COEFFS = [82, 0, 0, 0, 0, 0, -18, 0, 0, 0, 0, 0]   # w^12 = 18 w^6 - 82


class FQ12:
    __slots__ = ("c",)

    def __init__(self, c):
        self.c = [x % P for x in c]
        assert len(self.c) == 12

    @staticmethod
    def one(): return FQ12([1] + [0] * 11)
    @staticmethod
    def zero(): return FQ12([0] * 12)
    @staticmethod
    def scalar(x): return FQ12([x] + [0] * 11)

    def __add__(self, o): return FQ12([a + b for a, b in zip(self.c, o.c)])
    def __sub__(self, o): return FQ12([a - b for a, b in zip(self.c, o.c)])
    def __neg__(self): return FQ12([-a for a in self.c])
    def __eq__(self, o): return self.c == o.c

    def __mul__(self, o):
        if isinstance(o, int):
            return FQ12([a * o for a in self.c])
        b = [0] * 23
        for i, x in enumerate(self.c):
            if x:
                for j, y in enumerate(o.c):
                    b[i + j] += x * y
        for exp in range(22, 11, -1):
            top = b[exp]
            if top:
                b[exp] = 0
                b[exp - 6] += 18 * top
                b[exp - 12] -= 82 * top
        return FQ12(b[:12])

    def inv(self):
        """Extended Euclid on polynomials over Fp."""
        lm, hm = [1] + [0] * 12, [0] * 13
        low, high = self.c + [0], _MOD_COEFFS + [1]
        def deg(p):
            d = len(p) - 1
            while d and p[d] % P == 0:
                d -= 1
            return d
        def poly_div(a, b):
            dega, degb = deg(a), deg(b)
            temp = list(a)
            o = [0] * len(a)
            for i in range(dega - degb, -1, -1):
                q = temp[degb + i] * inv(b[degb], P) % P
                o[i] = (o[i] + q) % P
                for c in range(degb + 1):
                    temp[c + i] = (temp[c + i] - q * b[c]) % P
            return o[:deg(o) + 1]
        while deg(low):
            r = poly_div(high, low)
            r += [0] * (13 - len(r))
            nm, new = list(hm), list(high)
            for i in range(13):
                for j in range(13 - i):
                    nm[i + j] -= lm[i] * r[j]
                    new[i + j] -= low[i] * r[j]
            nm = [x % P for x in nm]
            new = [x % P for x in new]
            lm, low, hm, high = nm, new, lm, low
        li = inv(low[0], P)
        return FQ12([x * li for x in lm[:12]])

    def __truediv__(self, o): return self * o.inv()

    def __pow__(self, e):
        acc, base = FQ12.one(), self
        while e:
            if e & 1:
                acc = acc * base
            base = base * base
            e >>= 1
        return acc


_W = FQ12([0, 1] + [0] * 10)
_W2, _W3 = _W * _W, _W * _W * _W


def _twist(q):
    """G2 point over Fp2 -> point on y^2 = x^3 + 3 over Fp12."""
    (x0, x1), (y0, y1) = q
    nx = FQ12([x0 - 9 * x1] + [0] * 5 + [x1] + [0] * 5)
    ny = FQ12([y0 - 9 * y1] + [0] * 5 + [y1] + [0] * 5)
    return (nx * _W2, ny * _W3)


def _cast_g1(p):
    return (FQ12.scalar(p[0]), FQ12.scalar(p[1]))


def _double(pt):
    x, y = pt
    m = (x * x * 3) / (y * 2)
    nx = m * m - x * 2
    return (nx, m * (x - nx) - y)


def _add(p1, p2):
    if p1 is None: return p2
    if p2 is None: return p1
    x1, y1 = p1
    x2, y2 = p2
    if x1 == x2:
        return _double(p1) if y1 == y2 else None
    m = (y2 - y1) / (x2 - x1)
    nx = m * m - x1 - x2
    return (nx, m * (x1 - nx) - y1)


def _linefunc(p1, p2, t):
    x1, y1 = p1
    x2, y2 = p2
    xt, yt = t
    if not x1 == x2:
        m = (y2 - y1) / (x2 - x1)
        return m * (xt - x1) - (yt - y1)
    if y1 == y2:
        m = (x1 * x1 * 3) / (y1 * 2)
        return m * (xt - x1) - (yt - y1)
    return xt - x1


def miller_loop(q, p):
    """q in G2 (affine over Fp2), p in G1 (affine); None = infinity -> 1."""
    if q is None or p is None:
        return FQ12.one()
    Q, Pt = _twist(q), _cast_g1(p)
    Rr, f = Q, FQ12.one()
    for i in range(LOG_ATE_LOOP_COUNT, -1, -1):
        f = f * f * _linefunc(Rr, Rr, Pt)
        Rr = _double(Rr)
        if ATE_LOOP_COUNT & (1 << i):
            f = f * _linefunc(Rr, Q, Pt)
            Rr = _add(Rr, Q)
    Q1 = (Q[0] ** P, Q[1] ** P)
    nQ2 = (Q1[0] ** P, -(Q1[1] ** P))
    f = f * _linefunc(Rr, Q1, Pt)
    Rr = _add(Rr, Q1)
    f = f * _linefunc(Rr, nQ2, Pt)
    return f


def final_exponentiation(f):
    return f ** ((P ** 12 - 1) // R)


def pairing(q, p):
    return final_exponentiation(miller_loop(q, p))


def pairing_product_is_one(pairs):
    """prod e(P_i, Q_i) == 1 for pairs (P_i in G1, Q_i in G2); one shared final exponentiation."""
    f = FQ12.one()
    for p, q in pairs:
        f = f * miller_loop(q, p)
    return final_exponentiation(f) == FQ12.one()

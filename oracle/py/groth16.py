"""Groth16 over BN254 exactly as gnark v0.14.0 runs it -- TEST ORACLE ONLY (see bn254.py).

CPU restatement of the algorithm behind `sunspot prove <acir> <witness.gz> <ccs> <pk>`
(/root/reference/client/proof.helper.ts:58-66, /root/reference/noir_circuit/prove_linux.sh:83),
i.e. gnark `backend/groth16/bn254/{setup,prove,verify,marshal}.go`, `constraint/bn254/solver.go`,
`constraint/solver/hint.go`, gnark-crypto `fr/pedersen`, `fr/hash_to_field` (third-party Go,
not in /root/reference; versions pinned by the `.ccs` artifact -- SURVEY.md 8c).

PARITY UNPINNED: the reference tree holds no proving key, witness or proof
(/root/reference/.MISSING_LARGE_BLOBS, /root/reference/.gitignore:51-54), so byte equality with
gnark itself cannot be checked here.  What IS pinned by the reference (tests/test_oracle_kat.py):
the `.ccs` decoding, the `.vk` byte layout (both committed .vk files re-serialise identically),
the 388-byte proof / 12+32n-byte public-witness framing
(shielded_pool_program/src/instructions/withdraw.rs:13-16, submit_audit.rs:18-21) and the `.pw`
built from client/prover-params.toml.  Soundness of the restatement is checked by pairing
verification and by the closed-form check against the known toxic waste.
"""
import hashlib
import struct

import bn254 as B
from bn254 import R, P, inv
import ccs as ccsmod
import serialize as S

COMMITMENT_DST = b"bsb22-commitment"      # gnark constraint.CommitmentDst
FOLD_DST = b"G16-BSB22"


# ------------------------------------------------------------------------------------------
# RFC 9380 expand_message_xmd (SHA-256) and gnark-crypto fr.Hash
# ------------------------------------------------------------------------------------------
def expand_message_xmd(msg, dst, length):
    h = hashlib.sha256
    b_in_bytes, r_in_bytes = 32, 64
    ell = (length + b_in_bytes - 1) // b_in_bytes
    assert ell <= 255 and len(dst) <= 255
    dst_prime = dst + bytes([len(dst)])
    b0 = h(b"\x00" * r_in_bytes + msg + struct.pack(">H", length) + b"\x00" + dst_prime).digest()
    b1 = h(b0 + b"\x01" + dst_prime).digest()
    out, prev = b1, b1
    for i in range(2, ell + 1):
        prev = h(bytes(x ^ y for x, y in zip(b0, prev)) + bytes([i]) + dst_prime).digest()
        out += prev
    return out[:length]


def hash_to_field(msg, dst, count=1):
    """gnark-crypto fr.Hash: 48 pseudo-random bytes per element, big-endian, reduced mod r."""
    L = 48
    raw = expand_message_xmd(msg, dst, L * count)
    return [int.from_bytes(raw[L * i:L * (i + 1)], "big") % R for i in range(count)]


# ------------------------------------------------------------------------------------------
# hints (gnark std)
# ------------------------------------------------------------------------------------------
class HintError(Exception):
    pass


def _hint_nbits(ins, nout):
    return [(ins[0] >> i) & 1 for i in range(nout)]


def _hint_invzero(ins, nout):
    return [0 if ins[0] % R == 0 else inv(ins[0], R)]


def _hint_decompose(ins, nout):
    # rangecheck.DecomposeHint(varSize, limbSize, value)
    limb = ins[1]
    return [(ins[2] >> (i * limb)) & ((1 << limb) - 1) for i in range(nout)]


def _hint_count(ins, nout, tolerant=False):
    # logderivarg.countHint(tableSize, nbCols, table rows..., query rows...)
    size, cols = ins[0], ins[1]
    table = [tuple(ins[2 + r * cols:2 + (r + 1) * cols]) for r in range(size)]
    q0 = 2 + size * cols
    index = {}
    for j, row in enumerate(table):
        index.setdefault(row, j)
    out = [0] * nout
    for k in range(q0, len(ins), cols):
        row = tuple(ins[k:k + cols])
        if row not in index:
            if tolerant:
                continue             # diagnostic solve: count what is there
            raise HintError("countHint: query not in table")
        out[index[row]] += 1
    return out


def _recompose(limbs, nbits):
    return sum(int(v) << (nbits * i) for i, v in enumerate(limbs))


def _decompose_exact(x, nbits, n, what):
    if x < 0 or x >> (nbits * n):
        raise HintError("%s does not fit %d limbs of %d bits" % (what, n, nbits))
    return [(x >> (nbits * i)) & ((1 << nbits) - 1) for i in range(n)]


def _hint_emulated_mul(ins, nout):
    """gnark std/math/emulated mulHint (v0.14): inputs nbBits, nbLimbs, len(a), len(quo), modulus limbs, a limbs,
    b limbs; outputs quo | rem | carries with  a(X) b(X) = rem(X) + quo(X) p(X) + (2^nbBits - X) carry(X)
    over the integers -- the identity gnark's deferred multiplication check evaluates at the commitment challenge.
    Quotient, remainder and carries are determined by that identity, so there is nothing to pin beyond it."""
    nbits, nlimbs, na, nquo = (int(v) for v in ins[:4])
    pl = ins[4:4 + nlimbs]
    al = ins[4 + nlimbs:4 + nlimbs + na]
    bl = ins[4 + nlimbs + na:]
    nb = len(bl)
    if nb < 1 or nlimbs < 1:
        raise HintError("emulated.mulHint: malformed inputs")
    ncarry = max(na + nb - 1, nquo + nlimbs - 1) - 1
    nrem = nout - nquo - ncarry
    if nrem not in (0, nlimbs):
        raise HintError("emulated.mulHint: %d outputs do not match quo %d + rem + carries %d" % (nout, nquo, ncarry))
    p, a, b = _recompose(pl, nbits), _recompose(al, nbits), _recompose(bl, nbits)
    if p == 0:
        raise HintError("emulated.mulHint: zero modulus")
    quo, rem = divmod(a * b, p)
    if nrem == 0 and rem:
        raise HintError("emulated.mulHint: product is not a multiple of the modulus")
    ql = _decompose_exact(quo, nbits, nquo, "quotient")
    rl = _decompose_exact(rem, nbits, nrem, "remainder") if nrem else []
    xp = [0] * (na + nb - 1)
    yp = [0] * (nquo + nlimbs - 1)
    for i in range(na):
        for j in range(nb):
            xp[i + j] += int(al[i]) * int(bl[j])
    for i in range(nlimbs):
        if i < nrem:
            yp[i] += rl[i]
        for j in range(nquo):
            yp[i + j] += ql[j] * int(pl[i])
    carries, carry = [], 0
    for i in range(ncarry):
        if i < len(xp):
            carry += xp[i]
        if i < len(yp):
            carry -= yp[i]
        carry >>= nbits          # exact: the low limb of the difference cancels
        carries.append(carry % R)
    return ql + rl + carries


# Grumpkin (y^2 = x^3 - 17 over BN254 Fr): scalar field = BN254 Fq; LAMBDA = the cube root of unity the
# withdraw circuit multiplies s2 by (the constants 6296954981786320894, 15344436770043026511, 6476857749317913516
# of instruction 20 of shielded_pool_verifier.ccs); B1, B2 = a reduced basis of {(u, v): u = LAMBDA v mod q}.
GRUMPKIN_Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583
GRUMPKIN_LAMBDA = 2203960485148121921418603742825762020974279258880205651966
GRUMPKIN_B1 = (9931322734385697762, 147946756881789319000765030803803410729)
GRUMPKIN_B2 = (147946756881789319010696353538189108491, -9931322734385697762)


def glv_split_nonneg(s, bits=127):
    """(s1, s2) with s1 = s + LAMBDA*s2 (mod q) and 0 <= s1, s2 < 2^bits -- the relation and the range the withdraw
    circuit enforces on the outputs of sunspot's `sw-grumpkin.decomposeScalar` (instructions 18-21, 150 of the
    .ccs).  The box usually holds more than one such pair; sunspot's own choice is not observable offline, so this
    restatement takes the pair with the smallest max(s1, s2), then the smallest s2 (UNPINNED choice: every pair
    satisfies the circuit, but only sunspot's gives sunspot's wire values)."""
    (a1, b1), (a2, b2) = GRUMPKIN_B1, GRUMPKIN_B2
    det = a1 * b2 - b1 * a2
    half = 1 << (bits - 1)
    tx, ty = s - half, -half                      # (s, 0) - centre of the box, in basis coordinates (floored)
    c1 = (tx * b2 - ty * a2) // det
    c2 = (ty * a1 - tx * b1) // det
    best = None
    for d1 in range(-2, 4):
        for d2 in range(-2, 4):
            x = s - (c1 + d1) * a1 - (c2 + d2) * a2
            y = -(c1 + d1) * b1 - (c2 + d2) * b2
            if 0 <= x < (1 << bits) and 0 <= y < (1 << bits):
                key = (max(x, y), y)
                if best is None or key < best[0]:
                    best = (key, x, y)
    if best is None:
        raise HintError("decomposeScalar: no decomposition in range")
    return best[1], best[2]


def _hint_grumpkin_decompose_scalar(ins, nout):
    # generic emulated-hint framing: counts (native in, emulated in x2, emulated out, ...), the native scalar,
    # then nbLimbs, nbBits and the limbs of the emulated modulus
    if len(ins) < 9 or int(ins[0]) != 1 or int(ins[3]) * int(ins[7]) != nout:
        raise HintError("sw-grumpkin.decomposeScalar: unexpected input framing")
    s, nlimbs, nbits = int(ins[6]), int(ins[7]), int(ins[8])
    if _recompose(ins[9:9 + nlimbs], nbits) != GRUMPKIN_Q or int(ins[3]) != 2:
        raise HintError("sw-grumpkin.decomposeScalar: not the Grumpkin scalar field")
    s1, s2 = glv_split_nonneg(s % GRUMPKIN_Q)
    return _decompose_exact(s1, nbits, nlimbs, "s1") + _decompose_exact(s2, nbits, nlimbs, "s2")


def _hint_grumpkin_decompose(ins, nout):
    # the native scalar as nout limbs of 64 bits (its representation in the emulated field)
    return _decompose_exact(int(ins[0]), 64, nout, "scalar")


HINTS = {
    "github.com/consensys/gnark/std/math/emulated.mulHint": _hint_emulated_mul,
    "sunspot/go/sw-grumpkin.decomposeScalar": _hint_grumpkin_decompose_scalar,
    "sunspot/go/sw-grumpkin.decompose": _hint_grumpkin_decompose,
    "github.com/consensys/gnark/std/math/bits.nBits": _hint_nbits,
    "github.com/consensys/gnark/constraint/solver.InvZeroHint": _hint_invzero,
    "github.com/consensys/gnark/std/rangecheck.DecomposeHint": _hint_decompose,
    "github.com/consensys/gnark/std/internal/logderivarg.countHint": _hint_count,
}
HINT_RANDOMIZE = "github.com/consensys/gnark/internal/hints.Randomize"
HINT_COMMIT = "github.com/consensys/gnark/frontend/cs.Bsb22CommitmentComputePlaceholder"


# ------------------------------------------------------------------------------------------
# solver (gnark constraint/bn254/solver.go)
# ------------------------------------------------------------------------------------------
class Unsatisfied(Exception):
    pass


def solve(c, assignment, pk=None, blinder=None, failed_rows=None):
    """Extends `assignment` (public values without the ONE wire, then secret values) to every wire.

    Returns (wires, commitments) with commitments = list of (G1 point, [committed values]).
    `failed_rows` (a list) switches to diagnostic mode: unsatisfied rows are collected instead of raised.
    """
    npub, nsec = c.nb_public, c.nb_secret
    if len(assignment) != npub - 1 + nsec:
        raise ValueError("assignment has %d values, circuit wants %d" % (len(assignment), npub - 1 + nsec))
    w = [None] * c.nb_wires
    w[0] = 1
    for i, v in enumerate(assignment):
        w[1 + i] = v % R
    names = c.body["MHintsDependencies"]
    infos = c.commitments
    commitments = []

    def term(coeff_id, wire):
        if wire == ccsmod.CONST_WIRE:
            return c.coeffs[coeff_id]
        if w[wire] is None:
            raise Unsatisfied("wire %d used before it is solved" % wire)
        return c.coeffs[coeff_id] * w[wire] % R

    def lin(expr):
        return sum(term(cid, wid) for cid, wid in expr) % R

    for level in c.levels:
        for i in level:
            if c.blueprint[i] == ccsmod.BLUEPRINT_HINT:
                hid, ins, o0, o1 = c.hint(i)
                name = names[hid]
                vals = [lin(e) for e in ins]
                if name == HINT_RANDOMIZE:
                    if blinder is None:
                        raise HintError("hints.Randomize needs an injected blinder")
                    outs = [blinder % R] * (o1 - o0)
                elif name == HINT_COMMIT:
                    info = infos[len(commitments)]
                    npc = len(info["PublicAndCommitmentCommitted"])
                    hashed, committed = vals[1:1 + npc], vals[1 + npc:]
                    key = pk["commitment_keys"][len(commitments)]
                    pt = B.g1_msm(key["basis"], committed)
                    msg = S.g1_to_bytes(pt) + b"".join(S.fr_to_bytes(x) for x in hashed)
                    outs = [hash_to_field(msg, COMMITMENT_DST)[0]]
                    commitments.append((pt, committed))
                elif name in HINTS:
                    try:
                        if failed_rows is not None and HINTS[name] is _hint_count:
                            outs = _hint_count(vals, o1 - o0, tolerant=True)
                        else:
                            outs = HINTS[name](vals, o1 - o0)
                    except HintError:
                        if failed_rows is None:
                            raise
                        outs = [0] * (o1 - o0)          # diagnostic mode: keep going (e.g. a lookup of a garbage value)
                else:
                    raise HintError("hint %s is not restated (semantics unknown offline)" % name)
                for k, v in enumerate(outs):
                    w[o0 + k] = v % R
            else:
                L, Rr, O = c.r1c(i)
                unknown = None
                for side, expr in enumerate((L, Rr, O)):
                    for cid, wid in expr:
                        if wid != ccsmod.CONST_WIRE and w[wid] is None:
                            if unknown is not None and unknown != (side, cid, wid):
                                raise Unsatisfied("row %d has two unknown wires" % c.constraint_offset[i])
                            unknown = (side, cid, wid)
                if unknown is not None:
                    side, cid, wid = unknown
                    w[wid] = 0
                    a, b, cc = lin(L), lin(Rr), lin(O)
                    coeff = c.coeffs[cid]
                    if side == 2:
                        val = (a * b - cc) * inv(coeff, R)
                    elif side == 0:
                        # gnark solveR1C: a zero divisor leaves the wire at 0 and only checks a*b == c
                        # (the DivUnchecked(0, 0) = 0 convention); the row check below rejects c != 0
                        val = 0 if b == 0 else (cc * inv(b, R) - a) * inv(coeff, R)
                    else:
                        val = 0 if a == 0 else (cc * inv(a, R) - b) * inv(coeff, R)
                    w[wid] = val % R
                if lin(L) * lin(Rr) % R != lin(O):
                    if failed_rows is None:
                        raise Unsatisfied("constraint #%d is not satisfied" % c.constraint_offset[i])
                    failed_rows.append(c.constraint_offset[i])
    if any(v is None for v in w):
        raise Unsatisfied("unsolved wires remain")
    return w, commitments


def evaluate_abc(c, w):
    """a = A.w, b = B.w, c = C.w over the rows, zero-padded to the domain size."""
    n = domain_size(c)
    a, b, cc = [0] * n, [0] * n, [0] * n
    def lin(expr):
        return sum(c.coeffs[cid] * (1 if wid == ccsmod.CONST_WIRE else w[wid]) for cid, wid in expr) % R
    for k, (L, Rr, O) in enumerate(c.rows()):
        a[k], b[k], cc[k] = lin(L), lin(Rr), lin(O)
    return a, b, cc


def domain_size(c):
    n = 1
    while n < c.nb_constraints:
        n <<= 1
    return n


# ------------------------------------------------------------------------------------------
# setup with known toxic waste (gnark backend/groth16/bn254/setup.go)
# ------------------------------------------------------------------------------------------
def toxic_from_seed(seed):
    names = ["tau", "alpha", "beta", "gamma", "delta", "sigma", "g2k"]
    out = {}
    for nme in names:
        v = 0
        ctr = 0
        while v == 0:
            v = int.from_bytes(hashlib.sha256(b"g16b200/setup/%s/%s/%d" % (nme.encode(), seed, ctr)).digest() +
                               hashlib.sha256(b"g16b200/setup2/%s/%s/%d" % (nme.encode(), seed, ctr)).digest(),
                               "big") % R
            ctr += 1
        out[nme] = v
    return out


def wire_polys_at_tau(c, tau):
    """A_i(tau), B_i(tau), C_i(tau) for every wire i."""
    n = domain_size(c)
    w = B.fr_root(n)
    # Lagrange basis at tau: L_k(tau) = (tau^n - 1) * w^k / (n * (tau - w^k))
    zt = (pow(tau, n, R) - 1) % R
    ninv = inv(n, R)
    lag = []
    wk = 1
    for k in range(c.nb_constraints):
        lag.append(zt * wk % R * ninv % R * inv(tau - wk, R) % R)
        wk = wk * w % R
    nw = c.nb_wires
    A, Bv, C = [0] * nw, [0] * nw, [0] * nw
    for k, (L, Rr, O) in enumerate(c.rows()):
        for dst, expr in ((A, L), (Bv, Rr), (C, O)):
            for cid, wid in expr:
                wi = 0 if wid == ccsmod.CONST_WIRE else wid
                dst[wi] = (dst[wi] + c.coeffs[cid] * lag[k]) % R
    return A, Bv, C


class _FastFixedBase:
    """[k]G for lists of k through the C oracle (oracle/c, OpenMP) -- same results as bn254.FixedBase, used for
    circuit sizes where the pure-Python table walk would take minutes (audit_like: ~170 K points)."""
    def __init__(self, group):
        import coracle
        self.group, self.co = group, coracle
        self.base = S.g1_to_bytes(B.G1_GEN) if group == "g1" else S.g2_to_bytes(B.G2_GEN)
        self.dec = S.g1_from_bytes if group == "g1" else S.g2_from_bytes
    def muls(self, ks):
        return [self.dec(b) for b in self.co.fixed_base(self.base, list(ks), self.group)]
    def mul(self, k):
        return self.muls([k])[0]


class _PyFixedBase:
    def __init__(self, fb):
        self.fb = fb
    def muls(self, ks):
        return [self.fb.mul(k) for k in ks]
    def mul(self, k):
        return self.fb.mul(k)


def setup(c, seed=b"oracle", fast=False):
    """-> (pk, vk, toxic).  Points are affine tuples / None; layout mirrors gnark's ProvingKey.
    fast=True computes the key points with the C oracle's fixed-base multiplier."""
    tx = toxic_from_seed(seed)
    tau, alpha, beta, gamma, delta = (tx[k] for k in ("tau", "alpha", "beta", "gamma", "delta"))
    n = domain_size(c)
    A, Bv, C = wire_polys_at_tau(c, tau)
    if fast:
        fb1, fb2 = _FastFixedBase("g1"), _FastFixedBase("g2")
    else:
        fb1, fb2 = _PyFixedBase(B.g1_fixed_base()), _PyFixedBase(B.g2_fixed_base())
    nw, npub = c.nb_wires, c.nb_public
    infos = c.commitments
    committed_private = [list(info["PrivateCommitted"]) for info in infos]
    commitment_wires = [info["CommitmentIndex"] for info in infos]
    in_commit = {wid for lst in committed_private for wid in lst}
    ginv, dinv = inv(gamma, R), inv(delta, R)

    inf_a = [x == 0 for x in A]
    inf_b = [x == 0 for x in Bv]
    pk = {
        "domain": n,
        "alpha1": fb1.mul(alpha), "beta1": fb1.mul(beta), "delta1": fb1.mul(delta),
        "beta2": fb2.mul(beta), "delta2": fb2.mul(delta),
        "A": fb1.muls([x for x in A if x]), "B1": fb1.muls([x for x in Bv if x]),
        "B2": fb2.muls([x for x in Bv if x]),
        "infinity_a": inf_a, "infinity_b": inf_b, "nb_wires": nw,
    }
    t = [(beta * A[i] + alpha * Bv[i] + C[i]) % R for i in range(nw)]
    k_private = [i for i in range(npub, nw) if i not in in_commit and i not in commitment_wires]
    pk["K"] = fb1.muls([t[i] * dinv % R for i in k_private])
    pk["k_wires"] = k_private
    zdt = (pow(tau, n, R) - 1) * dinv % R
    z_nat = []
    acc = zdt
    for _ in range(n):
        z_nat.append(acc)
        acc = acc * tau % R
    logn = n.bit_length() - 1
    z_br = [z_nat[B.bitrev(p, logn)] for p in range(n)]
    pk["Z"] = fb1.muls(z_br[:n - 1])      # bit-reversed order, n-1 entries
    vk_wires = list(range(npub)) + commitment_wires
    vk = {
        "alpha1": pk["alpha1"], "beta1": pk["beta1"], "beta2": pk["beta2"],
        "gamma2": fb2.mul(gamma), "delta1": pk["delta1"], "delta2": pk["delta2"],
        "K": fb1.muls([t[i] * ginv % R for i in vk_wires]),
        "public_and_commitment_committed": [list(info["PublicAndCommitmentCommitted"]) for info in infos],
    }
    # Pedersen keys (gnark-crypto fr/pedersen Setup): basis = gamma-scaled K of the committed wires
    sigma = tx["sigma"]
    g2 = fb2.mul(tx["g2k"])
    keys = []
    for lst in committed_private:
        basis = fb1.muls([t[i] * ginv % R for i in lst])
        keys.append({"basis": basis, "basis_exp_sigma": fb1.muls([t[i] * ginv % R * sigma % R for i in lst])})
    pk["commitment_keys"] = keys
    vk["commitment_keys"] = [{"g": g2, "g_sigma_neg": B.g2_mul(g2, (-sigma) % R)} for _ in keys]
    return pk, vk, tx


# ------------------------------------------------------------------------------------------
# serialisation (gnark marshal.go, raw/uncompressed)
# ------------------------------------------------------------------------------------------
def _g1_slice(pts): return struct.pack(">I", len(pts)) + b"".join(S.g1_to_bytes(p) for p in pts)
def _g2_slice(pts): return struct.pack(">I", len(pts)) + b"".join(S.g2_to_bytes(p) for p in pts)


def write_pk(pk):
    """ProvingKey.WriteRawTo layout as recalled in SURVEY.md 9.2 (UNVERIFIED against a real .pk)."""
    n = pk["domain"]
    w = B.fr_root(n)
    out = struct.pack(">Q", n)
    for x in (inv(n, R), w, inv(w, R), B.COSET_GEN, inv(B.COSET_GEN, R)):
        out += S.fr_to_bytes(x)
    out += b"\x00"                                    # withPrecompute = false
    out += S.g1_to_bytes(pk["alpha1"]) + S.g1_to_bytes(pk["beta1"]) + S.g1_to_bytes(pk["delta1"])
    out += _g1_slice(pk["A"]) + _g1_slice(pk["B1"]) + _g1_slice(pk["Z"]) + _g1_slice(pk["K"])
    out += S.g2_to_bytes(pk["beta2"]) + S.g2_to_bytes(pk["delta2"]) + _g2_slice(pk["B2"])
    nw = pk["nb_wires"]
    out += struct.pack(">QQQ", nw, sum(pk["infinity_a"]), sum(pk["infinity_b"]))
    out += struct.pack(">I", nw) + bytes(int(x) for x in pk["infinity_a"])
    out += struct.pack(">I", nw) + bytes(int(x) for x in pk["infinity_b"])
    out += struct.pack(">I", len(pk["commitment_keys"]))
    for key in pk["commitment_keys"]:
        out += _g1_slice(key["basis"]) + _g1_slice(key["basis_exp_sigma"])
    return out


def read_pk(buf, c):
    """Inverse of write_pk (needs the Ccs for the wire bookkeeping gnark recomputes at prove time)."""
    off = 0
    def u32():
        nonlocal off
        v = struct.unpack_from(">I", buf, off)[0]; off += 4; return v
    def u64():
        nonlocal off
        v = struct.unpack_from(">Q", buf, off)[0]; off += 8; return v
    def g1():
        nonlocal off
        p = S.g1_from_bytes(buf[off:off + 64]); off += 64; return p
    def g2():
        nonlocal off
        p = S.g2_from_bytes(buf[off:off + 128]); off += 128; return p
    pk = {"domain": u64()}
    off += 5 * 32 + 1
    pk["alpha1"], pk["beta1"], pk["delta1"] = g1(), g1(), g1()
    pk["A"] = [g1() for _ in range(u32())]
    pk["B1"] = [g1() for _ in range(u32())]
    pk["Z"] = [g1() for _ in range(u32())]
    pk["K"] = [g1() for _ in range(u32())]
    pk["beta2"], pk["delta2"] = g2(), g2()
    pk["B2"] = [g2() for _ in range(u32())]
    nw, _, _ = u64(), u64(), u64()
    pk["nb_wires"] = nw
    n = u32(); pk["infinity_a"] = [bool(x) for x in buf[off:off + n]]; off += n
    n = u32(); pk["infinity_b"] = [bool(x) for x in buf[off:off + n]]; off += n
    keys = []
    for _ in range(u32()):
        basis = [g1() for _ in range(u32())]
        keys.append({"basis": basis, "basis_exp_sigma": [g1() for _ in range(u32())]})
    pk["commitment_keys"] = keys
    assert off == len(buf)
    skip = set()
    for info in c.commitments:
        skip.add(info["CommitmentIndex"]); skip.update(info["PrivateCommitted"])
    pk["k_wires"] = [i for i in range(c.nb_public, nw) if i not in skip]
    return pk


def write_vk(vk):
    """VerifyingKey.WriteRawTo -- layout verified on the reference's committed .vk files."""
    out = (S.g1_to_bytes(vk["alpha1"]) + S.g1_to_bytes(vk["beta1"]) + S.g2_to_bytes(vk["beta2"]) +
           S.g2_to_bytes(vk["gamma2"]) + S.g1_to_bytes(vk["delta1"]) + S.g2_to_bytes(vk["delta2"]))
    out += _g1_slice(vk["K"])
    pacc = vk["public_and_commitment_committed"]
    out += struct.pack(">I", len(pacc))
    for lst in pacc:
        out += struct.pack(">I", len(lst)) + b"".join(struct.pack(">Q", x) for x in lst)
    out += struct.pack(">I", len(vk["commitment_keys"]))
    for key in vk["commitment_keys"]:
        out += S.g2_to_bytes(key["g"]) + S.g2_to_bytes(key["g_sigma_neg"])
    return out


def read_vk(buf):
    off = 0
    def g1():
        nonlocal off
        p = S.g1_from_bytes(buf[off:off + 64]); off += 64; return p
    def g2():
        nonlocal off
        p = S.g2_from_bytes(buf[off:off + 128]); off += 128; return p
    def u32():
        nonlocal off
        v = struct.unpack_from(">I", buf, off)[0]; off += 4; return v
    vk = {"alpha1": g1(), "beta1": g1(), "beta2": g2(), "gamma2": g2(), "delta1": g1(), "delta2": g2()}
    vk["K"] = [g1() for _ in range(u32())]
    pacc = []
    for _ in range(u32()):
        ln = u32()
        pacc.append([struct.unpack_from(">Q", buf, off + 8 * i)[0] for i in range(ln)])
        off += 8 * ln
    vk["public_and_commitment_committed"] = pacc
    vk["commitment_keys"] = [{"g": g2(), "g_sigma_neg": g2()} for _ in range(u32())]
    assert off == len(buf), (off, len(buf))
    return vk


def write_proof(proof):
    """Proof.WriteRawTo: Ar | Bs | Krs | u32 nCommitments | commitments | CommitmentPok."""
    out = S.g1_to_bytes(proof["ar"]) + S.g2_to_bytes(proof["bs"]) + S.g1_to_bytes(proof["krs"])
    out += struct.pack(">I", len(proof["commitments"]))
    out += b"".join(S.g1_to_bytes(p) for p in proof["commitments"])
    out += S.g1_to_bytes(proof["commitment_pok"])
    return out


def read_proof(buf):
    ncom = struct.unpack_from(">I", buf, 256)[0]
    p = {"ar": S.g1_from_bytes(buf[0:64]), "bs": S.g2_from_bytes(buf[64:192]), "krs": S.g1_from_bytes(buf[192:256])}
    p["commitments"] = [S.g1_from_bytes(buf[260 + 64 * i:324 + 64 * i]) for i in range(ncom)]
    p["commitment_pok"] = S.g1_from_bytes(buf[260 + 64 * ncom:324 + 64 * ncom])
    assert len(buf) == 324 + 64 * ncom
    return p


def write_public_witness(public_values):
    """witness.WriteTo of the public part: u32 nPublic | u32 nSecret(=0) | u32 len | values."""
    n = len(public_values)
    return struct.pack(">III", n, 0, n) + b"".join(S.fr_to_bytes(v) for v in public_values)


def read_public_witness(buf):
    n, nsec, ln = struct.unpack_from(">III", buf, 0)
    assert nsec == 0 and ln == n and len(buf) == 12 + 32 * n
    return [int.from_bytes(buf[12 + 32 * i:44 + 32 * i], "big") for i in range(n)]


# ------------------------------------------------------------------------------------------
# prove (gnark backend/groth16/bn254/prove.go)
# ------------------------------------------------------------------------------------------
def prove_from_wires(c, pk, w, commitments, r, s):
    """Steps 2..6 of SURVEY.md 3.2 from a full wire vector."""
    n = pk["domain"]
    logn = n.bit_length() - 1
    a, b, cc = evaluate_abc(c, w)
    h = B.bitrev_permute(B.quotient_h(a, b, cc))          # gnark leaves H bit-reversed
    wa = [w[i] for i in range(len(w)) if not pk["infinity_a"][i]]
    wb = [w[i] for i in range(len(w)) if not pk["infinity_b"][i]]
    d1, d2 = pk["delta1"], pk["delta2"]
    ar = B.g1_add(B.g1_add(B.g1_msm(pk["A"], wa), pk["alpha1"]), B.g1_mul(d1, r))
    bs1 = B.g1_add(B.g1_add(B.g1_msm(pk["B1"], wb), pk["beta1"]), B.g1_mul(d1, s))
    bs = B.g2_add(B.g2_add(B.g2_msm(pk["B2"], wb), pk["beta2"]), B.g2_mul(d2, s))
    krs = B.g1_msm(pk["K"], [w[i] for i in pk["k_wires"]])
    krs = B.g1_add(krs, B.g1_msm(pk["Z"], h[:n - 1]))
    krs = B.g1_add(krs, B.g1_mul(d1, (-r * s) % R))
    krs = B.g1_add(krs, B.g1_mul(ar, s))
    krs = B.g1_add(krs, B.g1_mul(bs1, r))
    # proof of knowledge of the committed values, folded (coefficient 1 for a single commitment)
    pok = None
    if commitments:
        ser = b"".join(S.g1_to_bytes(pt) for pt, _ in commitments)
        rho = hash_to_field(ser, FOLD_DST)[0] if len(commitments) > 1 else 1
        coef = 1
        for key, (_, vals) in zip(pk["commitment_keys"], commitments):
            pok = B.g1_add(pok, B.g1_mul(B.g1_msm(key["basis_exp_sigma"], vals), coef))
            coef = coef * rho % R
    else:
        pok = None
    proof = {"ar": ar, "bs": bs, "krs": krs, "commitments": [pt for pt, _ in commitments], "commitment_pok": pok}
    return proof, {"a": a, "b": b, "c": cc, "h": h}


def prove(c, pk, assignment, r, s, blinder=None):
    w, commitments = solve(c, assignment, pk, blinder)
    proof, aux = prove_from_wires(c, pk, w, commitments, r, s)
    aux["wires"] = w
    return write_proof(proof), write_public_witness(w[1:c.nb_public]), aux


# ------------------------------------------------------------------------------------------
# verify (gnark backend/groth16/bn254/verify.go)
# ------------------------------------------------------------------------------------------
def verify(vk, proof_bytes, pw_bytes):
    import pairing as Pg
    try:
        proof = read_proof(proof_bytes)
        pub = read_public_witness(pw_bytes)
    except (AssertionError, ValueError, IndexError, struct.error):
        return False                      # undecodable encoding == rejected, as gnark's ReadFrom would
    ncom = len(vk["commitment_keys"])
    if len(proof["commitments"]) != ncom or len(vk["K"]) != 1 + len(pub) + ncom:
        return False
    for p in (proof["ar"], proof["krs"], proof["commitment_pok"], *proof["commitments"]):
        if not B.g1_on_curve(p):
            return False
    if not B.g2_on_curve(proof["bs"]):
        return False
    vkx = vk["K"][0]
    for v, k in zip(pub, vk["K"][1:]):
        vkx = B.g1_add(vkx, B.g1_mul(k, v))
    full = [1] + pub
    for i, cm in enumerate(proof["commitments"]):
        hashed = [full[j] for j in vk["public_and_commitment_committed"][i]]
        msg = S.g1_to_bytes(cm) + b"".join(S.fr_to_bytes(x) for x in hashed)
        chal = hash_to_field(msg, COMMITMENT_DST)[0]
        vkx = B.g1_add(vkx, B.g1_mul(vk["K"][1 + len(pub) + i], chal))
        vkx = B.g1_add(vkx, cm)
    if ncom:
        # folded Pedersen proof of knowledge: e(sum rho^i C_i, GSigmaNeg) * e(PoK, G) == 1
        ser = b"".join(S.g1_to_bytes(cm) for cm in proof["commitments"])
        rho = hash_to_field(ser, FOLD_DST)[0] if ncom > 1 else 1
        folded, coef = None, 1
        for cm in proof["commitments"]:
            folded = B.g1_add(folded, B.g1_mul(cm, coef))
            coef = coef * rho % R
        key = vk["commitment_keys"][0]
        if not Pg.pairing_product_is_one([(folded, key["g_sigma_neg"]), (proof["commitment_pok"], key["g"])]):
            return False
    # e(A,B) == e(alpha,beta) e(vk_x,gamma) e(Krs,delta)
    return Pg.pairing_product_is_one([
        (proof["ar"], proof["bs"]),
        (B.g1_neg(vk["alpha1"]), vk["beta2"]),
        (B.g1_neg(vkx), vk["gamma2"]),
        (B.g1_neg(proof["krs"]), vk["delta2"]),
    ])


def closed_form_check(c, tx, w, proof_bytes, r, s):
    """With the toxic waste known every proof element is a known multiple of the generator."""
    proof = read_proof(proof_bytes)
    tau, alpha, beta, delta = tx["tau"], tx["alpha"], tx["beta"], tx["delta"]
    A, Bv, C = wire_polys_at_tau(c, tau)
    at = sum(x * y for x, y in zip(A, w)) % R
    bt = sum(x * y for x, y in zip(Bv, w)) % R
    a_scalar = (alpha + at + r * delta) % R
    b_scalar = (beta + bt + s * delta) % R
    return (proof["ar"] == B.g1_mul(B.G1_GEN, a_scalar) and proof["bs"] == B.g2_mul(B.G2_GEN, b_scalar))

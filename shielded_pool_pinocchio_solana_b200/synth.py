"""Synthetic gnark-format circuits for tests and benchmarks.

The reference ships only ONE constraint system (noir_circuit/target/shielded_pool_verifier.ccs);
the audit circuit's `.ccs` is a missing blob (/root/reference/.MISSING_LARGE_BLOBS) and no witness
exists for either.  This module writes satisfiable stand-ins in the same gnark v0.14 `.ccs`
container so that they enter the prover through exactly the same parser, solver and kernels:

  * `audit_like`  : 26,000 constraints (domain 2^15), 2 public inputs, 1 BSB22 commitment, row
                    densities of the withdraw circuit (avg nnz/row A 1.02, B 5.06, C 2.31)
  * any size      : `build(n_constraints=...)`

Row k defines a fresh internal wire:  (a * w_p) * (sum_j b_j * w_qj) = w_new + [c * w_r].
A handful of rows exercise the gnark solver hints the withdraw circuit uses (InvZero, nBits,
rangecheck.DecomposeHint, logderivarg.countHint, hints.Randomize, Bsb22 commitment) and a few
assertion rows make wrong witnesses fail.
"""
import random
import struct

import cbor2

R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
CONST_WIRE = 0xFFFFFFFF
H_DECOMPOSE = 467796477
H_RANDOMIZE = 1774611027
H_COUNT = 2138922168
H_INVZERO = 2534161455
H_NBITS = 4115454955
H_COMMIT = 4156202267
MIX_RECENT = 0.575  # probability that an operand of a value_mix row is a recent wire
MIX_SPAN = 8       # value_mix rows: how far back a 'recent' operand reaches (sets the dependency depth)
HINT_NAMES = {
    H_DECOMPOSE: "github.com/consensys/gnark/std/rangecheck.DecomposeHint",
    H_RANDOMIZE: "github.com/consensys/gnark/internal/hints.Randomize",
    H_COUNT: "github.com/consensys/gnark/std/internal/logderivarg.countHint",
    H_INVZERO: "github.com/consensys/gnark/constraint/solver.InvZeroHint",
    H_NBITS: "github.com/consensys/gnark/std/math/bits.nBits",
    H_COMMIT: "github.com/consensys/gnark/frontend/cs.Bsb22CommitmentComputePlaceholder",
}


def _varbyte_stream(values, wbits):
    raw = bytearray()
    prev = 0
    for v in values:
        d = v - prev
        assert d >= 0
        prev = v
        while True:
            b = d & 0x7F
            d >>= 7
            if d:
                raw.append(b | 0x80)
            else:
                raw.append(b)
                break
    wb = wbits // 8
    while len(raw) % wb:
        raw.append(0x80)
    data = [int.from_bytes(raw[i:i + wb], "big") for i in range(0, len(raw), wb)]
    if wbits == 32:
        tail = [len(values), len(data) + 2] + data
    else:
        tail = [len(values) | ((len(data) + 1) << 32)] + data
    words = tail + [len(tail)]
    return struct.pack("<Q%d%s" % (len(words), "I" if wbits == 32 else "Q"), len(words), *words)


def _bitpacked_stream32(values):
    """intcomp bit-packed section (zigzag deltas), for streams that are not monotone."""
    vals = list(values)
    while len(vals) % 128:
        vals.append(vals[-1] if vals else 0)
    body = []
    prev = 0
    for blk in range(0, len(vals), 128):
        descs, packed = [], []
        for sb in range(4):
            chunk = vals[blk + 32 * sb: blk + 32 * sb + 32]
            deltas = []
            for v in chunk:
                deltas.append(v - prev)
                prev = v
            zz = any(d < 0 for d in deltas)
            enc = [((d << 1) ^ (d >> 63)) & 0xFFFFFFFF if zz else d for d in deltas]
            width = max(e.bit_length() for e in enc)
            descs.append((int(zz) << 7) | width)
            bits = 0
            for i, e in enumerate(enc):
                bits |= e << (width * i)
            packed += [(bits >> (32 * k)) & 0xFFFFFFFF for k in range(width)]
        body.append((descs[0] << 24) | (descs[1] << 16) | (descs[2] << 8) | descs[3])
        body += packed
    words = [len(vals), len(body) + 3, 0] + body + [0]
    return struct.pack("<Q%dI" % len(words), len(words), *words)


class SyntheticCircuit:
    """A generated circuit: `.ccs` bytes plus a recipe for satisfying assignments."""

    def __init__(self):
        self.ccs = b""
        self.nb_public = 0       # including the ONE wire
        self.nb_secret = 0
        self.nb_wires = 0
        self.nb_constraints = 0
        self.commitment = False
        self._assign = None

    def assignment(self, seed):
        """-> list of ints: public inputs (without ONE) then secret inputs, satisfying the circuit."""
        return self._assign(seed)

    def assignment_bytes(self, seed):
        return b"".join(v.to_bytes(32, "big") for v in self.assignment(seed))

    def witness_gz(self, seed):
        """The same assignment as a Noir witness file (`target/<name>.gz` of `nargo execute`):
        gzip(bincode WitnessStack), field elements as 32 big-endian bytes."""
        import gzip
        vals = self.assignment(seed)
        body = struct.pack("<Q", 1) + struct.pack("<I", 0) + struct.pack("<Q", len(vals))
        for k, v in enumerate(vals):
            body += struct.pack("<I", k) + struct.pack("<Q", 32) + v.to_bytes(32, "big")
        return gzip.compress(body)


def build(n_constraints, n_public=2, n_secret=64, commitment=True, n_committed=48, seed=0xA0D17,
          dens_b=5, frac_c2=0.31, n_coeffs=256, div_zero=False, value_mix=None, n_public_committed=0):
    """`value_mix` = (zero, one, small, uniform) fractions: body rows are then drawn from four kinds whose SOLVED
    wire is a balanced bit (xor of two bits), a zero (b * (1 - b)), a byte (sum of 2^k * bit) or a uniform field
    element, and the secret inputs past the first 8 are bits / bytes / uniform values in matching proportions --
    the witness-like value distribution of SURVEY.md 8(d) arrived at through the solver, not injected."""
    rng = random.Random(seed)
    assert n_public >= 2 and n_secret >= 8
    coeffs = [0, 1, 2, R - 1, R - 2] + [rng.randrange(1, R) for _ in range(n_coeffs)]
    for k in range(13):                      # 2^k for recompositions
        coeffs.append(1 << k)
    pow2 = {k: len(coeffs) - 13 + k for k in range(13)}

    def const_id(v):
        if v not in coeffs:
            coeffs.append(v)
        return coeffs.index(v)
    npub = n_public + 1                     # + ONE
    first_secret = npub
    first_internal = npub + n_secret
    next_wire = first_internal
    wire_level = {w: -1 for w in range(first_internal)}   # inputs are known before level 0

    instr = []      # (blueprint, calldata list, level)
    row_count = [0]
    def lvl(wires):
        return 1 + max([wire_level[w] for w in wires if w != CONST_WIRE] + [-1])

    def add_r1c(L, Rr, O, defines=None):
        cd = [0, len(L), len(Rr), len(O)]
        for expr in (L, Rr, O):
            for cid, wid in expr:
                cd += [cid, wid]
        cd[0] = len(cd)
        used = [w for expr in (L, Rr, O) for _, w in expr if w != defines]
        level = lvl(used)
        if defines is not None:
            wire_level[defines] = level
        instr.append((1, cd, level))
        row_count[0] += 1

    def add_hint(hid, inputs, nout):
        nonlocal next_wire
        cd = [0, hid, len(inputs)]
        used = []
        for expr in inputs:
            cd.append(len(expr))
            for cid, wid in expr:
                cd += [cid, wid]
                used.append(wid)
        o0 = next_wire
        next_wire += nout
        cd += [o0, o0 + nout]
        cd[0] = len(cd)
        level = lvl(used)
        for w in range(o0, o0 + nout):
            wire_level[w] = level
        instr.append((0, cd, level))
        return list(range(o0, o0 + nout))

    S = lambda i: first_secret + i          # secret wire i
    PUB = lambda i: 1 + i                   # public wire i (0-based among the real publics)
    ONE = (1, 0)

    # ---- hint-exercising prologue ------------------------------------------------------------
    # secret 1 is non-zero: s1 * inv = 1
    (iz,) = add_hint(H_INVZERO, [[(1, S(1))]], 1)
    add_r1c([(1, S(1))], [(1, iz)], [ONE])
    # secret 2 < 2^12: bits, booleanity, recomposition
    bits = add_hint(H_NBITS, [[(1, S(2))]], 12)
    for b in bits:
        add_r1c([(1, b)], [(1, 0), (3, b)], [(0, 0)])
    add_r1c([(pow2[k], b) for k, b in enumerate(bits)], [ONE], [(1, S(2))])
    # the same value as two 6-bit limbs (rangecheck) and their multiplicities in a 64-row table
    limbs = add_hint(H_DECOMPOSE, [[(const_id(12), CONST_WIRE)], [(const_id(6), CONST_WIRE)], [(1, S(2))]], 2)
    add_r1c([(1, limbs[0]), (pow2[6], limbs[1])], [ONE], [(1, S(2))])
    mult = add_hint(H_COUNT, [[(const_id(64), CONST_WIRE)], [(1, CONST_WIRE)]] +
                    [[(const_id(v), CONST_WIRE)] for v in range(64)] + [[(1, limbs[0])], [(1, limbs[1])]], 64)
    add_r1c([(1, m) for m in mult], [ONE], [(2, 0)])
    # public 0 == secret 0 squared ; public 1 == secret 0 * secret 3
    add_r1c([(1, S(0))], [(1, S(0))], [(1, PUB(0))])
    add_r1c([(1, S(0))], [(1, S(3))], [(1, PUB(1))])

    known = list(range(0, first_internal)) + [iz] + bits + limbs + mult
    if div_zero:
        # q * (s4 - s5) = (s6 - s7) with s4 == s5 and s6 == s7: gnark's solver leaves q = 0
        # (DivUnchecked(0, 0) = 0) instead of failing on the zero divisor
        dq = next_wire
        next_wire += 1
        add_r1c([(1, dq)], [(1, S(4)), (3, S(5))], [(1, S(6)), (3, S(7))], defines=dq)
        known.append(dq)
    n_rows_target = n_constraints
    rows_so_far = lambda: row_count[0]

    # ---- commitment (Randomize blinder + BSB22 placeholder) -----------------------------------
    commit_info = None
    n_rows_before_commit = max(8, min(n_committed, n_constraints // 4))
    def add_body_row():
        nonlocal next_wire
        new = next_wire
        next_wire += 1
        nk = len(known)
        pick = lambda: known[rng.randrange(nk)]
        recent = lambda span: known[rng.randrange(max(0, nk - span), nk)]
        # coefficient mix measured on the reference's withdraw circuit
        # (noir_circuit/target/shielded_pool_verifier.ccs): A 99.9 % ones; B 69 % general, 19 % one,
        # 12 % minus-one; C mostly 1 / 2 / -1 with a few general ones
        def coef_b():
            u = rng.random()
            return rng.randrange(5, 5 + n_coeffs) if u < 0.69 else (1 if u < 0.88 else 3)
        def coef_c():
            u = rng.random()
            return 1 if u < 0.45 else (2 if u < 0.75 else (3 if u < 0.95 else rng.randrange(5, 5 + n_coeffs)))
        p = recent(4096) if rng.random() < 0.5 else pick()
        L = [(1, p)]
        if rng.random() < 0.02:
            L.append((rng.randrange(5, 5 + n_coeffs), pick()))
        Rr = [(coef_b(), recent(64) if rng.random() < 0.3 else pick()) for _ in range(dens_b)]
        if rng.random() < 0.06:
            Rr.append((1, 0))
        O = [(1, new), (coef_c(), pick())]
        if rng.random() < frac_c2:
            O.append((coef_c(), pick()))
        add_r1c(L, Rr, O, defines=new)
        known.append(new)
        return new

    # ---- witness-like value mix: pools of wires whose solved value is a bit / zero / byte -----
    sec_kind = {}
    if value_mix is not None:
        f_zero, f_one, f_small, f_uni = value_mix
        f_bit = 2.0 * f_one                         # balanced bits: half of them are ones
        f_z = max(0.0, f_zero - f_one)              # the zeros that are not just 0-valued bits
        assert abs(f_bit + f_z + f_small + f_uni - 1.0) < 1e-9
        pools = {"bit": [], "zero": [], "small": [], "uni": list(known)}
        for i in range(8, n_secret):
            u = rng.random()
            kind = "bit" if u < f_bit + f_z else ("small" if u < f_bit + f_z + f_small else "uni")
            sec_kind[i] = kind
            if kind != "uni":
                pools["uni"].remove(S(i))
                pools[kind].append(S(i))
        assert len(pools["bit"]) >= 16
        known = pools["uni"]
        uniform_row = add_body_row

        def pool_pick(kind, span=4096):
            # half of the operands are recent wires: the dependency depth (solver levels per row) then matches the
            # withdraw circuit's (657 levels for 12,452 rows)
            pl = pools[kind]
            return pl[rng.randrange(max(0, len(pl) - span), len(pl))] if rng.random() < MIX_RECENT else pl[rng.randrange(len(pl))]

        def add_body_row():                          # noqa: F811 -- replaces the uniform-only row generator
            nonlocal next_wire
            u = rng.random()
            if not pools["zero"]:
                kind = "zero"                        # the xor rows pad their B side with zero-valued wires
            elif u < f_bit:
                kind = "bit"
            elif u < f_bit + f_z:
                kind = "zero"
            elif u < f_bit + f_z + f_small:
                kind = "small"
            else:
                return uniform_row()                 # draws its operands from `known` = the uniform wires only
            new = next_wire
            next_wire += 1
            pad = lambda k: [((rng.randrange(5, 5 + n_coeffs) if rng.random() < 0.69 else 1), pool_pick("zero")) for _ in range(k)]
            if kind == "bit":        # 2a * (b + zeros) = a + b - new   (xor)
                a, b = pool_pick("bit", MIX_SPAN), pool_pick("bit", MIX_SPAN)
                add_r1c([(2, a)], [(1, b)] + pad(dens_b - 1), [(1, a), (1, b), (3, new)], defines=new)
            elif kind == "zero":     # a * (1 - a) = new
                a = pool_pick("bit", MIX_SPAN)
                add_r1c([(1, a)], [(1, 0), (3, a)], [(1, new)], defines=new)
            else:                    # 1 * sum 2^k b_k = new   (a byte)
                add_r1c([ONE], [(pow2[k], pool_pick("bit", MIX_SPAN)) for k in range(8)], [(1, new)], defines=new)
            pools[kind].append(new)
            return new

    body_wires = []
    while len(body_wires) < n_rows_before_commit and rows_so_far() < n_rows_target:
        body_wires.append(add_body_row())
    if commitment:
        (blinder,) = add_hint(H_RANDOMIZE, [], 1)
        committed = body_wires[:n_committed] + [blinder]
        # gnark api.Commit over public wires too: their values are hashed into the challenge next to the commitment
        pub_committed = [PUB(i) for i in range(min(n_public_committed, n_public))]
        (cw,) = add_hint(H_COMMIT, [[(0, CONST_WIRE)]] + [[(1, w)] for w in pub_committed] + [[(1, w)] for w in committed], 1)
        known.append(cw)
        # the challenge must feed later rows or it would be dead weight
        commit_info = {"CommitmentIndex": cw, "PrivateCommitted": committed, "NbPublicCommitted": len(pub_committed),
                       "PublicAndCommitmentCommitted": pub_committed}
        new = next_wire
        next_wire += 1
        add_r1c([(1, cw)], [(1, cw), (rng.randrange(5, 5 + n_coeffs), body_wires[0])], [(1, new)], defines=new)
        known.append(new)
    while rows_so_far() < n_rows_target:
        add_body_row()

    # ---- container ------------------------------------------------------------------------------
    nlevels = 1 + max(level for _, _, level in instr)
    levels = [[] for _ in range(nlevels)]
    for i, (_, _, level) in enumerate(instr):
        levels[level].append(i)
    blueprint, coff, woff, start, calldata = [], [], [], [], []
    nrow = 0
    wire_cursor = first_internal
    for bp, cd, _ in instr:
        blueprint.append(bp)
        coff.append(nrow)
        woff.append(wire_cursor)
        start.append(len(calldata))
        calldata += cd
        if bp == 1:
            nrow += 1
            for cid, wid in zip(cd[4::2], cd[5::2]):
                if wid != CONST_WIRE and wid >= wire_cursor:
                    wire_cursor = wid + 1
        else:
            wire_cursor = max(wire_cursor, cd[-1])
    nb_wires = next_wire
    body = {
        "Type": 1,
        "Public": ["1"] + ["pub_%d" % i for i in range(n_public)],
        # named like sunspot names ACIR witnesses, so the Noir-witness front door maps them
        "Secret": ["__witness_%d" % (n_public + i) for i in range(n_secret)],
        "NbInternalVariables": nb_wires - first_internal,
        "NbConstraints": nrow,
        "ScalarField": "30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001",
        "GnarkVersion": "0.14.0",
        "Blueprints": [cbor2.CBORTag(5309735, {}), cbor2.CBORTag(5309736, {})],
        "CommitmentInfo": cbor2.CBORTag(5309737, [commit_info] if commit_info else []),
        "MHintsDependencies": HINT_NAMES,
        "GkrInfo": None, "Logs": [], "DebugInfo": [], "MDebug": {}, "SymbolTable": None,
    }
    lv = struct.pack("<Q", len(levels)) + b"".join(_varbyte_stream(l, 32) for l in levels)
    ins = (_bitpacked_stream32(blueprint) + _varbyte_stream(coff, 32) + _varbyte_stream(woff, 32) +
           _varbyte_stream(start, 64))
    cd = bytearray(struct.pack("<Q", len(calldata)))
    for v in calldata:
        while True:
            b = v & 0x7F
            v >>= 7
            if v:
                cd.append(b | 0x80)
            else:
                cd.append(b)
                break
    body_b = cbor2.dumps(body)
    co = bytearray(struct.pack("<Q", len(coeffs)))
    for x in coeffs:
        m = (x << 256) % R
        co += struct.pack("<4Q", *[(m >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)])
    payload = struct.pack("<4Q", len(lv), len(ins), len(cd), len(body_b)) + lv + ins + bytes(cd) + body_b + bytes(co)

    out = SyntheticCircuit()
    out.ccs = struct.pack("<4Q", len(payload), 0, 14, 0) + payload
    out.nb_public, out.nb_secret, out.nb_wires, out.nb_constraints = npub, n_secret, nb_wires, nrow
    out.commitment = commitment

    def assign(aseed):
        r2 = random.Random((seed << 20) ^ aseed)
        sec = [r2.randrange(1, R) for _ in range(n_secret)]
        for i, kind in sec_kind.items():
            if kind == "bit":
                sec[i] = r2.randrange(2)
            elif kind == "small":
                sec[i] = r2.randrange(256)
        sec[2] = r2.randrange(1 << 12)
        pub = [r2.randrange(R) for _ in range(n_public)]
        pub[0] = sec[0] * sec[0] % R
        pub[1] = sec[0] * sec[3] % R
        if div_zero:
            sec[5], sec[7] = sec[4], sec[6]
        return pub + sec
    out._assign = assign
    return out


def audit_like(uniform=False):
    """Stand-in for the missing audit circuit: ~26K constraints, 2 public inputs, 1 commitment
    (/root/reference/README.md:49; rlwe_audit.vk has 4 K points = ONE + 2 public + commitment).
    Default: the solved witness has the witness-like value mix of SURVEY.md 8(d) (40 % zero, 30 % one, 20 % < 2^8,
    10 % uniform -- an RLWE audit circuit is range checks and small-coefficient arithmetic); `uniform=True` is the
    round-1 variant whose every wire is a uniform field element (the worst case for the MSMs)."""
    return build(26000, n_public=2, n_secret=2400, commitment=True, n_committed=490, seed=0xA0D17,
                 value_mix=None if uniform else (0.4, 0.3, 0.2, 0.1))

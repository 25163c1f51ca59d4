"""The three solver hints the withdraw circuit needs beyond gnark's common ones:
gnark `std/math/emulated.mulHint`, sunspot `sw-grumpkin.decomposeScalar` and `sw-grumpkin.decompose`.

Their semantics are derived from the constraints of /root/reference/noir_circuit/target/shielded_pool_verifier.ccs
(committed as tests/golden/shielded_pool_verifier.ccs) that consume their outputs -- instructions 18-21, 150 and the
deferred multiplication check in the last rows.  No Noir witness of that circuit exists offline (every ACIR witness is
a secret input, and there is no ACVM here), so the hints are driven with random inputs in the solver's diagnostic
mode: rows that do not touch the gadget fail, rows that do must hold.  CPU only."""
import os
import random
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "py"))
import ccs as occs            # noqa: E402
import groth16 as G           # noqa: E402

R = G.R
REAL = os.path.join(ROOT, "tests", "golden", "shielded_pool_verifier.ccs")


def test_glv_split_relation_and_range():
    rng = random.Random(11)
    q, lam = G.GRUMPKIN_Q, G.GRUMPKIN_LAMBDA
    assert (lam * lam + lam + 1) % q == 0
    for a, b in (G.GRUMPKIN_B1, G.GRUMPKIN_B2):
        assert (a - lam * b) % q == 0                      # both basis vectors lie in the lattice u = lambda v
    for k in range(3000):
        s = [0, 1, q - 1, (1 << 127) - 1, 1 << 127, (1 << 128) - 1][k] if k < 6 else (
            rng.randrange(1 << 128) if k % 2 else rng.randrange(q))
        s1, s2 = G.glv_split_nonneg(s)
        assert 0 <= s1 < 1 << 127 and 0 <= s2 < 1 << 127
        assert (s1 - s - lam * s2) % q == 0


def _poly(limbs, x):
    acc = 0
    for v in reversed(limbs):
        acc = (acc * x + v) % R
    return acc


@pytest.mark.parametrize("na,nb,with_rem", [(4, 4, True), (6, 1, True), (7, 4, True), (4, 4, False)])
def test_emulated_mul_hint_identity(na, nb, with_rem):
    """quo, rem, carries satisfy a(X) b(X) = rem(X) + quo(X) p(X) + (2^64 - X) carry(X) at random points of Fr --
    the identity gnark's deferred multiplication check evaluates at the commitment challenge."""
    rng = random.Random(na * 100 + nb)
    nbits, nlimbs = 64, 4
    p = G.GRUMPKIN_Q
    pl = [(p >> (64 * i)) & (2**64 - 1) for i in range(4)]
    for _ in range(20):
        al = [rng.randrange(1 << rng.choice((64, 70, 130))) for _ in range(na)]
        bl = [rng.randrange(1 << 64) for _ in range(nb)]
        a, b = G._recompose(al, 64), G._recompose(bl, 64)
        if not with_rem:                                   # the zero-check form: a*b must be a multiple of p
            al[0] += (-(a * b) * pow(b, -1, p)) % p if b % p else 0
            a = G._recompose(al, 64)
            if (a * b) % p:
                continue
        nquo = max(1, ((a * b) // p).bit_length() + 63) // 64 if a * b else 1
        ncarry = max(na + nb - 1, nquo + nlimbs - 1) - 1
        nout = nquo + (nlimbs if with_rem else 0) + ncarry
        out = G._hint_emulated_mul([nbits, nlimbs, na, nquo] + pl + al + bl, nout)
        ql, rl, cl = out[:nquo], out[nquo:nout - ncarry], out[nout - ncarry:]
        assert G._recompose(ql, 64) * p + G._recompose(rl, 64) == a * b
        for _x in range(3):
            x = rng.randrange(R)
            lhs = _poly(al, x) * _poly(bl, x) % R
            rhs = (_poly(rl, x) + _poly(ql, x) * _poly(pl, x) + ((1 << 64) - x) * _poly(cl, x)) % R
            assert lhs == rhs


@pytest.fixture(scope="module")
def real():
    c = occs.parse_ccs(open(REAL, "rb").read())
    pk, _vk, _ = G.setup(c, b"hint-test", fast=True)
    return c, pk


def _assignment(c, seed):
    rng = random.Random(seed)
    asg = [rng.randrange(R) for _ in range(c.nb_public - 1 + c.nb_secret)]
    asg[27] = rng.randrange(1 << 128) if seed % 2 == 0 else rng.randrange(R)     # wire 28: the scalar
    return asg


def test_withdraw_circuit_gadget_rows_hold(real):
    """Diagnostic solve of the reference's own constraint system with random inputs: the solver gets past the three
    hints, and every row that reads a wire they define (incl. the deferred checks in the last 16 rows) is satisfied."""
    c, pk = real
    names = c.body["MHintsDependencies"]
    gadget = set()
    for ins, bp in enumerate(c.blueprint):
        if bp == occs.BLUEPRINT_HINT:
            hid, _ins, o0, o1 = c.hint(ins)
            if names[hid].endswith(("emulated.mulHint", "sw-grumpkin.decomposeScalar", "sw-grumpkin.decompose")):
                gadget |= set(range(o0, o1))
    assert gadget == set(range(6196, 6222))
    touching = [k for k, rows in enumerate(c.rows()) if {w for e in rows for _cid, w in e} & gadget]
    assert len(touching) == 34 and max(touching) == c.nb_constraints - 1
    for seed in range(3):
        failed = []
        w, _ = G.solve(c, _assignment(c, seed), pk=pk, blinder=5, failed_rows=failed)
        assert failed, "random inputs cannot satisfy the whole circuit"
        assert not set(failed) & set(touching)
        s1 = G._recompose(w[6196:6200], 64)
        s2 = G._recompose(w[6200:6204], 64)
        assert G._recompose(w[6204:6208], 64) == w[28]
        assert (s1 - w[28] - G.GRUMPKIN_LAMBDA * s2) % G.GRUMPKIN_Q == 0 and max(s1, s2) < 1 << 127


def test_host_solver_hints_match_the_restatement(real, monkeypatch):
    """C++ host solver (libg16b200, no GPU needed) == oracle/py on every wire of the withdraw circuit, diagnostic mode."""
    import shielded_pool_pinocchio_solana_b200 as g16
    monkeypatch.setenv("G16_SOLVER_DIAG", "1")
    c, pk = real
    raw = open(REAL, "rb").read()
    cw = c.commitments[0]["CommitmentIndex"]
    for seed in range(2):
        asg = _assignment(c, seed)
        w, _ = G.solve(c, asg, pk=pk, blinder=7, failed_rows=[])
        wires_be, _ = g16.solve_assignment(raw, b"".join(v.to_bytes(32, "big") for v in asg), c.nb_wires,
                                           blinder_be=(7).to_bytes(32, "big"), challenges_be=w[cw].to_bytes(32, "big"),
                                           n_committed=len(c.commitments[0]["PrivateCommitted"]))
        got = [int.from_bytes(wires_be[32 * i:32 * i + 32], "big") for i in range(c.nb_wires)]
        assert got == w

"""Pins the C restatement (oracle/c, the CPU baseline) to the big-integer oracle -- CPU only."""
import json
import os
import random
import subprocess

import pytest

import bn254 as B
import serialize as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def co():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    import coracle
    coracle.set_threads(4)
    return coracle


def test_c_msm_matches_golden(co):
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "msm_small.json")))
    for grp, psize in (("g1", 64), ("g2", 128)):
        case = g[grp]
        pts, sc, res = (bytes.fromhex(case[k]) for k in ("points", "scalars", "results"))
        n = case["n"]
        for b in range(case["batch"]):
            got = co.msm(pts, sc[32 * n * b:32 * n * (b + 1)], grp)
            assert got == res[psize * b:psize * (b + 1)]


@pytest.mark.parametrize("logn", [1, 4, 9])
def test_c_ntt_and_h_match_bigint(co, logn):
    n = 1 << logn
    rng = random.Random(logn)
    enc = lambda v: b"".join(S.fr_to_bytes(x) for x in v)
    dec = lambda b: [int.from_bytes(b[i:i + 32], "big") for i in range(0, len(b), 32)]
    x = [rng.randrange(B.R) for _ in range(n)]
    for coset in (False, True):
        f = dec(co.ntt(enc(x), logn, False, coset))
        assert f == B.bitrev_permute(B.ntt_natural(x, coset=B.COSET_GEN if coset else None))
        assert dec(co.ntt(enc(f), logn, True, coset)) == x
    a = [rng.randrange(B.R) for _ in range(n)]
    b = [rng.randrange(B.R) for _ in range(n)]
    c = [p * q % B.R for p, q in zip(a, b)]
    assert dec(co.compute_h(enc(a) + enc(b) + enc(c), logn)) == B.bitrev_permute(B.quotient_h(a, b, c))


def test_c_prove_matches_bigint_proof(co):
    """Whole prove-from-wires in C equals the proof points of the golden (big-integer) proofs."""
    import ccs
    import groth16 as G
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "prove_small.json")))
    c = ccs.parse_ccs(bytes.fromhex(g["ccs"]))
    pk, vk, tx = G.setup(c, b"golden-small")
    assert G.write_pk(pk).hex() == g["pk"]
    cc = co.CCircuit(c, pk)
    for case in g["cases"]:
        rnd = bytes.fromhex(case["rnd"])
        r, s = int.from_bytes(rnd[:32], "big"), int.from_bytes(rnd[32:64], "big")
        out = cc.prove_from_wires(bytes.fromhex(case["wires"]), r, s)
        proof = bytes.fromhex(case["proof"])
        assert out[:256] == proof[:256]            # Ar | Bs | Krs
        assert out[256:320] == proof[324:388]      # PoK
    cc.close()

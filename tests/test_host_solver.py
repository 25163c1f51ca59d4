"""C++ host logic (no GPU): `.ccs` parser + witness solver vs the oracle and the reference artifact."""
import json
import os

import pytest

import shielded_pool_pinocchio_solana_b200 as g16

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "prove_small.json")
REAL_CCS = os.path.join(os.path.dirname(__file__), "golden", "shielded_pool_verifier.ccs")   # copy of the reference file


def test_solver_matches_oracle_wires():
    g = json.load(open(GOLDEN))
    ccs = bytes.fromhex(g["ccs"])
    for case in g["cases"]:
        rnd = bytes.fromhex(case["rnd"])
        wires, committed = g16.solve_assignment(ccs, bytes.fromhex(case["assignment"]), g["nb_wires"],
                                                blinder_be=rnd[64:96], challenges_be=bytes.fromhex(case["challenge"]),
                                                n_committed=g["n_committed"])
        assert wires.hex() == case["wires"]
        assert len(committed) == 32 * g["n_committed"]


def test_solver_rejects_bad_witness():
    g = json.load(open(GOLDEN))
    case = g["cases"][0]
    asg = bytearray(bytes.fromhex(case["assignment"]))
    asg[31] ^= 1
    with pytest.raises(g16.G16Error) as e:
        g16.solve_assignment(bytes.fromhex(g["ccs"]), bytes(asg), g["nb_wires"], bytes.fromhex(case["rnd"])[64:],
                             bytes.fromhex(case["challenge"]), g["n_committed"])
    assert e.value.code == 3


def test_parser_rejects_garbage():
    g = json.load(open(GOLDEN))
    ccs = bytes.fromhex(g["ccs"])
    for bad in (ccs[:100], b"\x00" * 64, ccs[:-1], ccs + b"\x00"):
        with pytest.raises(g16.G16Error) as e:
            g16.solve_assignment(bad, b"", 1)
        assert e.value.code == 2


def test_real_withdraw_ccs_parses_and_reports_missing_hints():
    """The committed withdraw circuit goes through the C++ parser; with an all-zero witness the
    first row fails, which proves the 12,452-row instruction stream was decoded and walked."""
    real = open(REAL_CCS, "rb").read()
    with pytest.raises(g16.G16Error) as e:
        g16.solve_assignment(real, b"\x00" * 32 * 6189, 12939, b"\x00" * 31 + b"\x01", b"\x00" * 31 + b"\x01", 490)
    assert e.value.code == 3 and "constraint #0" in str(e.value)
    with pytest.raises(g16.G16Error) as e:
        g16.solve_assignment(real, b"\x00" * 32 * 10, 12939)
    assert "6189" in str(e.value)


def _div_zero_circuit():
    from shielded_pool_pinocchio_solana_b200 import synth
    return synth.build(60, n_public=2, n_secret=16, commitment=False, seed=21, div_zero=True)


def test_zero_divisor_follows_gnark_div_unchecked():
    """gnark's solveR1C (constraint/bn254/solver.go) does not fail on a zero divisor: the unknown
    wire stays 0 and the row is only checked (a*b == c).  0/0 is accepted, x/0 (x != 0) is UNSAT.
    Host C++ solver == oracle/py solver on both."""
    import ccs as occs
    import groth16 as G
    import serialize as S
    sc = _div_zero_circuit()
    c = occs.parse_ccs(sc.ccs)
    asg = sc.assignment(3)
    w_oracle, _ = G.solve(c, asg)
    wires, _ = g16.solve_assignment(sc.ccs, sc.assignment_bytes(3), sc.nb_wires)
    assert wires == b"".join(S.fr_to_bytes(x) for x in w_oracle)
    # s6 != s7 with s4 == s5: q * 0 = nonzero has no solution
    n_pub = 2
    bad = list(asg)
    bad[n_pub + 7] = (bad[n_pub + 7] + 1) % synth_R()
    with pytest.raises(G.Unsatisfied):
        G.solve(c, bad)
    with pytest.raises(g16.G16Error) as e:
        g16.solve_assignment(sc.ccs, b"".join(v.to_bytes(32, "big") for v in bad), sc.nb_wires)
    assert e.value.code == 3


def synth_R():
    from shielded_pool_pinocchio_solana_b200 import synth
    return synth.R

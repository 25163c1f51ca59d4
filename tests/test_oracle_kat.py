"""Pins the oracle (and the C++ parser through it) on the reference's OWN committed artifacts.

Nothing in the reference pins proof bytes, but these do pin formats and constants:
  * noir_circuit/target/shielded_pool_verifier.ccs  -- container, streams, calldata, coefficient table
  * noir_circuit/target/shielded_pool_verifier.vk, audit_circuit/target/rlwe_audit.vk -- VK layout
  * client/prover-params.toml                       -- a complete public input set -> the .pw bytes
  * shielded_pool_program/src/instructions/{withdraw,submit_audit}.rs -- 388 / 172 / 76 byte framing
The artifacts are committed copies under tests/golden/ (byte-identical to the reference files; checked
against /root/reference whenever that tree is mounted), so the suite also runs on the GPU box.
"""
import os
import re

import pytest

import bn254 as B
import ccs
import groth16 as G

REF = "/root/reference"
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIXTURES = {"shielded_pool_verifier.ccs": "/noir_circuit/target/shielded_pool_verifier.ccs",
            "shielded_pool_verifier.vk": "/noir_circuit/target/shielded_pool_verifier.vk",
            "rlwe_audit.vk": "/audit_circuit/target/rlwe_audit.vk",
            "audit_circuit.vk": "/audit_circuit/target/audit_circuit.vk",
            "prover-params.toml": "/client/prover-params.toml"}


def fixture_bytes(name):
    return open(os.path.join(GOLD, name), "rb").read()


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")
def test_committed_fixtures_are_the_reference_files():
    for name, rel in FIXTURES.items():
        assert fixture_bytes(name) == open(REF + rel, "rb").read(), name


@pytest.fixture(scope="module")
def withdraw():
    return ccs.parse_ccs(fixture_bytes("shielded_pool_verifier.ccs"))


def test_ccs_header_sizes_and_constants(withdraw):
    c = withdraw
    assert (c.nb_constraints, c.nb_wires, c.nb_public, c.nb_secret, c.nb_internal) == (12452, 12939, 6, 6184, 6749)
    assert int(c.body["ScalarField"], 16) == B.R and c.body["GnarkVersion"] == "0.14.0"
    assert c.body["Public"] == ["1", "root", "nullifier", "recipient", "amount", "wa_commitment"]
    assert c.body["Secret"][0] == "__witness_5" and c.body["Secret"][-1] == "__witness_23643"
    assert len(c.levels) == 657 and len(c.blueprint) == 12493 and len(c.calldata) == 262332 and len(c.coeffs) == 1629
    # Montgomery coefficient table: ids 0..4 = 0, 1, 2, -1, -2
    assert c.coeffs[:5] == [0, 1, 2, B.R - 1, B.R - 2]
    assert [len(l) for l in c.levels[:13]] == [6250, 4643, 398, 6, 2, 492, 6, 6, 5, 5, 4, 2, 2]


def test_ccs_decoder_self_checks(withdraw):
    c = withdraw
    n = len(c.blueprint)
    assert sorted(i for l in c.levels for i in l) == list(range(n))
    assert all(l == sorted(l) for l in c.levels)
    assert sum(1 for b in c.blueprint if b == 1) == c.nb_constraints
    run = 0
    for i in range(n):
        assert c.constraint_offset[i] == run
        run += c.blueprint[i] == 1
    assert all(a <= b for a, b in zip(c.wire_offset, c.wire_offset[1:]))
    assert c.wire_offset[0] == 6190
    for i in range(n - 1):
        assert c.start_calldata[i + 1] - c.start_calldata[i] == c.calldata[c.start_calldata[i]]
    # walking the levels, every row has at most one unknown wire and every hint input is known
    known = [False] * c.nb_wires
    for i in range(c.nb_public + c.nb_secret):
        known[i] = True
    for level in c.levels:
        new = []
        for i in level:
            if c.blueprint[i] == 1:
                unk = {w for side in c.r1c(i) for _, w in side if w != ccs.CONST_WIRE and not known[w]}
                assert len(unk) <= 1
                new += list(unk)
            else:
                _, ins, o0, o1 = c.hint(i)
                assert all(known[w] for e in ins for _, w in e if w != ccs.CONST_WIRE)
                new += list(range(o0, o1))
        for w in new:
            known[w] = True
    assert all(known)


def test_ccs_row_shapes(withdraw):
    c = withdraw
    assert c.r1c(0) == ([(1, 0)], [(5, 0), (3, 28), (3, 29)], [(0, 0)])
    assert c.coeffs[5] == 0xffffffffffffffffffffffffffffffff
    assert c.r1c(3) == ([(1, 27)], [(1, 31)], [(1, 6190)])
    for row_instr in range(22, 149):                      # booleanity of the 127 nBits outputs
        L, R, O = c.r1c(row_instr)
        (c1, w), = L
        assert c1 == 1 and 6222 <= w <= 6348 and R == [(1, 0), (3, w)] and O == [(0, 0)]
    names = c.body["MHintsDependencies"]
    hid, ins, o0, o1 = c.hint(18)
    assert names[hid].endswith("sw-grumpkin.decomposeScalar") and (o0, o1) == (6196, 6204) and len(ins) == 13


def test_msm_sizes_derived_from_the_ccs(withdraw):
    """|A| = 4,175, |B| = 12,701, |K| = 12,442, |Z| = 16,383, commit = 490."""
    c = withdraw
    in_a, in_b = set(), set()
    for L, R, O in c.rows():
        in_a.update(0 if w == ccs.CONST_WIRE else w for _, w in L)
        in_b.update(0 if w == ccs.CONST_WIRE else w for _, w in R)
    info = c.commitments[0]
    assert len(in_a) == 4175 and len(in_b) == 12701
    assert len(info["PrivateCommitted"]) == 490 and info["CommitmentIndex"] == 12426
    assert c.nb_wires - c.nb_public - 490 - 1 == 12442
    assert G.domain_size(c) == 16384


@pytest.mark.parametrize("path,nk,size", [("shielded_pool_verifier.vk", 7, 1296), ("rlwe_audit.vk", 4, 1104),
                                          ("audit_circuit.vk", 4, 1104)])
def test_vk_layout_roundtrip_and_curve_membership(path, nk, size):
    raw = fixture_bytes(path)
    assert len(raw) == size
    vk = G.read_vk(raw)
    assert len(vk["K"]) == nk and len(vk["commitment_keys"]) == 1 and vk["public_and_commitment_committed"] == [[]]
    assert G.write_vk(vk) == raw
    for p in [vk["alpha1"], vk["beta1"], vk["delta1"]] + vk["K"]:
        assert B.g1_on_curve(p)
    for q in [vk["beta2"], vk["gamma2"], vk["delta2"], vk["commitment_keys"][0]["g"], vk["commitment_keys"][0]["g_sigma_neg"]]:
        assert B.g2_on_curve(q)


def test_public_witness_bytes_from_prover_params():
    """client/prover-params.toml holds a full input set: the .pw is header + five 32-byte BE values
    (withdraw.rs:14-16, 71-90: 12-byte header, 32-byte inputs, amount in the last 8 bytes of input 3)."""
    txt = fixture_bytes("prover-params.toml").decode()
    val = lambda k: re.search(r"^%s\s*=\s*\"?(0x[0-9a-f]+|\d+)\"?" % k, txt, re.M).group(1)
    pub = [int(val(k), 0) for k in ("root", "nullifier", "recipient", "amount", "wa_commitment")]
    pw = G.write_public_witness(pub)
    assert len(pw) == 172 and pw[:12] == bytes.fromhex("000000050000000000000005")
    assert int.from_bytes(pw[12 + 32 * 3 + 24:12 + 32 * 4], "big") == 10000000
    assert G.read_public_witness(pw) == pub
    assert 388 + len(pw) == 560                          # payroll-demo.ts:357 "~560 bytes"


def test_cpp_parser_agrees_with_the_oracle_on_the_real_circuit():
    """The product's C++ `.ccs` parser + solver walk the same instruction stream: feeding it a witness
    that satisfies row 0..2 but not a later row reports exactly the first unsatisfied row."""
    import shielded_pool_pinocchio_solana_b200 as g16
    real = fixture_bytes("shielded_pool_verifier.ccs")
    with pytest.raises(g16.G16Error) as e:
        g16.solve_assignment(real, b"\x00" * 32 * 6189, 12939, b"\x00" * 31 + b"\x01", b"\x00" * 31 + b"\x01", 490)
    assert e.value.code == 3 and "constraint #0 " in str(e.value)

"""The reference's withdraw circuit with REAL witnesses (BASELINE.json configs[0]).

No witness file exists in the reference tree and there is no ACVM here, but the circuit's R1CS determines its own
intermediate witnesses: oracle/py/witness_completion.py rebuilds a satisfying assignment from the ABI inputs of
/root/reference/client/prover-params.toml (committed copy: tests/golden/prover-params.toml).  Given only the PRIVATE
inputs it re-derives that file's public key and public inputs -- a known-answer test, on the reference's own vector,
of the `.ccs` decoder, the solver semantics and the three sunspot / gnark integer hints.  CPU only."""
import json
import os
import re
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "py"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ccs as occs                                   # noqa: E402
import groth16 as G                                  # noqa: E402
import gen_golden_withdraw_witness as gen            # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
R = G.R


@pytest.fixture(scope="module")
def circuit():
    return occs.parse_ccs(open(os.path.join(GOLD, "shielded_pool_verifier.ccs"), "rb").read())


@pytest.fixture(scope="module")
def fixtures():
    meta = json.load(open(os.path.join(GOLD, "withdraw_assignments.json")))
    blob = open(os.path.join(GOLD, "withdraw_assignments.bin"), "rb").read()
    nb = meta["n_values"] * 32
    return meta, [blob[i * nb:(i + 1) * nb] for i in range(meta["n"])]


def toml_values():
    txt = open(os.path.join(GOLD, "prover-params.toml")).read()
    val = lambda k: int(re.search(r"^%s\s*=\s*\"?(0x[0-9a-f]+|\d+)\"?" % k, txt, re.M).group(1), 0)   # noqa: E731
    out = {k: val(k) for k in ("root", "nullifier", "recipient", "amount", "wa_commitment", "secret_key", "owner_x",
                               "owner_y", "randomness", "index")}
    sib = re.findall(r"\"(0x[0-9a-f]{64})\",", txt.split("siblings", 1)[1])
    assert len(sib) == 16
    for i, s in enumerate(sib):
        out["sibling_%d" % i] = int(s, 16)
    return out


def test_private_inputs_of_prover_params_rederive_its_public_inputs(circuit, fixtures):
    """secret_key, randomness, index, siblings (+ recipient, amount) -> owner_x, owner_y = secret_key * G on Grumpkin
    (sunspot's scalar-mul gadget driven by this repo's sw-grumpkin / emulated hints), wa_commitment, nullifier and the
    Merkle root (Poseidon, as constraints) -- all equal to what the reference committed."""
    want = toml_values()
    inputs = {k: v for k, v in want.items() if k not in gen.DERIVED}
    w, asg = gen.complete_from_private(circuit, inputs)
    wmap = gen.wire_of_abi(circuit)
    for k, name in enumerate(gen.ABI):
        assert w[wmap[k]] == want[name], name
    meta, blobs = fixtures
    assert b"".join(v.to_bytes(32, "big") for v in asg) == blobs[0]          # the committed fixture is this assignment
    assert {k: int(v, 16) for k, v in meta["prover_params_toml"].items()} == want
    # the public witness file the on-chain program reads (withdraw.rs:14-16) is these five values
    pw = G.write_public_witness([want[k] for k in ("root", "nullifier", "recipient", "amount", "wa_commitment")])
    assert pw[12:] == blobs[0][:160]


def test_fixture_witnesses_satisfy_every_row_and_both_solvers_agree(circuit, fixtures):
    """Strict gnark-order solve (oracle/py) of every committed assignment; the product's C++ host solver gives the same
    wires (commitment challenge injected)."""
    import shielded_pool_pinocchio_solana_b200 as g16
    meta, blobs = fixtures
    raw = open(os.path.join(GOLD, "shielded_pool_verifier.ccs"), "rb").read()
    pk, _vk, _ = G.setup(circuit, b"withdraw-real", fast=True)
    cw = circuit.commitments[0]["CommitmentIndex"]
    for n, blob in enumerate(blobs):
        asg = [int.from_bytes(blob[32 * i:32 * i + 32], "big") for i in range(meta["n_values"])]
        wires, _ = G.solve(circuit, asg, pk=pk, blinder=7 + n)
        if n < 2:
            for L, Rr, O in circuit.rows():
                ev = lambda e: sum(circuit.coeffs[cid] * (1 if wid == occs.CONST_WIRE else wires[wid]) for cid, wid in e) % R   # noqa: E731
                assert ev(L) * ev(Rr) % R == ev(O)
        got, _ = g16.solve_assignment(raw, blob, circuit.nb_wires, blinder_be=(7 + n).to_bytes(32, "big"),
                                      challenges_be=wires[cw].to_bytes(32, "big"),
                                      n_committed=len(circuit.commitments[0]["PrivateCommitted"]))
        assert got == b"".join(v.to_bytes(32, "big") for v in wires), "witness %d" % n
    # distinct witnesses, and a real witness is dominated by full-size field elements (Poseidon), not by zeros and ones
    assert len(set(blobs)) == len(blobs)
    big = sum(1 for v in wires if v >> 128) / len(wires)
    assert big > 0.85


def test_a_wrong_private_input_is_rejected(circuit, fixtures):
    meta, blobs = fixtures
    asg = [int.from_bytes(blobs[0][32 * i:32 * i + 32], "big") for i in range(meta["n_values"])]
    asg[5] ^= 1                                        # secret_key
    pk, _vk, _ = G.setup(circuit, b"withdraw-real", fast=True)
    with pytest.raises(G.Unsatisfied):
        G.solve(circuit, asg, pk=pk, blinder=1)


def test_product_completion_equals_the_fixtures(circuit, fixtures):
    """g16_complete_assignment (C++, host only): from the private ABI inputs alone, byte-identical to every committed
    assignment (which the Python restatement generated)."""
    import shielded_pool_pinocchio_solana_b200 as g16
    meta, blobs = fixtures
    raw = open(os.path.join(GOLD, "shielded_pool_verifier.ccs"), "rb").read()
    wmap = gen.wire_of_abi(circuit)
    for n, blob in enumerate(blobs):
        abi = {name: int(v, 16) for name, v in meta["witnesses"][n].items()}
        known = {wmap[k]: abi[name] for k, name in enumerate(gen.ABI) if name not in gen.DERIVED}
        assert g16.complete_assignment(raw, known) == blob, "witness %d" % n


def test_prover_toml_front_door_and_rejections(circuit, fixtures):
    """Prover.toml + the program's ABI (target/<name>.json) -> assignment: what `nargo execute` + sunspot's witness
    ingest do for this circuit (noir_circuit/prove_linux.sh:62-83), without an ACVM."""
    import tomllib
    import shielded_pool_pinocchio_solana_b200 as g16
    meta, blobs = fixtures
    raw = open(os.path.join(GOLD, "shielded_pool_verifier.ccs"), "rb").read()
    abi_json = open(os.path.join(GOLD, "shielded_pool_verifier.abi.json"), "rb").read()
    toml = tomllib.load(open(os.path.join(GOLD, "prover-params.toml"), "rb"))
    names = circuit.body["Secret"]
    # every input given (as a caller holding Prover.toml would): completion also CHECKS the public inputs
    known = g16.abi_inputs_to_wires(abi_json, names, circuit.nb_public, toml)
    assert len(known) == 26
    assert g16.complete_assignment(raw, known) == blobs[0]
    # private inputs only: the public ones are derived
    private = {k: v for k, v in toml.items() if k not in ("root", "nullifier", "wa_commitment", "owner_x", "owner_y")}
    assert g16.complete_assignment(raw, g16.abi_inputs_to_wires(abi_json, names, circuit.nb_public, private)) == blobs[0]
    # a wrong Merkle root contradicts the constraints (propagation runs backwards from it into a failing row or a
    # failing range-check lookup, whichever comes first)
    bad = dict(toml, root="0x%064x" % (int(toml["root"], 16) ^ 1))
    with pytest.raises(g16.G16Error) as e:
        g16.complete_assignment(raw, g16.abi_inputs_to_wires(abi_json, names, circuit.nb_public, bad))
    assert e.value.code in (3, 5)
    # without the secret key nothing determines the key pair: rejected as well
    short = {k: v for k, v in private.items() if k != "secret_key"}
    with pytest.raises(g16.G16Error) as e:
        g16.complete_assignment(raw, g16.abi_inputs_to_wires(abi_json, names, circuit.nb_public, short))
    assert e.value.code in (3, 5)


def test_execute_c_abi_and_cli(fixtures, tmp_path):
    """g16_execute / `g16prove execute`: Prover.toml + ABI + .ccs -> the witness file (`nargo execute`,
    prove_linux.sh:62), host only; reading it back through the witness ingest gives the committed assignment."""
    import subprocess
    import shielded_pool_pinocchio_solana_b200 as g16
    meta, blobs = fixtures
    raw = open(os.path.join(GOLD, "shielded_pool_verifier.ccs"), "rb").read()
    abi_json = open(os.path.join(GOLD, "shielded_pool_verifier.abi.json"), "rb").read()
    toml = open(os.path.join(GOLD, "prover-params.toml"), "rb").read()
    gz = g16.execute(raw, abi_json, toml)
    assert gz[:2] == b"\x1f\x8b" and g16.witness_to_assignment(raw, gz) == blobs[0]
    exe = os.path.join(ROOT, "shielded_pool_pinocchio_solana_b200", "g16prove")
    out = tmp_path / "w.gz"
    r = subprocess.run([exe, "execute", os.path.join(GOLD, "shielded_pool_verifier.abi.json"),
                        os.path.join(GOLD, "prover-params.toml"), os.path.join(GOLD, "shielded_pool_verifier.ccs"), str(out)],
                       capture_output=True)
    assert r.returncode == 0, r.stderr
    assert g16.witness_to_assignment(raw, out.read_bytes()) == blobs[0]
    # a Prover.toml whose secret key does not match its public key is refused with a non-zero exit
    bad = tmp_path / "bad.toml"
    bad.write_bytes(toml.replace(b"43f5147fe5a665df", b"43f5147fe5a665de"))
    r = subprocess.run([exe, "execute", os.path.join(GOLD, "shielded_pool_verifier.abi.json"), str(bad),
                        os.path.join(GOLD, "shielded_pool_verifier.ccs"), str(out)], capture_output=True)
    assert r.returncode != 0 and b"execute" in r.stderr

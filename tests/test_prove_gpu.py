"""End-to-end CUDA prover vs the oracle, through the C ABI -- proof bytes must be identical.

Reference behaviour: `sunspot prove` = gnark groth16.Prove + Proof.WriteRawTo / witness.WriteTo
(/root/reference/client/proof.helper.ts:58-71; byte framing pinned by
shielded_pool_program/src/instructions/withdraw.rs:13-16 and submit_audit.rs:18-21).
"""
import json
import os

import pytest

import shielded_pool_pinocchio_solana_b200 as g16

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "prove_small.json")


@pytest.fixture(scope="module")
def golden():
    return json.load(open(GOLDEN))


@pytest.fixture(scope="module")
def circuit(ctx, golden):
    c = g16.Circuit(ctx, bytes.fromhex(golden["ccs"]), bytes.fromhex(golden["pk"]))
    yield c
    c.free()


def test_circuit_info(circuit, golden):
    assert circuit.info["nb_wires"] == golden["nb_wires"]
    assert circuit.info["nb_commitments"] == 1 and circuit.info["n_committed"] == golden["n_committed"]
    assert circuit.proof_len == 388            # withdraw.rs:13, submit_audit.rs:18
    assert circuit.pw_len == 76                # 2 public inputs, like the audit circuit (submit_audit.rs:19-21)


def test_single_proofs_bit_exact(circuit, golden):
    for case in golden["cases"]:
        proof, pw = circuit.prove_assignment(bytes.fromhex(case["assignment"]), bytes.fromhex(case["rnd"]))
        assert pw.hex() == case["pw"]
        assert proof.hex() == case["proof"]


def test_batch_bit_exact(circuit, golden):
    cases = golden["cases"] * 5           # 15 proofs in one device batch
    asg = b"".join(bytes.fromhex(c["assignment"]) for c in cases)
    rnd = b"".join(bytes.fromhex(c["rnd"]) for c in cases)
    proofs, pws = circuit.prove_batch(asg, len(cases), rnd)
    assert [p.hex() for p in proofs] == [c["proof"] for c in cases]
    assert [p.hex() for p in pws] == [c["pw"] for c in cases]


def test_prove_wires_bit_exact(circuit, golden):
    """Bypassing the solver with the oracle's wire vector gives the same proof."""
    case = golden["cases"][1]
    (proof,) = circuit.prove_wires(bytes.fromhex(case["wires"]), 1, bytes.fromhex(case["rnd"]))
    assert proof.hex() == case["proof"]


def test_random_blinding_verifies(circuit, golden):
    """With r, s and the blinder drawn by the library every proof must verify under the
    reference verifier's equations (oracle restatement of `sunspot verify`)."""
    import groth16 as G
    vk = G.read_vk(bytes.fromhex(golden["vk"]))
    case = golden["cases"][0]
    proof, pw = circuit.prove_assignment(bytes.fromhex(case["assignment"]), None)
    assert proof.hex() != case["proof"]
    assert G.verify(vk, proof, pw)
    # fault injection as in client/test-shielded-pool.ts:386-392 (proof byte 0 XOR 0xff)
    bad = bytes([proof[0] ^ 0xFF]) + proof[1:]
    assert not G.verify(vk, bad, pw)


def test_unsatisfied_witness_is_rejected(circuit, golden):
    case = golden["cases"][0]
    asg = bytearray(bytes.fromhex(case["assignment"]))
    asg[31] ^= 1                               # public input 0 no longer equals secret_0^2
    with pytest.raises(g16.G16Error) as e:
        circuit.prove_assignment(bytes(asg), bytes.fromhex(case["rnd"]))
    assert e.value.code == 3 and "not satisfied" in str(e.value)


def test_mismatched_key_is_rejected(ctx, golden):
    from shielded_pool_pinocchio_solana_b200 import synth
    other = synth.build(900, n_public=2, n_secret=16, n_committed=12, seed=8)
    with pytest.raises(g16.G16Error):
        g16.Circuit(ctx, other.ccs, bytes.fromhex(golden["pk"]))


def test_gpu_setup_matches_oracle_key(ctx, golden):
    """`sunspot setup` stand-in: the GPU-built pk / vk equal the oracle's byte for byte."""
    pk, vk = ctx.setup(bytes.fromhex(golden["ccs"]), b"golden-small")
    assert vk.hex() == golden["vk"]
    assert pk.hex() == golden["pk"]


def test_audit_like_circuit_proves_and_verifies(ctx):
    """Config 2 stand-in (26,000 constraints, domain 2^15, 2 public inputs, 1 commitment): GPU
    setup + GPU proofs, checked by the oracle's pairing verifier and byte framing."""
    import groth16 as G
    from shielded_pool_pinocchio_solana_b200 import synth
    sc = synth.audit_like()
    pk, vk = ctx.setup(sc.ccs, b"audit-like")
    circ = ctx.load_circuit(sc.ccs, pk)
    assert circ.info["domain"] == 1 << 15 and circ.info["nb_constraints"] == 26000
    n = 3
    asg = b"".join(sc.assignment_bytes(i) for i in range(n))
    proofs, pws = circ.prove_batch(asg, n)
    assert all(len(p) == 388 for p in proofs) and all(len(w) == 76 for w in pws)
    vkd = G.read_vk(vk)
    assert len(vkd["K"]) == 4          # like audit_circuit/target/rlwe_audit.vk
    assert G.verify(vkd, proofs[0], pws[0])
    assert G.verify(vkd, proofs[2], pws[2])
    assert not G.verify(vkd, proofs[1], pws[0][:-1] + bytes([pws[0][-1] ^ 1]))
    circ.free()


def test_commitment_over_public_wires_device_path(ctx):
    """A commitment that also covers public inputs takes the DEVICE solver: the committed public values are hashed
    into the challenge on the host from the caller's assignment.  GPU setup = oracle setup (same seed), so the proof
    bytes must equal the oracle's; the product's verifier accepts them and rejects a changed committed input."""
    import hashlib
    import bn254 as B
    import ccs
    import groth16 as G
    import serialize as S
    from shielded_pool_pinocchio_solana_b200 import synth
    sc = synth.build(120, n_public=3, n_secret=16, n_committed=8, seed=11, n_public_committed=2)
    c = ccs.parse_ccs(sc.ccs)
    opk, ovk, _tx = G.setup(c, b"pubcommit")
    pk, vk = ctx.setup(sc.ccs, b"pubcommit")
    assert pk == G.write_pk(opk) and vk == G.write_vk(ovk)
    circ = ctx.load_circuit(sc.ccs, pk)
    assert circ.solver == "gpu", circ.solver
    enc = lambda v: b"".join(S.fr_to_bytes(x) for x in v)
    n = 3
    asgs = [sc.assignment(5 + i) for i in range(n)]
    rnds = [[int.from_bytes(hashlib.sha256(b"pubcommit/%d/%d" % (i, k)).digest(), "big") % B.R for k in range(3)] for i in range(n)]
    proofs, pws = circ.prove_batch(b"".join(enc(a) for a in asgs), n, b"".join(enc(x) for x in rnds))
    for i in range(n):
        want_proof, want_pw, _aux = G.prove(c, opk, asgs[i], *rnds[i])
        assert proofs[i] == want_proof and pws[i] == want_pw, "proof %d" % i
        assert g16.verify(vk, proofs[i], pws[i]) is True
    bad = bytearray(pws[0])
    bad[12 + 31] ^= 1
    assert g16.verify(vk, proofs[0], bytes(bad)) is False
    circ.free()


def test_gpu_and_host_solver_agree(ctx, golden, circuit, monkeypatch):
    """The batched device solver (default) and the C++ host solver give the same proof bytes."""
    assert circuit.solver == "gpu"
    monkeypatch.setenv("G16_HOST_SOLVER", "1")
    host = g16.Circuit(ctx, bytes.fromhex(golden["ccs"]), bytes.fromhex(golden["pk"]))
    monkeypatch.delenv("G16_HOST_SOLVER")
    assert host.solver != "gpu"
    cases = golden["cases"]
    asg = b"".join(bytes.fromhex(c["assignment"]) for c in cases)
    rnd = b"".join(bytes.fromhex(c["rnd"]) for c in cases)
    ph, wh = host.prove_batch(asg, len(cases), rnd)
    pg, wg = circuit.prove_batch(asg, len(cases), rnd)
    assert ph == pg and wh == wg
    assert [p.hex() for p in ph] == [c["proof"] for c in cases]
    host.free()


def test_many_chunks_pipeline(ctx, golden, circuit):
    """More proofs than one device batch: chunks are pipelined through two slots, order preserved."""
    n = 2 * circuit.info["max_batch"] + 5
    cases = [golden["cases"][i % 3] for i in range(n)]
    asg = b"".join(bytes.fromhex(c["assignment"]) for c in cases)
    rnd = b"".join(bytes.fromhex(c["rnd"]) for c in cases)
    proofs, pws = circuit.prove_batch(asg, n, rnd)
    assert [p.hex() for p in proofs] == [c["proof"] for c in cases]
    assert [p.hex() for p in pws] == [c["pw"] for c in cases]
    # a bad witness in the LAST chunk is reported with its global index
    bad = bytearray(asg)
    bad[(n - 1) * circuit.n_values * 32 + 31] ^= 1
    with pytest.raises(g16.G16Error) as e:
        circuit.prove_batch(bytes(bad), n, rnd)
    assert e.value.code == 3 and "proof %d" % (n - 1) in str(e.value)


def test_noir_witness_front_door_and_cli(ctx, tmp_path):
    """`sunspot prove <acir> <witness.gz> <ccs> <pk>` drop-in: g16_prove on a Noir-format witness and
    the g16prove CLI with sunspot's argv write the .proof/.pw files proof.helper.ts:68-69 reads; both
    verify, and the library call equals the assignment call under the same randomness."""
    import subprocess
    import groth16 as G
    from shielded_pool_pinocchio_solana_b200 import synth
    sc = synth.build(400, n_public=3, n_secret=12, n_committed=10, seed=11)
    pk, vk = ctx.setup(sc.ccs, b"front-door")
    circ = ctx.load_circuit(sc.ccs, pk)
    rnd = bytes(range(1, 97))
    p1, w1 = circ.prove(sc.witness_gz(5), rnd)
    p2, w2 = circ.prove_assignment(sc.assignment_bytes(5), rnd)
    assert (p1, w1) == (p2, w2) and len(w1) == 12 + 32 * 3
    assert G.verify(G.read_vk(vk), p1, w1)
    with pytest.raises(g16.G16Error) as e:
        circ.prove(b"not a gzip stream", rnd)
    assert e.value.code == 2
    circ.free()
    # CLI, same argv as sunspot (client/proof.helper.ts:64)
    d = tmp_path / "target"
    d.mkdir()
    for name, data in (("c.json", b"{}"), ("c.gz", sc.witness_gz(6)), ("c.ccs", sc.ccs), ("c.pk", pk)):
        (d / name).write_bytes(data)
    exe = os.path.join(os.path.dirname(g16.LIB_PATH), "g16prove")
    subprocess.check_call([exe, "prove", str(d / "c.json"), str(d / "c.gz"), str(d / "c.ccs"), str(d / "c.pk")])
    proof, pw = (d / "c.proof").read_bytes(), (d / "c.pw").read_bytes()
    assert len(proof) == 388 and G.verify(G.read_vk(vk), proof, pw)
    bad = subprocess.run([exe, "prove", str(d / "c.json"), str(d / "c.ccs"), str(d / "c.ccs"), str(d / "c.pk")], capture_output=True)
    assert bad.returncode != 0 and b"witness" in bad.stderr


def test_native_synth_circuit_and_split_context_world1(ctx):
    """BASELINE.json configs[3] plumbing on one GPU: the natively generated synthetic circuit (g16_synth_ccs) is
    set up, solved, proved and verified; a context that joined a (1-rank) NCCL communicator takes the split-proof
    path (slice = everything, all-gather of one rank) and returns the same bytes."""
    import random
    ccs = g16.synth_ccs(1 << 12, 2, 64, 99)
    pk, vk = ctx.setup(ccs, b"split-test")
    circ = ctx.load_circuit(ccs, pk)
    assert circ.info["nb_constraints"] == 4096 and circ.info["nb_commitments"] == 0 and circ.proof_len == 324
    rng = random.Random(5)
    asg = b"".join(rng.randrange(synth_R()).to_bytes(32, "big") for _ in range(66))
    rnd = bytes(range(3, 99))
    proof, pw = circ.prove_assignment(asg, rnd)
    assert g16.verify(vk, proof, pw)
    wires, _ = g16.solve_assignment(ccs, asg, circ.info["nb_wires"])
    assert circ.prove_wires(wires, 1, rnd) == [proof]
    circ.free()
    split = g16.Context(0)
    split.comm_init(g16.comm_unique_id(), 0, 1)
    c1 = split.load_circuit(ccs, pk)
    assert c1.prove_wires(wires, 1, rnd) == [proof]
    c1.free()
    split.close()


def synth_R():
    from shielded_pool_pinocchio_solana_b200 import synth
    return synth.R


def test_pipelined_groups_wires_and_batch_bit_exact(circuit, golden):
    """More proofs than one witness-solve group (4 device batches): g16_prove_wires and g16_prove_batch go through
    the two-stage pipeline (worker-thread staging, device conversion / solver, pinned result ring, deferred
    serialisation) across several groups and ragged last chunks; every proof must still be the golden one."""
    cases = golden["cases"]
    n = 2 * 4 * circuit.info["max_batch"] + 37
    pick = [cases[(7 * i + i // 5) % len(cases)] for i in range(n)]
    rnd = b"".join(bytes.fromhex(c["rnd"]) for c in pick)
    proofs = circuit.prove_wires(b"".join(bytes.fromhex(c["wires"]) for c in pick), n, rnd)
    assert [p.hex() for p in proofs] == [c["proof"] for c in pick]
    proofs, pws = circuit.prove_batch(b"".join(bytes.fromhex(c["assignment"]) for c in pick), n, rnd)
    assert [p.hex() for p in proofs] == [c["proof"] for c in pick]
    assert [p.hex() for p in pws] == [c["pw"] for c in pick]
    # library-drawn blinding on the wires path: different bytes, valid proofs
    import groth16 as G
    vk = G.read_vk(bytes.fromhex(golden["vk"]))
    (p0,) = circuit.prove_wires(bytes.fromhex(cases[0]["wires"]), 1, None)
    assert p0.hex() != cases[0]["proof"] and G.verify(vk, p0, bytes.fromhex(cases[0]["pw"]))

"""Byte parity AT THE BENCHMARKED SIZES: the CUDA prover against oracle/c (the C restatement of gnark's
CPU prover, loaded into this process) on the exact configurations bench.py measures.

  * audit_like (BASELINE.json configs[1]: 26,000 rows, domain 2^15, 1 commitment), a FULL device batch
    (64 proofs, the windows / side streams / strided NTT passes of the benchmark), injected (r, s, blinder),
    through g16_prove_batch (device solver), g16_prove_wires (host wires) and g16_prove_wires_dev
    (device-resident wires): Ar | Bs | Krs | Commitment | PoK must equal the oracle's bytes.
  * the same with the witness-like scalar mix of SURVEY.md 8(d) (40 % zero, 30 % one, 20 % < 2^8,
    10 % uniform) -- hot buckets, zero/one fast paths.
  * the reference's own withdraw circuit (tests/golden/shielded_pool_verifier.ccs, a copy of
    /root/reference/noir_circuit/target/shielded_pool_verifier.ccs; MSM sizes 4,175 / 12,701 / 12,442 /
    16,383 / 490): GPU setup, load, prove from seeded wire vectors (client/proof.helper.ts:64-69 proves
    exactly this circuit), bytes = oracle/c.

The oracle is the checker only; the product path never touches it.
"""
import hashlib
import os
import random

import pytest

import shielded_pool_pinocchio_solana_b200 as g16
from shielded_pool_pinocchio_solana_b200 import synth

pytestmark = pytest.mark.gpu
R = synth.R
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rnd_for(i):
    """(r, s, blinder) = SHA-256("g16b200/rnd/" || i || k) mod r  """
    return b"".join((int.from_bytes(hashlib.sha256(b"g16b200/rnd/%d/%d" % (i, k)).digest(), "big") % R).to_bytes(32, "big")
                    for k in range(3))


def witness_like_wires(nw, seed):
    """One wire vector drawn from the SURVEY 8(d) mix; wire 0 = 1."""
    rng = random.Random(seed)
    out = bytearray(32 * nw)
    out[31] = 1
    for i in range(1, nw):
        u = rng.random()
        if u < 0.4:
            continue
        v = 1 if u < 0.7 else (rng.randrange(256) if u < 0.9 else rng.randrange(R))
        out[32 * i:32 * i + 32] = v.to_bytes(32, "big")
    return bytes(out)


class Oracle:
    """oracle/c prover for one (ccs, pk): expected proof bytes from full wire vectors."""

    def __init__(self, ccs_bytes, pk_bytes):
        import ccs as occs
        import coracle
        import groth16 as G
        import serialize as S
        coracle.set_threads(os.cpu_count() or 1)
        self.c = occs.parse_ccs(ccs_bytes)
        pk = self.pk = G.read_pk(pk_bytes, self.c)
        self.cc = coracle.CCircuit(self.c, pk)
        self.coracle = coracle
        info = self.c.commitments[0]
        self.committed = list(info["PrivateCommitted"])
        self.basis = b"".join(S.g1_to_bytes(p) for p in pk["commitment_keys"][0]["basis"])
        self.nw = self.c.nb_wires

    def proof(self, wires_be, rnd96):
        """-> the 388 bytes gnark's Proof.WriteRawTo would produce for these wires and (r, s)."""
        r, s = int.from_bytes(rnd96[:32], "big"), int.from_bytes(rnd96[32:64], "big")
        pts = self.cc.prove_from_wires(wires_be, r, s)            # Ar | Bs | Krs | PoK
        cv = b"".join(wires_be[32 * w:32 * w + 32] for w in self.committed)
        commitment = self.coracle.msm(self.basis, cv, "g1")
        return pts[:256] + b"\x00\x00\x00\x01" + commitment + pts[256:320]


def dev_points_to_proof(raw320, commitment):
    """g16_prove_wires_dev output (canonical LE limbs: Ar | Bs.x.c0,c1,y.c0,c1 | Krs | PoK) -> proof bytes."""
    fe = lambda k: raw320[32 * k:32 * k + 32][::-1]
    ar = fe(0) + fe(1)
    bs = fe(3) + fe(2) + fe(5) + fe(4)            # gnark order X.A1 | X.A0 | Y.A1 | Y.A0
    krs = fe(6) + fe(7)
    pok = fe(8) + fe(9)
    return ar + bs + krs + b"\x00\x00\x00\x01" + commitment + pok


@pytest.fixture(scope="module")
def audit(ctx):
    sc = synth.audit_like()
    pk, vk = ctx.setup(sc.ccs, b"parity-at-size")
    circ = ctx.load_circuit(sc.ccs, pk)
    assert circ.info["domain"] == 1 << 15 and circ.info["max_batch"] == 64
    yield sc, circ, Oracle(sc.ccs, pk), vk
    circ.free()


def test_audit_like_full_batch_prove_batch_matches_oracle(audit):
    """64 proofs from assignments (device witness solver, commitment MSM, pipelined groups) = oracle bytes."""
    import ccs as occs
    import groth16 as G
    import serialize as S
    sc, circ, orc, vk = audit
    B = circ.info["max_batch"]
    asg = b"".join(sc.assignment_bytes(7000 + i) for i in range(B))
    rnd = b"".join(rnd_for(i) for i in range(B))
    assert circ.solver == "gpu"
    proofs, pws = circ.prove_batch(asg, B, rnd)
    # wires from the HOST solver (independent of the device solver used above) feed the oracle prover
    wires = circ.witness_batch(asg, B, rnd)
    nwb = orc.nw * 32
    # ... and the host solver itself is pinned on the big-integer oracle for one proof of the batch
    w0, _ = G.solve(orc.c, sc.assignment(7000), orc.pk, blinder=int.from_bytes(rnd[64:96], "big"))
    assert wires[:nwb] == b"".join(S.fr_to_bytes(x) for x in w0)
    for i in range(B):
        assert proofs[i] == orc.proof(wires[i * nwb:(i + 1) * nwb], rnd[96 * i:96 * i + 96]), "proof %d" % i
    assert G.verify(G.read_vk(vk), proofs[B - 1], pws[B - 1])


def test_audit_like_witness_mix_wires_host_and_dev(audit, ctx):
    """Witness-like scalar mix (hot buckets: ~30 % ones, ~40 % zeros) through g16_prove_wires and
    g16_prove_wires_dev, full batch, non-zero (r, s): bytes = oracle."""
    import torch
    sc, circ, orc, vk = audit
    B = circ.info["max_batch"]
    nw = orc.nw
    wires = b"".join(witness_like_wires(nw, i) for i in range(B))
    rnd = b"".join(rnd_for(100 + i) for i in range(B))
    expect = [orc.proof(wires[i * nw * 32:(i + 1) * nw * 32], rnd[96 * i:96 * i + 96]) for i in range(B)]
    got = circ.prove_wires(wires, B, rnd)
    for i in range(B):
        assert got[i] == expect[i], "prove_wires proof %d" % i
    d_w = torch.empty((B * nw, 8), dtype=torch.int32, device="cuda")
    ctx.fr_to_device(wires, d_w.data_ptr())
    d_out = torch.empty((B, 80), dtype=torch.int32, device="cuda")
    circ.prove_wires_dev(d_w.data_ptr(), B, d_out.data_ptr(), rnd)
    ctx.sync()
    raw = d_out.cpu().numpy().tobytes()
    for i in range(B):
        assert dev_points_to_proof(raw[320 * i:320 * i + 320], expect[i][260:324]) == expect[i], "prove_wires_dev proof %d" % i


def test_real_withdraw_circuit_shape_on_gpu(ctx):
    """The reference's withdraw .ccs: GPU `setup`, load, and proofs from seeded wire vectors equal the
    oracle's bytes (the three sunspot/gnark-private hints are only needed to SOLVE it, not to prove it)."""
    real = open(os.path.join(GOLD, "shielded_pool_verifier.ccs"), "rb").read()
    pk, vk = ctx.setup(real, b"withdraw-shape")
    circ = ctx.load_circuit(real, pk)
    info = circ.info
    assert (info["nb_constraints"], info["nb_wires"], info["domain"]) == (12452, 12939, 16384)
    assert (info["n_a"], info["n_b"], info["n_k"], info["n_z"], info["n_committed"]) == (4175, 12701, 12442, 16383, 490)
    assert len(vk) == 1296                                   # = noir_circuit/target/shielded_pool_verifier.vk
    orc = Oracle(real, pk)
    B = 16
    nw = info["nb_wires"]
    wires = b"".join(witness_like_wires(nw, 500 + i) if i % 2 else
                     b"".join(random.Random(900 + i).randrange(R).to_bytes(32, "big") for _ in range(nw)) for i in range(B))
    rnd = b"".join(rnd_for(200 + i) for i in range(B))
    got = circ.prove_wires(wires, B, rnd)
    for i in range(B):
        assert got[i] == orc.proof(wires[i * nw * 32:(i + 1) * nw * 32], rnd[96 * i:96 * i + 96]), "proof %d" % i
    circ.free()


def test_device_solver_equals_host_solver_audit_like(audit):
    """The batched device witness solver (what g16_prove_batch uses) and the C++ host solver agree on
    every wire of a full group of audit_like witnesses (two-phase solve, commitment MSM and challenge in between)."""
    sc, circ, orc, vk = audit
    n = 70                                              # more than one device batch, ragged
    asg = b"".join(sc.assignment_bytes(9000 + i) for i in range(n))
    rnd = b"".join(rnd_for(300 + i) for i in range(n))
    assert circ.solver == "gpu"
    assert circ.witness_batch_dev(asg, n, rnd) == circ.witness_batch(asg, n, rnd)


def test_device_solver_kernels_agree(ctx, audit, monkeypatch):
    """Every device solver kernel yields the same wires as the host solver, on both benchmark circuits: the default
    (k_solve_2p for wide levels + k_solve_narrow for runs of thin levels -- the division chain at the end of the
    withdraw circuit), without the thread-per-proof kernel, the lanes-per-row kernel (k_solve_tpi) and the round-1
    thread-per-row kernel (k_solve_levels)."""
    import json
    sc, circ_a, _orc, _vk = audit
    n = 37                                              # ragged: not a multiple of a warp
    asg_a = b"".join(sc.assignment_bytes(9100 + i) for i in range(n))
    raw = open(os.path.join(GOLD, "shielded_pool_verifier.ccs"), "rb").read()
    meta = json.load(open(os.path.join(GOLD, "withdraw_assignments.json")))
    blob = open(os.path.join(GOLD, "withdraw_assignments.bin"), "rb").read()
    nb = meta["n_values"] * 32
    asg_w = b"".join(blob[(i % meta["n"]) * nb:(i % meta["n"] + 1) * nb] for i in range(n))
    pk, _ = ctx.setup(raw, b"withdraw-solver-kernels")
    circ_w = ctx.load_circuit(raw, pk)
    rnd = b"".join(rnd_for(900 + i) for i in range(n))
    for circ, asg in ((circ_a, asg_a), (circ_w, asg_w)):
        assert circ.solver == "gpu"
        want = circ.witness_batch(asg, n, rnd)          # host solver
        for env in ({}, {"G16_SOLVER_NARROW": "0"}, {"G16_SOLVER_TPI": "1"}, {"G16_SOLVER_TPI": "1", "G16_SOLVER_NARROW": "0"},
                    {"G16_SOLVER_CTA": "1"}):
            for k in ("G16_SOLVER_NARROW", "G16_SOLVER_TPI", "G16_SOLVER_CTA"):
                monkeypatch.delenv(k, raising=False)
            for k, v in env.items():
                monkeypatch.setenv(k, v)
            assert circ.witness_batch_dev(asg, n, rnd) == want, env
    circ_w.free()


def test_rows_longer_than_the_product_buffer(ctx, monkeypatch):
    """Rows with more non-unit terms (1,100) than k_solve_2p's shared product buffer holds (1,024) multiply in place:
    the device solver's wires equal the host solver's, with and without the thread-per-proof kernel."""
    from shielded_pool_pinocchio_solana_b200 import synth
    sc = synth.build(60, n_public=2, n_secret=16, n_committed=8, seed=3, dens_b=1100)
    pk, _ = ctx.setup(sc.ccs, b"long-rows")
    circ = ctx.load_circuit(sc.ccs, pk)
    assert circ.solver == "gpu", circ.solver
    n = 5
    asg = b"".join(sc.assignment_bytes(40 + i) for i in range(n))
    rnd = b"".join(rnd_for(950 + i) for i in range(n))
    want = circ.witness_batch(asg, n, rnd)
    assert circ.witness_batch_dev(asg, n, rnd) == want
    monkeypatch.setenv("G16_SOLVER_NARROW", "0")
    assert circ.witness_batch_dev(asg, n, rnd) == want
    circ.free()


def test_withdraw_circuit_solves_on_the_device(ctx, monkeypatch):
    """The reference's withdraw circuit takes the DEVICE solver: its three integer hints (Grumpkin scalar split, limb
    decomposition, emulated product) hang off an input wire, are evaluated on the host per proof and scattered into the
    wire vectors before level 0.  No satisfying witness exists offline, so random inputs in diagnostic mode: every wire
    must equal the host solver's (itself equal to oracle/py, tests/test_hints.py), given the same commitment challenge."""
    import ccs as occs
    raw = open(os.path.join(GOLD, "shielded_pool_verifier.ccs"), "rb").read()
    c = occs.parse_ccs(raw)
    pk, _vk = ctx.setup(raw, b"withdraw-device-solver")
    circ = ctx.load_circuit(raw, pk)
    assert circ.solver == "gpu", circ.solver
    monkeypatch.setenv("G16_SOLVER_DIAG", "1")
    rng = random.Random(17)
    nin = c.nb_public - 1 + c.nb_secret
    n = 3
    asgs = []
    for b in range(n):
        a = [rng.randrange(R) for _ in range(nin)]
        a[27] = rng.randrange(1 << 128) if b % 2 == 0 else rng.randrange(R)
        asgs.append(b"".join(v.to_bytes(32, "big") for v in a))
    rnd = b"".join(rnd_for(500 + i) for i in range(n))
    dev = circ.witness_batch_dev(b"".join(asgs), n, rnd)
    nwb = c.nb_wires * 32
    cw = c.commitments[0]["CommitmentIndex"]
    for b in range(n):
        w_dev = dev[b * nwb:(b + 1) * nwb]
        host, _ = g16.solve_assignment(raw, asgs[b], c.nb_wires, blinder_be=rnd[96 * b + 64:96 * b + 96],
                                       challenges_be=w_dev[32 * cw:32 * cw + 32],
                                       n_committed=len(c.commitments[0]["PrivateCommitted"]))
        assert w_dev == host, "proof %d" % b
    circ.free()


def test_withdraw_real_witnesses_prove_bit_exact_and_verify(ctx):
    """BASELINE.json configs[0] on the GPU: the reference's withdraw circuit, REAL witnesses (tests/golden/
    withdraw_assignments.bin: witness 0 = client/prover-params.toml), from assignments through the DEVICE solver
    (host-evaluated integer hints), a ragged multi-chunk batch.  Proof bytes = oracle/c from the same wires and (r, s);
    every proof passes the product's verifier and the oracle's pairing check; the .pw is the toml's five public inputs."""
    import json
    import groth16 as G
    raw = open(os.path.join(GOLD, "shielded_pool_verifier.ccs"), "rb").read()
    meta = json.load(open(os.path.join(GOLD, "withdraw_assignments.json")))
    blob = open(os.path.join(GOLD, "withdraw_assignments.bin"), "rb").read()
    nb = meta["n_values"] * 32
    fixtures = [blob[i * nb:(i + 1) * nb] for i in range(meta["n"])]
    pk, vk = ctx.setup(raw, b"withdraw-real-gpu")
    circ = ctx.load_circuit(raw, pk)
    assert circ.solver == "gpu" and circ.proof_len == 388 and circ.pw_len == 172
    n = 70                                                   # two device chunks, the second ragged
    asg = b"".join(fixtures[i % len(fixtures)] for i in range(n))
    rnd = b"".join(rnd_for(700 + i) for i in range(n))
    proofs, pws = circ.prove_batch(asg, n, rnd)
    orc = Oracle(raw, pk)
    wires = circ.witness_batch(asg[:len(fixtures) * nb], len(fixtures), rnd[:96 * len(fixtures)])     # host solver
    nwb = orc.nw * 32
    assert circ.witness_batch_dev(asg[:len(fixtures) * nb], len(fixtures), rnd[:96 * len(fixtures)]) == wires
    for i in (0, 1, 2, 3):
        assert proofs[i] == orc.proof(wires[i * nwb:(i + 1) * nwb], rnd[96 * i:96 * i + 96]), "proof %d" % i
    ovk = G.read_vk(vk)
    for i in (0, 3, 64, 69):
        assert g16.verify(vk, proofs[i], pws[i]) and G.verify(ovk, proofs[i], pws[i]), "proof %d" % i
    assert not g16.verify(vk, proofs[0], pws[1])             # another witness's public inputs
    want = meta["prover_params_toml"]
    assert pws[0] == bytes.fromhex("000000050000000000000005") + b"".join(
        bytes.fromhex(want[k]) for k in ("root", "nullifier", "recipient", "amount", "wa_commitment"))
    # the witness-file front door: the full ACIR witness map, as `nargo execute` writes it
    import gzip
    import struct
    import ccs as occs
    c = occs.parse_ccs(raw)
    secret = [int(name.rsplit("_", 1)[1]) for name in c.body["Secret"]]
    vals = [fixtures[0][32 * i:32 * i + 32] for i in range(meta["n_values"])]
    acir = {k: vals[k] for k in range(5)}
    acir.update({k: vals[5 + pos] for pos, k in enumerate(secret)})
    body = struct.pack("<QIQ", 1, 0, len(acir)) + b"".join(struct.pack("<IQ", k, 32) + v for k, v in sorted(acir.items()))
    p0, w0 = circ.prove(gzip.compress(body), rnd[:96])
    assert p0 == proofs[0] and w0 == pws[0]
    # ... and straight from Prover.toml: g16_execute rebuilds the witness file `nargo execute` would write
    toml = open(os.path.join(GOLD, "prover-params.toml"), "rb").read()
    abi_json = open(os.path.join(GOLD, "shielded_pool_verifier.abi.json"), "rb").read()
    p1, w1 = circ.prove(g16.execute(raw, abi_json, toml), rnd[:96])
    assert p1 == proofs[0] and w1 == pws[0]
    circ.free()
    # the CLI with sunspot's argv, Prover.toml in place of the witness file (prove_linux.sh:62-87 without nargo)
    import shutil
    import subprocess
    import tempfile
    exe = os.path.join(os.path.dirname(g16.LIB_PATH), "g16prove")
    with tempfile.TemporaryDirectory() as d:
        for name in ("shielded_pool_verifier.abi.json", "prover-params.toml", "shielded_pool_verifier.ccs"):
            shutil.copy(os.path.join(GOLD, name), d)
        open(os.path.join(d, "shielded_pool_verifier.pk"), "wb").write(pk)
        open(os.path.join(d, "shielded_pool_verifier.vk"), "wb").write(vk)
        j = lambda n: os.path.join(d, n)     # noqa: E731
        r = subprocess.run([exe, "prove", j("shielded_pool_verifier.abi.json"), j("prover-params.toml"),
                            j("shielded_pool_verifier.ccs"), j("shielded_pool_verifier.pk")], capture_output=True)
        assert r.returncode == 0, r.stderr
        assert open(j("shielded_pool_verifier.pw"), "rb").read() == pws[0]
        r = subprocess.run([exe, "verify", j("shielded_pool_verifier.vk"), j("shielded_pool_verifier.proof"),
                            j("shielded_pool_verifier.pw")], capture_output=True)
        assert r.returncode == 0 and b"accepted" in r.stdout

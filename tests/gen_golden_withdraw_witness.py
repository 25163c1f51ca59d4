"""Generates tests/golden/withdraw_assignments.{bin,json}: satisfying assignments of the reference's withdraw circuit
(tests/golden/shielded_pool_verifier.ccs = /root/reference/noir_circuit/target/shielded_pool_verifier.ccs).

    python tests/gen_golden_withdraw_witness.py          # needs /root/reference for witness 0 (prover-params.toml)

Witness 0 uses the PRIVATE inputs of /root/reference/client/prover-params.toml (secret_key, randomness, index,
siblings) plus recipient and amount; its public inputs root / nullifier / wa_commitment and the public key
(owner_x, owner_y) are DERIVED by oracle/py/witness_completion.py and asserted equal to the values that file commits.
Witnesses 1.. use seeded random private inputs.  Every witness is checked row by row and by the strict solver.

.bin : N x 6189 x 32 bytes, big-endian: the assignment (public inputs then secret inputs, `.ccs` order)
.json: N, the ABI inputs of every witness (hex), the toml's committed values"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "py"))
import ccs as occs                    # noqa: E402
import groth16 as G                   # noqa: E402
import witness_completion as WC       # noqa: E402

R = G.R
N = 4
# ABI order (noir_circuit/src/main.nr:39-53): ACIR witnesses 0..25
ABI = ["root", "nullifier", "recipient", "amount", "wa_commitment", "secret_key", "owner_x", "owner_y", "randomness",
       "index"] + ["sibling_%d" % i for i in range(16)]
DERIVED = ("root", "nullifier", "wa_commitment", "owner_x", "owner_y")


def wire_of_abi(c):
    """ACIR witness k -> gnark wire: publics are witnesses 0..4 (wires 1..5), secrets are named __witness_k."""
    secret = [int(n.rsplit("_", 1)[1]) for n in c.body["Secret"]]
    m = {k: 1 + k for k in range(c.nb_public - 1)}
    for pos, k in enumerate(secret):
        m[k] = c.nb_public + pos
    return m


def complete_from_private(c, inputs):
    """inputs: {abi name: int} for everything but DERIVED -> (full wires incl. gnark-internal ones up to the commitment,
    assignment list)"""
    wmap = wire_of_abi(c)
    known = {wmap[k]: inputs[name] % R for k, name in enumerate(ABI) if name in inputs}
    w, unknown, failed, _stuck, _rows, _com = WC.complete(c, known)
    nin = c.nb_public - 1 + c.nb_secret
    assert not failed, failed
    assert all(w[i] is not None for i in range(1 + nin)), "assignment incomplete"
    return w, w[1:1 + nin]


def main():
    c = occs.parse_ccs(open(os.path.join(ROOT, "tests", "golden", "shielded_pool_verifier.ccs"), "rb").read())
    import tomllib
    toml = tomllib.load(open("/root/reference/client/prover-params.toml", "rb"))
    h = lambda x: int(x, 16) if isinstance(x, str) else int(x)     # noqa: E731
    committed = {k: h(toml[k]) for k in ("root", "nullifier", "recipient", "amount", "wa_commitment", "secret_key",
                                         "owner_x", "owner_y", "randomness", "index")}
    for i, s in enumerate(toml["siblings"]):
        committed["sibling_%d" % i] = h(s)
    wmap = wire_of_abi(c)
    metas, blobs = [], []
    for n in range(N):
        if n == 0:
            inputs = {k: v for k, v in committed.items() if k not in DERIVED}
        else:
            rng = random.Random(0xA11CE + n)
            inputs = {"recipient": rng.randrange(1, R), "amount": rng.randrange(1, 1 << 40), "secret_key": rng.randrange(1, 1 << 250),
                      "randomness": rng.randrange(R), "index": rng.randrange(1 << 16)}
            for i in range(16):
                inputs["sibling_%d" % i] = rng.randrange(R)
        w, asg = complete_from_private(c, inputs)
        abi_vals = {name: w[wmap[k]] for k, name in enumerate(ABI)}
        if n == 0:
            for name in DERIVED:
                assert abi_vals[name] == committed[name], "derived %s differs from prover-params.toml" % name
        # strict gnark-order solve (needs a key for the commitment hint) + every row
        pk, _vk, _ = G.setup(c, b"golden-withdraw", fast=True)
        wires, _ = G.solve(c, asg, pk=pk, blinder=7)
        for L, Rr, O in c.rows():
            ev = lambda e: sum(c.coeffs[cid] * (1 if wid == occs.CONST_WIRE else wires[wid]) for cid, wid in e) % R   # noqa: E731
            assert ev(L) * ev(Rr) % R == ev(O)
        metas.append({name: "%064x" % v for name, v in abi_vals.items()})
        blobs.append(b"".join(v.to_bytes(32, "big") for v in asg))
        print("witness %d ok (%d values)" % (n, len(asg)))
    out = os.path.join(ROOT, "tests", "golden")
    open(os.path.join(out, "withdraw_assignments.bin"), "wb").write(b"".join(blobs))
    json.dump({"n": N, "n_values": len(blobs[0]) // 32, "abi_order": ABI, "derived": list(DERIVED), "witnesses": metas,
               "prover_params_toml": {k: "%064x" % v for k, v in committed.items()},
               "source": "witness 0: private inputs of /root/reference/client/prover-params.toml"},
              open(os.path.join(out, "withdraw_assignments.json"), "w"), indent=1)


if __name__ == "__main__":
    main()

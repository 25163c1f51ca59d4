"""Witness ingest: `target/<name>.gz` of `nargo execute` -> the circuit's assignment.

No witness file exists in the reference tree and `nargo` is absent, so the reader cannot be pinned on real bytes; the
formats are written here by INDEPENDENT encoders (struct.pack for bincode, the `msgpack` package for MessagePack) in
every framing the ACVM has used -- plain bincode, a format byte + bincode, a format byte + MessagePack with structs as
maps or as arrays ("compact"), field elements as 32 raw bytes / 64 hex characters / byte arrays -- and must all give
the same assignment, mapped through the `.ccs` Public / Secret name lists.  CPU only."""
import gzip
import os
import random
import struct
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import shielded_pool_pinocchio_solana_b200 as g16             # noqa: E402
from shielded_pool_pinocchio_solana_b200 import synth         # noqa: E402

msgpack = pytest.importorskip("msgpack")
R = synth.R


@pytest.fixture(scope="module")
def circuit():
    return synth.build(60, n_public=3, n_secret=12, commitment=False, seed=4)


def values(n, seed):
    rng = random.Random(seed)
    return [rng.randrange(R) for _ in range(n)]


def bincode(vals, field, extra_items=0):
    out = struct.pack("<Q", 1 + extra_items)
    for item in range(extra_items):                      # an inner call's witness map comes BEFORE main's
        out += struct.pack("<IQ", 1 + item, 2)
        for k in range(2):
            out += struct.pack("<IQ", k, 32) + (7 + k).to_bytes(32, "big")
    out += struct.pack("<IQ", 0, len(vals))
    for k, v in enumerate(vals):
        enc = v.to_bytes(32, "big") if field == "raw" else v.to_bytes(32, "big").hex().encode()
        out += struct.pack("<IQ", k, len(enc)) + enc
    return out


def mp(vals, field, compact):
    def enc(v):
        b = v.to_bytes(32, "big")
        return b if field == "raw" else (b.hex() if field == "hex" else list(b))
    wmap = {k: enc(v) for k, v in enumerate(vals)}
    inner = {7: enc(1), 9: enc(2)}
    if compact:
        doc = [[[1, inner], [0, wmap]]]
    else:
        doc = {"stack": [{"index": 1, "witness": inner}, {"index": 0, "witness": wmap}]}
    return msgpack.packb(doc, use_bin_type=True, strict_types=False)


@pytest.mark.parametrize("framing", ["bincode-raw", "bincode-hex", "bincode-raw-2items", "fmt0-bincode", "fmt1-bincode",
                                     "fmt2-msgpack-raw", "fmt2-msgpack-hex", "fmt3-compact-raw", "fmt3-compact-array",
                                     "bare-msgpack"])
def test_every_framing_gives_the_same_assignment(circuit, framing):
    n = circuit.nb_public - 1 + circuit.nb_secret
    vals = values(n, 3)
    want = b"".join(v.to_bytes(32, "big") for v in vals)      # synth names witnesses 0.. in public-then-secret order
    body = {
        "bincode-raw": lambda: bincode(vals, "raw"),
        "bincode-hex": lambda: bincode(vals, "hex"),
        "bincode-raw-2items": lambda: bincode(vals, "raw", extra_items=2),
        "fmt0-bincode": lambda: b"\x00" + bincode(vals, "raw"),
        "fmt1-bincode": lambda: b"\x01" + bincode(vals, "hex"),
        "fmt2-msgpack-raw": lambda: b"\x02" + mp(vals, "raw", False),
        "fmt2-msgpack-hex": lambda: b"\x02" + mp(vals, "hex", False),
        "fmt3-compact-raw": lambda: b"\x03" + mp(vals, "raw", True),
        "fmt3-compact-array": lambda: b"\x03" + mp(vals, "array", True),
        "bare-msgpack": lambda: mp(vals, "raw", False),
    }[framing]()
    assert g16.witness_to_assignment(circuit.ccs, gzip.compress(body)) == want
    assert circuit.witness_gz(0)[:2] == b"\x1f\x8b"            # the repo's own writer is gzip too


def test_rejections(circuit):
    n = circuit.nb_public - 1 + circuit.nb_secret
    vals = values(n, 5)
    with pytest.raises(g16.G16Error) as e:
        g16.witness_to_assignment(circuit.ccs, b"not gzip")
    assert e.value.code == 2
    with pytest.raises(g16.G16Error) as e:                     # a witness the circuit needs is absent
        g16.witness_to_assignment(circuit.ccs, gzip.compress(bincode(vals[:-1], "raw")))
    assert e.value.code == 2 and "missing" in str(e.value)
    with pytest.raises(g16.G16Error):
        g16.witness_to_assignment(circuit.ccs, gzip.compress(b"\x02" + msgpack.packb({"stack": []})))
    with pytest.raises(g16.G16Error):                          # truncated
        g16.witness_to_assignment(circuit.ccs, gzip.compress(bincode(vals, "raw")[:-5]))


def test_public_inputs_are_the_witnesses_not_named_secret():
    """The withdraw circuit names its 6,184 secrets after the ACIR witnesses they are (__witness_5 ... __witness_23643,
    with gaps: sunspot passes only the witnesses its constraints read); the five witnesses below the first secret
    (root, nullifier, recipient, amount, wa_commitment -- noir_circuit/target/shielded_pool_verifier.json `abi`) are
    the public inputs, in that order."""
    sys.path.insert(0, os.path.join(ROOT, "oracle", "py"))
    import ccs as occs
    real = open(os.path.join(ROOT, "tests", "golden", "shielded_pool_verifier.ccs"), "rb").read()
    c = occs.parse_ccs(real)
    secret = [int(name.rsplit("_", 1)[1]) for name in c.body["Secret"]]
    assert len(secret) == 6184 and secret[0] == 5 and secret[-1] == 23643 and c.body["Public"][1:] == [
        "root", "nullifier", "recipient", "amount", "wa_commitment"]
    vals = values(secret[-1] + 1, 9)                     # a full ACIR witness map, as nargo writes it
    asg = g16.witness_to_assignment(real, gzip.compress(b"\x03" + mp(vals, "raw", True)))
    want = [vals[k] for k in range(5)] + [vals[k] for k in secret]
    assert asg == b"".join(v.to_bytes(32, "big") for v in want)

"""gnark wire encodings (raw / uncompressed) -- TEST ORACLE ONLY (see bn254.py header).

Follows gnark-crypto `ecc/bn254/marshal.go` (third-party, absent; pinned by the byte layout of
the reference's committed /root/reference/noir_circuit/target/shielded_pool_verifier.vk and
/root/reference/audit_circuit/target/rlwe_audit.vk -- SURVEY.md 8a rows a13-a15):
  Fr  : 32 B big-endian canonical
  G1  : X || Y                        (64 B)   infinity: all zeros (RawBytes does not flag it;
                                               0b01 in the top bits is the 32-B COMPRESSED infinity)
  G2  : X.A1 || X.A0 || Y.A1 || Y.A0  (128 B)  infinity: all zeros
"""
from bn254 import P, R


def fr_to_bytes(x): return (x % R).to_bytes(32, 'big')
def fp_to_bytes(x): return (x % P).to_bytes(32, 'big')
def fr_from_bytes(b): return int.from_bytes(b[:32], 'big')


def g1_to_bytes(p):
    if p is None:
        return b'\x00' * 64
    return fp_to_bytes(p[0]) + fp_to_bytes(p[1])


def g1_from_bytes(b):
    if b[0] & 0xc0 == 0x40 or not any(b[:64]):     # 0x40 still accepted on read (round-1 files)
        return None
    assert b[0] & 0xc0 == 0, "compressed encodings are not expected in raw artifacts"
    return (int.from_bytes(b[:32], 'big'), int.from_bytes(b[32:64], 'big'))


def g2_to_bytes(p):
    if p is None:
        return b'\x00' * 128
    (x0, x1), (y0, y1) = p
    return fp_to_bytes(x1) + fp_to_bytes(x0) + fp_to_bytes(y1) + fp_to_bytes(y0)


def g2_from_bytes(b):
    if b[0] & 0xc0 == 0x40 or not any(b[:128]):
        return None
    assert b[0] & 0xc0 == 0
    x1, x0, y1, y0 = (int.from_bytes(b[32 * i:32 * i + 32], 'big') for i in range(4))
    return ((x0, x1), (y0, y1))

"""gnark v0.14.0 R1CS (`.ccs`) container reader/writer -- TEST ORACLE ONLY (see bn254.py header).

Restates gnark `constraint/marshal.go` + `github.com/ronanh/intcomp` section compression
(third-party, absent from /root/reference; pinned by `GnarkVersion: '0.14.0'` inside
/root/reference/noir_circuit/target/shielded_pool_verifier.ccs).  The layout was reverse-
engineered on that committed file and every self-check listed there is
a unit test in tests/test_ccs_format.py.  PARITY: pinned on the reference's own artifact for
the reader; the writer is pinned by round-tripping through the reader.

File layout (little-endian):
  [0,32)   4 x u64: totalLen (= fileSize-32), 0, 14, 0
  [32,64)  4 x u64: levelsLen, instructionsLen, calldataLen, bodyLen
  levels        u64 nLevels, then nLevels C32 streams (ascending instruction indices)
  instructions  C32 blueprintID[], C32 constraintOffset[], C32 wireOffset[], C64 startCallData[]
  calldata      u64 count, then count LEB128 uvarints (u32; 0xFFFFFFFF as wire = constant term)
  body          CBOR map
  coefficients  u64 count, count x 32 B: four LE u64 limbs, Montgomery form (R = 2^256)
"""
import struct

import cbor2

from bn254 import R, MONT_R, inv

CONST_WIRE = 0xFFFFFFFF
BLUEPRINT_HINT = 0
BLUEPRINT_R1C = 1
TAG_HINT, TAG_R1C, TAG_COMMIT = 5309735, 5309736, 5309737


# ------------------------------------------------------------------------------------------
# intcomp streams
# ------------------------------------------------------------------------------------------
def _unzigzag(v):
    return (v >> 1) ^ -(v & 1)


def _decode_stream(words, wbits):
    """words: list of u32 (wbits=32) or u64 (wbits=64) of ONE compressed stream."""
    mask = (1 << wbits) - 1
    if not words:
        return []
    T = words[-1]
    body = words[:len(words) - 1 - T]
    tail = words[len(words) - 1 - T:len(words) - 1]
    out = []
    # --- bit-packed part
    p = 0
    group = 32 if wbits == 32 else 64     # ints per sub-block
    while p < len(body):
        if wbits == 32:
            N, section, init = body[p], body[p + 1], body[p + 2]
            q = p + 3
        else:
            N, section = body[p] & 0xFFFFFFFF, body[p] >> 32
            init = body[p + 1]
            q = p + 2
        end = p + section
        cur = init
        produced = 0
        while produced < N:
            header = body[q] & 0xFFFFFFFF
            q += 1
            for sb in range(4):
                desc = (header >> (24 - 8 * sb)) & 0xFF
                width, zz = desc & 0x7F, desc >> 7
                # `width` words hold `group` deltas of `width` bits each, LSB first
                bits = 0
                for k in range(width):
                    bits |= body[q + k] << (wbits * k)
                q += width
                for i in range(group):
                    dlt = (bits >> (width * i)) & ((1 << width) - 1) if width else 0
                    if zz:
                        dlt = _unzigzag(dlt)
                    cur = (cur + dlt) & mask
                    out.append(cur)
                produced += group
        assert q == end, (q, end)
        p = end
    # --- varbyte part: deltas from 0, bytes big-endian inside each word, LEB128
    if T:
        if wbits == 32:
            count, section = tail[0], tail[1]
            data = tail[2:]
        else:
            count, section = tail[0] & 0xFFFFFFFF, tail[0] >> 32
            data = tail[1:]
        raw = b''.join(w.to_bytes(wbits // 8, 'big') for w in data)
        cur, i = 0, 0
        for _ in range(count):
            v, shift = 0, 0
            while True:
                byte = raw[i]
                i += 1
                v |= (byte & 0x7F) << shift
                shift += 7
                if not byte & 0x80:
                    break
            cur = (cur + v) & mask
            out.append(cur)
    return out


def _read_stream(buf, off, wbits):
    """u64 nWords, then nWords words.  Returns (values, new offset)."""
    n = struct.unpack_from('<Q', buf, off)[0]
    off += 8
    fmt = '<%d%s' % (n, 'I' if wbits == 32 else 'Q')
    words = list(struct.unpack_from(fmt, buf, off))
    off += n * (wbits // 8)
    return _decode_stream(words, wbits), off


def _encode_stream(values, wbits):
    """Simplest valid encoding: everything in the varbyte section (non-negative deltas only)."""
    raw = bytearray()
    prev = 0
    for v in values:
        d = v - prev
        assert d >= 0, "writer only supports non-decreasing streams"
        prev = v
        while True:
            b = d & 0x7F
            d >>= 7
            if d:
                raw.append(b | 0x80)
            else:
                raw.append(b)
                break
    wb = wbits // 8
    while len(raw) % wb:
        raw.append(0x80)
    data = [int.from_bytes(raw[i:i + wb], 'big') for i in range(0, len(raw), wb)]
    if wbits == 32:
        tail = [len(values), len(data) + 2] + data
    else:
        tail = [len(values) | ((len(data) + 1) << 32)] + data
    words = tail + [len(tail)]
    fmt = '<Q%d%s' % (len(words), 'I' if wbits == 32 else 'Q')
    return struct.pack(fmt, len(words), *words)


# ------------------------------------------------------------------------------------------
# container
# ------------------------------------------------------------------------------------------
class Ccs:
    """Decoded constraint system.  All coefficients are canonical integers mod r."""

    def __init__(self):
        self.levels = []          # list of lists of instruction indices
        self.blueprint = []       # per instruction
        self.constraint_offset = []
        self.wire_offset = []
        self.start_calldata = []
        self.calldata = []
        self.body = {}
        self.coeffs = []

    # -- derived sizes
    @property
    def nb_public(self): return len(self.body['Public'])          # includes the ONE wire
    @property
    def nb_secret(self): return len(self.body['Secret'])
    @property
    def nb_internal(self): return self.body['NbInternalVariables']
    @property
    def nb_wires(self): return self.nb_public + self.nb_secret + self.nb_internal
    @property
    def nb_constraints(self): return self.body['NbConstraints']
    @property
    def commitments(self):
        ci = self.body.get('CommitmentInfo')
        if ci is None:
            return []
        v = ci.value if hasattr(ci, 'value') else ci
        return v or []

    def instruction_calldata(self, i):
        s = self.start_calldata[i]
        return self.calldata[s:s + self.calldata[s]]

    def r1c(self, i):
        """-> (L, R, O) lists of (coeffID, wireID) for an R1C instruction."""
        cd = self.instruction_calldata(i)
        nl, nr, no = cd[1], cd[2], cd[3]
        p = 4
        def take(k):
            nonlocal p
            t = [(cd[p + 2 * j], cd[p + 2 * j + 1]) for j in range(k)]
            p += 2 * k
            return t
        L, Rr, O = take(nl), take(nr), take(no)
        assert p == cd[0]
        return L, Rr, O

    def hint(self, i):
        """-> (hintID, inputs [list of linear expressions], out_start, out_end)."""
        cd = self.instruction_calldata(i)
        hint_id, nin = cd[1], cd[2]
        p = 3
        inputs = []
        for _ in range(nin):
            ln = cd[p]
            p += 1
            inputs.append([(cd[p + 2 * j], cd[p + 2 * j + 1]) for j in range(ln)])
            p += 2 * ln
        out_start, out_end = cd[p], cd[p + 1]
        assert p + 2 == cd[0]
        return hint_id, inputs, out_start, out_end

    def rows(self):
        """All R1C rows in constraint order: list of (L, R, O)."""
        out = [None] * self.nb_constraints
        for i, bp in enumerate(self.blueprint):
            if bp == BLUEPRINT_R1C:
                out[self.constraint_offset[i]] = self.r1c(i)
        return out


def parse_ccs(buf):
    c = Ccs()
    total, v0, v1, v2 = struct.unpack_from('<4Q', buf, 0)
    if total != len(buf) - 32 or (v0, v1, v2) != (0, 14, 0):
        raise ValueError('not a gnark 0.14 .ccs (header %r, size %d)' % ((total, v0, v1, v2), len(buf)))
    lv_len, ins_len, cd_len, body_len = struct.unpack_from('<4Q', buf, 32)
    off = 64
    # levels
    end = off + lv_len
    nlev = struct.unpack_from('<Q', buf, off)[0]
    off += 8
    for _ in range(nlev):
        vals, off = _read_stream(buf, off, 32)
        c.levels.append(vals)
    assert off == end, (off, end)
    # instructions
    end = off + ins_len
    c.blueprint, off = _read_stream(buf, off, 32)
    c.constraint_offset, off = _read_stream(buf, off, 32)
    c.wire_offset, off = _read_stream(buf, off, 32)
    c.start_calldata, off = _read_stream(buf, off, 64)
    assert off == end, (off, end)
    # streams are padded to a multiple of the block size: trim to the true instruction count
    ninstr = sum(len(l) for l in c.levels)
    for name in ('blueprint', 'constraint_offset', 'wire_offset', 'start_calldata'):
        setattr(c, name, getattr(c, name)[:ninstr])
    # calldata
    end = off + cd_len
    count = struct.unpack_from('<Q', buf, off)[0]
    off += 8
    cd = []
    for _ in range(count):
        v, shift = 0, 0
        while True:
            b = buf[off]
            off += 1
            v |= (b & 0x7F) << shift
            shift += 7
            if not b & 0x80:
                break
        cd.append(v)
    assert off == end, (off, end)
    c.calldata = cd
    # body
    c.body = cbor2.loads(buf[off:off + body_len])
    off += body_len
    # coefficient table
    ncoef = struct.unpack_from('<Q', buf, off)[0]
    off += 8
    rinv = inv(MONT_R, R)
    for _ in range(ncoef):
        limbs = struct.unpack_from('<4Q', buf, off)
        off += 32
        m = limbs[0] | (limbs[1] << 64) | (limbs[2] << 128) | (limbs[3] << 192)
        c.coeffs.append(m * rinv % R)
    assert off == len(buf), (off, len(buf))
    return c


def write_ccs(c):
    """Serialise a Ccs (same container; streams use the varbyte-only encoding)."""
    lv = struct.pack('<Q', len(c.levels)) + b''.join(_encode_stream(l, 32) for l in c.levels)
    ins = (_encode_stream(c.blueprint, 32) + _encode_stream(c.constraint_offset, 32) +
           _encode_stream(c.wire_offset, 32) + _encode_stream(c.start_calldata, 64))
    cd = bytearray(struct.pack('<Q', len(c.calldata)))
    for v in c.calldata:
        while True:
            b = v & 0x7F
            v >>= 7
            if v:
                cd.append(b | 0x80)
            else:
                cd.append(b)
                break
    body = cbor2.dumps(c.body)
    co = bytearray(struct.pack('<Q', len(c.coeffs)))
    for x in c.coeffs:
        m = x * MONT_R % R
        co += struct.pack('<4Q', *[(m >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)])
    payload = struct.pack('<4Q', len(lv), len(ins), len(cd), len(body)) + lv + ins + bytes(cd) + body + bytes(co)
    return struct.pack('<4Q', len(payload), 0, 14, 0) + payload

"""ctypes loader for oracle/c/libg16oracle.so -- TEST ORACLE / CPU BASELINE ONLY (see g16_oracle.c)."""
import ctypes
import os
import struct

R_MOD = 21888242871839275222246405745257275088548364400416034343698204186575808495617
_HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(_HERE, "..", "c", "libg16oracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(PATH)
        L.oracle_set_threads.restype = ctypes.c_int
        L.oracle_circuit_new.restype = ctypes.c_void_p
        L.oracle_circuit_free.argtypes = [ctypes.c_void_p]
        L.oracle_prove_from_wires.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p]
        _lib = L
    return _lib


def set_threads(t=0):
    return lib().oracle_set_threads(int(t))


def msm(points_be, scalars_be, group="g1"):
    size = 64 if group == "g1" else 128
    n = len(points_be) // size
    out = ctypes.create_string_buffer(size)
    fn = lib().oracle_msm_g1 if group == "g1" else lib().oracle_msm_g2
    fn(points_be, scalars_be, ctypes.c_size_t(n), out)
    return out.raw


def fixed_base(base_be, scalars, group="g1"):
    """[k]base for every k in `scalars` (ints) -> list of point byte strings (gnark raw encoding)."""
    size = 64 if group == "g1" else 128
    n = len(scalars)
    out = ctypes.create_string_buffer(max(1, n) * size)
    sc = b"".join((k % R_MOD).to_bytes(32, "big") for k in scalars)
    fn = lib().oracle_fixed_base_g1 if group == "g1" else lib().oracle_fixed_base_g2
    fn(base_be, sc, ctypes.c_size_t(n), out, ctypes.c_size_t(size))
    return [out.raw[size * i:size * (i + 1)] for i in range(n)]


def ntt(values_be, logn, inverse=False, coset=False):
    buf = ctypes.create_string_buffer(values_be, len(values_be))
    lib().oracle_ntt(buf, ctypes.c_uint(logn), int(inverse), int(coset))
    return buf.raw


def compute_h(abc_be, logn):
    out = ctypes.create_string_buffer(32 << logn)
    lib().oracle_compute_h(abc_be, ctypes.c_uint(logn), out)
    return out.raw


class CCircuit:
    """Arrays of a parsed circuit + proving key handed to the C oracle (prove from full wires)."""

    def __init__(self, c, pk):
        """c: ccs.Ccs, pk: dict as produced by groth16.setup / read from a .pk by the oracle."""
        import serialize as S
        from ccs import CONST_WIRE
        rows = c.rows()
        mats = []
        for side in range(3):
            rowptr, wires, coeff = [0], [], bytearray()
            for row in rows:
                for cid, wid in row[side]:
                    wires.append(0 if wid == CONST_WIRE else wid)
                    coeff += S.fr_to_bytes(c.coeffs[cid])
                rowptr.append(len(wires))
            mats.append((rowptr, wires, bytes(coeff)))
        u32 = lambda v: (ctypes.c_uint32 * max(1, len(v)))(*v)
        g1s = lambda pts: b"".join(S.g1_to_bytes(p) for p in pts)
        nw = c.nb_wires
        mapA = [i for i in range(nw) if not pk["infinity_a"][i]]
        mapB = [i for i in range(nw) if not pk["infinity_b"][i]]
        mapK = list(pk["k_wires"])
        keys = pk["commitment_keys"]
        pok = keys[0]["basis_exp_sigma"] if keys else []
        mapP = list(c.commitments[0]["PrivateCommitted"]) if keys else []
        logn = pk["domain"].bit_length() - 1
        self._keep = [u32(m[0]) for m in mats] + [u32(m[1]) for m in mats] + [u32(mapA), u32(mapB), u32(mapK), u32(mapP)]
        k = self._keep
        self.handle = lib().oracle_circuit_new(
            ctypes.c_uint32(nw), ctypes.c_uint32(c.nb_constraints), ctypes.c_uint32(logn), ctypes.c_uint32(c.nb_public),
            k[0], k[3], mats[0][2], k[1], k[4], mats[1][2], k[2], k[5], mats[2][2],
            g1s(pk["A"]), k[6], ctypes.c_uint32(len(mapA)),
            g1s(pk["B1"]), b"".join(S.g2_to_bytes(p) for p in pk["B2"]), k[7], ctypes.c_uint32(len(mapB)),
            g1s(pk["K"]), k[8], ctypes.c_uint32(len(mapK)),
            g1s(pk["Z"]), ctypes.c_uint32(len(pk["Z"])),
            g1s(pok), k[9], ctypes.c_uint32(len(mapP)),
            S.g1_to_bytes(pk["alpha1"]), S.g1_to_bytes(pk["beta1"]), S.g1_to_bytes(pk["delta1"]),
            S.g2_to_bytes(pk["beta2"]), S.g2_to_bytes(pk["delta2"]))

    def prove_from_wires(self, wires_be, r, s):
        """-> 320 bytes: Ar | Bs | Krs | PoK (gnark raw encodings)"""
        out = ctypes.create_string_buffer(320)
        rs = r.to_bytes(32, "big") + s.to_bytes(32, "big")
        lib().oracle_prove_from_wires(ctypes.c_void_p(self.handle), wires_be, rs, out)
        return out.raw

    def close(self):
        if self.handle:
            lib().oracle_circuit_free(ctypes.c_void_p(self.handle))
            self.handle = None

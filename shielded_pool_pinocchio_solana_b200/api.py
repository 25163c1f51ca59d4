"""Thin object layer over the C ABI (include/g16b200.h) for tests, bench.py and Python callers.

Byte formats are gnark's (Fr 32 B BE, G1 64 B raw, G2 128 B raw), i.e. what
`sunspot prove` reads and writes (/root/reference/client/proof.helper.ts:58-71).
"""
import ctypes
import os

from . import _lib
from ._lib import check


class Context:
    """One GPU, one context (one process per GPU)."""

    def __init__(self, device=0):
        self.lib = _lib.load()
        self.handle = ctypes.c_void_p()
        ids = (ctypes.c_int * 1)(device)
        check(self.lib.g16_init(ids, 1, ctypes.byref(self.handle)))
        self.device = device

    def close(self):
        if self.handle:
            self.lib.g16_shutdown(self.handle)
            self.handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def comm_init(self, unique_id: bytes, rank, world):
        """Join the NCCL communicator of a single-proof-across-GPUs run (g16_comm_init)."""
        assert len(unique_id) == 128
        _preload_nccl()
        check(self.lib.g16_comm_init(self.handle, unique_id, rank, world))

    def set_deferred_join(self, on=True):
        """prove_wires_dev calls stop joining the context stream (they overlap); call join() before using outputs."""
        check(self.lib.g16_set_deferred_join(self.handle, int(bool(on))))

    def join(self):
        check(self.lib.g16_join(self.handle))

    def set_stream(self, cuda_stream):
        check(self.lib.g16_set_stream(self.handle, ctypes.c_void_p(cuda_stream)))

    def sync(self):
        check(self.lib.g16_sync(self.handle))

    def last_launches(self):
        return self.lib.g16_last_launches(self.handle)

    def measure_imad_peak(self, kind=1):
        out = ctypes.c_double()
        check(self.lib.g16_measure_imad_peak(self.handle, kind, ctypes.byref(out)))
        return out.value

    def profile_enable(self, on=True):
        check(self.lib.g16_profile_enable(self.handle, int(on)))

    def profile_read(self):
        """-> {slot: (ms, launches, units)}; slot 0 = G1 bucket accumulation, 1 = G2."""
        ms, ln, un = ((ctypes.c_double * 8)() for _ in range(3))
        check(self.lib.g16_profile_read(self.handle, ms, ln, un))
        return {i: (ms[i], ln[i], un[i]) for i in range(8)}

    def fr_to_device(self, values_be: bytes, d_ptr):
        check(self.lib.g16_fr_to_device(self.handle, values_be, len(values_be) // 32, ctypes.c_void_p(d_ptr)))

    def generate_points(self, n, seed=0xB200, group="g1") -> bytes:
        """Synthetic bases P_i = [k_i]G computed on the device (gnark raw bytes)."""
        size = 64 if group == "g1" else 128
        out = ctypes.create_string_buffer(n * size)
        check(self.lib.g16_generate_points(self.handle, 0 if group == "g1" else 1, seed, n, out))
        return out.raw

    def ntt(self, values_be: bytes, logn, batch=1, inverse=False, coset=False) -> bytes:
        """Fr NTT on host buffers: forward = natural -> bit-reversed (DIF), inverse = bit-reversed ->
        natural (DIT, scaled by 1/n); coset=True shifts by g = 5 (fr/fft OnCoset)."""
        assert len(values_be) == batch * (32 << logn)
        buf = ctypes.create_string_buffer(values_be, len(values_be))
        check(self.lib.g16_ntt(self.handle, buf, logn, batch, int(inverse), int(coset)))
        return buf.raw

    def ntt_dev(self, d_ptr, logn, batch=1, inverse=False, coset=False):
        check(self.lib.g16_ntt_dev(self.handle, ctypes.c_void_p(d_ptr), logn, batch, int(inverse), int(coset)))

    def compute_h(self, abc_be: bytes, logn, nproofs=1) -> bytes:
        assert len(abc_be) == nproofs * 3 * (32 << logn)
        out = ctypes.create_string_buffer(nproofs * (32 << logn))
        check(self.lib.g16_compute_h(self.handle, abc_be, logn, nproofs, out))
        return out.raw

    def compute_h_dev(self, d_abc_ptr, logn, nproofs=1):
        check(self.lib.g16_compute_h_dev(self.handle, ctypes.c_void_p(d_abc_ptr), logn, nproofs))

    def setup(self, ccs: bytes, seed: bytes):
        """`sunspot setup` on the GPU: -> (pk bytes, vk bytes) in gnark raw format."""
        pl, vl = ctypes.c_size_t(0), ctypes.c_size_t(0)
        check(self.lib.g16_setup(self.handle, ccs, len(ccs), seed, len(seed), None, ctypes.byref(pl), None,
                                 ctypes.byref(vl)))
        pk, vk = ctypes.create_string_buffer(pl.value), ctypes.create_string_buffer(vl.value)
        check(self.lib.g16_setup(self.handle, ccs, len(ccs), seed, len(seed), pk, ctypes.byref(pl), vk,
                                 ctypes.byref(vl)))
        return pk.raw, vk.raw

    def load_circuit(self, ccs: bytes, pk: bytes):
        return Circuit(self, ccs, pk)

    def load_bases(self, points_be: bytes, group="g1", window=0, batch_hint=1):
        return Bases(self, points_be, group, window, batch_hint)


class Bases:
    """Device-resident, window-expanded MSM bases (pk.G1.A / B / K / Z, pk.G2.B ...)."""

    def __init__(self, ctx, points_be, group, window, batch_hint):
        self.ctx = ctx
        self.group = group
        self.point_size = 64 if group == "g1" else 128
        assert len(points_be) % self.point_size == 0
        self.n = len(points_be) // self.point_size
        self.handle = ctypes.c_void_p()
        fn = ctx.lib.g16_bases_load_g1 if group == "g1" else ctx.lib.g16_bases_load_g2
        check(fn(ctx.handle, points_be, self.n, window, batch_hint, ctypes.byref(self.handle)))

    @property
    def window(self):
        return self.ctx.lib.g16_bases_window(self.handle)

    def free(self):
        if self.handle:
            self.ctx.lib.g16_bases_free(self.handle)
            self.handle = ctypes.c_void_p()

    def msm(self, scalars_be: bytes, batch=1) -> bytes:
        """out[b] = sum_i s[b][i] P_i ; host buffers in, gnark raw point bytes out."""
        assert len(scalars_be) == batch * self.n * 32
        out = ctypes.create_string_buffer(batch * self.point_size)
        fn = self.ctx.lib.g16_msm_g1 if self.group == "g1" else self.ctx.lib.g16_msm_g2
        check(fn(self.ctx.handle, self.handle, scalars_be, batch, out))
        return out.raw

    def msm_dev(self, d_scalars_ptr, batch, d_out_ptr, montgomery=True):
        check(self.ctx.lib.g16_msm_dev(self.ctx.handle, self.handle, ctypes.c_void_p(d_scalars_ptr),
                                       1 if montgomery else 0, batch, ctypes.c_void_p(d_out_ptr)))


class Circuit:
    """A constraint system + proving key resident on one GPU (parsed and uploaded once)."""

    INFO = ("nb_constraints", "nb_wires", "nb_public", "nb_secret", "domain", "nb_commitments", "n_a", "n_b", "n_k",
            "n_z", "n_committed", "max_batch", "window_a", "window_b1", "window_kz", "window_b2")

    def __init__(self, ctx, ccs: bytes, pk: bytes, acir_json=None):
        self.ctx = ctx
        self.handle = ctypes.c_void_p()
        check(ctx.lib.g16_circuit_load(ctx.handle, ccs, len(ccs), pk, len(pk), acir_json, ctypes.byref(self.handle)))
        arr = (ctypes.c_uint64 * 16)()
        check(ctx.lib.g16_circuit_info(self.handle, arr))
        self.info = dict(zip(self.INFO, [int(x) for x in arr]))
        self.proof_len = 388 if self.info["nb_commitments"] else 324
        self.pw_len = 12 + 32 * (self.info["nb_public"] - 1)
        self.n_values = self.info["nb_public"] - 1 + self.info["nb_secret"]
        self.solver = ctx.lib.g16_circuit_solver(self.handle).decode()

    def free(self):
        if self.handle:
            self.ctx.lib.g16_circuit_free(self.handle)
            self.handle = ctypes.c_void_p()

    def prove_batch(self, assignments_be: bytes, n, rnd: bytes = None):
        """n proofs from n assignments (public then secret values, 32 B BE each).
        -> (list of proof bytes, list of public-witness bytes)."""
        assert len(assignments_be) == n * self.n_values * 32
        assert rnd is None or len(rnd) == 96 * n
        proofs = ctypes.create_string_buffer(n * self.proof_len)
        pws = ctypes.create_string_buffer(n * self.pw_len)
        check(self.ctx.lib.g16_prove_batch(self.handle, n, assignments_be, self.n_values, rnd, proofs, pws, self.pw_len))
        P, W = proofs.raw, pws.raw
        return ([P[i * self.proof_len:(i + 1) * self.proof_len] for i in range(n)],
                [W[i * self.pw_len:(i + 1) * self.pw_len] for i in range(n)])

    def prove_assignment(self, assignment_be: bytes, rnd: bytes = None):
        p, w = self.prove_batch(assignment_be, 1, rnd)
        return p[0], w[0]

    def prove(self, witness_gz: bytes, rnd: bytes = None):
        """Drop-in for `sunspot prove`: Noir witness (.gz) in, (proof, public witness) bytes out."""
        proof = ctypes.create_string_buffer(self.proof_len)
        pw = ctypes.create_string_buffer(self.pw_len)
        pl, wl = ctypes.c_size_t(self.proof_len), ctypes.c_size_t(self.pw_len)
        check(self.ctx.lib.g16_prove(self.handle, witness_gz, len(witness_gz), rnd, proof, ctypes.byref(pl), pw,
                                     ctypes.byref(wl)))
        return proof.raw[:pl.value], pw.raw[:wl.value]

    def prove_wires(self, wires_be: bytes, n, rnd: bytes = None):
        assert len(wires_be) == n * self.info["nb_wires"] * 32
        proofs = ctypes.create_string_buffer(n * self.proof_len)
        check(self.ctx.lib.g16_prove_wires(self.handle, n, wires_be, rnd, proofs))
        return [proofs.raw[i * self.proof_len:(i + 1) * self.proof_len] for i in range(n)]

    def witness_batch(self, assignments_be: bytes, n, rnd: bytes = None) -> bytes:
        """Full wire vectors (n * nb_wires * 32 B BE): the R1CS solve only."""
        out = ctypes.create_string_buffer(n * self.info["nb_wires"] * 32)
        check(self.ctx.lib.g16_witness_batch(self.handle, n, assignments_be, self.n_values, rnd, out))
        return out.raw

    def witness_batch_dev(self, assignments_be: bytes, n, rnd: bytes = None) -> bytes:
        """Full wire vectors from the batched DEVICE solver (G16Error code 4 when the circuit needs the host solver)."""
        out = ctypes.create_string_buffer(n * self.info["nb_wires"] * 32)
        check(self.ctx.lib.g16_witness_batch_dev(self.handle, n, assignments_be, self.n_values, rnd, out))
        return out.raw

    def prove_wires_dev(self, d_wires_ptr, n, d_out_ptr, rnd: bytes = None):
        """Device-resident wires in, 320-byte proof points out (device); rnd = n*96 B or None (CSPRNG)."""
        assert rnd is None or len(rnd) == 96 * n
        check(self.ctx.lib.g16_prove_wires_dev(self.handle, n, ctypes.c_void_p(d_wires_ptr), rnd,
                                               ctypes.c_void_p(d_out_ptr)))


def solve_assignment(ccs: bytes, assignment_be: bytes, nb_wires, blinder_be=None, challenges_be=b"", n_committed=0):
    """Host-only witness solver (no GPU): -> (wires_be, committed_be)."""
    lib = _lib.load()
    wires = ctypes.create_string_buffer(nb_wires * 32)
    committed = ctypes.create_string_buffer(max(1, n_committed) * 32)
    check(lib.g16_solve_assignment(ccs, len(ccs), assignment_be, len(assignment_be) // 32, blinder_be, challenges_be,
                                   len(challenges_be) // 32, wires, len(wires), committed, n_committed * 32))
    return wires.raw, committed.raw[:n_committed * 32]


def abi_inputs_to_wires(acir_json: bytes, secret_names, n_public, inputs: dict) -> dict:
    """Noir ABI inputs (Prover.toml as a dict: hex / decimal strings, ints, lists) -> {gnark wire id: value}.

    ABI parameters are ACIR witnesses 0, 1, ... in declaration order, arrays flattened (`abi.parameters` of
    target/<name>.json); sunspot maps the public ones to gnark wires 1.. and names the secret ones __witness_<k>
    (`Secret` list of the .ccs).  Parameters missing from `inputs` are left for g16_complete_assignment to derive."""
    import json
    abi = json.loads(acir_json)["abi"]["parameters"]
    wire_of = {k: 1 + k for k in range(n_public - 1)}
    for pos, name in enumerate(secret_names):
        wire_of[int(name.rsplit("_", 1)[1])] = n_public + pos
    to_int = lambda v: int(v, 0) if isinstance(v, str) else int(v)     # noqa: E731
    known, k = {}, 0
    for p in abi:
        length = p["type"].get("length") if p["type"]["kind"] == "array" else None
        vals = inputs.get(p["name"])
        for i in range(length or 1):
            if vals is not None and k in wire_of:
                known[wire_of[k]] = to_int(vals[i] if length else vals)
            k += 1
    return known


def complete_assignment(ccs: bytes, known: dict) -> bytes:
    """Host only ("ACVM-lite"): {wire id: int value} for a subset of the input wires (the program's ABI inputs) -> the full
    public + secret assignment, when the circuit's constraints determine the rest (include/g16b200.h)."""
    lib = _lib.load()
    ids = (ctypes.c_uint32 * len(known))(*known.keys())
    vals = b"".join(int(v).to_bytes(32, "big") for v in known.values())
    n = ctypes.c_size_t(0)
    check(lib.g16_complete_assignment(ccs, len(ccs), ids, vals, len(known), None, ctypes.byref(n)))
    buf = ctypes.create_string_buffer(32 * n.value)
    check(lib.g16_complete_assignment(ccs, len(ccs), ids, vals, len(known), buf, ctypes.byref(n)))
    return buf.raw


def execute(ccs: bytes, acir_json: bytes, prover_toml: bytes) -> bytes:
    """Host only: Prover.toml + program ABI + .ccs -> witness file (gzip), for circuits whose constraints determine their
    witnesses (`nargo execute` stand-in; include/g16b200.h g16_execute)."""
    lib = _lib.load()
    n = ctypes.c_size_t(0)
    check(lib.g16_execute(ccs, len(ccs), acir_json, len(acir_json), prover_toml, len(prover_toml), None, ctypes.byref(n)))
    buf = ctypes.create_string_buffer(n.value)
    check(lib.g16_execute(ccs, len(ccs), acir_json, len(acir_json), prover_toml, len(prover_toml), buf, ctypes.byref(n)))
    return buf.raw[:n.value]


def witness_to_assignment(ccs: bytes, witness_gz: bytes) -> bytes:
    """Host only: Noir witness file -> public + secret assignment (32 B big-endian values, `.ccs` order)."""
    lib = _lib.load()
    n = ctypes.c_size_t(0)
    check(lib.g16_witness_to_assignment(ccs, len(ccs), witness_gz, len(witness_gz), None, ctypes.byref(n)))
    buf = ctypes.create_string_buffer(32 * n.value)
    check(lib.g16_witness_to_assignment(ccs, len(ccs), witness_gz, len(witness_gz), buf, ctypes.byref(n)))
    return buf.raw


def _preload_nccl():
    """Load the NCCL build PyTorch ships (nvidia-nccl wheel) before libg16b200 dlopens "libnccl.so.2": otherwise the
    system NCCL would be bound first and a LATER `import torch` would resolve its newer symbols against it and fail."""
    import importlib.util
    try:
        spec = importlib.util.find_spec("nvidia.nccl")
        for root in (spec.submodule_search_locations if spec else []):
            path = os.path.join(root, "lib", "libnccl.so.2")
            if os.path.exists(path):
                ctypes.CDLL(path, mode=ctypes.RTLD_GLOBAL)
                return
    except Exception:                               # noqa: BLE001 -- fall back to the library's own search
        pass


def comm_unique_id() -> bytes:
    """NCCL unique id (rank 0 draws it, every rank passes it to Context.comm_init)."""
    _preload_nccl()
    buf = ctypes.create_string_buffer(128)
    check(_lib.load().g16_comm_unique_id(buf))
    return buf.raw


def synth_ccs(n_constraints, n_public=2, n_secret=64, seed=0x5EED22) -> bytes:
    """Large synthetic circuit in the gnark .ccs container, generated natively."""
    lib = _lib.load()
    ln = ctypes.c_size_t(0)
    check(lib.g16_synth_ccs(n_constraints, n_public, n_secret, seed, None, ctypes.byref(ln)))
    buf = ctypes.create_string_buffer(ln.value)
    check(lib.g16_synth_ccs(n_constraints, n_public, n_secret, seed, buf, ctypes.byref(ln)))
    return buf.raw[:ln.value]


def verify(vk: bytes, proof: bytes, pw: bytes) -> bool:
    """`sunspot verify <vk> <proof> <pw>` (host only): True iff the proof is accepted.  Undecodable inputs
    raise G16Error (code 2), like the non-zero exit of the reference binary."""
    lib = _lib.load()
    ok = ctypes.c_int(0)
    check(lib.g16_verify(vk, len(vk), proof, len(proof), pw, len(pw), ctypes.byref(ok)))
    return bool(ok.value)

"""Thin object layer over the C ABI (include/g16b200.h) for tests, bench.py and Python callers.

Byte formats are gnark's (Fr 32 B BE, G1 64 B raw, G2 128 B raw), i.e. what
`sunspot prove` reads and writes (/root/reference/client/proof.helper.ts:58-71).
"""
import ctypes

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

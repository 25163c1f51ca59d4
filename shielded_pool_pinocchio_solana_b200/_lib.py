"""ctypes binding of libg16b200.so (include/g16b200.h).

The library is the product; this module only loads it and declares prototypes.  There is no
Python or CPU fallback: if the shared object is missing the import of any compute entry point
raises, loudly.
"""
import ctypes
import os

# One circuit handle drives ~20 CUDA streams; with the default of 8 hardware connections they alias and the witness
# stage of the next group queues behind kernels of the current one (csrc/capi.cu g16_init).  The variable is read when
# the CUDA context is created, so it is set here, before torch or the library touch the device.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libg16b200.so")

c_u8p = ctypes.POINTER(ctypes.c_uint8)
c_void_pp = ctypes.POINTER(ctypes.c_void_p)

# name -> (restype, argtypes); mirrors include/g16b200.h one to one
PROTOTYPES = {
    "g16_init": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int), ctypes.c_int, c_void_pp]),
    "g16_shutdown": (None, [ctypes.c_void_p]),
    "g16_last_error": (ctypes.c_char_p, []),
    "g16_set_stream": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "g16_sync": (ctypes.c_int, [ctypes.c_void_p]),
    "g16_last_launches": (ctypes.c_int, [ctypes.c_void_p]),
    "g16_measure_imad_peak": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_double)]),
    "g16_bases_load_g1": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int,
                                         ctypes.c_size_t, c_void_pp]),
    "g16_bases_load_g2": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int,
                                         ctypes.c_size_t, c_void_pp]),
    "g16_bases_free": (None, [ctypes.c_void_p]),
    "g16_bases_window": (ctypes.c_int, [ctypes.c_void_p]),
    "g16_msm_g1": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t,
                                  ctypes.c_char_p]),
    "g16_msm_g2": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t,
                                  ctypes.c_char_p]),
    "g16_generate_points": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64, ctypes.c_size_t,
                                           ctypes.c_char_p]),
    "g16_ntt": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_uint, ctypes.c_size_t, ctypes.c_int,
                               ctypes.c_int]),
    "g16_ntt_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint, ctypes.c_size_t, ctypes.c_int,
                                   ctypes.c_int]),
    "g16_compute_h": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_uint, ctypes.c_size_t,
                                     ctypes.c_char_p]),
    "g16_compute_h_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint, ctypes.c_size_t]),
    "g16_msm_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                   ctypes.c_size_t, ctypes.c_void_p]),
    "g16_circuit_load": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p,
                                        ctypes.c_size_t, ctypes.c_char_p, c_void_pp]),
    "g16_setup": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t,
                                 ctypes.c_char_p, ctypes.POINTER(ctypes.c_size_t), ctypes.c_char_p,
                                 ctypes.POINTER(ctypes.c_size_t)]),
    "g16_profile_enable": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "g16_profile_read": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_double),
                                        ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]),
    "g16_fr_to_device": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_void_p]),
    "g16_witness_batch": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t,
                                         ctypes.c_char_p, ctypes.c_char_p]),
    "g16_circuit_solver": (ctypes.c_char_p, [ctypes.c_void_p]),
    "g16_circuit_free": (None, [ctypes.c_void_p]),
    "g16_circuit_info": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64)]),
    "g16_prove": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_char_p,
                                 ctypes.POINTER(ctypes.c_size_t), ctypes.c_char_p, ctypes.POINTER(ctypes.c_size_t)]),
    "g16_prove_assignment": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p,
                                            ctypes.c_char_p, ctypes.POINTER(ctypes.c_size_t), ctypes.c_char_p,
                                            ctypes.POINTER(ctypes.c_size_t)]),
    "g16_prove_batch": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t,
                                       ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t]),
    "g16_prove_wires": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_char_p,
                                       ctypes.c_char_p]),
    "g16_prove_wires_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_char_p,
                                           ctypes.c_void_p]),
    "g16_witness_batch_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t,
                                             ctypes.c_char_p, ctypes.c_char_p]),
    "g16_complete_assignment": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_uint32), ctypes.c_char_p,
                                               ctypes.c_size_t, ctypes.c_char_p, ctypes.POINTER(ctypes.c_size_t)]),
    "g16_execute": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p,
                                   ctypes.c_size_t, ctypes.c_char_p, ctypes.POINTER(ctypes.c_size_t)]),
    "g16_witness_to_assignment": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t,
                                                 ctypes.c_char_p, ctypes.POINTER(ctypes.c_size_t)]),
    "g16_set_deferred_join": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "g16_join": (ctypes.c_int, [ctypes.c_void_p]),
    "g16_comm_unique_id": (ctypes.c_int, [ctypes.c_char_p]),
    "g16_comm_init": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int]),
    "g16_synth_ccs": (ctypes.c_int, [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_char_p,
                                     ctypes.POINTER(ctypes.c_size_t)]),
    "g16_verify": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p,
                                  ctypes.c_size_t, ctypes.POINTER(ctypes.c_int)]),
    "g16_solve_assignment": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t,
                                            ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p,
                                            ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t]),
}

_lib = None


class G16Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"g16b200 error {code}: {msg}")
        self.code = code


def load():
    """Load libg16b200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). This package has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library drift
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().g16_last_error()
        raise G16Error(rc, msg.decode() if msg else "")

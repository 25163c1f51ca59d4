"""g16b200 -- B200-native Groth16/BN254 prover behind the `sunspot prove` boundary.

Product code lives in csrc/ (CUDA, sm_100a) and is reached through the C ABI in
include/g16b200.h; this package is the ctypes mirror used by tests, bench.py and Python callers.
"""
from ._lib import G16Error, LIB_PATH, PROTOTYPES, load  # noqa: F401
from .api import (Bases, Circuit, Context, abi_inputs_to_wires, comm_unique_id, complete_assignment, execute, solve_assignment, synth_ccs,  # noqa: F401
                  verify, witness_to_assignment)

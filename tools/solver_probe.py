"""Runs g16_prove_batch on audit_like for a few groups (target for ncu on the device witness solver)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import shielded_pool_pinocchio_solana_b200 as g16
from shielded_pool_pinocchio_solana_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ctx = g16.Context(0)
sc = synth.audit_like()
pk, _ = ctx.setup(sc.ccs, b"solver-probe")
circ = ctx.load_circuit(sc.ccs, pk)
asg = b"".join(sc.assignment_bytes(i) for i in range(64)) * (n // 64)
circ.prove_batch(asg, n)
print("ok", circ.solver)

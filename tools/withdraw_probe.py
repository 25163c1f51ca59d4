"""g16_prove_batch on the real withdraw circuit with the fixture witnesses (ncu target for stage A + stage B)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import shielded_pool_pinocchio_solana_b200 as g16
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ctx = g16.Context(0)
ccs = open(bench.REAL_CCS, "rb").read()
pk, _ = ctx.setup(ccs, b"withdraw-probe")
circ = ctx.load_circuit(ccs, pk)
asg, nv = bench.withdraw_assignments(n)
circ.prove_batch(asg, n)
print("ok", circ.solver, circ.info)

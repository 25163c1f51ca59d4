"""Host-side timeline of the end-to-end calls (G16_TRACE=1 python tools/e2e_trace.py [chunks]).
Prints wall-clock per call; the library's trace lines go to stderr."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import shielded_pool_pinocchio_solana_b200 as g16
from shielded_pool_pinocchio_solana_b200 import synth
sys.path.insert(0, ROOT)
import bench

chunks = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ctx = g16.Context(0)
real_ccs = open(bench.REAL_CCS, "rb").read()
sc_a = synth.audit_like()
pk_w, _ = ctx.setup(real_ccs, b"bench-withdraw")
pk_a, _ = ctx.setup(sc_a.ccs, b"bench-audit-like")
circ_w = ctx.load_circuit(real_ccs, pk_w)
circ_a = ctx.load_circuit(sc_a.ccs, pk_a)
B = 64
asg = b"".join(sc_a.assignment_bytes(i) for i in range(B))
asg_all = asg * chunks
wires = bench.witness_like_wires(B, circ_w.info["nb_wires"], 3)
rnd = bench.bench_rnd(b"w", B)
circ_a.prove_batch(asg_all, B * chunks)
wires_all, rnd_all = wires * chunks, rnd * chunks
circ_w.prove_wires(wires_all, B * chunks, rnd_all)
torch.cuda.synchronize()
print("=== traced prove_batch", file=sys.stderr, flush=True)
t0 = time.time(); circ_a.prove_batch(asg_all, B * chunks); t1 = time.time()
print("prove_batch: %.1f ms per 64-proof chunk" % ((t1 - t0) * 1e3 / chunks))
print("=== traced prove_wires", file=sys.stderr, flush=True)
t0 = time.time(); circ_w.prove_wires(wires_all, B * chunks, rnd_all); t1 = time.time()
print("prove_wires: %.1f ms per 64-proof chunk" % ((t1 - t0) * 1e3 / chunks))
t0 = time.time(); circ_w.prove_wires(wires, B, rnd); t1 = time.time()
print("prove_wires, one chunk alone: %.1f ms" % ((t1 - t0) * 1e3))
asg_w, _nv = bench.withdraw_assignments(B * chunks)
circ_w.prove_batch(asg_w, B * chunks)
print("=== traced withdraw prove_batch (real witnesses)", file=sys.stderr, flush=True)
t0 = time.time(); circ_w.prove_batch(asg_w, B * chunks); t1 = time.time()
print("withdraw prove_batch (real witnesses, device solver): %.1f ms per 64-proof chunk" % ((t1 - t0) * 1e3 / chunks))
wires_real = circ_w.witness_batch(asg_w[:B * _nv * 32], B)
t0 = time.time(); circ_w.prove_wires(wires_real * chunks, B * chunks, rnd_all); t1 = time.time()
print("withdraw prove_wires (real wires): %.1f ms per 64-proof chunk" % ((t1 - t0) * 1e3 / chunks))
os.environ["G16_SOLVE_OVERLAP"] = "0"
print("=== traced prove_batch, no overlap", file=sys.stderr, flush=True)
t0 = time.time(); circ_a.prove_batch(asg_all, B * chunks); t1 = time.time()
print("prove_batch (solver not overlapped): %.1f ms per 64-proof chunk" % ((t1 - t0) * 1e3 / chunks))

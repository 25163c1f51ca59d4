"""Device-resident proving time per 64-proof chunk (g16_prove_wires_dev) for the two bench circuits, witness-like and
uniform wires: python tools/dev_step.py [reps].  Honors G16_WINDOW_DELTA (window experiments)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import shielded_pool_pinocchio_solana_b200 as g16
from shielded_pool_pinocchio_solana_b200 import synth
import bench

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
ctx = g16.Context(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)
ctx.set_deferred_join(os.environ.get("G16_DEFER", "1") != "0")
if os.environ.get("G16_SERIAL"):
    ctx.profile_enable(True)        # single stream, no overlap: clean per-phase durations in the timeline
real_ccs = open(bench.REAL_CCS, "rb").read()
sc_a = synth.audit_like()
B = 64


def uniform_wires(n_vectors, nw, seed):
    rng = np.random.default_rng(seed)
    limbs = rng.integers(0, 1 << 32, size=(n_vectors * nw, 8), dtype=np.uint64).astype(np.uint32)
    limbs[:, 7] &= 0x0FFFFFFF
    limbs[::nw] = 0
    limbs[::nw, 0] = 1
    return np.ascontiguousarray(limbs[:, ::-1]).astype(">u4").tobytes()


out = {}
for name, ccs in (("withdraw", real_ccs), ("audit_like", sc_a.ccs)):
    pk, _ = ctx.setup(ccs, b"dev-step-" + name.encode())
    circ = ctx.load_circuit(ccs, pk)
    nw = circ.info["nb_wires"]
    rnd = bench.bench_rnd(name.encode(), B)
    obuf = torch.empty((B, 80), dtype=torch.int32, device="cuda")
    for dist, gen in (("mix", bench.witness_like_wires), ("uniform", uniform_wires)):
        sets = []
        for s in range(2):
            t = torch.empty((B * nw, 8), dtype=torch.int32, device="cuda")
            ctx.fr_to_device(gen(B, nw, 10 * s + 3), t.data_ptr())
            sets.append(t)
        for i in range(3):
            circ.prove_wires_dev(sets[i % 2].data_ptr(), B, obuf.data_ptr(), rnd)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(reps):
            circ.prove_wires_dev(sets[i % 2].data_ptr(), B, obuf.data_ptr(), rnd)
        ctx.join()
        e1.record(stream)
        torch.cuda.synchronize()
        out[(name, dist)] = e0.elapsed_time(e1) / reps
        print("=== timeline %s %s" % (name, dist), file=sys.stderr, flush=True)
        ctx.profile_read()      # G16_TIMELINE=1: dumps the event timeline of the loop above
    print(name, {k: circ.info[k] for k in ("window_a", "window_b1", "window_kz", "window_b2")}, flush=True)
    circ.free()
print("delta=%s" % os.environ.get("G16_WINDOW_DELTA", "0"), {"%s/%s" % k: round(v, 2) for k, v in out.items()},
      "pair mix: %.2f ms" % (out[("withdraw", "mix")] + out[("audit_like", "mix")]))

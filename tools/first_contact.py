"""First GPU contact: IMAD peak, MSM parity on the golden vectors, rough MSM timings."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import shielded_pool_pinocchio_solana_b200 as g16

ctx = g16.Context(0)
for kind in (0, 1):
    print("imad peak kind", kind, "%.3e instr/s" % ctx.measure_imad_peak(kind), flush=True)
case = json.load(open(os.path.join(ROOT, "tests/golden/msm_small.json")))
for grp in ("g1", "g2"):
    c = case[grp]
    b = ctx.load_bases(bytes.fromhex(c["points"]), grp, window=0, batch_hint=c["batch"])
    got = b.msm(bytes.fromhex(c["scalars"]), batch=c["batch"])
    print(grp, "golden parity:", got.hex() == c["results"], "window", b.window, flush=True)

stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)
assert stream.cuda_stream != 0
def time_msm(logn, batch, group="g1", reps=5):
    n = 1 << logn
    t0 = time.time()
    pts = ctx.generate_points(n, 0xB200, group)
    t1 = time.time()
    bases = ctx.load_bases(pts, group, batch_hint=batch)
    t2 = time.time()
    g = torch.Generator(device="cuda").manual_seed(1)
    sc = torch.randint(0, 2**31 - 1, (batch * n, 8), dtype=torch.int32, device="cuda", generator=g)
    sc[:, 7] &= 0x0fffffff
    out = torch.empty((batch, 64 if group == "g2" else 32), dtype=torch.int32, device="cuda")
    for _ in range(2):
        bases.msm_dev(sc.data_ptr(), batch, out.data_ptr(), montgomery=False)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(reps):
        bases.msm_dev(sc.data_ptr(), batch, out.data_ptr(), montgomery=False)
    ev[1].record(); torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / reps
    print(f"{group} n=2^{logn} batch={batch} c={bases.window}: {ms:.3f} ms  {batch*n/ms/1e3:.1f} Mpts/s  "
          f"(gen {t1-t0:.2f}s load {t2-t1:.2f}s)", flush=True)
    bases.free()

for logn, batch in ((14, 1), (14, 64), (16, 1), (18, 1), (20, 1), (22, 1)):
    time_msm(logn, batch)
time_msm(14, 16, "g2")
time_msm(18, 1, "g2")

def time_ntt(logn, batch, reps=5):
    n = 1 << logn
    x = torch.randint(0, 2**31 - 1, (batch * n, 8), dtype=torch.int32, device="cuda")
    x[:, 7] &= 0x0fffffff
    for _ in range(2):
        ctx.ntt_dev(x.data_ptr(), logn, batch)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(reps):
        ctx.ntt_dev(x.data_ptr(), logn, batch)
    ev[1].record(); torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / reps
    bf = batch * (n // 2) * logn
    print(f"ntt n=2^{logn} batch={batch}: {ms:.3f} ms  {bf/ms/1e6:.2f} G butterflies/s  {batch*n*64*ctx.last_launches()/ms/1e6:.0f} GB/s", flush=True)
for logn, batch in ((14, 192), (15, 96), (18, 8), (20, 4), (22, 1), (24, 1)):
    time_ntt(logn, batch)

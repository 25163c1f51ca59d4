"""Runs a few MSM launches for one (logn, batch, group) -- target for ncu launch lists."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import shielded_pool_pinocchio_solana_b200 as g16

logn, batch = int(sys.argv[1]), int(sys.argv[2])
group = sys.argv[3] if len(sys.argv) > 3 else "g1"
window = int(sys.argv[4]) if len(sys.argv) > 4 else 0
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
ctx = g16.Context(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
n = 1 << logn
bases = ctx.load_bases(ctx.generate_points(n, 0xB200, group), group, window=window, batch_hint=batch)
g = torch.Generator(device="cuda").manual_seed(1)
sc = torch.randint(-2**31, 2**31 - 1, (batch * n, 8), dtype=torch.int32, device="cuda", generator=g)
sc[:, 7] &= 0x0fffffff
out = torch.empty((batch, 64 if group == "g2" else 32), dtype=torch.int32, device="cuda")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
bases.msm_dev(sc.data_ptr(), batch, out.data_ptr(), montgomery=False)
torch.cuda.synchronize()
ev[0].record()
for _ in range(reps):
    bases.msm_dev(sc.data_ptr(), batch, out.data_ptr(), montgomery=False)
ev[1].record(); torch.cuda.synchronize()
print(f"{group} n=2^{logn} batch={batch} c={bases.window}: {ev[0].elapsed_time(ev[1])/reps:.3f} ms/launch", flush=True)

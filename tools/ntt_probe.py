"""Fr NTT timing: python tools/ntt_probe.py [logn batch reps] -> G butterflies/s (forward DIF, device-resident)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import shielded_pool_pinocchio_solana_b200 as g16

logn = int(sys.argv[1]) if len(sys.argv) > 1 else 15
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 192
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
ctx = g16.Context(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
n = 1 << logn
v = torch.randint(-2**31, 2**31 - 1, (batch * n, 8), dtype=torch.int32, device="cuda")
v[:, 7] &= 0x0FFFFFFF
for _ in range(3):
    ctx.ntt_dev(v.data_ptr(), logn, batch, inverse=False, coset=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(reps):
    ctx.ntt_dev(v.data_ptr(), logn, batch, inverse=False, coset=False)
e1.record(stream); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print("ntt 2^%d x %d: %.3f ms, %.1f G butterflies/s" % (logn, batch, ms, batch * (n / 2) * logn / ms / 1e6))

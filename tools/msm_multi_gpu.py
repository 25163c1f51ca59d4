"""torchrun --nproc-per-node N tools/msm_multi_gpu.py [logn]: one G1 MSM with its point set split over
N GPUs (point ranges resident per rank, one all-gather of the 64-byte partial sums over NCCL)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import shielded_pool_pinocchio_solana_b200 as g16
from shielded_pool_pinocchio_solana_b200.dist import msm_sharded, max_over_ranks

logn = int(sys.argv[1]) if len(sys.argv) > 1 else 18
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = g16.Context(local)
n = 1 << logn
pts = ctx.generate_points(n, 0xB200, "g1")                 # every rank derives the same synthetic key
import random
rng = random.Random(7)
sc = b"".join(rng.randrange(1 << 253).to_bytes(32, "big") for _ in range(n))
cache = {}
out = msm_sharded(ctx, pts, sc, "g1", cache)               # loads this rank's slice
torch.cuda.synchronize()
t0 = time.time()
out2 = msm_sharded(ctx, pts, sc, "g1", cache)
torch.cuda.synchronize()
dt = max_over_ranks(time.time() - t0, "cuda")
if world > 1:
    ref = [None]
    if rank == 0:
        full = ctx.load_bases(pts, "g1")
        ref[0] = full.msm(sc)
    dist.broadcast_object_list(ref, 0)
    assert out == ref[0] == out2, "sharded MSM differs from the single-GPU result"
if rank == 0:
    print("sharded G1 MSM 2^%d over %d GPU(s): ok, %.1f ms per call incl. host byte conversion" % (logn, world, dt * 1e3))
if world > 1:
    dist.destroy_process_group()

"""Single proof split across ranks vs the same proof on one GPU (run under torchrun, one rank per GPU):
   python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/single_proof_multi_gpu.py [logn]
Checks that the sharded proof bytes equal the unsharded ones and that the proof verifies."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import random
import torch
import torch.distributed as dist
import shielded_pool_pinocchio_solana_b200 as g16
from shielded_pool_pinocchio_solana_b200 import dist as gd

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 14
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.cuda.set_device(local)
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
ccs = g16.synth_ccs(1 << logn, 2, 64, 99)
rng = random.Random(5)
asg = b"".join(rng.randrange(R).to_bytes(32, "big") for _ in range(66))
rnd = bytes(range(3, 99))
# unsharded reference proof on this rank's GPU
plain = g16.Context(local)
pk, vk = plain.setup(ccs, b"multi-gpu-test")
c0 = plain.load_circuit(ccs, pk)
nw = c0.info["nb_wires"]
wires, _ = g16.solve_assignment(ccs, asg, nw)
(want,) = c0.prove_wires(wires, 1, rnd)
c0.free(); plain.close()
# sharded
ctx = g16.Context(local)
ctx.comm_init(gd.share_unique_id(rank, world), rank, world)
c1 = ctx.load_circuit(ccs, pk)
(got,) = c1.prove_wires(wires, 1, rnd)
pw = (2).to_bytes(4, "big") + (0).to_bytes(4, "big") + (2).to_bytes(4, "big") + asg[:64]
ok = g16.verify(vk, got, pw)
print("rank %d/%d logn=%d: sharded == unsharded: %s, verifies: %s" % (rank, world, logn, got == want, ok), flush=True)
assert got == want and ok
stream = torch.cuda.Stream()          # a real stream: handle 0 would mean 'the context's own stream'
torch.cuda.set_stream(stream)
res = gd.bench_single_proof(None, torch, stream, rank, world, logn=logn, reps=2)
if rank == 0:
    print(res, flush=True)
c1.free(); ctx.close()
if world > 1:
    dist.destroy_process_group()

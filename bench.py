#!/usr/bin/env python
"""bench.py -- Groth16 proofs/sec for the `sunspot prove` hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our CUDA prover (one rank per GPU)
    python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the reference prover

Workload = BASELINE.json configs[2] (the configuration the proofs/s metric is quoted on): a batch of
withdraw + audit proof PAIRS, sharded one-proof-per-GPU-slot, no data-path collective (weak scaling).
  * withdraw : the reference's own constraint system (tests/golden/shielded_pool_verifier.ccs = /root/reference/
               noir_circuit/target/shielded_pool_verifier.ccs: 12,452 rows, domain 2^14, MSM sizes 4,175 / 12,701 /
               12,442 / 16,383 / 490) with REAL witnesses (tests/golden/withdraw_assignments.bin: witness 0 =
               client/prover-params.toml, three more from seeded private inputs; tests/gen_golden_withdraw_witness.py).
               A real withdraw witness is 92 % full-size field elements (Poseidon), i.e. the worst case for the MSMs,
               not the witness-like mix SURVEY.md 8(d) assumed before a witness existed.
  * audit    : `audit_like` (26,000 rows, domain 2^15, 2 public inputs, 1 BSB22 commitment) -- the audit circuit's
               own .ccs/.pk are missing blobs in the reference.
One step = one device batch per circuit = 64 pairs = 128 proofs; 4096 pairs = 64 steps (--steps 64).

 value : proofs/s, wire vectors already resident in HBM (g16_prove_wires_dev, non-zero blinding)
 e2e   : proofs/s through the reference-facing call with HOST buffers inside the timed region: g16_prove_batch for
         both circuits (host assignments -> H2D -> device witness solver -> commitment -> proof), proof bytes back on
         the host
 roofline : the dominant kernel k_msm_accumulate (G1 + G2), CUDA events on the launching stream; integer-pipe
         bound; work = SURVEY.md 8(d): 23,936 IMAD per G1 point, 71,808 per G2 point; peak = this run's own
         IMAD microbenchmark (MEASURED_PEAKS.json has no integer number)
 cpu_baseline : oracle/c (C + OpenMP restatement of gnark's CPU prover) on this box's cores, bounded sample
 sweep : BASELINE.json configs[4]: standalone G1 / G2 MSM and Fr NTT, 2^16.., uniform and skewed scalars, each with
         its fraction of the IMAD and HBM roofs
 single_proof_2p22 : BASELINE.json configs[3] (see --single): one 2^22-constraint proof, MSM point sets split
         across the ranks, partial sums all-gathered with NCCL inside the library
"""
import argparse
import hashlib
import json
import os

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # before CUDA is initialised: see _lib.py
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "groth16_proofs_per_sec"
UNIT = "proofs/s"
WORKLOAD = ("pairs: withdraw (reference shielded_pool_verifier.ccs, 12452 constraints, domain 2^14, REAL witnesses) + "
            "audit_like (26000 constraints, domain 2^15, 2 public inputs, 1 BSB22 commitment) -- BASELINE.json configs[2]; "
            "one step = 64 pairs per GPU")
IMAD_PER_G1_POINT = 23936.0   # 16 windows x 11 modmul x 136 IMAD
IMAD_PER_G2_POINT = 71808.0
IMAD_PER_BUTTERFLY = 136.0
R_MOD = 21888242871839275222246405745257275088548364400416034343698204186575808495617
REAL_CCS = os.path.join(ROOT, "tests", "golden", "shielded_pool_verifier.ccs")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    self.rows.append([x.strip() for x in line.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = sorted(float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for k, nme in enumerate(names):
                if len(r) > 5 + k and r[5 + k].lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


_REAL_STDOUT = None


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_accumulate_dram_traffic.json")) as f:
            return json.load(f)["dram_bytes_per_launch_avg"]
    except (OSError, KeyError, ValueError):
        return None


def quiet_stdout():
    """stdout carries exactly ONE line (the JSON record): anything a library prints there at the C level (NCCL's
    version banner at communicator creation, for one) is sent to stderr instead."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def witness_like_wires(n_vectors, nw, seed):
    """n_vectors wire vectors of the SURVEY 8(d) mix as big-endian bytes (wire 0 = 1)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    n = n_vectors * nw
    u = rng.random(n)
    limbs = np.zeros((n, 8), dtype=np.uint32)
    limbs[(u >= 0.4) & (u < 0.7), 0] = 1
    small = (u >= 0.7) & (u < 0.9)
    limbs[small, 0] = rng.integers(0, 256, size=int(small.sum()), dtype=np.uint32)
    big = u >= 0.9
    nb = int(big.sum())
    limbs[big] = rng.integers(0, 1 << 32, size=(nb, 8), dtype=np.uint64).astype(np.uint32)
    limbs[big, 7] &= 0x0FFFFFFF                                 # < 2^252 < r
    limbs[::nw] = 0
    limbs[::nw, 0] = 1
    return np.ascontiguousarray(limbs[:, ::-1]).astype(">u4").tobytes()


def withdraw_assignments(n, rotate=0):
    """n real withdraw assignments (the committed fixtures, tiled), big-endian."""
    meta = json.load(open(os.path.join(ROOT, "tests", "golden", "withdraw_assignments.json")))
    blob = open(os.path.join(ROOT, "tests", "golden", "withdraw_assignments.bin"), "rb").read()
    nb = meta["n_values"] * 32
    fx = [blob[i * nb:(i + 1) * nb] for i in range(meta["n"])]
    return b"".join(fx[(i + rotate) % len(fx)] for i in range(n)), meta["n_values"]


def bench_rnd(tag, n):
    return b"".join(hashlib.sha256(b"g16b200/bench/%s/%d" % (tag, k)).digest()[:31].rjust(32, b"\0") for k in range(3 * n))


# ------------------------------------------------------------------------------------------------------
# reference arm: the CPU restatement of the reference prover, independent of libg16b200.so
# ------------------------------------------------------------------------------------------------------
def _oracle_imports():
    sys.path.insert(0, os.path.join(ROOT, "oracle", "py"))
    import ccs as occs
    import coracle
    import groth16 as G
    return occs, coracle, G


def _synth_module():
    """synth.py loaded by path: the reference arm must not import the package (it would map the CUDA library)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("g16_synth", os.path.join(ROOT, "shielded_pool_pinocchio_solana_b200", "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def reference_workload(cache_dir):
    """-> [(name, ccs bytes, pk bytes)] built WITHOUT the product library: keys from the oracle's own setup (C
    fixed-base multiplier), cached on disk between the latency and the sustained legs."""
    occs, coracle, G = _oracle_imports()
    synth = _synth_module()
    out = []
    for name, ccs_bytes in (("withdraw", open(REAL_CCS, "rb").read()), ("audit_like", synth.audit_like().ccs)):
        path = os.path.join(cache_dir, "g16ref_%s_%s.pk" % (name, hashlib.sha256(ccs_bytes).hexdigest()[:12]))
        if os.path.exists(path):
            pk_bytes = open(path, "rb").read()
        else:
            pk, _, _ = G.setup(occs.parse_ccs(ccs_bytes), b"bench-reference-" + name.encode(), fast=True)
            pk_bytes = G.write_pk(pk)
            with open(path + ".tmp%d" % os.getpid(), "wb") as f:
                f.write(pk_bytes)
            os.replace(path + ".tmp%d" % os.getpid(), path)
        out.append((name, ccs_bytes, pk_bytes))
    return out


class CpuProver:
    """oracle/c prove-from-wires for a list of (ccs, pk): one `pair` = one proof of every circuit."""

    def __init__(self, workload, threads):
        occs, coracle, G = _oracle_imports()
        self.cores = coracle.set_threads(threads)
        self.items = []
        for k, (name, ccs_bytes, pk_bytes) in enumerate(workload):
            c = occs.parse_ccs(ccs_bytes)
            cc = coracle.CCircuit(c, G.read_pk(pk_bytes, c))
            nb = c.nb_wires * 32
            if name == "withdraw":      # real witnesses, solved by the oracle's own solver
                asg, nv = withdraw_assignments(4)
                pk = G.read_pk(pk_bytes, c)
                wl = []
                for i in range(4):
                    vals = [int.from_bytes(asg[(i * nv + j) * 32:(i * nv + j + 1) * 32], "big") for j in range(nv)]
                    w, _ = G.solve(c, vals, pk=pk, blinder=1 + i)
                    wl.append(b"".join(v.to_bytes(32, "big") for v in w))
            else:
                wires = witness_like_wires(4, c.nb_wires, 77 + k)
                wl = [wires[i * nb:(i + 1) * nb] for i in range(4)]
            self.items.append((cc, wl))
        self.k = 0

    def prove_pair(self):
        for cc, wl in self.items:
            cc.prove_from_wires(wl[self.k % len(wl)], 12345 + self.k, 67890 + self.k)
        self.k += 1
        return len(self.items)

    def run(self, budget_s, max_pairs=None):
        """-> (proofs done, seconds)"""
        done, t0 = 0, time.time()
        while (done == 0 or time.time() - t0 < budget_s) and (max_pairs is None or done < max_pairs * len(self.items)):
            done += self.prove_pair()
        return done, time.time() - t0


def _sustained_worker(args):
    cache_dir, budget_s, start_at = args
    prover = CpuProver(reference_workload(cache_dir), 1)
    prover.prove_pair()                                  # warm-up
    while time.time() < start_at:
        time.sleep(0.01)
    return prover.run(budget_s)


def cpu_sustained(cache_dir, budget_s, workers):
    """One single-threaded prover per core, all running at once (BASELINE.md 4 'sustained')."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    start_at = time.time() + 30.0                        # workers parse keys and build their circuits first
    with ctx.Pool(workers) as pool:
        res = pool.map(_sustained_worker, [(cache_dir, budget_s, start_at)] * workers)
    proofs = sum(d for d, _ in res)
    span = max(t for _, t in res)
    return proofs, span


def run_reference(args):
    cache_dir = os.environ.get("G16_BENCH_CACHE", "/tmp")
    workload = reference_workload(cache_dir)
    ncores = os.cpu_count() or 1
    per_step = max(1.0, args.cpu_budget / max(1, args.steps))
    prover = CpuProver(workload, ncores)
    for _ in range(max(1, args.warmup)):
        prover.prove_pair()
    lat_n, lat_t = 0, 0.0
    for _ in range(args.steps):                          # each step is a bounded sample of the 64-pair batch
        d, t = prover.run(per_step)
        lat_n += d
        lat_t += t
    latency_rate = lat_n / lat_t
    cores = prover.cores
    del prover
    try:
        sus_n, sus_t = cpu_sustained(cache_dir, args.cpu_budget, ncores)
        sustained_rate = sus_n / sus_t
    except Exception as e:                               # noqa: BLE001 -- keep the arm alive, say what happened
        sus_n, sus_t, sustained_rate = 0, 0.0, 0.0
        sys.stderr.write("sustained leg failed: %r\n" % (e,))
    value = max(latency_rate, sustained_rate)
    sample = ("latency mode: %d proofs in %.1f s with all %d threads on one proof at a time (%.2f proofs/s); sustained mode: "
              "%d proofs in %.1f s with one single-threaded prover per core (%.2f proofs/s); prove-from-wires "
              "(SpMV, H, 6 MSMs) of withdraw (real witnesses) + audit_like pairs, witness solve excluded; gnark-algorithm CPU "
              "restatement (oracle/c), not gnark" % (lat_n, lat_t, cores, latency_rate, sus_n, sus_t, sustained_rate))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * lat_t / max(1, args.steps), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u256 (4x64 Montgomery)", "data": "synthetic",
            "config": {"workload": WORKLOAD, "proofs_per_step_sample": lat_n / max(1, args.steps)},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "latency_mode": latency_rate, "sustained_mode": sustained_rate},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------------
# kernel sweep (BASELINE.json configs[4])
# ------------------------------------------------------------------------------------------------------
def device_scalars(torch, n, dist, seed):
    """n canonical scalars (8 LE limbs) on the device: 'uniform' (< 2^252) or the witness-like 'skewed' mix."""
    gen = torch.Generator(device="cuda").manual_seed(seed)
    sc = torch.randint(-2**31, 2**31 - 1, (n, 8), dtype=torch.int32, device="cuda", generator=gen)
    sc[:, 7] &= 0x0FFFFFFF
    if dist == "skewed":
        u = torch.rand(n, device="cuda", generator=gen)
        small = torch.randint(0, 256, (n,), dtype=torch.int32, device="cuda", generator=gen)
        lo = torch.where(u < 0.4, torch.zeros_like(small), torch.where(u < 0.7, torch.ones_like(small), small))
        keep = (u >= 0.9).unsqueeze(1)
        sc = torch.where(keep, sc, torch.zeros_like(sc))
        sc[:, 0] = torch.where(u >= 0.9, sc[:, 0], lo)
    return sc.contiguous()


def timed_median(torch, stream, fn, warm=3, reps=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def run_sweep(torch, ctx, stream, imad_peak, hbm_gbs, mode):
    out = []
    g1_logs = [16, 18, 20, 22, 24] + ([26] if mode == "full" else [])
    g2_logs = [16, 18, 20, 22] + ([24] if mode == "full" else [])
    ntt_logs = [16, 18, 20, 22, 24, 26]
    for group, logs, per_pt, ptb in (("g1", g1_logs, IMAD_PER_G1_POINT, 64), ("g2", g2_logs, IMAD_PER_G2_POINT, 128)):
        for logn in logs:
            n = 1 << logn
            try:
                bases = ctx.load_bases(ctx.generate_points(n, 0xB200, group), group)
                res = torch.empty((1, 64 if group == "g2" else 32), dtype=torch.int32, device="cuda")
                for dist in ("uniform", "skewed"):
                    sc = device_scalars(torch, n, dist, 0xB201)
                    ms = timed_median(torch, stream, lambda: bases.msm_dev(sc.data_ptr(), 1, res.data_ptr(), montgomery=False),
                                      reps=10 if logn <= 22 else 5)
                    imad = n * per_pt / (ms * 1e-3)
                    out.append({"kernel": "msm_" + group, "logn": logn, "scalars": dist, "window": bases.window, "ms": ms,
                                "mpoints_s": n / ms / 1e3, "frac_imad": imad / imad_peak,
                                "frac_hbm": n * (ptb + 32) / (ms * 1e-3) / (hbm_gbs * 1e9)})
                    del sc
                bases.free()
            except Exception as e:                       # noqa: BLE001
                out.append({"kernel": "msm_" + group, "logn": logn, "error": repr(e)})
    for logn in ntt_logs:
        n = 1 << logn
        try:
            v = device_scalars(torch, n, "uniform", 0xB202)
            ms = timed_median(torch, stream, lambda: ctx.ntt_dev(v.data_ptr(), logn, 1, inverse=False, coset=False),
                              reps=10 if logn <= 24 else 5)
            bfly = n // 2 * logn
            passes = 2 if logn <= 22 else 3
            out.append({"kernel": "ntt_fr", "logn": logn, "ms": ms, "gbutterflies_s": bfly / ms / 1e6,
                        "frac_imad": bfly * IMAD_PER_BUTTERFLY / (ms * 1e-3) / imad_peak,
                        "frac_hbm": passes * 2 * 32 * n / (ms * 1e-3) / (hbm_gbs * 1e9)})
            del v
        except Exception as e:                           # noqa: BLE001
            out.append({"kernel": "ntt_fr", "logn": logn, "error": repr(e)})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    ap.add_argument("--sweep", default="default", choices=["off", "default", "full"])
    ap.add_argument("--single", default="auto", choices=["auto", "on", "off"],
                    help="2^22 single proof split across the ranks (configs[3]); auto = only when N > 1")
    ap.add_argument("--single-log", type=int, default=22)
    args = ap.parse_args()
    quiet_stdout()
    rank, world, local = dist_env()
    if world != args.gpus and world != 1:
        args.gpus = world
    if args.impl == "reference":
        return run_reference(args) if rank == 0 else 0      # rank 0 alone runs the CPU arm

    import torch
    import shielded_pool_pinocchio_solana_b200 as g16
    from shielded_pool_pinocchio_solana_b200 import synth

    use_dist = world > 1
    if use_dist:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    os.environ.setdefault("G16_HOST_THREADS", str(max(1, (os.cpu_count() or 1) // max(1, world))))
    ctx = g16.Context(local)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    ctx.set_deferred_join(True)          # consecutive g16_prove_wires_dev calls overlap; timed() joins before its end event

    # ---- workload: circuits, keys (GPU setup, untimed), wires ---------------------------------------------
    real_ccs = open(REAL_CCS, "rb").read()
    sc_a = synth.audit_like()
    pk_w, _ = ctx.setup(real_ccs, b"bench-withdraw")
    pk_a, _ = ctx.setup(sc_a.ccs, b"bench-audit-like")
    circ_w = ctx.load_circuit(real_ccs, pk_w)
    circ_a = ctx.load_circuit(sc_a.ccs, pk_a)
    B = min(circ_w.info["max_batch"], circ_a.info["max_batch"])
    n_sets = 2                                                # alternate inputs between steps
    dev, asg_w, asg_a, out_buf = {}, [], [], {}
    for name, circ in (("w", circ_w), ("a", circ_a)):
        nw = circ.info["nb_wires"]
        dev[name] = []
        for s in range(n_sets):
            if name == "w":             # real witnesses: assignments -> full wire vectors with the host solver (untimed)
                asg, _nv = withdraw_assignments(B, rotate=s + rank)
                asg_w.append(asg)
                wires = circ.witness_batch(asg, B, bench_rnd(b"wit%d" % s, B))
            else:
                wires = witness_like_wires(B, nw, 1000 * rank + 10 * s + 5)
            t = torch.empty((B * nw, 8), dtype=torch.int32, device="cuda")
            ctx.fr_to_device(wires, t.data_ptr())
            dev[name].append(t)
        out_buf[name] = torch.empty((B, 80), dtype=torch.int32, device="cuda")
    for s in range(n_sets):
        asg_a.append(b"".join(sc_a.assignment_bytes(1000 * rank + 100 * s + i) for i in range(B)))
    rnd_w, rnd_a = bench_rnd(b"w%d" % rank, B), bench_rnd(b"a%d" % rank, B)

    def barrier():
        torch.cuda.synchronize()
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    def step_dev(i, which="wa"):
        n = 0
        if "w" in which:
            circ_w.prove_wires_dev(dev["w"][i % n_sets].data_ptr(), B, out_buf["w"].data_ptr(), rnd_w)
            n += ctx.last_launches()
        if "a" in which:
            circ_a.prove_wires_dev(dev["a"][i % n_sets].data_ptr(), B, out_buf["a"].data_ptr(), rnd_a)
            n += ctx.last_launches()
        return n

    def run_dev(which="wa"):
        return sum(step_dev(i, which) for i in range(args.steps))

    asg_all = b"".join(asg_a[i % n_sets] for i in range(args.steps))
    asg_w_all = b"".join(asg_w[i % n_sets] for i in range(args.steps))

    def run_e2e(which="wa"):
        # ONE g16_prove_batch call per circuit for all K steps' proofs (what a caller with K*B pending proofs does):
        # the library pipelines stage A (H2D of the assignments, device witness solver incl. the host-evaluated
        # integer hints of the withdraw circuit, commitment) of one group with the proving of the previous one
        n = 0
        if "a" in which:
            circ_a.prove_batch(asg_all, B * args.steps)
            n += ctx.last_launches()
        if "w" in which:
            circ_w.prove_batch(asg_w_all, B * args.steps)
            n += ctx.last_launches()
        return n

    def timed(fn):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        launches = fn()
        ctx.join()                       # deferred joins: the context stream now waits for every chunk's assembly
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if use_dist:
            tt = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms, launches

    imad_peak = ctx.measure_imad_peak(0)
    imadw_peak = ctx.measure_imad_peak(1)
    dfma_peak = ctx.measure_imad_peak(2)
    results = {}
    for i in range(args.warmup):
        step_dev(i)
    sampler = ClockSampler(local)
    sampler.start()
    results["dev"], results["launches"] = timed(run_dev)
    sampler.stop_flag = True
    results["dev_audit"], _ = timed(lambda: run_dev("a"))
    results["dev_withdraw"], _ = timed(lambda: run_dev("w"))
    # second pass of the same K steps with per-kernel CUDA events on (single stream, no overlap): roofline line
    ctx.profile_enable(True)
    ctx.profile_read()
    results["dev_serial"], _ = timed(run_dev)
    results["prof"] = ctx.profile_read()
    ctx.profile_enable(False)
    # warm-up with the group size the timed call uses (scratch buffers grow on first use)
    nwarm = min(args.steps, 12)          # 12 chunks = the 4-chunk first group + one full 8-chunk group
    circ_a.prove_batch(asg_all[:nwarm * B * circ_a.n_values * 32], nwarm * B)
    circ_w.prove_batch(asg_w_all[:nwarm * B * circ_w.n_values * 32], nwarm * B)
    results["e2e"], _ = timed(run_e2e)
    results["e2e_audit"], _ = timed(lambda: run_e2e("a"))

    single = None
    want_single = args.single == "on" or (args.single == "auto" and world > 1)
    if want_single:
        try:
            from shielded_pool_pinocchio_solana_b200 import dist as g16dist
            single = g16dist.bench_single_proof(ctx, torch, stream, rank, world, logn=args.single_log)
        except Exception as e:                           # noqa: BLE001
            single = {"error": repr(e)}

    if rank == 0:
        pairs = B * args.steps * args.gpus
        value = 2 * pairs / (results["dev"] * 1e-3)
        e2e = 2 * pairs / (results["e2e"] * 1e-3)
        prof = results["prof"]
        acc_ms = prof[0][0] + prof[1][0]
        acc_launches = prof[0][1] + prof[1][1]
        work = prof[0][2] * IMAD_PER_G1_POINT + prof[1][2] * IMAD_PER_G2_POINT
        achieved = work / (acc_ms * 1e-3) / 1e12 if acc_ms > 0 else 0.0
        h2d = B * (circ_w.n_values * 32 + 96 + 32) + B * (circ_a.n_values * 32 + 96 + 32)
        d2h = 2 * B * (320 + 64 + 4)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": results["dev"] / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u256 (8x32-bit Montgomery limbs, IMAD.WIDE)",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "pairs_per_step_per_gpu": B, "proofs_per_step_per_gpu": 2 * B,
                       "audit_witness_solver": circ_a.solver, "withdraw_witness_solver": circ_w.solver,
                       "withdraw_witnesses": "real (tests/golden/withdraw_assignments.bin, witness 0 = client/prover-params.toml)",
                       "audit_like_witnesses": "solved wires follow the SURVEY 8(d) mix (40 % zero, 30 % one, 20 % < 2^8, 10 % uniform)",
                       "l2": "inputs larger than L2 (per-step working set > 400 MB), two input sets alternated",
                       "windows_withdraw": {k: circ_w.info[k] for k in ("window_a", "window_b1", "window_kz", "window_b2")},
                       "windows_audit": {k: circ_a.info[k] for k in ("window_a", "window_b1", "window_kz", "window_b2")}},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": results["e2e"] / args.steps},
            "audit_only": {"value": pairs / (results["dev_audit"] * 1e-3), "e2e": pairs / (results["e2e_audit"] * 1e-3),
                           "unit": UNIT, "note": "BASELINE.json configs[1] alone (round-1 headline configuration)"},
            "withdraw_only": {"value": pairs / (results["dev_withdraw"] * 1e-3), "unit": UNIT},
            "gpu_launches": results["launches"],
            "roofline": {"bound": "imad", "kernel": "k_msm_accumulate (G1+G2 bucket accumulation)",
                         "achieved": achieved, "peak": imad_peak / 1e12, "unit": "TIMAD/s",
                         "frac": achieved / (imad_peak / 1e12) if imad_peak else None, "traffic": ncu_traffic(),
                         "traffic_source": "profiles/r02_accumulate_dram_traffic.json: average dram read+write bytes per accumulate "
                                           "launch of one 64-proof withdraw chunk, one ncu --set full capture (not live)",
                         "peak_source": "measured in this run: g16_measure_imad_peak(IMAD); IMAD.WIDE.U32.X issues at %.2f T/s, "
                                        "FP64 FMA at %.2f T/s" % (imadw_peak / 1e12, dfma_peak / 1e12),
                         "kernel_ms_per_step": acc_ms / args.steps, "kernel_share_of_step": acc_ms / results["dev_serial"],
                         "measured": "second pass of the same K steps, per-kernel CUDA events, streams serialised (%.2f ms/step)"
                                     % (results["dev_serial"] / args.steps),
                         "launches_timed": acc_launches},
            "clocks": sampler.summary(),
        }
        if single is not None:
            line["single_proof_2p22"] = single
        hbm_gbs = 6548.5
        try:
            hbm_gbs = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        except Exception:                                # noqa: BLE001
            pass
        if args.gpus == 1 and args.sweep != "off":
            circ_w.free()
            circ_a.free()
            torch.cuda.empty_cache()
            sweep = run_sweep(torch, ctx, stream, imad_peak, hbm_gbs, args.sweep)
            line["sweep"] = sweep
            for row in sweep:                            # second headline of BASELINE.json
                if row.get("kernel") == "msm_g1" and row.get("logn") == 22 and row.get("scalars") == "uniform":
                    line["msm_g1_2p22"] = {"metric": "BN254 G1 MSM Mpoints/s at 2^22", "value": row["mpoints_s"],
                                           "unit": "Mpoints/s", "ms": row["ms"], "window": row["window"],
                                           "frac_of_imad_peak": row["frac_imad"],
                                           "frac_of_imad_wide_roof": row["frac_imad"] * imad_peak / (imadw_peak * 136.0 / 128.0),
                                           "note": "fixed bases (window tables resident), uniform 252-bit scalars resident in HBM, "
                                                   "whole MSM (digits, sort, accumulate, reduce); 23,936 IMAD per point (SURVEY.md 8d)"}
        # CPU baseline on this box's host cores, bounded sample (rank 0, N=1 only)
        if args.gpus == 1:
            try:
                prover = CpuProver([("withdraw", real_ccs, pk_w), ("audit_like", sc_a.ccs, pk_a)], os.cpu_count() or 1)
                prover.prove_pair()
                done, dt = prover.run(args.cpu_budget)
                line["cpu_baseline"] = {"value": done / dt, "unit": UNIT, "cores": prover.cores, "kind": "port",
                                        "sample": "%d proofs (withdraw + audit_like pairs, prove-from-wires, witness solve "
                                                  "excluded) in %.1f s with oracle/c, all threads on one proof at a time "
                                                  "(gnark-algorithm CPU restatement, not gnark); `--impl reference` also runs "
                                                  "the sustained one-prover-per-core mode" % (done, dt)}
            except Exception as e:                       # noqa: BLE001 -- a reported baseline must not sink the GPU number
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed: %r" % (e,)}
        emit(line)
    ctx.close()
    if use_dist:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

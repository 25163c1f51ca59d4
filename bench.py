#!/usr/bin/env python
"""bench.py -- Groth16 proofs/sec for the `sunspot prove` hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # our CUDA prover (one rank per GPU)
    python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the reference prover

Workload (BASELINE.json configs[1]): the audit-circuit stand-in `audit_like` -- 26,000 constraints,
FFT domain 2^15, 2 public inputs, one BSB22 commitment, gnark `.ccs` container (the real
audit_circuit .ccs/.pk are missing blobs in the reference).  One step = one device batch of
`max_batch` independent proofs of that circuit; every rank runs the same number of proofs per step
(weak scaling, one proof per GPU slot, no data-path collective).

 value : proofs/s with the full wire vectors already resident in HBM (g16_prove_wires_dev)
 e2e   : proofs/s through the reference-facing call g16_prove_batch with HOST assignments:
         R1CS solve on the host cores, commitment MSM on the GPU, H2D of the wire vectors, device
         pipeline, D2H of the proof points, gnark serialisation -- all inside the timed region.
 roofline : the dominant kernel k_msm_accumulate (G1 and G2 bucket accumulation), timed live
         with CUDA events on the launching stream; integer-pipe bound ("imad"), work = SURVEY.md
         8(d) figures (23,936 IMAD per G1 point, 71,808 per G2 point), peak = this run's own IMAD
         microbenchmark (MEASURED_PEAKS.json holds no integer-pipe number).
 cpu_baseline : oracle/c (C + OpenMP restatement of gnark's CPU prover, "port") on this box's
         host cores, prove-from-wires of the same circuit, bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "groth16_proofs_per_sec"
UNIT = "proofs/s"
WORKLOAD = "audit_like: 26000 constraints, domain 2^15, 2 public inputs, 1 BSB22 commitment (BASELINE.json configs[1])"
IMAD_PER_G1_POINT = 23936.0   # 16 windows x 11 modmul x 136 IMAD
IMAD_PER_G2_POINT = 71808.0


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


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


class CpuReference:
    """oracle/c prove-from-wires (all host threads) -- the CPU restatement of the reference prover."""

    def __init__(self, ccs_bytes, pk_bytes):
        sys.path.insert(0, os.path.join(ROOT, "oracle", "py"))
        import ccs as occs
        import coracle
        import groth16 as G
        # torchrun exports OMP_NUM_THREADS=1: ask for every host core explicitly
        self.cores = coracle.set_threads(os.cpu_count() or 1)
        c = occs.parse_ccs(ccs_bytes)
        self.cc = coracle.CCircuit(c, G.read_pk(pk_bytes, c))
        self.k = 0

    def run(self, wires_list, budget_s, max_proofs=None):
        """-> (proofs done, seconds)"""
        done, t0 = 0, time.time()
        while (done == 0 or time.time() - t0 < budget_s) and (max_proofs is None or done < max_proofs):
            self.cc.prove_from_wires(wires_list[self.k % len(wires_list)], 12345 + self.k, 67890 + self.k)
            self.k += 1
            done += 1
        return done, time.time() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="proofs per step and GPU (default: circuit max_batch)")
    ap.add_argument("--cpu-budget", type=float, default=25.0)
    ap.add_argument("--skip-msm", action="store_true", help="skip the standalone 2^22 G1 MSM measurement")
    args = ap.parse_args()
    rank, world, local = dist_env()
    if world != args.gpus and world != 1:
        args.gpus = world

    import torch
    import shielded_pool_pinocchio_solana_b200 as g16
    from shielded_pool_pinocchio_solana_b200 import synth

    if args.impl == "reference" and rank != 0:
        return 0                                              # rank 0 alone runs the CPU arm
    use_dist = world > 1 and args.impl == "ours"
    if use_dist:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    # host solver threads per rank: the box's cores are shared by all ranks
    os.environ.setdefault("G16_HOST_THREADS", str(max(1, (os.cpu_count() or 1) // max(1, world))))
    ctx = g16.Context(local)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)

    # ---- workload: circuit, key (GPU setup, untimed), assignments, witnesses ----------------------
    sc = synth.audit_like()
    pk, vk = ctx.setup(sc.ccs, b"bench-audit-like")
    circ = ctx.load_circuit(sc.ccs, pk)
    B = args.batch or circ.info["max_batch"]
    B = min(B, circ.info["max_batch"])
    nw = circ.info["nb_wires"]
    n_sets = 2                                                # alternate inputs between steps
    asg_sets = [b"".join(sc.assignment_bytes(1000 * rank + 100 * s + i) for i in range(B)) for s in range(n_sets)]

    if args.impl == "reference":
        wires = circ.witness_batch(asg_sets[0], B)
        wl = [wires[i * nw * 32:(i + 1) * nw * 32] for i in range(B)]
        # each "step" is a bounded sample: the CPU proves as many of the batch as fit the budget
        per_step = max(1.0, args.cpu_budget / max(1, args.steps))
        cpu = CpuReference(sc.ccs, pk)
        cores = cpu.cores
        rates = []
        for it in range(args.warmup + args.steps):
            if it < args.warmup:
                cpu.run(wl, 0.0, max_proofs=1)
            else:
                rates.append(cpu.run(wl, per_step))
        tot_n = sum(d for d, _ in rates)
        tot_t = sum(t for _, t in rates)
        value = tot_n / tot_t
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(1, args.steps),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u256 (4x64 Montgomery)",
                "data": "synthetic", "config": {"workload": WORKLOAD, "proofs_per_step_sample": tot_n / max(1, args.steps)},
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                                 "sample": "%d proofs (prove-from-wires: SpMV, H, 6 MSMs; witness solve excluded) in %.1f s; "
                                           "gnark-algorithm CPU restatement (oracle/c), not gnark" % (tot_n, tot_t)},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # device-resident wire vectors for the kernel-only number
    d_wires = []
    for s in range(n_sets):
        wires = circ.witness_batch(asg_sets[s], B)
        t = torch.empty((B * nw, 8), dtype=torch.int32, device="cuda")
        ctx.fr_to_device(wires, t.data_ptr())
        d_wires.append(t)
    d_out = torch.empty((B, 80), dtype=torch.int32, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    import hashlib
    # non-zero blinding in the device-resident step too: k_finalize's two 254-bit scalar multiplications run
    rnd_dev = b"".join(hashlib.sha256(b"g16b200/bench/%d/%d" % (rank, k)).digest()[:31].rjust(32, b"\0") for k in range(3 * B))

    def step_dev(i):
        circ.prove_wires_dev(d_wires[i % n_sets].data_ptr(), B, d_out.data_ptr(), rnd_dev)

    def step_e2e(i):
        circ.prove_batch(asg_sets[i % n_sets], B)

    imad_peak = ctx.measure_imad_peak(0)
    imadw_peak = ctx.measure_imad_peak(1)

    results = {}
    sampler = ClockSampler(local)
    # e2e goes through ONE g16_prove_batch call for all K steps' proofs: the library pipelines the
    # chunks (host solve of chunk k+1 overlaps the device work of chunk k), exactly what a caller
    # with K*B pending proofs would do.
    asg_all = b"".join(asg_sets[i % n_sets] for i in range(args.steps))

    def run_dev():
        n = 0
        for i in range(args.steps):
            step_dev(i)
            n += ctx.last_launches()
        return n

    def run_e2e():
        circ.prove_batch(asg_all, B * args.steps)
        return ctx.last_launches()

    def timed(fn):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        launches = fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if use_dist:
            tt = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms, launches

    for i in range(args.warmup):
        step_dev(i)
    sampler.start()
    results["dev"], results["launches"] = timed(run_dev)
    sampler.stop_flag = True
    # second pass of the same K steps with per-kernel CUDA events on (single stream, no overlap):
    # the roofline line's kernel time
    ctx.profile_enable(True)
    ctx.profile_read()
    results["dev_serial"], _ = timed(run_dev)
    results["prof"] = ctx.profile_read()
    ctx.profile_enable(False)
    # warm-up with the same group size the timed call uses (scratch buffers grow on first use)
    circ.prove_batch(asg_all[:min(args.steps, 8) * B * circ.n_values * 32], min(args.steps, 8) * B)
    results["e2e"], _ = timed(run_e2e)

    if rank == 0:
        total_proofs = B * args.steps * args.gpus
        value = total_proofs / (results["dev"] * 1e-3)
        e2e = total_proofs / (results["e2e"] * 1e-3)
        prof = results["prof"]
        acc_ms = prof[0][0] + prof[1][0]
        acc_launches = prof[0][1] + prof[1][1]
        work = prof[0][2] * IMAD_PER_G1_POINT + prof[1][2] * IMAD_PER_G2_POINT
        achieved = work / (acc_ms * 1e-3) / 1e12 if acc_ms > 0 else 0.0
        wstride = nw + 8
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": results["dev"] / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u256 (8x32-bit Montgomery limbs, IMAD.WIDE)",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "proofs_per_step_per_gpu": B, "witness_solver": circ.solver,
                       "l2": "inputs larger than L2 (per-step working set > 400 MB), two input sets alternated",
                       "windows": {k: circ.info[k] for k in ("window_a", "window_b1", "window_kz", "window_b2")}},
            "e2e": {"value": e2e, "unit": UNIT,
                    # GPU witness solver: only the assignments, (r,s,blinder) and the challenges go up;
                    # proof points, commitment points and solver status come back
                    "h2d_bytes_per_step": B * (circ.n_values * 32 + 96 + 32) if circ.solver == "gpu"
                    else B * (wstride * 32 + circ.info["n_committed"] * 32),
                    "d2h_bytes_per_step": B * (320 + 64 + 4), "ms_per_step": results["e2e"] / args.steps},
            "gpu_launches": results["launches"],
            "roofline": {"bound": "imad", "kernel": "k_msm_accumulate (G1+G2 bucket accumulation)",
                         "achieved": achieved, "peak": imad_peak / 1e12, "unit": "TIMAD/s",
                         "frac": achieved / (imad_peak / 1e12) if imad_peak else None, "traffic": 280.9e6,
                         "peak_source": "measured in this run: g16_measure_imad_peak(IMAD); IMAD.WIDE.U32.X issues at %.2f T/s"
                                        % (imadw_peak / 1e12),
                         "kernel_ms_per_step": acc_ms / args.steps, "kernel_share_of_step": acc_ms / results["dev_serial"],
                         "measured": "second pass of the same K steps, per-kernel CUDA events, streams serialised (%.2f ms/step)"
                                     % (results["dev_serial"] / args.steps),
                         "ncu": "profiles/r01_ncu_full_k_msm_accumulate_g1_raw_selected.csv: sm__pipe_fmaheavy_cycles_active 82.7 %%, "
                                "dram read+write 281 MB per launch",
                         "launches_timed": acc_launches},
            "clocks": sampler.summary(),
        }
        # second headline of BASELINE.json: standalone BN254 G1 MSM at 2^22 points (N=1 only)
        if args.gpus == 1 and not args.skip_msm:
            try:
                n22 = 1 << 22
                bases = ctx.load_bases(ctx.generate_points(n22, 0xB200, "g1"), "g1")
                gen = torch.Generator(device="cuda").manual_seed(0xB201)
                sc22 = torch.randint(-2**31, 2**31 - 1, (n22, 8), dtype=torch.int32, device="cuda", generator=gen)
                sc22[:, 7] &= 0x0fffffff                          # canonical scalars < 2^252
                out22 = torch.empty((1, 16), dtype=torch.int32, device="cuda")
                for _ in range(3):
                    bases.msm_dev(sc22.data_ptr(), 1, out22.data_ptr(), montgomery=False)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record(stream)
                for _ in range(10):
                    bases.msm_dev(sc22.data_ptr(), 1, out22.data_ptr(), montgomery=False)
                e1.record(stream)
                torch.cuda.synchronize()
                ms22 = e0.elapsed_time(e1) / 10
                imad22 = n22 * IMAD_PER_G1_POINT / (ms22 * 1e-3) / 1e12
                line["msm_g1_2p22"] = {"metric": "BN254 G1 MSM Mpoints/s at 2^22", "value": n22 / ms22 / 1e3, "unit": "Mpoints/s",
                                       "ms": ms22, "window": bases.window, "achieved_timad_s": imad22,
                                       "frac_of_imad_peak": imad22 / (imad_peak / 1e12),
                                       "frac_of_imad_wide_roof": imad22 / (imadw_peak * 136.0 / 128.0 / 1e12),
                                       "note": "fixed bases (window tables resident), uniform 252-bit scalars resident in HBM, whole MSM "
                                               "(digits, sort, accumulate, reduce); work unit 23,936 IMAD per point (SURVEY.md 8d)"}
                bases.free()
                del sc22
            except Exception as e:
                line["msm_g1_2p22"] = {"error": repr(e)}
        # CPU baseline on this box's host cores, bounded sample (rank 0, N=1 only)
        if args.gpus == 1:
            try:
                wires = circ.witness_batch(asg_sets[0], min(B, 16))
                wl = [wires[i * nw * 32:(i + 1) * nw * 32] for i in range(min(B, 16))]
                cpu = CpuReference(sc.ccs, pk)
                cpu.run(wl, 0.0, max_proofs=1)
                done, dt = cpu.run(wl, args.cpu_budget)
                line["cpu_baseline"] = {"value": done / dt, "unit": UNIT, "cores": cpu.cores, "kind": "port",
                                        "sample": "%d proofs of the same circuit (prove-from-wires, witness solve excluded) "
                                                  "in %.1f s with oracle/c (gnark-algorithm CPU restatement, not gnark)" % (done, dt)}
            except Exception as e:  # the oracle is a reported baseline; never let it sink the GPU number
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed: %r" % (e,)}
        print(json.dumps(line))
    circ.free()
    ctx.close()
    if use_dist:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

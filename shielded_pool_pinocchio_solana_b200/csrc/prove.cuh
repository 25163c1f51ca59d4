// prove.cuh -- device-resident circuit handle and the batched Groth16 proving pipeline.
//
// Replaces gnark `groth16.Prove` (backend/groth16/bn254/prove.go) as reached through
// `sunspot prove` (/root/reference/client/proof.helper.ts:64) -- SURVEY.md 3.2.
#pragma once
#include <memory>
#include <vector>

#include "capi.cuh"
#include "ccs.hpp"
#include "gpusolver.cuh"
#include "solver.hpp"

namespace g16 {

// host copy of a parsed gnark ProvingKey (points in canonical little-endian limb form)
struct ProvingKeyHost {
    uint64_t domain = 0;
    G1Affine alpha1, beta1, delta1;
    G2Affine beta2, delta2;
    std::vector<G1Affine> A, B1, Z, K;
    std::vector<G2Affine> B2;
    std::vector<uint8_t> infinity_a, infinity_b;
    struct CommitKey {
        std::vector<G1Affine> basis, basis_exp_sigma;
    };
    std::vector<CommitKey> commitment_keys;
};

int parse_pk(const uint8_t* buf, size_t len, ProvingKeyHost* out);


// per-proof results of the device pipeline (canonical little-endian limbs)
struct ProofPoints {
    G1Affine ar;
    G2Affine bs;
    G1Affine krs;
    G1Affine pok;
};

}  // namespace g16

struct g16_circuit {
    g16_ctx* ctx = nullptr;
    g16::Circuit circ;
    unsigned logn = 0;
    size_t n = 0;           // domain size
    size_t nw = 0;          // wires
    size_t wstride = 0;     // nw + X_COUNT
    size_t max_batch = 0;      // proofs per device proving batch
    size_t solve_batch = 0;    // proofs per witness-solve batch (a multiple of max_batch)
    size_t first_group = 0;   // proofs in the first group of a call (<= solve_batch)
    bool has_commitment = false;
    int unit_ids = 0;       // coefficient ids 0/1/3 are 0/+1/-1 (gnark's fixed table prefix)
    size_t n_committed = 0;        // committed wires held by THIS rank (all of them unless the circuit is split)
    size_t n_committed_total = 0;
    int world = 1;                 // > 1: MSM point sets are split across ranks
    std::vector<uint32_t> committed_wires;
    // R1CS on the device: three CSR matrices sharing one coefficient table
    uint32_t* d_rowptr[3] = {nullptr, nullptr, nullptr};
    uint32_t* d_cid[3] = {nullptr, nullptr, nullptr};
    uint32_t* d_wid[3] = {nullptr, nullptr, nullptr};
    g16::Fr* d_coeffs = nullptr;
    // MSM bases (window-expanded) and scalar maps
    g16::MsmBases<g16::Fp> bA, bB1, bKZ, bCommit, bPok;
    g16::MsmBases<g16::Fp2> bB2;
    uint32_t *d_mapA = nullptr, *d_mapB = nullptr, *d_mapKZ = nullptr, *d_mapPok = nullptr;
    size_t nA = 0, nB = 0, nK = 0, nZ = 0;
    // Scratch sized for max_batch proofs, TWO sets used alternately (`parity`): consecutive chunks are independent on
    // the device, so the latency-bound end of chunk k (K|Z bucket reduction, assembly: ~2 ms on a few hundred
    // threads) runs under the start of chunk k+1.  ev_fin[p] = the chunk that used set p has been assembled
    // (and its outputs copied wherever the caller wanted them); set p is reused only behind it.
    int parity = 0;
    g16::DeviceBuf d_abc[2], d_out[2], d_wdev[2];
    g16::G1Affine* d_tmp_g1[2] = {nullptr, nullptr};   // [4][max_batch]: A, B1, KZ, PoK results (Montgomery)
    g16::G2Affine* d_tmp_g2[2] = {nullptr, nullptr};   // [max_batch]
    g16::G1XYZZ* d_parts[2] = {nullptr, nullptr};      // [2][max_batch]: s*Ar, r*Bs1
    cudaStream_t fin_stream = nullptr;                 // joins the MSM streams; assembly and output copies
    cudaEvent_t ev_fin[2] = {nullptr, nullptr};
    g16::MsmRunner<g16::Fp> g1_kz;                     // K|Z MSM (own scratch: chunks of different circuits overlap)
    // Two pipeline slots: while the device proves chunk k out of slot k%2, the host solves chunk
    // k+1 into the other slot (its commitment MSM runs on `aux_stream` with its own MSM scratch).
    struct Slot {
        g16::DeviceBuf d_wires, d_commit_vals, d_commit_out, d_asg_be, d_rnd_be, d_err, d_chal;
        g16::DeviceBuf d_pre;            // host-evaluated hint outputs (GpuSolverPlan::host_wires)
        g16::DeviceBuf d_wires_be;       // g16_prove_wires: the caller's big-endian wire vectors (grown on first use)
        void* h_stage = nullptr;         // pinned: assignments | rnd | challenges | err
        void* h_wires = nullptr;         // pinned, max_batch * wstride Fr
        cudaEvent_t ready = nullptr;     // wires of this slot are in d_wires
        std::vector<g16::G1Affine> commits;
    } slots[2];
    // results of the proving chunks come back through two pinned buffers: the host serialises chunk k-1 while
    // the device proves chunk k
    g16::ProofPoints* h_pts[2] = {nullptr, nullptr};
    cudaEvent_t ev_done[2] = {nullptr, nullptr};
    cudaStream_t aux_stream = nullptr;
    g16::MsmRunner<g16::Fp> g1_aux;
    g16::GpuSolverPlan plan;             // valid => witnesses are solved on the GPU
    std::string host_solver_reason;
    uint32_t* d_map_commit = nullptr;    // committed wire ids (scalar map of the commitment MSM)
    // The MSMs that only need the wires (A, B1, PoK, B2) run on one side stream EACH next to
    // SpMV -> H -> K|Z on the context stream, and the two scalar multiplications of the Krs assembly on a
    // fifth one behind A and B1: every MSM ends in latency-bound kernels (scan, ordering, bucket reduction,
    // ~1-3 ms of dependent point additions on few threads) that only cost time when nothing runs beside them.
    // The side streams have a higher priority than the stream all bucket accumulations go to (acc_stream).
    enum { SIDE_A = 0, SIDE_B1 = 1, SIDE_POK = 2, SIDE_B2 = 3, SIDE_SM = 4, SIDE_KZ = 5, N_SIDE = 6 };
    cudaStream_t side[N_SIDE] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_h = nullptr, ev_join[N_SIDE] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    g16::MsmRunner<g16::Fp> g1_side[3];   // A, B1, PoK (own scratch each: they run concurrently)
    g16::MsmRunner<g16::Fp2> g2_side;
    cudaStream_t acc_stream = nullptr;    // low priority: every bucket accumulation of this circuit (MsmRunner::acc_stream)
    int last_launches = 0;
    ~g16_circuit();
};

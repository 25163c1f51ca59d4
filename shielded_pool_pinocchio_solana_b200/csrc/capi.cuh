// capi.cuh -- private definitions behind the opaque handles of include/g16b200.h.
#pragma once
#include <mutex>

#include "../../include/g16b200.h"
#include "common.cuh"
#include "msm.cuh"
#include "ntt.cuh"

namespace g16 {

struct DeviceBuf {
    void* ptr = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes);
    ~DeviceBuf();
};

void be32_to_limbs(const uint8_t* be, uint32_t* limbs);
void limbs_to_be32(const uint32_t* limbs, uint8_t* be);
void g1_from_be(const uint8_t* be, G1Affine* out);
void g1_to_be(const G1Affine& p, uint8_t* be);
void g2_from_be(const uint8_t* be, G2Affine* out);
void g2_to_be(const G2Affine& p, uint8_t* be);
// [lo, hi) of rank `rank` among `world` over n items
void shard_range(size_t n, int rank, int world, size_t* lo, size_t* hi);

}  // namespace g16

// Every entry point that touches a context takes its (recursive) mutex: MSM / NTT scratch and the
// streams belong to the context, so concurrent callers on one context are serialised, not corrupted
// (SURVEY.md 8b "Threading").  For parallel host threads that must not wait on each other, use one
// context per thread.
#define G16_LOCK(ctxptr) std::lock_guard<std::recursive_mutex> _g16_lock((ctxptr)->mu)

struct g16_circuit;
struct g16_ctx {
    std::recursive_mutex mu;
    int device = 0;
    int sm_count = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    int last_launches = 0;
    g16::MsmRunner<g16::Fp> g1;
    g16::MsmRunner<g16::Fp2> g2;
    g16::NttEngine ntt;
    g16::KernelProfiler prof;
    g16::DeviceBuf scratch, scalars, results;
    // multi-GPU single proof (comm.cu): NCCL communicator (opaque), this context's slice of every MSM
    void* comm = nullptr;
    int rank = 0, world = 1;
    g16::DeviceBuf comm_recv;
    // one large standalone MSM (g16_msm_dev, batch 1): point slices run as a software pipeline on MSM_PARTS streams
    // (digit sort of one slice under the accumulation of another, bucket reduction tails hidden)
    enum { MSM_PARTS = 4 };
    g16::MsmRunner<g16::Fp> g1_part[MSM_PARTS];
    g16::MsmRunner<g16::Fp2> g2_part[MSM_PARTS];
    cudaStream_t part_stream[MSM_PARTS] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t part_fork = nullptr, part_done[MSM_PARTS] = {nullptr, nullptr, nullptr, nullptr};
    g16::DeviceBuf part_out;
    // g16_set_deferred_join: device-side calls (g16_prove_wires_dev) return without making the context stream wait
    // for their assembly; g16_join / g16_sync / the next use of the same buffers do.  Lets consecutive calls overlap.
    bool deferred_join = false;
    std::vector<g16_circuit*> circuits;   // loaded on this context (g16_join walks them)
};

namespace g16 {
// d_pts[i] <- sum over the ranks of d_pts[i] (NCCL all-gather + add); no-op for world == 1
template <class F>
int comm_sum_points(g16_ctx* ctx, Affine<F>* d_pts, size_t count, cudaStream_t st);
void comm_release(g16_ctx* ctx);
}  // namespace g16

struct g16_bases {
    int g2 = 0;
    g16::MsmBases<g16::Fp> b1;
    g16::MsmBases<g16::Fp2> b2;
};

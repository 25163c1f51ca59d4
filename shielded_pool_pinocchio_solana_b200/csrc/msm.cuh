// msm.cuh -- batched, fixed-base Pippenger multi-scalar multiplication on BN254 G1 / G2.
//
// Replaces gnark-crypto `G1Jac.MultiExp` / `G2Jac.MultiExp` as called five times (+2
// Pedersen commitments) by gnark's groth16.Prove -- SURVEY.md 3.2 step 5, 8a rows
// a4,a5,a7-a11 (third-party Go; the reference invokes it through `sunspot prove`,
// /root/reference/client/proof.helper.ts:64).
//
// B200-first design (not gnark's goroutine-per-window layout):
//  * The bases of a Groth16 proving key never change, and a B200 has 180 GB of HBM.  So at
//    load time every base P_i is expanded ONCE into its window multiples 2^(c*j) P_i
//    (j < W, affine, Montgomery) -- `table[j*n + i]`.  All W signed-digit windows of a
//    scalar then feed ONE bucket set: no per-window bucket reduction and no window-
//    combination doublings remain on the per-proof path.
//  * A launch processes a BATCH of independent scalar vectors over the same bases (one per
//    proof); bucket id = batch*NB + |digit|-1.
//  * Pipeline per launch:
//      digits/histogram (warp-aggregated atomics) -> scan -> scatter (entry ids sorted by bucket)
//      -> buckets cut into "virtual buckets" of at most CAP entries: the hot buckets of a skewed,
//         witness-like scalar vector (30 % ones -> one bucket holds 30 % of all entries) become many
//         bounded chains whose partial sums are joined afterwards
//      -> virtual buckets ordered by decreasing size (counting sort over 4096 size classes), so the
//         32 lanes of a warp run chains of equal length and the longest chains start first
//      -> accumulate: thread per virtual bucket, XYZZ += affine (8M + 2S), next entry + point
//         gathered before the add (integer-pipe bound, ~80 % of the time)
//      -> join the pieces of split buckets -> bucket reduction sum_k k*B_k by segmented running
//         sums (multi-CTA), one affine normalisation per result.
//  * A batched-affine tree formulation (pairwise rounds, shared inversions: ~6 modmul per addition)
//    was built and measured in round 2 (git history: "Batched-affine tree MSM"): bit-exact, but the
//    operands of every addition have to be touched twice (denominator pass, addition pass), which
//    makes the randomly gathered first round DRAM-access bound (~20 G random 64-B accesses/s on
//    this part); end to end it lost to the chain below, which hides its single gather under 1.4 K
//    integer instructions per addition.  See DESIGN.md 3.
//  * Results are group elements, so they do not depend on the order in which a bucket's
//    entries are added: the unordered atomic scatter keeps the output bit-exact.
#pragma once
#include <vector>

#include "common.cuh"
#include "ec.cuh"

namespace g16 {

struct MsmConfig {
    int c;         // window bits (signed digits: magnitudes 1..2^(c-1))
    int W;         // number of windows = ceil(254 / c)
    uint32_t nb;   // buckets per batch element = 2^(c-1)
};

static inline MsmConfig msm_config(int c) {
    MsmConfig m;
    m.c = c;
    m.W = (254 + c - 1) / c;
    m.nb = 1u << (c - 1);
    return m;
}

// Window size minimising accumulate + bucket-reduction work for n points per batch element.
int msm_pick_window(size_t n, size_t batch);

// ---------------------------------------------------------------------------------------------
// digit extraction
// ---------------------------------------------------------------------------------------------
constexpr int MSM_DIGIT_THREADS = 256;

// pass 0: count  /  pass 1: scatter (entry = window*n + point, sign in bit 31).   grid = (ceil(n/256), batch)
// Canonical scalar limbs are parked in shared memory ([limb][thread], conflict-free) so the
// window loop can index them dynamically without spilling a register array to local memory.
// Atomics are warp-aggregated (match.any on the bucket id): a witness-like scalar vector puts
// ~30 % of its entries into bucket |digit| = 1, which would otherwise serialise on one address.
template <int PASS>
__global__ void __launch_bounds__(MSM_DIGIT_THREADS)
k_msm_digits(const Fr* __restrict__ scalars, size_t scalar_stride, const Fr* __restrict__ scalars1,
             size_t scalar_stride1, const uint32_t* __restrict__ map, uint32_t n, uint32_t lo, uint32_t n_total,
             int montgomery, MsmConfig cfg, uint32_t* __restrict__ counts_or_cursor,
             uint32_t* __restrict__ entries) {
    __shared__ uint32_t sk[8][MSM_DIGIT_THREADS];
    __shared__ uint32_t big_list[MSM_DIGIT_THREADS];
    __shared__ uint32_t big_count;
    const uint32_t i = blockIdx.x * MSM_DIGIT_THREADS + threadIdx.x;
    const uint32_t b = blockIdx.y;
    const bool live = i < n;
    if (threadIdx.x == 0) big_count = 0;
    Fr k = Fr::zero();
    if (live) {
        // map entries with bit 31 set read the second scalar source (e.g. the quotient H next to the wires)
        uint32_t si = map ? map[lo + i] : lo + i;   // this launch covers points [lo, lo + n) of n_total
        k = (si >> 31) ? scalars1[(size_t)b * scalar_stride1 + (si & 0x7fffffffu)]
                       : scalars[(size_t)b * scalar_stride + si];
        // most wires of a real witness are 0 or 1: skip the Montgomery reduction for those
        if (montgomery && !k.is_zero()) k = (k == Fr::one()) ? Fr{{1u, 0u, 0u, 0u, 0u, 0u, 0u, 0u}} : k.from_mont();
    }
#pragma unroll
    for (int l = 0; l < 8; l++) sk[l][threadIdx.x] = k.v[l];
    uint32_t* base = counts_or_cursor + (size_t)b * cfg.nb;
    const int c = cfg.c;
    const uint32_t mask = (1u << c) - 1u, half = 1u << (c - 1);
    const uint32_t lane = threadIdx.x & 31, lt = (1u << lane) - 1u;
    // one warp-aggregated histogram / scatter step: lanes with the same bucket share one atomic
    auto emit = [&](bool valid, uint32_t key, uint32_t entry) {
        const uint32_t peers = __match_any_sync(0xffffffffu, valid ? key : 0xffffffffu);
        const uint32_t leader = __ffs(peers) - 1, rank = __popc(peers & lt);
        uint32_t pos = 0;
        if (valid && lane == leader) pos = atomicAdd(base + key, (uint32_t)__popc(peers));
        pos = __shfl_sync(0xffffffffu, pos, leader);
        if (PASS == 1 && valid) entries[pos + rank] = entry;
    };
    // ---- scalars below 2^(c-1) (zero, one, bytes: ~90 % of a witness) have a single digit, in window 0 ----------
    const bool small = (k.v[1] | k.v[2] | k.v[3] | k.v[4] | k.v[5] | k.v[6] | k.v[7]) == 0 && k.v[0] <= half;
    emit(live && small && k.v[0] != 0, k.v[0] - 1, lo + i);
    // ---- the others are compacted and walked window by window by full warps ---------------------------------------
    __syncthreads();
    if (live && !small) big_list[atomicAdd(&big_count, 1u)] = threadIdx.x;
    __syncthreads();
    const uint32_t nbig = big_count;
    if ((threadIdx.x & ~31u) >= nbig) return;   // whole warp idle
    const bool mine = threadIdx.x < nbig;
    const uint32_t src = mine ? big_list[threadIdx.x] : 0;
    const uint32_t gi = lo + blockIdx.x * MSM_DIGIT_THREADS + src;
    const uint32_t* s = &sk[0][src];
    uint32_t carry = 0;
    for (int j = 0; j < cfg.W; j++) {
        int bit = j * c;
        int idx = bit >> 5, sh = bit & 31;
        uint64_t w = s[idx * MSM_DIGIT_THREADS];
        if (idx < 7) w |= (uint64_t)s[(idx + 1) * MSM_DIGIT_THREADS] << 32;
        uint32_t d = ((uint32_t)(w >> sh) & mask) + carry;
        uint32_t neg = 0;
        carry = 0;
        if (d > half) {
            d = (1u << c) - d;
            neg = 1;
            carry = 1;
        }
        emit(mine && d != 0, d - 1, ((uint32_t)j * n_total + gi) | (neg << 31));
    }
}

// ---------------------------------------------------------------------------------------------
// exclusive scan of uint32 (3 small kernels)
// ---------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;  // per thread
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total) {
    __shared__ uint32_t warp_sums[SCAN_THREADS / 32];
    __shared__ uint32_t block_total;
    uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
        uint32_t s = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < SCAN_THREADS / 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += y;
        }
        if (lane < SCAN_THREADS / 32) warp_sums[lane] = s;
        if (lane == SCAN_THREADS / 32 - 1) block_total = s;
    }
    __syncthreads();
    uint32_t prefix = wid ? warp_sums[wid - 1] : 0;
    *total = block_total;
    __syncthreads();   // the shared words are reused by the caller's next scan
    return prefix + x - v;
}

// phase A: per-tile totals
static __global__ void __launch_bounds__(SCAN_THREADS) k_scan_tile_sums(const uint32_t* __restrict__ in, size_t n,
                                                                         uint32_t* __restrict__ tile_sums) {
    size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++)
        if (base + k < n) s += in[base + k];
    uint32_t total;
    block_exclusive_scan(s, &total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}
// phase B: scan the tile totals in place (single CTA, loops)
static __global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(uint32_t* tile_sums, uint32_t ntiles) {
    uint32_t running = 0;
    for (uint32_t base = 0; base < ntiles; base += SCAN_THREADS) {
        uint32_t i = base + threadIdx.x;
        uint32_t v = i < ntiles ? tile_sums[i] : 0;
        uint32_t total;
        uint32_t ex = block_exclusive_scan(v, &total);
        if (i < ntiles) tile_sums[i] = running + ex;
        running += total;
    }
}
// phase C: final offsets, written to both `starts` and `cursor`
// (`in` and `cursor` may alias: every thread reads its own items before it writes them)
static __global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(const uint32_t* in, size_t n,
                                                                     const uint32_t* __restrict__ tile_sums,
                                                                     uint32_t* __restrict__ starts, uint32_t* cursor) {
    size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        v[k] = base + k < n ? in[base + k] : 0;
        s += v[k];
    }
    uint32_t total;
    uint32_t ex = block_exclusive_scan(s, &total) + tile_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        if (base + k < n) {
            starts[base + k] = ex;
            if (cursor) cursor[base + k] = ex;
        }
        ex += v[k];
    }
}

// ---------------------------------------------------------------------------------------------
// virtual buckets and size classes
// ---------------------------------------------------------------------------------------------
// A virtual bucket is a run of at most `cap` consecutive entries of one real bucket.  Class index
// d = 4095 - size, so ascending d walks the sizes downwards (largest buckets first).
constexpr int VB_CLASSES = 4096;       // sizes 0..4095
constexpr int VB_MAX_CAP = 4095;
constexpr int MSM_VB_THREADS = 256;

// nv[k] = ceil(size_k / cap); nv[nbuckets] = 0 (so that the exclusive scan ends with the total)
static __global__ void __launch_bounds__(MSM_VB_THREADS)
k_vb_count(const uint32_t* __restrict__ starts, const uint32_t* __restrict__ ends, uint32_t nbuckets, uint32_t cap,
           uint32_t* __restrict__ nv) {
    uint32_t k = blockIdx.x * MSM_VB_THREADS + threadIdx.x;
    if (k > nbuckets) return;
    nv[k] = k < nbuckets ? (ends[k] - starts[k] + cap - 1) / cap : 0;
}

// descriptors of the virtual buckets of every real bucket, the size histogram, and the list of
// real buckets that were split (their partial sums are joined by k_vb_join)
static __global__ void __launch_bounds__(MSM_VB_THREADS)
k_vb_fill(const uint32_t* __restrict__ starts, const uint32_t* __restrict__ ends, const uint32_t* __restrict__ vbase,
          uint32_t nbuckets, uint32_t cap, uint32_t* __restrict__ vb_start, uint32_t* __restrict__ vb_size,
          uint32_t* __restrict__ hist, uint32_t* __restrict__ hot_count, uint32_t* __restrict__ hot_list, uint32_t hot_cap) {
    __shared__ uint32_t lh[VB_CLASSES];
    for (int i = threadIdx.x; i < VB_CLASSES; i += MSM_VB_THREADS) lh[i] = 0;
    __syncthreads();
    uint32_t k = blockIdx.x * MSM_VB_THREADS + threadIdx.x;
    if (k < nbuckets) {
        uint32_t s = starts[k], m = ends[k] - s, v = vbase[k];
        uint32_t nvk = (m + cap - 1) / cap;
        for (uint32_t q = 0; q < nvk; q++) {
            uint32_t sz = min(cap, m - q * cap);
            vb_start[v + q] = s + q * cap;
            vb_size[v + q] = sz;
            atomicAdd(&lh[VB_CLASSES - 1 - sz], 1u);
        }
        if (nvk > 1) {
            uint32_t slot = atomicAdd(hot_count, 1u);
            if (slot < hot_cap) hot_list[slot] = k;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < VB_CLASSES; i += MSM_VB_THREADS)
        if (lh[i]) atomicAdd(&hist[i], lh[i]);
}

// first[d] = number of virtual buckets in classes < d (the sorted position of class d); 1024 threads x 4
// classes; first[VB_CLASSES] = total
static __global__ void __launch_bounds__(1024) k_class_first(const uint32_t* __restrict__ hist, uint32_t* __restrict__ first) {
    __shared__ uint32_t sh[1024];
    const int t = threadIdx.x;
    uint32_t w[4], s = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        w[k] = hist[4 * t + k];
        s += w[k];
    }
    sh[t] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        uint32_t v = t >= o ? sh[t - o] : 0;
        __syncthreads();
        sh[t] += v;
        __syncthreads();
    }
    uint32_t ex = sh[t] - s;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        first[4 * t + k] = ex;
        ex += w[k];
    }
    if (t == 1023) first[VB_CLASSES] = ex;
}

// order[first[d] + rank] = v : virtual buckets sorted by decreasing size (counting sort)
static __global__ void __launch_bounds__(MSM_VB_THREADS)
k_vb_order(const uint32_t* __restrict__ vb_size, const uint32_t* __restrict__ nv_total, const uint32_t* __restrict__ first,
           uint32_t* __restrict__ cursor, uint32_t* __restrict__ order) {
    __shared__ uint32_t lh[VB_CLASSES], lbase[VB_CLASSES];
    for (int i = threadIdx.x; i < VB_CLASSES; i += MSM_VB_THREADS) lh[i] = 0;
    __syncthreads();
    const uint32_t nvb = *nv_total;
    uint32_t v = blockIdx.x * MSM_VB_THREADS + threadIdx.x;
    uint32_t d = 0, rank = 0;
    if (v < nvb) {
        d = VB_CLASSES - 1 - vb_size[v];
        rank = atomicAdd(&lh[d], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < VB_CLASSES; i += MSM_VB_THREADS)
        if (lh[i]) lbase[i] = atomicAdd(&cursor[i], lh[i]);
    __syncthreads();
    if (v < nvb) order[first[d] + lbase[d] + rank] = v;
}

// ---------------------------------------------------------------------------------------------
// bucket accumulation: one thread per virtual bucket, in order of decreasing size
// ---------------------------------------------------------------------------------------------
#ifndef MSM_ACC_MIN_CTAS
#define MSM_ACC_MIN_CTAS 4   // 128 registers; 5 (96 registers, small spills) measured no faster: the wide-multiplier pipe is the limit, not occupancy
#endif
#ifndef MSM_ACC_G2_MIN_CTAS
#define MSM_ACC_G2_MIN_CTAS 1
#endif
template <class F>
__global__ void __launch_bounds__(128, sizeof(F) == sizeof(Fp) ? MSM_ACC_MIN_CTAS : MSM_ACC_G2_MIN_CTAS)
k_msm_accumulate(const Affine<F>* __restrict__ table, const uint32_t* __restrict__ entries,
                 const uint32_t* __restrict__ vb_start, const uint32_t* __restrict__ vb_size,
                 const uint32_t* __restrict__ order, const uint32_t* __restrict__ nv_total,
                 XYZZ<F>* __restrict__ result_vb) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= *nv_total) return;
    const uint32_t v = order[q];
    const uint32_t s = vb_start[v], e = s + vb_size[v];   // size >= 1
    XYZZ<F> acc = XYZZ<F>::inf();
    uint32_t en = entries[s];
    Affine<F> p = table[en & 0x7fffffffu];
    for (uint32_t it = s; it < e; it++) {
        uint32_t en_next = 0;
        Affine<F> pn = p;
        if (it + 1 < e) {  // issue the next gather before the ~1.4k-IMAD add
            en_next = entries[it + 1];
            pn = table[en_next & 0x7fffffffu];
        }
        if (en >> 31) p.y = p.y.neg();
        acc.madd(p);
        en = en_next;
        p = pn;
    }
    result_vb[v] = acc;
}

// ---------------------------------------------------------------------------------------------
// batched-affine bucket accumulation
// ---------------------------------------------------------------------------------------------
// The chain kernel above pays 10 modmul per addition because its accumulator is projective.  Here a thread
// owns BA_K virtual buckets (consecutive in the size-sorted order, so equally long) and advances all of them
// one entry per step in AFFINE coordinates: the BA_K slopes of a step share one field inversion (Montgomery's
// trick, 3 modmul per element) and an addition costs 6 modmul.  The inversion is a binary GCD on the ALU pipe
// (no multiplier), so it runs beside the IMAD.WIDE work of the other warps of the SM.
//   pass 1 (k up)  : d_k = x_P - x_acc;  prefix_k = d_0 ... d_(k-1)   (only the x coordinates are read)
//   inversion      : inv = (d_0 ... d_(K-1))^-1
//   pass 2 (k down): 1/d_k = inv * prefix_k;  inv *= d_k;  lambda = (y_P - y_acc)/d_k;  x3, y3
// Per-thread state (accumulators, prefixes, chain descriptors) lives in global memory in [k][thread] order:
// every access is coalesced across the warp; the points themselves are gathered from the table twice (the
// second gather finds a window-table of a proving key in L2).  Exceptional steps (an operand at infinity,
// equal or opposite x, an x coordinate of zero) are left out of the batch and done by the complete XYZZ
// formulas with their own inversion: never taken on real inputs, exercised by the golden vectors.
constexpr int BA_K = 64;
constexpr int BA_THREADS = 128;

template <class F>
__device__ __noinline__ Affine<F> ba_slow_add(const Affine<F>& a, const Affine<F>& p) {
    XYZZ<F> t = XYZZ<F>::from_affine(a);
    t.madd(p);
    return t.to_affine();
}

template <class F>
__global__ void __launch_bounds__(BA_THREADS)
k_msm_accumulate_ba(const Affine<F>* __restrict__ table, const uint32_t* __restrict__ entries,
                    const uint32_t* __restrict__ vb_start, const uint32_t* __restrict__ vb_size,
                    const uint32_t* __restrict__ order, const uint32_t* __restrict__ nv_total, uint32_t nlanes,
                    F* __restrict__ accx, F* __restrict__ accy, F* __restrict__ prefix, uint32_t* __restrict__ cs,
                    uint32_t* __restrict__ cm, XYZZ<F>* __restrict__ result_vb) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t nv = *nv_total;
    const uint32_t q0 = t * BA_K;
    if (t >= nlanes || q0 >= nv) return;
    const uint32_t kn = min((uint32_t)BA_K, nv - q0);
    // ---- step 0: every chain starts at its first entry --------------------------------------------
    uint32_t steps = 0;
    for (uint32_t k = 0; k < kn; k++) {
        const uint32_t v = order[q0 + k];
        const uint32_t s = vb_start[v], m = vb_size[v];
        const size_t o = (size_t)k * nlanes + t;
        cs[o] = s;
        cm[o] = m;
        steps = max(steps, m);
        const uint32_t en = entries[s];
        Affine<F> p = table[en & 0x7fffffffu];
        if (en >> 31) p.y = p.y.neg();
        accx[o] = p.x;
        accy[o] = p.y;
    }
    // ---- steps 1 .. longest chain - 1 ----------------------------------------------------------------
    for (uint32_t j = 1; j < steps; j++) {
        F run = F::one();
        bool any = false;
        for (uint32_t k = 0; k < kn; k++) {
            const size_t o = (size_t)k * nlanes + t;
            if (j >= cm[o]) continue;
            const uint32_t en = entries[cs[o] + j];
            const F px = table[en & 0x7fffffffu].x;
            const F ax = accx[o];
            const F d = px - ax;
            if (d.is_zero() || ax.is_zero() || px.is_zero()) continue;   // exceptional: handled in pass 2
            prefix[o] = run;
            run = run * d;
            any = true;
        }
        F inv = any ? run.inverse() : run;
        for (int k = (int)kn - 1; k >= 0; k--) {
            const size_t o = (size_t)k * nlanes + t;
            if (j >= cm[o]) continue;
            const uint32_t en = entries[cs[o] + j];
            Affine<F> p = table[en & 0x7fffffffu];
            if (en >> 31) p.y = p.y.neg();
            const F ax = accx[o], ay = accy[o];
            const F d = p.x - ax;
            if (d.is_zero() || ax.is_zero() || p.x.is_zero()) {
                Affine<F> a = {ax, ay};
                a = ba_slow_add(a, p);
                accx[o] = a.x;
                accy[o] = a.y;
                continue;
            }
            const F dinv = inv * prefix[o];
            inv = inv * d;
            const F lam = (p.y - ay) * dinv;
            const F x3 = lam.sqr() - ax - p.x;
            const F y3 = lam * (ax - x3) - ay;
            accx[o] = x3;
            accy[o] = y3;
        }
    }
    // ---- results (XYZZ with ZZ = ZZZ = 1, what the join / reduction kernels read) ---------------------------------
    for (uint32_t k = 0; k < kn; k++) {
        const size_t o = (size_t)k * nlanes + t;
        Affine<F> a = {accx[o], accy[o]};
        result_vb[order[q0 + k]] = XYZZ<F>::from_affine(a);
    }
}

template <class P>
__device__ __forceinline__ P shfl_down_point(const P& p, int delta) {
    P r;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&p);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(P) / 4); i++) dst[i] = __shfl_down_sync(0xffffffffu, src[i], delta);
    return r;
}

// Joins the partial sums of split (hot) buckets: result_vb[vbase[k]] = sum of the bucket's pieces.
// Typical case (the buckets fed by a short top window, a few pieces each): one WARP per bucket.
// Extreme skew (one bucket holding 30 % of a 2^22-point MSM is thousands of pieces): one CTA per
// bucket (WIDE = true), launched over the same list; each kernel skips the other's buckets.
constexpr int MSM_JOIN_THREADS = 128;
constexpr uint32_t MSM_JOIN_WIDE = 128;   // pieces above which a bucket is joined by a whole CTA
template <class F, bool WIDE>
__global__ void __launch_bounds__(MSM_JOIN_THREADS)
k_vb_join(const uint32_t* __restrict__ hot_count, const uint32_t* __restrict__ hot_list, uint32_t hot_cap,
          const uint32_t* __restrict__ vbase, XYZZ<F>* __restrict__ result_vb) {
    __shared__ XYZZ<F> sh[MSM_JOIN_THREADS / 32];
    const uint32_t nhot = min(*hot_count, hot_cap);
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t unit = WIDE ? blockIdx.x : blockIdx.x * (MSM_JOIN_THREADS / 32) + wid;
    const uint32_t nunits = WIDE ? gridDim.x : gridDim.x * (MSM_JOIN_THREADS / 32);
    const uint32_t tid = WIDE ? threadIdx.x : lane, step = WIDE ? MSM_JOIN_THREADS : 32;
    for (uint32_t h = unit; h < nhot; h += nunits) {
        const uint32_t k = hot_list[h];
        const uint32_t v0 = vbase[k], v1 = vbase[k + 1];
        if ((v1 - v0 > MSM_JOIN_WIDE) != WIDE) continue;   // uniform over the unit
        XYZZ<F> acc = XYZZ<F>::inf();
        for (uint32_t v = v0 + tid; v < v1; v += step) acc.add(result_vb[v]);
        const int top = WIDE ? 16 : (v1 - v0 > 16 ? 16 : (v1 - v0 > 8 ? 8 : (v1 - v0 > 4 ? 4 : (v1 - v0 > 2 ? 2 : 1))));
#pragma unroll 1
        for (int dlt = top; dlt >= 1; dlt >>= 1) {
            XYZZ<F> o = shfl_down_point(acc, dlt);
            acc.add(o);
        }
        if (!WIDE) {
            if (lane == 0) result_vb[v0] = acc;
        } else {
            if (lane == 0) sh[wid] = acc;
            __syncthreads();
            if (threadIdx.x == 0) {
                for (int w = 1; w < MSM_JOIN_THREADS / 32; w++) acc.add(sh[w]);
                result_vb[v0] = acc;
            }
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------------------
// bucket reduction:  sum_{idx} (idx+1) * B[idx]
// ---------------------------------------------------------------------------------------------
// reduce1: thread (b,t) folds `seg` consecutive buckets into
//   run = sum B[t*seg+i],  acc = sum (i+1) * B[t*seg+i]
// Bucket k's point is result_vb[vbase[k]] (empty bucket: vbase[k+1] == vbase[k]).
template <class F>
__global__ void __launch_bounds__(128)
k_msm_reduce1(const XYZZ<F>* __restrict__ result_vb, const uint32_t* __restrict__ vbase, uint32_t seg,
              uint32_t nseg_total, XYZZ<F>* __restrict__ seg_acc, XYZZ<F>* __restrict__ seg_run) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;  // = b*(nb/seg) + t
    if (g >= nseg_total) return;
    const uint32_t k0 = g * seg;
    XYZZ<F> run = XYZZ<F>::inf(), acc = XYZZ<F>::inf();
    uint32_t vhi = vbase[k0 + seg];
    for (int i = (int)seg - 1; i >= 0; i--) {
        uint32_t vlo = vbase[k0 + i];
        if (vlo != vhi) run.add(result_vb[vlo]);
        acc.add(run);
        vhi = vlo;
    }
    seg_acc[g] = acc;
    seg_run[g] = run;
}

// [k]P for a small non-negative integer k (MSB-first double-and-add from the top set bit)
template <class F>
__device__ XYZZ<F> small_mul(const XYZZ<F>& p, uint32_t k) {
    XYZZ<F> r = XYZZ<F>::inf();
    if (k == 0) return r;
    for (int bit = 31 - __clz(k); bit >= 0; bit--) {
        r = r.dbl();
        if ((k >> bit) & 1u) r.add(p);
    }
    return r;
}

// Threads per part of reduce2.  The kernel is latency-bound, but the step is issue-bound: every warp pays the
// small_mul and the shuffle tree whatever it owns, so FEWER, longer threads issue less (measured on the device step:
// 256 -> 64 threads: withdraw 25.4 -> 24.4 ms, audit_like 16.4 -> 15.8 ms; a standalone 2^22 MSM is unchanged).
#ifndef MSM_R2_THREADS_N
#define MSM_R2_THREADS_N 64
#endif
constexpr int MSM_R2_THREADS = MSM_R2_THREADS_N;

// reduce2: `parts` CTAs per batch element.  With t the segment index,
//   result = sum_t acc_t + seg * sum_t t*run_t
// Thread u of part p owns segments [t0, t0 + per): local 0-based weighted sum + t0 * (sum of its
// runs); warp-shuffle tree, then a tree over the warp leaders in shared memory.
// part_out[(b*parts + p)*2 + {0,1}] = (sum acc, sum t*run) of the part.
template <class F>
__global__ void __launch_bounds__(MSM_R2_THREADS)
k_msm_reduce2(const XYZZ<F>* __restrict__ seg_acc, const XYZZ<F>* __restrict__ seg_run, uint32_t nseg, uint32_t parts,
              uint32_t per, XYZZ<F>* __restrict__ part_out) {
    __shared__ XYZZ<F> sh[2][MSM_R2_THREADS / 32];
    const uint32_t b = blockIdx.x / parts, p = blockIdx.x % parts, u = threadIdx.x;
    const uint32_t t0 = (p * MSM_R2_THREADS + u) * per, t1 = min(t0 + per, nseg);
    const XYZZ<F>* A = seg_acc + (size_t)b * nseg;
    const XYZZ<F>* Rn = seg_run + (size_t)b * nseg;
    XYZZ<F> asum = XYZZ<F>::inf(), run = XYZZ<F>::inf(), wsum = XYZZ<F>::inf();
    if (t0 < t1) {
        for (int t = (int)t1 - 1; t >= (int)t0; t--) {
            asum.add(A[t]);
            wsum.add(run);     // 0-based weights: segment t gets (t - t0)
            run.add(Rn[t]);
        }
        if (t0) wsum.add(small_mul(run, t0));
    }
#pragma unroll 1
    for (int d = 16; d >= 1; d >>= 1) {
        XYZZ<F> o1 = shfl_down_point(asum, d);
        XYZZ<F> o2 = shfl_down_point(wsum, d);
        asum.add(o1);
        wsum.add(o2);
    }
    if ((u & 31) == 0) {
        sh[0][u >> 5] = asum;
        sh[1][u >> 5] = wsum;
    }
    __syncthreads();
    // two lanes of one warp each fold one of the two per-warp arrays
    if (u < 2) {
        const int which = (int)u;
        XYZZ<F> s = sh[which][0];
        for (int w = 1; w < MSM_R2_THREADS / 32; w++) s.add(sh[which][w]);
        part_out[((size_t)b * parts + p) * 2 + which] = s;
    }
}

// reduce3: one warp per batch element joins the parts, applies the segment factor and normalises.
template <class F>
__global__ void __launch_bounds__(32)
k_msm_reduce3(const XYZZ<F>* __restrict__ part_out, uint32_t parts, uint32_t seg, Affine<F>* __restrict__ out) {
    const uint32_t b = blockIdx.x, lane = threadIdx.x;
    XYZZ<F> asum = XYZZ<F>::inf(), wsum = XYZZ<F>::inf();
    for (uint32_t p = lane; p < parts; p += 32) {
        asum.add(part_out[((size_t)b * parts + p) * 2]);
        wsum.add(part_out[((size_t)b * parts + p) * 2 + 1]);
    }
#pragma unroll 1
    for (int d = 16; d >= 1; d >>= 1) {
        XYZZ<F> o1 = shfl_down_point(asum, d);
        XYZZ<F> o2 = shfl_down_point(wsum, d);
        asum.add(o1);
        wsum.add(o2);
    }
    if (lane == 0) {
        for (uint32_t s = seg; s > 1; s >>= 1) wsum = wsum.dbl();
        asum.add(wsum);
        out[b] = asum.to_affine();
    }
}

// reduce3 when every batch element has ONE part (small MSMs, the batched prover): a thread per batch element --
// a warp per element would issue the same instructions 32 times over.
template <class F>
__global__ void __launch_bounds__(32)
k_msm_reduce3_flat(const XYZZ<F>* __restrict__ part_out, uint32_t batch, uint32_t seg, Affine<F>* __restrict__ out) {
    const uint32_t b = blockIdx.x * 32 + threadIdx.x;
    if (b >= batch) return;
    XYZZ<F> asum = part_out[(size_t)b * 2], wsum = part_out[(size_t)b * 2 + 1];
    for (uint32_t s = seg; s > 1; s >>= 1) wsum = wsum.dbl();
    asum.add(wsum);
    out[b] = asum.to_affine();
}

// ---------------------------------------------------------------------------------------------
// load-time expansion of the bases into window multiples
// ---------------------------------------------------------------------------------------------
// table[0*n + i] = P_i (already there); table[j*n + i] = 2^c * table[(j-1)*n + i]
template <class F>
__global__ void __launch_bounds__(128) k_msm_expand(Affine<F>* table, uint32_t n, int c, int W) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Affine<F> p = table[i];
    for (int j = 1; j < W; j++) {
        XYZZ<F> a = XYZZ<F>::dbl_affine(p);
        for (int k = 1; k < c; k++) a = a.dbl();
        p = a.to_affine();
        table[(size_t)j * n + i] = p;
    }
}

// canonical (little-endian limb) base-field elements -> Montgomery form, in place
static __global__ void k_fp_to_mont(Fp* v, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = v[i].to_mont();
}
static __global__ void k_fp_from_mont(Fp* v, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = v[i].from_mont();
}

// ---------------------------------------------------------------------------------------------
// host-side engine
// ---------------------------------------------------------------------------------------------
template <class F>
class MsmBases {
   public:
    MsmBases() {}
    ~MsmBases() { release(); }
    // `host_pts`: n affine points in device layout (Affine<F>, little-endian limbs); infinity = all-zero.
    // `canonical` != 0: coordinates are plain integers and are converted to Montgomery on the device.
    int load(const Affine<F>* host_pts, size_t n, int c, int canonical, cudaStream_t st);
    void release();

    size_t n = 0;
    MsmConfig cfg{};
    Affine<F>* table = nullptr;  // device, [W][n]
};

// growable device allocation owned by a runner
struct MsmScratch {
    void* ptr = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes);
    void release();
};

template <class F>
class MsmRunner {
   public:
    ~MsmRunner() { release(); }
    // Computes out[b] = sum_i scalars[b*stride + map[i]] * P_i  for b < batch.  `out` is a DEVICE
    // array of `batch` affine points (Montgomery).  All work is enqueued on `st`.
    // `lo`, `cnt` (cnt != 0): only the points [lo, lo + cnt) of the bases take part (a slice of one large MSM;
    // slices run as a software pipeline on several streams, see g16_msm_dev).
    int run(const MsmBases<F>& bases, const Fr* d_scalars, size_t stride, const uint32_t* d_map, int montgomery,
            size_t batch, Affine<F>* d_out, cudaStream_t st, const Fr* d_scalars1 = nullptr, size_t stride1 = 0,
            size_t lo = 0, size_t cnt = 0);
    void release();
    // kernels launched by the last run() (for bench.py's gpu_launches)
    int launches = 0;
    KernelProfiler* prof = nullptr;   // optional: times the accumulate kernel
    const char* label = "msm";        // name in the G16_TIMELINE dump
    // Optional: the bucket accumulation (the one throughput-bound kernel; its grid fills every SM's register file
    // for milliseconds) is launched on this LOW-priority stream, so that the short latency-bound kernels of the
    // other MSMs running beside it (scans, ordering, bucket reduction: high-priority `st`) get the CTA slots
    // that free up instead of queueing behind the whole grid.
    cudaStream_t acc_stream = nullptr;

   private:
    enum { S_COUNTS, S_STARTS, S_TILES, S_ENTRIES, S_NV, S_VBASE, S_VBSTART, S_VBSIZE, S_ORDER, S_TABLES, S_HOT,
           S_RESULT, S_SEGACC, S_SEGRUN, S_PARTS, S_BA_X, S_BA_Y, S_BA_PREFIX, S_BA_CS, S_BA_CM, S_COUNT };
    MsmScratch s[S_COUNT];
    cudaEvent_t ev_sorted = nullptr, ev_accumulated = nullptr;
};

}  // namespace g16

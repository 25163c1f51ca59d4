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
//  * Bucket accumulation is a BATCHED-AFFINE TREE REDUCTION, not a chain per bucket:
//      digits/histogram -> scan -> scatter (entries sorted by bucket)
//      -> buckets cut into "virtual buckets" of at most CAP entries (hot buckets of a skewed,
//         witness-like scalar vector are split; their partial sums are joined afterwards)
//      -> virtual buckets ordered by size; all buckets of one size share one tree shape, so the
//         work of a round is addressed by arithmetic on a 4096-entry class table (no per-round
//         scans, no per-element metadata)
//      -> R tree rounds; round r adds adjacent pairs of every bucket's current point list
//         (affine + affine = 2M + 1S + one shared inversion).  A round is three kernels:
//            A: denominators, per-thread prefix products      (1 modmul per addition)
//            B: Montgomery-trick inversion of the thread totals (amortised < 0.5 modmul)
//            C: back-substitution + the additions             (4M + 1S per addition)
//         ~6 modmul per addition instead of the 10 of an XYZZ mixed addition, and every pair of
//         a round is independent: a bucket holding 30 % of all entries costs the same as the
//         same entries spread out (no thread owns a bucket).
//      -> the <= 32 points left per virtual bucket are folded by a short XYZZ chain
//      -> bucket reduction sum_k k*B_k by segmented running sums, multi-CTA, one normalisation.
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

// pass 0: count  /  pass 1: scatter the window multiples 2^(c*j) P_i themselves (sign applied) into
// bucket order, so that every later pass streams over contiguous points.   grid = (ceil(n/256), batch)
// Canonical scalar limbs are parked in shared memory ([limb][thread], conflict-free) so the
// window loop can index them dynamically without spilling a register array to local memory.
// Atomics are warp-aggregated (match.any on the bucket id): a witness-like scalar vector puts
// ~30 % of its entries into bucket |digit| = 1, which would otherwise serialise on one address.
template <int PASS, class F>
__global__ void __launch_bounds__(MSM_DIGIT_THREADS)
k_msm_digits(const Fr* __restrict__ scalars, size_t scalar_stride, const Fr* __restrict__ scalars1,
             size_t scalar_stride1, const uint32_t* __restrict__ map, uint32_t n,
             int montgomery, MsmConfig cfg, uint32_t* __restrict__ counts_or_cursor,
             const Affine<F>* __restrict__ table, Affine<F>* __restrict__ sorted) {
    __shared__ uint32_t sk[8][MSM_DIGIT_THREADS];
    const uint32_t i = blockIdx.x * MSM_DIGIT_THREADS + threadIdx.x;
    const uint32_t b = blockIdx.y;
    const bool live = i < n;
    Fr k = Fr::zero();
    if (live) {
        // map entries with bit 31 set read the second scalar source (e.g. the quotient H next to the wires)
        uint32_t si = map ? map[i] : i;
        k = (si >> 31) ? scalars1[(size_t)b * scalar_stride1 + (si & 0x7fffffffu)]
                       : scalars[(size_t)b * scalar_stride + si];
        if (montgomery) k = k.from_mont();
    }
#pragma unroll
    for (int l = 0; l < 8; l++) sk[l][threadIdx.x] = k.v[l];
    uint32_t* base = counts_or_cursor + (size_t)b * cfg.nb;
    const uint32_t* s = &sk[0][threadIdx.x];
    const int c = cfg.c;
    const uint32_t mask = (1u << c) - 1u, half = 1u << (c - 1);
    const uint32_t lane = threadIdx.x & 31, lt = (1u << lane) - 1u;
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
        const bool valid = live && d != 0;
        const uint32_t key = valid ? d - 1 : 0xffffffffu;
        const uint32_t peers = __match_any_sync(0xffffffffu, key);
        const uint32_t leader = __ffs(peers) - 1, rank = __popc(peers & lt);
        uint32_t pos = 0;
        if (valid && lane == leader) pos = atomicAdd(base + key, (uint32_t)__popc(peers));
        pos = __shfl_sync(0xffffffffu, pos, leader);
        if (PASS == 1 && valid) {
            // the table read is coalesced (consecutive i); the write lands in the bucket's run
            Affine<F> p = table[(size_t)j * n + i];
            if (neg) p.y = p.y.neg();
            sorted[pos + rank] = p;
        }
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
constexpr int TREE_MAX_ROUNDS = 7;     // cap <= 2^(rounds+5) - 1 <= 4095
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

// Class tables (one CTA per table, 1024 threads x 4 classes):
//   block 0      : first[d]  = number of virtual buckets in classes < d  (sorted position of class d)
//   block 1 + r  : wp[r][d]  = number of round-r work items in classes < d; an item is one OUTPUT point
//                  of the round: a bucket of size c holds m_r = ceil(c / 2^r) points before round r and
//                  m_(r+1) after it.  Round 0 also passes single-entry buckets through.
// Every table has VB_CLASSES + 1 entries; the last one is the total.
__device__ __forceinline__ uint32_t tree_size_after(uint32_t c, int r) { return (c + (1u << r) - 1) >> r; }

static __global__ void __launch_bounds__(1024)
k_class_tables(const uint32_t* __restrict__ hist, uint32_t* __restrict__ first, uint32_t* __restrict__ wp, int rounds) {
    __shared__ uint32_t sh[1024];
    const int t = threadIdx.x;
    const int r = (int)blockIdx.x - 1;
    uint32_t w[4], s = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        int d = 4 * t + k;
        uint32_t c = VB_CLASSES - 1 - d, g = hist[d];
        if (r < 0) {
            w[k] = g;
        } else {
            uint32_t m0 = tree_size_after(c, r);
            bool active = r == 0 ? c >= 1 : m0 >= 2;
            w[k] = active ? g * ((m0 + 1) >> 1) : 0;
        }
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
    uint32_t* out = r < 0 ? first : wp + (size_t)r * (VB_CLASSES + 1);
    (void)rounds;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        out[4 * t + k] = ex;
        ex += w[k];
    }
    if (t == 1023) out[VB_CLASSES] = ex;
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
// tree rounds
// ---------------------------------------------------------------------------------------------
constexpr int TREE_THREADS = 256;

template <class F>
struct TreeArgs {
    const Affine<F>* sorted;     // round-0 input: points in bucket order (k_msm_digits<1>)
    const uint32_t* vb_start;
    const uint32_t* order;
    const uint32_t* first;       // [VB_CLASSES + 1]
    const uint32_t* wp_in;       // round r-1 table (input layout), unused in round 0
    const uint32_t* wp;          // round r table
    const Affine<F>* xin;        // input points of rounds >= 1
    Affine<F>* xout;             // output points (buckets that still hold >= 2 points)
    XYZZ<F>* result_vb;          // finished virtual buckets
    F* prefix;                   // one per work item
    F* totals;                   // one per thread-chunk
    int r;
};

// largest d in [0, VB_CLASSES) with tab[d] <= t  (tab ascending, tab[VB_CLASSES] > t)
__device__ __forceinline__ uint32_t class_of(const uint32_t* tab, uint32_t t) {
    uint32_t lo = 0, hi = VB_CLASSES;   // invariant: tab[lo] <= t < tab[hi]
#pragma unroll 1
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (tab[mid] <= t) lo = mid;
        else hi = mid;
    }
    return lo;
}

struct TreeItem {
    uint32_t in0;      // index of the first input point (in `sorted` for round 0, in `xin` later)
    uint32_t out;      // xout index, or the virtual bucket id when to_result
    bool has_b, to_result;
};

template <class F>
__device__ __forceinline__ TreeItem tree_map(const TreeArgs<F>& a, const uint32_t* wp_sh, uint32_t t) {
    TreeItem it;
    // consecutive lanes hold consecutive t: search once per warp (lane 0), then walk forward
    uint32_t d = __shfl_sync(0xffffffffu, class_of(wp_sh, __shfl_sync(0xffffffffu, t, 0)), 0);
    while (wp_sh[d + 1] <= t) d++;
    const uint32_t c = VB_CLASSES - 1 - d;
    const uint32_t m0 = tree_size_after(c, a.r), m1 = (m0 + 1) >> 1;
    const uint32_t local = t - wp_sh[d];
    const uint32_t g = local / m1, j = local - g * m1;
    const uint32_t q = a.first[d] + g;
    it.has_b = 2 * j + 1 < m0;
    it.to_result = m1 == 1;
    if (a.r == 0) it.in0 = a.vb_start[a.order[q]] + 2 * j;
    else it.in0 = a.wp_in[d] + g * m0 + 2 * j;
    it.out = it.to_result ? a.order[q] : t;
    return it;
}

template <class F>
__device__ __forceinline__ Affine<F> tree_load(const TreeArgs<F>& a, uint32_t idx) {
    return a.r == 0 ? a.sorted[idx] : a.xin[idx];
}

// what an (a, b) pair needs: 0 = no field work (a result that needs no division), 1 = chord, 2 = tangent
template <class F>
__device__ __forceinline__ int tree_classify(const Affine<F>& pa, const Affine<F>& pb, bool has_b, F* den) {
    if (!has_b || pa.is_inf() || pb.is_inf()) return 0;
    if (pa.x != pb.x) {
        *den = pb.x - pa.x;
        return 1;
    }
    if (pa.y == pb.y && !pa.y.is_zero()) {
        *den = pa.y.dbl();
        return 2;
    }
    return 0;   // P + (-P)
}

// Kernel A: per work item the denominator of its addition; per thread the running product of its
// K denominators (prefix[t] = product of the thread's earlier ones), thread total -> totals[].
template <class F, int K>
__global__ void __launch_bounds__(TREE_THREADS) k_tree_a(TreeArgs<F> a) {
    __shared__ uint32_t wp_sh[VB_CLASSES + 1];
    for (int i = threadIdx.x; i <= VB_CLASSES; i += TREE_THREADS) wp_sh[i] = a.wp[i];
    __syncthreads();
    const uint32_t T = wp_sh[VB_CLASSES];
    constexpr uint32_t CH = TREE_THREADS * K;
    for (uint32_t chunk = blockIdx.x; (uint64_t)chunk * CH < T; chunk += gridDim.x) {
        F run = F::one();
#pragma unroll 1
        for (int i = 0; i < K; i++) {
            uint32_t t = chunk * CH + i * TREE_THREADS + threadIdx.x;
            if (chunk * CH + i * TREE_THREADS >= T) break;          // uniform over the CTA
            const bool live = t < T;
            TreeItem it = tree_map(a, wp_sh, live ? t : T - 1);     // whole warps map together (shuffles inside)
            if (!live) continue;
            a.prefix[t] = run;
            if (!it.has_b) continue;
            Affine<F> pa = tree_load(a, it.in0), pb = tree_load(a, it.in0 + 1);
            F den;
            if (tree_classify(pa, pb, true, &den)) run = run * den;
        }
        a.totals[(size_t)chunk * TREE_THREADS + threadIdx.x] = run;
    }
}

// Kernel B: in-place inversion of `count` field elements, KB per thread with Montgomery's trick
// (one true inversion per KB elements).  Zeros cannot occur (denominators are non-zero).
constexpr int TREE_INV_THREADS = 128;
template <class F, int KB>
__global__ void __launch_bounds__(TREE_INV_THREADS)
k_tree_b(F* __restrict__ totals, F* __restrict__ scratch, const uint32_t* __restrict__ wp, uint32_t items_per_chunk) {
    const uint32_t T = wp[VB_CLASSES];
    const uint32_t nchunks = (T + items_per_chunk - 1) / items_per_chunk;
    const uint64_t count = (uint64_t)nchunks * TREE_THREADS;
    constexpr uint32_t CH = TREE_INV_THREADS * KB;
    for (uint64_t base = (uint64_t)blockIdx.x * CH; base < count; base += (uint64_t)gridDim.x * CH) {
        F run = F::one();
        int used = 0;
#pragma unroll 1
        for (int i = 0; i < KB; i++) {
            uint64_t idx = base + (uint64_t)i * TREE_INV_THREADS + threadIdx.x;
            if (idx >= count) break;
            scratch[idx] = run;
            run = run * totals[idx];
            used = i + 1;
        }
        if (used == 0) continue;
        F inv = run.inverse();
#pragma unroll 1
        for (int i = used - 1; i >= 0; i--) {
            uint64_t idx = base + (uint64_t)i * TREE_INV_THREADS + threadIdx.x;
            F x = totals[idx];
            totals[idx] = inv * scratch[idx];
            inv = inv * x;
        }
    }
}

// Kernel C: walks the thread's items backwards, peels 1/den off the inverted total and adds.
template <class F, int K>
__global__ void __launch_bounds__(TREE_THREADS) k_tree_c(TreeArgs<F> a) {
    __shared__ uint32_t wp_sh[VB_CLASSES + 1];
    for (int i = threadIdx.x; i <= VB_CLASSES; i += TREE_THREADS) wp_sh[i] = a.wp[i];
    __syncthreads();
    const uint32_t T = wp_sh[VB_CLASSES];
    constexpr uint32_t CH = TREE_THREADS * K;
    for (uint32_t chunk = blockIdx.x; (uint64_t)chunk * CH < T; chunk += gridDim.x) {
        F inv = a.totals[(size_t)chunk * TREE_THREADS + threadIdx.x];
        int last = -1;
        {
            uint64_t first_t = (uint64_t)chunk * CH + threadIdx.x;
            if (first_t < T) last = (int)min((uint64_t)(K - 1), (T - 1 - first_t) / TREE_THREADS);
        }
#pragma unroll 1
        for (int i = K - 1; i >= 0; i--) {
            uint32_t t = chunk * CH + i * TREE_THREADS + threadIdx.x;
            if (chunk * CH + i * TREE_THREADS >= T) continue;       // uniform over the CTA
            TreeItem it = tree_map(a, wp_sh, i <= last ? t : T - 1);
            if (i > last) continue;
            Affine<F> pa = tree_load(a, it.in0), pb = pa;
            if (it.has_b) pb = tree_load(a, it.in0 + 1);
            F den;
            int kind = tree_classify(pa, pb, it.has_b, &den);
            Affine<F> o;
            if (kind == 0) {
                if (!it.has_b || pb.is_inf()) o = pa;
                else if (pa.is_inf()) o = pb;
                else o = Affine<F>::inf();
            } else {
                F dinv = inv * a.prefix[t];
                inv = inv * den;
                F num;
                if (kind == 1) {
                    num = pb.y - pa.y;
                } else {
                    F xx = pa.x.sqr();
                    num = xx.dbl() + xx;
                }
                F lam = num * dinv;
                o.x = lam.sqr() - pa.x - pb.x;
                o.y = lam * (pa.x - o.x) - pa.y;
            }
            if (it.to_result) a.result_vb[it.out] = XYZZ<F>::from_affine(o);
            else a.xout[it.out] = o;
        }
    }
}

// After the last tree round every virtual bucket of size > 2^R still holds m_R <= 32 points (R = 0:
// every bucket, read through `entries`): fold them with a chain of XYZZ mixed additions.  Threads
// walk the buckets in sorted order, so the lanes of a warp run chains of equal length.
template <class F>
__global__ void __launch_bounds__(128) k_tree_finish(TreeArgs<F> a) {
    const int R = a.r;   // number of tree rounds done
    const uint32_t min_c = R == 0 ? 1u : (1u << R) + 1u;
    const uint32_t NQ = a.first[VB_CLASSES - min_c];   // buckets with size >= min_c
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < NQ; q += gridDim.x * blockDim.x) {
        const uint32_t d = class_of(a.first, q);
        const uint32_t c = VB_CLASSES - 1 - d, g = q - a.first[d];
        const uint32_t m = tree_size_after(c, R);
        const uint32_t v = a.order[q];
        const uint32_t base = R == 0 ? a.vb_start[v] : a.wp_in[d] + g * m;
        XYZZ<F> acc = XYZZ<F>::inf();
        Affine<F> p = tree_load(a, base);
        for (uint32_t i = 0; i < m; i++) {
            Affine<F> pn = p;
            if (i + 1 < m) pn = tree_load(a, base + i + 1);   // next load in flight during the addition
            acc.madd(p);
            p = pn;
        }
        a.result_vb[v] = acc;
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
constexpr int MSM_JOIN_THREADS = 128;
template <class F>
__global__ void __launch_bounds__(MSM_JOIN_THREADS)
k_vb_join(const uint32_t* __restrict__ hot_count, const uint32_t* __restrict__ hot_list, uint32_t hot_cap,
          const uint32_t* __restrict__ vbase, XYZZ<F>* __restrict__ result_vb) {
    __shared__ XYZZ<F> sh[MSM_JOIN_THREADS / 32];
    const uint32_t nhot = min(*hot_count, hot_cap);
    for (uint32_t h = blockIdx.x; h < nhot; h += gridDim.x) {
        const uint32_t k = hot_list[h];
        const uint32_t v0 = vbase[k], v1 = vbase[k + 1];
        XYZZ<F> acc = XYZZ<F>::inf();
        for (uint32_t v = v0 + threadIdx.x; v < v1; v += MSM_JOIN_THREADS) acc.add(result_vb[v]);
#pragma unroll 1
        for (int dlt = 16; dlt >= 1; dlt >>= 1) {
            XYZZ<F> o = shfl_down_point(acc, dlt);
            acc.add(o);
        }
        if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < MSM_JOIN_THREADS / 32; w++) acc.add(sh[w]);
            result_vb[v0] = acc;
        }
        __syncthreads();
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

constexpr int MSM_R2_THREADS = 256;

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
    // threads 0 and 32 each fold one of the two 8-entry arrays
    if (u == 0 || u == 32) {
        const int which = u >> 5;
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
    int run(const MsmBases<F>& bases, const Fr* d_scalars, size_t stride, const uint32_t* d_map, int montgomery,
            size_t batch, Affine<F>* d_out, cudaStream_t st, const Fr* d_scalars1 = nullptr, size_t stride1 = 0);
    void release();
    // kernels launched by the last run() (for bench.py's gpu_launches)
    int launches = 0;
    KernelProfiler* prof = nullptr;   // optional: times the accumulation (tree rounds + finish)

   private:
    enum { S_COUNTS, S_STARTS, S_TILES, S_ENTRIES, S_NV, S_VBASE, S_VBSTART, S_VBSIZE, S_ORDER, S_TABLES, S_HOT,
           S_X0, S_X1, S_PREFIX, S_TOTALS, S_TSCRATCH, S_RESULT, S_SEGACC, S_SEGRUN, S_PARTS, S_COUNT };
    MsmScratch s[S_COUNT];
};

}  // namespace g16

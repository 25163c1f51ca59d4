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
//    proof); bucket id = batch*NB + |digit|-1.  That is what fills 148 SMs when a real
//    circuit only has 2^12..2^15 points per MSM.
//  * Pipeline per launch group:   count (atomics histogram)  ->  exclusive scan  ->
//    scatter (counting sort of (window,point,sign) entries by bucket; HBM-bound)  ->
//    accumulate (thread per bucket, XYZZ += affine, prefetching one entry ahead; IMAD-bound,
//    >95 % of the time)  ->  reduce1 (segmented running sums)  ->  reduce2 (one CTA per batch
//    element: warp-shuffle + shared-memory tree, small-scalar fix-ups, affine normalisation).
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
    uint32_t seg;  // buckets per reduce1 thread (power of two)
};

static inline MsmConfig msm_config(int c) {
    MsmConfig m;
    m.c = c;
    m.W = (254 + c - 1) / c;
    m.nb = 1u << (c - 1);
    m.seg = m.nb >= (1u << 16) ? 64 : (m.nb >= 1024 ? 16 : 4);
    if (m.seg > m.nb) m.seg = m.nb;
    return m;
}

// Heuristic window size: minimise  n*W(c) + kappa*2^(c-1)  per batch element while keeping
// batch*2^(c-1) bucket threads >= ~2 per resident thread slot of the chip.
int msm_pick_window(size_t n, size_t batch);

// ---------------------------------------------------------------------------------------------
// digit extraction
// ---------------------------------------------------------------------------------------------
// Canonical scalar limbs are parked in shared memory ([limb][thread], conflict-free) so the
// window loop can index them dynamically without spilling a register array to local memory.
template <class Fn>
__device__ __forceinline__ void msm_for_each_digit(const uint32_t* sk /* &smem[0][tid] */, int stride, int c,
                                                   int W, Fn&& fn) {
    const uint32_t mask = (1u << c) - 1u;
    const uint32_t half = 1u << (c - 1);
    uint32_t carry = 0;
    for (int j = 0; j < W; j++) {
        int bit = j * c;
        int idx = bit >> 5, sh = bit & 31;
        uint64_t w = sk[idx * stride];
        if (idx < 7) w |= (uint64_t)sk[(idx + 1) * stride] << 32;
        uint32_t d = ((uint32_t)(w >> sh) & mask) + carry;
        uint32_t neg = 0;
        carry = 0;
        if (d > half) {
            d = (1u << c) - d;
            neg = 1;
            carry = 1;
        }
        if (d) fn(j, d, neg);
    }
}

constexpr int MSM_DIGIT_THREADS = 256;

// pass 0: count  /  pass 1: scatter.   grid = (ceil(n/256), batch)
template <int PASS>
__global__ void __launch_bounds__(MSM_DIGIT_THREADS)
k_msm_digits(const Fr* __restrict__ scalars, size_t scalar_stride, const Fr* __restrict__ scalars1,
             size_t scalar_stride1, const uint32_t* __restrict__ map, uint32_t n,
             int montgomery, MsmConfig cfg, uint32_t* __restrict__ counts_or_cursor,
             uint32_t* __restrict__ entries) {
    __shared__ uint32_t sk[8][MSM_DIGIT_THREADS];
    uint32_t i = blockIdx.x * MSM_DIGIT_THREADS + threadIdx.x;
    uint32_t b = blockIdx.y;
    if (i >= n) return;
    // map entries with bit 31 set read the second scalar source (e.g. the quotient H next to the wires)
    uint32_t si = map ? map[i] : i;
    Fr k = (si >> 31) ? scalars1[(size_t)b * scalar_stride1 + (si & 0x7fffffffu)]
                      : scalars[(size_t)b * scalar_stride + si];
    if (montgomery) k = k.from_mont();
#pragma unroll
    for (int l = 0; l < 8; l++) sk[l][threadIdx.x] = k.v[l];
    uint32_t* base = counts_or_cursor + (size_t)b * cfg.nb;
    msm_for_each_digit(&sk[0][threadIdx.x], MSM_DIGIT_THREADS, cfg.c, cfg.W, [&](int j, uint32_t mag, uint32_t neg) {
        if (PASS == 0) {
            atomicAdd(base + (mag - 1), 1u);
        } else {
            uint32_t pos = atomicAdd(base + (mag - 1), 1u);
            entries[pos] = ((uint32_t)j * n + i) | (neg << 31);
        }
    });
}

// ---------------------------------------------------------------------------------------------
// exclusive scan of uint32 (3 small kernels; n up to 2^24 * 4096)
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
        __syncthreads();
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
            cursor[base + k] = ex;
        }
        ex += v[k];
    }
}

// ---------------------------------------------------------------------------------------------
// schedule: buckets in order of decreasing size
// ---------------------------------------------------------------------------------------------
// A warp runs as long as its fullest bucket, and bucket sizes are Poisson-distributed (mean
// n*W/2^(c-1), often only 20-30), so lanes would idle a third of the time.  A counting sort of the
// bucket ids by size (1024 size classes, descending) gives every warp 32 buckets of (nearly)
// equal size and starts the heaviest buckets first (longest-processing-time-first).
constexpr int MSM_SIZE_BINS = 1024;
constexpr int MSM_ORDER_THREADS = 256;
constexpr int MSM_ORDER_ITEMS = 8;

static __global__ void __launch_bounds__(MSM_ORDER_THREADS)
k_msm_size_hist(const uint32_t* __restrict__ starts, const uint32_t* __restrict__ ends, uint32_t nbuckets,
                uint32_t* __restrict__ hist) {
    __shared__ uint32_t lh[MSM_SIZE_BINS];
    for (int i = threadIdx.x; i < MSM_SIZE_BINS; i += MSM_ORDER_THREADS) lh[i] = 0;
    __syncthreads();
    uint32_t base = blockIdx.x * MSM_ORDER_THREADS * MSM_ORDER_ITEMS;
#pragma unroll
    for (int it = 0; it < MSM_ORDER_ITEMS; it++) {
        uint32_t k = base + it * MSM_ORDER_THREADS + threadIdx.x;
        if (k < nbuckets) atomicAdd(&lh[min(ends[k] - starts[k], (uint32_t)MSM_SIZE_BINS - 1)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < MSM_SIZE_BINS; i += MSM_ORDER_THREADS)
        if (lh[i]) atomicAdd(&hist[i], lh[i]);
}
// cursor[s] = number of buckets strictly larger than s  (single CTA of MSM_SIZE_BINS threads)
static __global__ void __launch_bounds__(MSM_SIZE_BINS) k_msm_size_bins(const uint32_t* __restrict__ hist,
                                                                        uint32_t* __restrict__ cursor) {
    __shared__ uint32_t sh[MSM_SIZE_BINS];
    int t = threadIdx.x;
    sh[t] = hist[MSM_SIZE_BINS - 1 - t];  // descending size
    __syncthreads();
    for (int o = 1; o < MSM_SIZE_BINS; o <<= 1) {
        uint32_t v = t >= o ? sh[t - o] : 0;
        __syncthreads();
        sh[t] += v;
        __syncthreads();
    }
    cursor[MSM_SIZE_BINS - 1 - t] = sh[t] - hist[MSM_SIZE_BINS - 1 - t];
}
static __global__ void __launch_bounds__(MSM_ORDER_THREADS)
k_msm_order(const uint32_t* __restrict__ starts, const uint32_t* __restrict__ ends, uint32_t nbuckets,
            uint32_t* __restrict__ cursor, uint32_t* __restrict__ order) {
    __shared__ uint32_t lh[MSM_SIZE_BINS], lbase[MSM_SIZE_BINS];
    for (int i = threadIdx.x; i < MSM_SIZE_BINS; i += MSM_ORDER_THREADS) lh[i] = 0;
    __syncthreads();
    uint32_t base = blockIdx.x * MSM_ORDER_THREADS * MSM_ORDER_ITEMS;
    uint32_t sz[MSM_ORDER_ITEMS], rank[MSM_ORDER_ITEMS];
#pragma unroll
    for (int it = 0; it < MSM_ORDER_ITEMS; it++) {
        uint32_t k = base + it * MSM_ORDER_THREADS + threadIdx.x;
        if (k < nbuckets) {
            sz[it] = min(ends[k] - starts[k], (uint32_t)MSM_SIZE_BINS - 1);
            rank[it] = atomicAdd(&lh[sz[it]], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < MSM_SIZE_BINS; i += MSM_ORDER_THREADS)
        if (lh[i]) lbase[i] = atomicAdd(&cursor[i], lh[i]);
    __syncthreads();
#pragma unroll
    for (int it = 0; it < MSM_ORDER_ITEMS; it++) {
        uint32_t k = base + it * MSM_ORDER_THREADS + threadIdx.x;
        if (k < nbuckets) order[lbase[sz[it]] + rank[it]] = k;
    }
}

// ---------------------------------------------------------------------------------------------
// bucket accumulation: one thread per (batch, bucket)
// ---------------------------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(128)
k_msm_accumulate(const Affine<F>* __restrict__ table, const uint32_t* __restrict__ entries,
                 const uint32_t* __restrict__ starts, const uint32_t* __restrict__ ends,
                 const uint32_t* __restrict__ order, XYZZ<F>* __restrict__ buckets, uint32_t total_buckets) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total_buckets) return;
    uint32_t k = order[t];
    uint32_t s = starts[k], e = ends[k];
    XYZZ<F> acc = XYZZ<F>::inf();
    if (s < e) {
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
    }
    buckets[k] = acc;
}

// ---------------------------------------------------------------------------------------------
// bucket reduction:  sum_{idx} (idx+1) * B[idx]
// ---------------------------------------------------------------------------------------------
// reduce1: thread (b,t) folds `seg` consecutive buckets into
//   run = sum B[t*seg+i],  acc = sum (i+1) * B[t*seg+i]
template <class F>
__global__ void __launch_bounds__(128)
k_msm_reduce1(const XYZZ<F>* __restrict__ buckets, uint32_t nb, uint32_t seg, uint32_t nseg_total,
              XYZZ<F>* __restrict__ seg_acc, XYZZ<F>* __restrict__ seg_run) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;  // = b*(nb/seg) + t
    if (g >= nseg_total) return;
    const XYZZ<F>* B = buckets + (size_t)g * seg;
    XYZZ<F> run = XYZZ<F>::inf(), acc = XYZZ<F>::inf();
    for (int i = (int)seg - 1; i >= 0; i--) {
        run.add(B[i]);
        acc.add(run);
    }
    seg_acc[g] = acc;
    seg_run[g] = run;
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

// [k]P for a small non-negative integer k (MSB-first double-and-add)
template <class F>
__device__ XYZZ<F> small_mul(const XYZZ<F>& p, uint32_t k) {
    XYZZ<F> r = XYZZ<F>::inf();
    for (int bit = 31; bit >= 0; bit--) {
        r = r.dbl();
        if ((k >> bit) & 1u) r.add(p);
    }
    return r;
}

constexpr int MSM_R2_THREADS = 256;

// reduce2: one CTA per batch element.
//   result = sum_t acc_t + seg * sum_t t*run_t        (t = segment index)
// thread u owns segments [u*per, (u+1)*per); the CTA sum is a warp-shuffle tree followed by a
// shared-memory pass over the warp leaders.  Thread 0 normalises to affine (Montgomery form).
template <class F>
__global__ void __launch_bounds__(MSM_R2_THREADS)
k_msm_reduce2(const XYZZ<F>* __restrict__ seg_acc, const XYZZ<F>* __restrict__ seg_run, uint32_t nseg, uint32_t seg,
              Affine<F>* __restrict__ out) {
    __shared__ XYZZ<F> sh[2][MSM_R2_THREADS / 32];
    uint32_t b = blockIdx.x, u = threadIdx.x;
    uint32_t per = (nseg + MSM_R2_THREADS - 1) / MSM_R2_THREADS;
    uint32_t t0 = u * per, t1 = min(t0 + per, nseg);
    const XYZZ<F>* A = seg_acc + (size_t)b * nseg;
    const XYZZ<F>* Rn = seg_run + (size_t)b * nseg;
    XYZZ<F> asum = XYZZ<F>::inf(), run = XYZZ<F>::inf(), wsum = XYZZ<F>::inf();
    if (t0 < t1) {
        for (int t = (int)t1 - 1; t >= (int)t0; t--) {
            asum.add(A[t]);
            run.add(Rn[t]);
            wsum.add(run);
        }
        // wsum = sum (t - t0 + 1) run_t  ->  sum t*run_t = wsum + (t0 - 1) * run
        if (t0 == 0) wsum.add(run.neg());
        else if (t0 > 1) wsum.add(small_mul(run, t0 - 1));
    }
#pragma unroll
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
    if (u == 0) {
        for (int w = 1; w < MSM_R2_THREADS / 32; w++) {
            asum.add(sh[0][w]);
            wsum.add(sh[1][w]);
        }
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
    KernelProfiler* prof = nullptr;   // optional: times the accumulate kernel

   private:
    int reserve(const MsmBases<F>& bases, size_t batch);
    size_t cap_buckets = 0, cap_entries = 0, cap_segs = 0, cap_tiles = 0;
    uint32_t *counts = nullptr /* doubles as the scatter cursor */, *starts = nullptr, *tile_sums = nullptr,
             *entries = nullptr, *order = nullptr, *size_hist = nullptr /* [2][MSM_SIZE_BINS]: hist, cursor */;
    XYZZ<F>*buckets = nullptr, *seg_acc = nullptr, *seg_run = nullptr;
};

}  // namespace g16

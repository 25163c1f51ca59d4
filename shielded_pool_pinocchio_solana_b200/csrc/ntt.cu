// ntt.cu -- kernels and host orchestration for ntt.cuh.
#include "ntt.cuh"

namespace g16 {

constexpr int NTT_THREADS = 256;
#ifndef NTT_MIN_CTAS
#define NTT_MIN_CTAS 2   // 3 (80 registers, spills) measured 3 % slower: the wide-multiplier dependency chains are the limit
#endif
constexpr int NTT_TILE_LOG = 11;  // 2048 elements = 64 KiB of shared memory per CTA
constexpr int NTT_KC = 7;         // stages owned by the contiguous pass
constexpr int NTT_KMAX = 8;       // max stages per strided pass

// ---- shared-memory tile, limb-major ---------------------------------------------------------
// One padding word per 32 positions: the butterflies of a stage touch positions 2^s (radix 2) or 4 * 2^s (radix 4)
// apart, which without padding lands 2-8 lanes of a warp on the same bank (ncu, round 2: 1.4 conflicts per request
// in the contiguous pass).
__host__ __device__ __forceinline__ int tile_words(int T) { return T + (T >> 5); }
__device__ __forceinline__ Fr tile_ld(const uint32_t* sm, int T, int pos) {
    Fr r;
    const int TP = tile_words(T), pp = pos + (pos >> 5);
#pragma unroll
    for (int l = 0; l < 8; l++) r.v[l] = sm[l * TP + pp];
    return r;
}
__device__ __forceinline__ void tile_st(uint32_t* sm, int T, int pos, const Fr& x) {
    const int TP = tile_words(T), pp = pos + (pos >> 5);
#pragma unroll
    for (int l = 0; l < 8; l++) sm[l * TP + pp] = x.v[l];
}

template <bool DIT>
__device__ __forceinline__ void butterfly(uint32_t* sm, int T, int p0, int p1, const Fr& w) {
    Fr u = tile_ld(sm, T, p0), v = tile_ld(sm, T, p1);
    if (DIT) {
        v = v * w;
        tile_st(sm, T, p0, u + v);
        tile_st(sm, T, p1, u - v);
    } else {
        tile_st(sm, T, p0, u + v);
        tile_st(sm, T, p1, (u - v) * w);
    }
}

// Two stages at once on four elements held in registers (radix 4): halves the shared-memory round trips and the
// barriers of a pass.  p0 = position of the element whose bits `lo` and `hi = lo + 1` (of the tile-local row index)
// are both clear, d1 / d2 = position strides of those bits; wA, wB = stage-hi twiddles of the pairs (e0,e2) / (e1,e3),
// wC = stage-lo twiddle (shared by both of its pairs).  DIF runs stage hi first, DIT stage lo first.
template <bool DIT>
__device__ __forceinline__ void quad(uint32_t* sm, int T, int p0, int d1, int d2, const Fr& wA, const Fr& wB, const Fr& wC) {
    Fr e0 = tile_ld(sm, T, p0), e1 = tile_ld(sm, T, p0 + d1), e2 = tile_ld(sm, T, p0 + d2), e3 = tile_ld(sm, T, p0 + d1 + d2);
    if (DIT) {
        Fr t1 = e1 * wC, t3 = e3 * wC;
        Fr a0 = e0 + t1, a1 = e0 - t1, a2 = e2 + t3, a3 = e2 - t3;
        Fr u2 = a2 * wA, u3 = a3 * wB;
        e0 = a0 + u2; e2 = a0 - u2; e1 = a1 + u3; e3 = a1 - u3;
    } else {
        Fr a0 = e0 + e2, a2 = (e0 - e2) * wA, a1 = e1 + e3, a3 = (e1 - e3) * wB;
        e0 = a0 + a1; e1 = (a0 - a1) * wC; e2 = a2 + a3; e3 = (a2 - a3) * wC;
    }
    tile_st(sm, T, p0, e0);
    tile_st(sm, T, p0 + d1, e1);
    tile_st(sm, T, p0 + d2, e2);
    tile_st(sm, T, p0 + d1 + d2, e3);
}

// Strided pass: stages s_lo..s_hi (s_lo >= logC).  Tile = 2^k rows x C contiguous elements.
template <bool DIT>
__global__ void __launch_bounds__(NTT_THREADS, NTT_MIN_CTAS)
k_ntt_strided(Fr* __restrict__ data, size_t vec_stride, unsigned logn, int s_hi, int s_lo, int logC,
              const Fr* __restrict__ tw, const Fr* __restrict__ pre, const Fr* __restrict__ post) {
    extern __shared__ uint32_t sm[];
    const int k = s_hi - s_lo + 1, logT = k + logC, T = 1 << logT, C = 1 << logC;
    const unsigned tiles_per_vec = 1u << (logn - logT);
    const unsigned vec = blockIdx.x / tiles_per_vec, t = blockIdx.x % tiles_per_vec;
    const unsigned lo_tiles = 1u << (s_lo - logC);
    const unsigned lo0 = (t % lo_tiles) << logC, hi_part = t / lo_tiles;
    const size_t base = ((size_t)hi_part << (s_hi + 1)) + lo0;
    Fr* x = data + (size_t)vec * vec_stride;

    for (int e = threadIdx.x; e < T; e += NTT_THREADS) {
        int m = e >> logC, c = e & (C - 1);
        size_t idx = base + ((size_t)m << s_lo) + c;
        Fr v = x[idx];
        if (pre) v = v * pre[idx];
        tile_st(sm, T, e, v);
    }
    __syncthreads();
    auto single = [&](int ls) {
        const int s = s_lo + ls, hm = 1 << ls;
        for (int b = threadIdx.x; b < T / 2; b += NTT_THREADS) {
            int c = b & (C - 1), r = b >> logC;
            int m = ((r >> ls) << (ls + 1)) | (r & (hm - 1));
            int p0 = (m << logC) | c;
            size_t j = ((size_t)(r & (hm - 1)) << s_lo) | (lo0 + c);
            butterfly<DIT>(sm, T, p0, p0 + (hm << logC), tw[j << (logn - 1 - s)]);
        }
        __syncthreads();
    };
    auto pair = [&](int lo) {   // tile-local stages lo and lo + 1
        const int h1 = 1 << lo, hi = lo + 1, sh = s_lo + hi;
        for (int q = threadIdx.x; q < T / 4; q += NTT_THREADS) {
            int c = q & (C - 1), r = q >> logC;
            int low = r & (h1 - 1);
            int m0 = ((r >> lo) << (lo + 2)) | low;
            size_t jlow = ((size_t)low << s_lo) | (lo0 + c);
            const Fr wA = tw[jlow << (logn - 1 - sh)];
            const Fr wB = tw[(jlow + ((size_t)h1 << s_lo)) << (logn - 1 - sh)];
            const Fr wC = tw[jlow << (logn - sh)];
            quad<DIT>(sm, T, (m0 << logC) | c, h1 << logC, (h1 << 1) << logC, wA, wB, wC);
        }
        __syncthreads();
    };
    if (DIT) {
        int ls = 0;
        for (; ls + 1 < k; ls += 2) pair(ls);
        if (ls < k) single(ls);
    } else {
        int ls = k - 1;
        for (; ls >= 1; ls -= 2) pair(ls - 1);
        if (ls == 0) single(0);
    }
    for (int e = threadIdx.x; e < T; e += NTT_THREADS) {
        int m = e >> logC, c = e & (C - 1);
        size_t idx = base + ((size_t)m << s_lo) + c;
        Fr v = tile_ld(sm, T, e);
        if (post) v = v * post[idx];
        x[idx] = v;
    }
}

// Contiguous pass: stages 0..k-1 on T = 2^logT adjacent elements (2^(logT-k) groups of 2^k).
template <bool DIT>
__global__ void __launch_bounds__(NTT_THREADS, NTT_MIN_CTAS)
k_ntt_contig(Fr* __restrict__ data, size_t vec_stride, unsigned logn, int k, int logT, const Fr* __restrict__ tw,
             const Fr* __restrict__ pre, const Fr* __restrict__ post) {
    extern __shared__ uint32_t sm[];
    const int T = 1 << logT;
    const unsigned tiles_per_vec = 1u << (logn - logT);
    const unsigned vec = blockIdx.x / tiles_per_vec, t = blockIdx.x % tiles_per_vec;
    const size_t base = (size_t)t << logT;
    Fr* x = data + (size_t)vec * vec_stride;

    for (int e = threadIdx.x; e < T; e += NTT_THREADS) {
        Fr v = x[base + e];
        if (pre) v = v * pre[base + e];
        tile_st(sm, T, e, v);
    }
    __syncthreads();
    auto single = [&](int s) {
        const int h = 1 << s;
        for (int b = threadIdx.x; b < T / 2; b += NTT_THREADS) {
            int r = b & ((1 << (k - 1)) - 1), g = b >> (k - 1);
            int m = ((r >> s) << (s + 1)) | (r & (h - 1));
            int p0 = (g << k) | m;
            size_t j = (size_t)(r & (h - 1));
            butterfly<DIT>(sm, T, p0, p0 + h, tw[j << (logn - 1 - s)]);
        }
        __syncthreads();
    };
    auto pair = [&](int lo) {   // stages lo and lo + 1 (k >= 2)
        const int h1 = 1 << lo, sh = lo + 1;
        for (int q = threadIdx.x; q < T / 4; q += NTT_THREADS) {
            int r = q & ((1 << (k - 2)) - 1), g = q >> (k - 2);
            int low = r & (h1 - 1);
            int m0 = ((r >> lo) << (lo + 2)) | low;
            const Fr wA = tw[(size_t)low << (logn - 1 - sh)];
            const Fr wB = tw[(size_t)(low + h1) << (logn - 1 - sh)];
            const Fr wC = tw[(size_t)low << (logn - sh)];
            quad<DIT>(sm, T, (g << k) | m0, h1, h1 << 1, wA, wB, wC);
        }
        __syncthreads();
    };
    if (DIT) {
        int ls = 0;
        for (; ls + 1 < k; ls += 2) pair(ls);
        if (ls < k) single(ls);
    } else {
        int ls = k - 1;
        for (; ls >= 1; ls -= 2) pair(ls - 1);
        if (ls == 0) single(0);
    }
    for (int e = threadIdx.x; e < T; e += NTT_THREADS) {
        Fr v = tile_ld(sm, T, e);
        if (post) v = v * post[base + e];
        x[base + e] = v;
    }
}

// a := (a*b - c) * den   for every proof triple (a,b,c consecutive vectors of length n)
static __global__ void __launch_bounds__(256) k_h_pointwise(Fr* __restrict__ abc, size_t n, size_t total, Fr den) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    size_t p = i / n, e = i % n;
    Fr* a = abc + 3 * p * n;
    a[e] = (a[e] * a[n + e] - a[2 * n + e]) * den;
}

__device__ __forceinline__ uint32_t bitrev32(uint32_t x, unsigned bits) { return __brev(x) >> (32 - bits); }

// out[perm(i)] = scale * base^i, i < count; perm = bit reversal over `revbits` bits when revbits != 0
static __global__ void __launch_bounds__(256) k_powers(Fr* out, size_t count, Fr base, Fr scale, unsigned revbits) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    Fr acc = scale, b = base;
    for (size_t e = i; e; e >>= 1) {
        if (e & 1) acc = acc * b;
        b = b.sqr();
    }
    size_t pos = revbits ? bitrev32((uint32_t)i, revbits) : i;
    out[pos] = acc;
}

// ---- host side -------------------------------------------------------------------------------------
static Fr fr_from_u64(uint64_t x) {
    Fr f = Fr::zero();
    f.v[0] = (uint32_t)x;
    f.v[1] = (uint32_t)(x >> 32);
    return f.to_mont();
}

// gnark-crypto fr root of unity of order 2^28, canonical little-endian limbs
static const uint32_t ROOT_2_28[8] = {0x725b19f0u, 0x9bd61b6eu, 0x41112ed4u, 0x402d111eu,
                                      0x8ef62abcu, 0x00e0a7ebu, 0xa58a7e85u, 0x2a3c09f0u};

void NttEngine::release() {
    for (auto& kv : domains) {
        NttDomain& d = kv.second;
        cudaFree(d.tw_fwd); cudaFree(d.tw_inv); cudaFree(d.coset_nat); cudaFree(d.cosetinv_nat);
        cudaFree(d.coset_br); cudaFree(d.cosetinv_br); cudaFree(d.ninv_const);
    }
    domains.clear();
}

int NttEngine::domain(unsigned logn, cudaStream_t st, const NttDomain** out) {
    auto it = domains.find(logn);
    if (it != domains.end()) {
        *out = &it->second;
        return G16_OK;
    }
    if (logn < 1 || logn > 28) {
        set_error("ntt: domain size must be 2^1..2^28");
        return G16_E_ARG;
    }
    NttDomain d;
    d.logn = logn;
    d.n = (size_t)1 << logn;
    Fr w;
    for (int i = 0; i < 8; i++) w.v[i] = ROOT_2_28[i];
    w = w.to_mont();
    for (unsigned i = logn; i < 28; i++) w = w.sqr();
    Fr winv = w.inverse();
    Fr g = fr_from_u64(5), ginv = g.inverse();
    Fr ninv = fr_from_u64((uint64_t)d.n).inverse();
    Fr gn = g;
    for (unsigned i = 0; i < logn; i++) gn = gn.sqr();
    d.h_den = (gn - Fr::one()).inverse();

    size_t half = d.n / 2, n = d.n;
    G16_CUDA(cudaMalloc(&d.tw_fwd, sizeof(Fr) * (half ? half : 1)));
    G16_CUDA(cudaMalloc(&d.tw_inv, sizeof(Fr) * (half ? half : 1)));
    G16_CUDA(cudaMalloc(&d.coset_nat, sizeof(Fr) * n));
    G16_CUDA(cudaMalloc(&d.cosetinv_nat, sizeof(Fr) * n));
    G16_CUDA(cudaMalloc(&d.coset_br, sizeof(Fr) * n));
    G16_CUDA(cudaMalloc(&d.cosetinv_br, sizeof(Fr) * n));
    G16_CUDA(cudaMalloc(&d.ninv_const, sizeof(Fr) * n));
    Fr one = Fr::one(), zero_base = Fr::one();
    k_powers<<<cdiv(half, 256), 256, 0, st>>>(d.tw_fwd, half, w, one, 0);
    k_powers<<<cdiv(half, 256), 256, 0, st>>>(d.tw_inv, half, winv, one, 0);
    k_powers<<<cdiv(n, 256), 256, 0, st>>>(d.coset_nat, n, g, one, 0);
    k_powers<<<cdiv(n, 256), 256, 0, st>>>(d.cosetinv_nat, n, ginv, ninv, 0);
    k_powers<<<cdiv(n, 256), 256, 0, st>>>(d.coset_br, n, g, ninv, logn);
    k_powers<<<cdiv(n, 256), 256, 0, st>>>(d.cosetinv_br, n, ginv, ninv, logn);
    k_powers<<<cdiv(n, 256), 256, 0, st>>>(d.ninv_const, n, zero_base, ninv, 0);
    G16_CUDA(cudaGetLastError());
    auto ins = domains.emplace(logn, d);
    *out = &ins.first->second;
    return G16_OK;
}

namespace {
struct Pass {
    bool contig;
    int s_hi, s_lo, logC, logT;
};
// Stage plan, listed in DIF order (high stages first).
int plan(unsigned logn, Pass* out) {
    int np = 0;
    int kc = (int)logn < NTT_KC ? (int)logn : NTT_KC;
    int rem = (int)logn - kc;
    int nstr = (rem + NTT_KMAX - 1) / NTT_KMAX;
    int s_top = (int)logn - 1;
    for (int i = 0; i < nstr; i++) {
        int k = (rem + (nstr - i) - 1) / (nstr - i);  // even split, larger first
        Pass p;
        p.contig = false;
        p.s_hi = s_top;
        p.s_lo = s_top - k + 1;
        int logC = NTT_TILE_LOG - k;
        if (logC > p.s_lo) logC = p.s_lo;
        p.logC = logC;
        p.logT = k + logC;
        out[np++] = p;
        s_top -= k;
        rem -= k;
    }
    Pass c;
    c.contig = true;
    c.s_hi = kc - 1;
    c.s_lo = 0;
    c.logT = (int)logn < NTT_TILE_LOG ? (int)logn : NTT_TILE_LOG;
    c.logC = c.logT - kc;
    out[np++] = c;
    return np;
}
}  // namespace

int NttEngine::run(Fr* d_data, unsigned logn, size_t batch, NttDir dir, bool inverse_twiddles, const Fr* pre,
                   const Fr* post, cudaStream_t st) {
    return run_strided(d_data, (size_t)1 << logn, logn, batch, dir, inverse_twiddles, pre, post, st);
}

int NttEngine::run_strided(Fr* d_data, size_t vec_stride, unsigned logn, size_t batch, NttDir dir,
                           bool inverse_twiddles, const Fr* pre, const Fr* post, cudaStream_t st) {
    const NttDomain* d;
    G16_TRY(domain(logn, st, &d));
    if (batch == 0) return G16_OK;
    if (!attr_done) {   // per engine = per context = per device (the attribute is per device)
        int bytes = 32 * tile_words(1 << NTT_TILE_LOG);
        G16_CUDA(cudaFuncSetAttribute(k_ntt_strided<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        G16_CUDA(cudaFuncSetAttribute(k_ntt_strided<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        G16_CUDA(cudaFuncSetAttribute(k_ntt_contig<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        G16_CUDA(cudaFuncSetAttribute(k_ntt_contig<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        G16_CUDA(cudaFuncSetAttribute(k_ntt_strided<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        G16_CUDA(cudaFuncSetAttribute(k_ntt_strided<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        G16_CUDA(cudaFuncSetAttribute(k_ntt_contig<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        G16_CUDA(cudaFuncSetAttribute(k_ntt_contig<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        attr_done = true;
    }
    Pass passes[8];
    int np = plan(logn, passes);
    const Fr* tw = inverse_twiddles ? d->tw_inv : d->tw_fwd;
    for (int i = 0; i < np; i++) {
        const Pass& p = passes[dir == NTT_DIF ? i : np - 1 - i];
        const Fr* pre_i = (i == 0) ? pre : nullptr;
        const Fr* post_i = (i == np - 1) ? post : nullptr;
        size_t tiles = (((size_t)1 << logn) >> p.logT) * batch;
        if (tiles > 0x7fffffffull) {
            set_error("ntt: batch too large");
            return G16_E_ARG;
        }
        size_t smem = (size_t)32 * tile_words(1 << p.logT);
        if (p.contig) {
            if (dir == NTT_DIF)
                k_ntt_contig<false><<<(unsigned)tiles, NTT_THREADS, smem, st>>>(d_data, vec_stride, logn, p.s_hi + 1,
                                                                                 p.logT, tw, pre_i, post_i);
            else
                k_ntt_contig<true><<<(unsigned)tiles, NTT_THREADS, smem, st>>>(d_data, vec_stride, logn, p.s_hi + 1,
                                                                                p.logT, tw, pre_i, post_i);
        } else {
            if (dir == NTT_DIF)
                k_ntt_strided<false><<<(unsigned)tiles, NTT_THREADS, smem, st>>>(d_data, vec_stride, logn, p.s_hi,
                                                                                  p.s_lo, p.logC, tw, pre_i, post_i);
            else
                k_ntt_strided<true><<<(unsigned)tiles, NTT_THREADS, smem, st>>>(d_data, vec_stride, logn, p.s_hi,
                                                                                 p.s_lo, p.logC, tw, pre_i, post_i);
        }
        launches++;
    }
    G16_CUDA(cudaGetLastError());
    return G16_OK;
}

int NttEngine::compute_h(Fr* d_abc, unsigned logn, size_t nproofs, cudaStream_t st) {
    const NttDomain* d;
    G16_TRY(domain(logn, st, &d));
    const size_t n = d->n;
    // 1. interpolate a,b,c (DIF, inverse twiddles): bit-reversed coefficients, scaled by g^j / n
    G16_TRY(run_strided(d_abc, n, logn, 3 * nproofs, NTT_DIF, true, nullptr, d->coset_br, st));
    // 2. evaluate on the coset g*H (DIT, forward twiddles): natural order
    G16_TRY(run_strided(d_abc, n, logn, 3 * nproofs, NTT_DIT, false, nullptr, nullptr, st));
    // 3. h = (a*b - c) / (g^n - 1)
    size_t total = n * nproofs;
    k_h_pointwise<<<cdiv(total, 256), 256, 0, st>>>(d_abc, n, total, d->h_den);
    launches++;
    // 4. back to coefficients of H: DIF with inverse twiddles, times g^-j / n, bit-reversed order
    G16_TRY(run_strided(d_abc, 3 * n, logn, nproofs, NTT_DIF, true, nullptr, d->cosetinv_br, st));
    G16_CUDA(cudaGetLastError());
    return G16_OK;
}

}  // namespace g16

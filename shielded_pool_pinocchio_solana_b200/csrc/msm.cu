// msm.cu -- host-side orchestration of the batched fixed-base Pippenger MSM (see msm.cuh).
#include "msm.cuh"

#include <math.h>
#include <stdlib.h>

namespace g16 {

int msm_pick_window(size_t n, size_t batch) {
    // Work per batch element in modular multiplications: 10 per bucket addition (XYZZ += affine, n * W of
    // them for uniform scalars) plus the running-sum reduction, two full XYZZ additions (28) per bucket.  The
    // reduction runs as short dependent chains on few threads and reaches ~40 % of the multiplier throughput
    // the accumulation gets (profiles/r02_kernel_shares_device_step.csv), hence the weight of 70 -- and ties go
    // to the SMALLER window: witness-like scalars (mostly 0 / 1 / bytes) produce far fewer entries than
    // uniform ones, which moves their optimum further down.  G16_WINDOW_DELTA shifts the choice (experiments).
    (void)batch;
    int best = 8;
    double best_cost = 1e300;
    for (int c = 6; c <= 22; c++) {
        double W = ceil(254.0 / c);
        double nb = ldexp(1.0, c - 1);
        double cost = 10.0 * (double)n * W + 70.0 * nb;
        if (cost < best_cost * 0.995) {
            best_cost = cost;
            best = c;
        }
    }
    if (const char* e = getenv("G16_WINDOW_DELTA")) best += atoi(e);
    return best < 4 ? 4 : (best > 22 ? 22 : best);
}

template <class F>
void MsmBases<F>::release() {
    if (table) cudaFree(table);
    table = nullptr;
    n = 0;
}

template <class F>
int MsmBases<F>::load(const Affine<F>* host_pts, size_t n_, int c, int canonical, cudaStream_t st) {
    release();
    // windows below 3 bits are refused: with 254 % c == 0 the signed-digit carry out of the top window
    // would be lost (c = 2), and nothing is gained by such windows anyway
    if (n_ == 0 || c < 3 || c > 24) {
        set_error("MsmBases::load: bad n or window (3..24 bits)");
        return G16_E_ARG;
    }
    cfg = msm_config(c);
    if ((double)n_ * cfg.W >= 2147483648.0) {
        set_error("MsmBases::load: n*W exceeds 2^31 entry ids");
        return G16_E_ARG;
    }
    n = n_;
    G16_CUDA(cudaMalloc(&table, sizeof(Affine<F>) * n * cfg.W));
    G16_CUDA(cudaMemcpyAsync(table, host_pts, sizeof(Affine<F>) * n, cudaMemcpyHostToDevice, st));
    if (canonical) {
        size_t nfp = n * (sizeof(Affine<F>) / sizeof(Fp));
        k_fp_to_mont<<<cdiv(nfp, 256), 256, 0, st>>>(reinterpret_cast<Fp*>(table), nfp);
    }
    k_msm_expand<F><<<cdiv(n, 128), 128, 0, st>>>(table, (uint32_t)n, cfg.c, cfg.W);
    G16_CUDA(cudaGetLastError());
    G16_CUDA(cudaStreamSynchronize(st));
    return G16_OK;
}

int MsmScratch::ensure(size_t bytes) {
    if (bytes <= cap) return G16_OK;
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
    // grow with some slack so that slightly larger batches do not reallocate
    size_t want = bytes + bytes / 8 + 256;
    if (cudaMalloc(&ptr, want) != cudaSuccess) {
        cudaGetLastError();
        want = bytes;
        G16_CUDA(cudaMalloc(&ptr, want));
    }
    cap = want;
    return G16_OK;
}
void MsmScratch::release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
}

template <class F>
void MsmRunner<F>::release() {
    for (auto& x : s) x.release();
    if (ev_sorted) cudaEventDestroy(ev_sorted);
    if (ev_accumulated) cudaEventDestroy(ev_accumulated);
    ev_sorted = ev_accumulated = nullptr;
}

namespace {
constexpr int SM_COUNT = 148;   // B200
// exclusive scan of `n` uint32 (3 launches)
void scan_u32(const uint32_t* in, size_t n, uint32_t* tile_sums, uint32_t* out, uint32_t* out2, cudaStream_t st) {
    const uint32_t ntiles = cdiv(n, SCAN_TILE);
    k_scan_tile_sums<<<ntiles, SCAN_THREADS, 0, st>>>(in, n, tile_sums);
    k_scan_tiles<<<1, SCAN_THREADS, 0, st>>>(tile_sums, ntiles);
    k_scan_apply<<<ntiles, SCAN_THREADS, 0, st>>>(in, n, tile_sums, out, out2);
}
}  // namespace

template <class F>
int MsmRunner<F>::run(const MsmBases<F>& bases, const Fr* d_scalars, size_t stride, const uint32_t* d_map,
                      int montgomery, size_t batch, Affine<F>* d_out, cudaStream_t st, const Fr* d_scalars1,
                      size_t stride1, size_t lo, size_t cnt) {
    launches = 0;
    if (batch == 0) return G16_OK;
    if (!bases.table) {
        set_error("MsmRunner::run: bases not loaded");
        return G16_E_ARG;
    }
    const MsmConfig cfg = bases.cfg;
    if (cnt == 0) {
        lo = 0;
        cnt = bases.n;
    }
    if (lo + cnt > bases.n) {
        set_error("MsmRunner::run: point range out of bounds");
        return G16_E_ARG;
    }
    const uint32_t n = (uint32_t)cnt, n_total = (uint32_t)bases.n;
    const size_t nbuckets = batch * cfg.nb;
    const size_t nentries = batch * cnt * cfg.W;   // upper bound (zero digits produce no entry)
    if (nentries >= 4294967295.0 || nbuckets >= 2147483648.0) {
        set_error("MsmRunner: batch too large for 32-bit entry offsets");
        return G16_E_ARG;
    }
    // ---- virtual buckets ----------------------------------------------------------------------
    // A chain is at most `cap` additions long: ~2x the mean bucket size (Poisson tails stay unsplit),
    // so only the hot buckets of a skewed scalar vector are cut, into pieces that cost what a normal
    // bucket costs.
    const double mean = (double)cnt * cfg.W / cfg.nb;
    static const int ba_env = getenv("G16_MSM_BA") ? atoi(getenv("G16_MSM_BA")) : 0;
    static const int ba_cap = getenv("G16_BA_CAP") ? atoi(getenv("G16_BA_CAP")) : 16;
    const bool use_ba = ba_env != 0;
    uint32_t cap = 63;
    while (cap < VB_MAX_CAP && cap < 2.0 * mean) cap = 2 * cap + 1;
    // batched-affine chains are short: BA_K of them per thread have to add up to enough threads to fill the chip
    if (use_ba) cap = (uint32_t)(ba_cap < 2 ? 2 : (ba_cap > VB_MAX_CAP ? VB_MAX_CAP : ba_cap));
    const size_t nvmax = nbuckets + nentries / cap + 1;
    const size_t hot_cap = nentries / cap + 1;
    // ---- bucket reduction shape ----------------------------------------------------------------
    uint32_t seg = 64;
#ifndef MSM_MIN_SEG
#define MSM_MIN_SEG 4
#endif
    while (seg > MSM_MIN_SEG && nbuckets / seg < 65536) seg >>= 1;
    if (seg > cfg.nb) seg = cfg.nb;
    const uint32_t nseg = cfg.nb / seg;
    const size_t nseg_total = batch * nseg;
    uint32_t parts = cdiv(nseg, 1024);
    if (parts > 64) parts = 64;
    const uint32_t per = cdiv(nseg, parts * MSM_R2_THREADS);
    // ---- scratch ---------------------------------------------------------------------------------
    const size_t ntab = 2 * VB_CLASSES + (VB_CLASSES + 1) + 8;
    G16_TRY(s[S_COUNTS].ensure(4 * nbuckets));
    G16_TRY(s[S_STARTS].ensure(4 * nbuckets));
    G16_TRY(s[S_TILES].ensure(4 * ((size_t)cdiv(nbuckets + 1, SCAN_TILE) + 1)));
    G16_TRY(s[S_ENTRIES].ensure(4 * nentries));
    G16_TRY(s[S_NV].ensure(4 * (nbuckets + 1)));
    G16_TRY(s[S_VBASE].ensure(4 * (nbuckets + 1)));
    G16_TRY(s[S_VBSTART].ensure(4 * nvmax));
    G16_TRY(s[S_VBSIZE].ensure(4 * nvmax));
    G16_TRY(s[S_ORDER].ensure(4 * nvmax));
    G16_TRY(s[S_TABLES].ensure(4 * ntab));
    G16_TRY(s[S_HOT].ensure(4 * hot_cap));
    G16_TRY(s[S_RESULT].ensure(sizeof(XYZZ<F>) * nvmax));
    G16_TRY(s[S_SEGACC].ensure(sizeof(XYZZ<F>) * nseg_total));
    G16_TRY(s[S_SEGRUN].ensure(sizeof(XYZZ<F>) * nseg_total));
    G16_TRY(s[S_PARTS].ensure(sizeof(XYZZ<F>) * 2 * batch * parts));
    const uint32_t ba_lanes = cdiv(nvmax, BA_K);
    if (use_ba) {
        const size_t slots = (size_t)ba_lanes * BA_K;
        G16_TRY(s[S_BA_X].ensure(sizeof(F) * slots));
        G16_TRY(s[S_BA_Y].ensure(sizeof(F) * slots));
        G16_TRY(s[S_BA_PREFIX].ensure(sizeof(F) * slots));
        G16_TRY(s[S_BA_CS].ensure(4 * slots));
        G16_TRY(s[S_BA_CM].ensure(4 * slots));
    }
    uint32_t* counts = (uint32_t*)s[S_COUNTS].ptr;   // doubles as the scatter cursor
    uint32_t* starts = (uint32_t*)s[S_STARTS].ptr;
    uint32_t* tiles = (uint32_t*)s[S_TILES].ptr;
    uint32_t* entries = (uint32_t*)s[S_ENTRIES].ptr;
    uint32_t* nv = (uint32_t*)s[S_NV].ptr;
    uint32_t* vbase = (uint32_t*)s[S_VBASE].ptr;
    uint32_t* vb_start = (uint32_t*)s[S_VBSTART].ptr;
    uint32_t* vb_size = (uint32_t*)s[S_VBSIZE].ptr;
    uint32_t* order = (uint32_t*)s[S_ORDER].ptr;
    uint32_t* tab = (uint32_t*)s[S_TABLES].ptr;
    uint32_t* hist = tab;                            // [VB_CLASSES]
    uint32_t* cursor = tab + VB_CLASSES;             // [VB_CLASSES]
    uint32_t* hot_count = tab + 2 * VB_CLASSES;      // [1] (+7 pad)
    uint32_t* first = tab + 2 * VB_CLASSES + 8;      // [VB_CLASSES + 1]
    uint32_t* hot_list = (uint32_t*)s[S_HOT].ptr;
    XYZZ<F>* result_vb = (XYZZ<F>*)s[S_RESULT].ptr;
    XYZZ<F>* seg_acc = (XYZZ<F>*)s[S_SEGACC].ptr;
    XYZZ<F>* seg_run = (XYZZ<F>*)s[S_SEGRUN].ptr;
    XYZZ<F>* part_out = (XYZZ<F>*)s[S_PARTS].ptr;

    // ---- sort the signed digits by bucket ----------------------------------------------------------
    dim3 dgrid(cdiv(n, MSM_DIGIT_THREADS), (unsigned)batch);
    if (prof) prof->mark(label, "begin", st);
    G16_CUDA(cudaMemsetAsync(counts, 0, 4 * nbuckets, st));
    G16_CUDA(cudaMemsetAsync(tab, 0, 4 * (2 * VB_CLASSES + 8), st));
    k_msm_digits<0><<<dgrid, MSM_DIGIT_THREADS, 0, st>>>(d_scalars, stride, d_scalars1, stride1, d_map, n, (uint32_t)lo, n_total,
                                                         montgomery, cfg, counts, nullptr);
    scan_u32(counts, nbuckets, tiles, starts, counts, st);
    k_msm_digits<1><<<dgrid, MSM_DIGIT_THREADS, 0, st>>>(d_scalars, stride, d_scalars1, stride1, d_map, n, (uint32_t)lo, n_total,
                                                         montgomery, cfg, counts, entries);
    // after the scatter, counts[k] (the cursor) is the END of bucket k
    // ---- virtual buckets, size classes, schedule -------------------------------------------------------
    k_vb_count<<<cdiv(nbuckets + 1, MSM_VB_THREADS), MSM_VB_THREADS, 0, st>>>(starts, counts, (uint32_t)nbuckets, cap, nv);
    scan_u32(nv, nbuckets + 1, tiles, vbase, nullptr, st);
    k_vb_fill<<<cdiv(nbuckets, MSM_VB_THREADS), MSM_VB_THREADS, 0, st>>>(starts, counts, vbase, (uint32_t)nbuckets, cap,
                                                                         vb_start, vb_size, hist, hot_count, hot_list,
                                                                         (uint32_t)hot_cap);
    k_class_first<<<1, 1024, 0, st>>>(hist, first);
    k_vb_order<<<cdiv(nvmax, MSM_VB_THREADS), MSM_VB_THREADS, 0, st>>>(vb_size, vbase + nbuckets, first, cursor, order);
    launches += 12;
    // ---- bucket accumulation ---------------------------------------------------------------------------
    if (prof) prof->mark(label, "sorted", st);
    cudaStream_t sa = st;
    if (acc_stream && acc_stream != st && !(prof && prof->enabled)) {
        if (!ev_sorted) {
            G16_CUDA(cudaEventCreateWithFlags(&ev_sorted, cudaEventDisableTiming));
            G16_CUDA(cudaEventCreateWithFlags(&ev_accumulated, cudaEventDisableTiming));
        }
        sa = acc_stream;
        G16_CUDA(cudaEventRecord(ev_sorted, st));
        G16_CUDA(cudaStreamWaitEvent(sa, ev_sorted, 0));
    }
    if (prof) prof->begin(sizeof(F) == sizeof(Fp) ? PROF_MSM_ACC_G1 : PROF_MSM_ACC_G2, (double)batch * n, sa);
    if (use_ba)
        k_msm_accumulate_ba<F><<<cdiv(ba_lanes, BA_THREADS), BA_THREADS, 0, sa>>>(
            bases.table, entries, vb_start, vb_size, order, vbase + nbuckets, ba_lanes, (F*)s[S_BA_X].ptr, (F*)s[S_BA_Y].ptr,
            (F*)s[S_BA_PREFIX].ptr, (uint32_t*)s[S_BA_CS].ptr, (uint32_t*)s[S_BA_CM].ptr, result_vb);
    else
        k_msm_accumulate<F><<<cdiv(nvmax, 128), 128, 0, sa>>>(bases.table, entries, vb_start, vb_size, order,
                                                              vbase + nbuckets, result_vb);
    if (prof) prof->end(sa);
    if (sa != st) {
        G16_CUDA(cudaEventRecord(ev_accumulated, sa));
        G16_CUDA(cudaStreamWaitEvent(st, ev_accumulated, 0));
    }
    if (prof) prof->mark(label, "accumulated", st);
    {
        const unsigned jw = (unsigned)(cdiv(hot_cap, MSM_JOIN_THREADS / 32) < (unsigned)SM_COUNT * 8 ? cdiv(hot_cap, MSM_JOIN_THREADS / 32)
                                                                                                 : SM_COUNT * 8);
        const unsigned jc = (unsigned)(hot_cap / MSM_JOIN_WIDE + 1 < (size_t)SM_COUNT * 2 ? hot_cap / MSM_JOIN_WIDE + 1 : SM_COUNT * 2);
        k_vb_join<F, false><<<jw, MSM_JOIN_THREADS, 0, st>>>(hot_count, hot_list, (uint32_t)hot_cap, vbase, result_vb);
        k_vb_join<F, true><<<jc, MSM_JOIN_THREADS, 0, st>>>(hot_count, hot_list, (uint32_t)hot_cap, vbase, result_vb);
    }
    launches += 3;
    // ---- bucket reduction ------------------------------------------------------------------------------------
    k_msm_reduce1<F><<<cdiv(nseg_total, 128), 128, 0, st>>>(result_vb, vbase, seg, (uint32_t)nseg_total, seg_acc, seg_run);
    k_msm_reduce2<F><<<(unsigned)(batch * parts), MSM_R2_THREADS, 0, st>>>(seg_acc, seg_run, nseg, parts, per, part_out);
    if (parts == 1 && batch >= 32)
        k_msm_reduce3_flat<F><<<cdiv(batch, 32), 32, 0, st>>>(part_out, (uint32_t)batch, seg, d_out);
    else
        k_msm_reduce3<F><<<(unsigned)batch, 32, 0, st>>>(part_out, parts, seg, d_out);
    launches += 3;
    if (prof) prof->mark(label, "reduced", st);
    G16_CUDA(cudaGetLastError());
    return G16_OK;
}

template class MsmBases<Fp>;
template class MsmBases<Fp2>;
template class MsmRunner<Fp>;
template class MsmRunner<Fp2>;

}  // namespace g16

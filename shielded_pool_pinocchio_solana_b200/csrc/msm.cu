// msm.cu -- host-side orchestration of the batched fixed-base Pippenger MSM (see msm.cuh).
#include "msm.cuh"

#include <math.h>

namespace g16 {

int msm_pick_window(size_t n, size_t batch) {
    // cost per batch element in mixed-add equivalents: n*W(c) accumulate adds plus ~3.5 add-
    // equivalents per bucket for the two-level reduction (full XYZZ adds cost ~1.4x a mixed add).
    // Below ~64K bucket threads the accumulate kernel cannot fill 148 SMs, so small windows are
    // penalised by the occupancy they leave idle.
    const double kappa = 3.5;
    const double want_threads = 148.0 * 1024.0;
    int best = 8;
    double best_cost = 1e300;
    for (int c = 8; c <= 22; c++) {
        double W = ceil(254.0 / c);
        double nb = ldexp(1.0, c - 1);
        double cost = (double)n * W + kappa * nb;
        double threads = nb * (double)batch;
        if (threads < want_threads) cost *= want_threads / threads;  // idle lanes
        // a bucket needs a few entries on average or the warp diverges on empty buckets
        if (cost < best_cost) {
            best_cost = cost;
            best = c;
        }
    }
    return best;
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
    if (n_ == 0 || c < 2 || c > 24) {
        set_error("MsmBases::load: bad n or window");
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

template <class F>
void MsmRunner<F>::release() {
    cudaFree(counts); cudaFree(starts); cudaFree(tile_sums); cudaFree(entries); cudaFree(order); cudaFree(size_hist);
    order = size_hist = nullptr;
    cudaFree(buckets); cudaFree(seg_acc); cudaFree(seg_run);
    counts = starts = tile_sums = entries = nullptr;
    buckets = seg_acc = seg_run = nullptr;
    cap_buckets = cap_entries = cap_segs = cap_tiles = 0;
}

template <class F>
int MsmRunner<F>::reserve(const MsmBases<F>& bases, size_t batch) {
    size_t nbuckets = batch * bases.cfg.nb;
    size_t nentries = batch * bases.n * bases.cfg.W;
    size_t nsegs = nbuckets / bases.cfg.seg;
    size_t ntiles = cdiv(nbuckets, SCAN_TILE);
    if (nentries >= 4294967295.0 || nbuckets >= 2147483648.0) {
        set_error("MsmRunner: batch too large for 32-bit entry offsets");
        return G16_E_ARG;
    }
    if (nbuckets > cap_buckets) {
        cudaFree(counts); cudaFree(starts); cudaFree(buckets); cudaFree(order);
        counts = starts = order = nullptr; buckets = nullptr; cap_buckets = 0;
        G16_CUDA(cudaMalloc(&order, 4 * nbuckets));
        if (!size_hist) G16_CUDA(cudaMalloc(&size_hist, 4 * 2 * MSM_SIZE_BINS));
        G16_CUDA(cudaMalloc(&counts, 4 * nbuckets));
        G16_CUDA(cudaMalloc(&starts, 4 * nbuckets));
        G16_CUDA(cudaMalloc(&buckets, sizeof(XYZZ<F>) * nbuckets));
        cap_buckets = nbuckets;
    }
    if (nentries > cap_entries) {
        cudaFree(entries); entries = nullptr; cap_entries = 0;
        G16_CUDA(cudaMalloc(&entries, 4 * nentries));
        cap_entries = nentries;
    }
    if (nsegs > cap_segs) {
        cudaFree(seg_acc); cudaFree(seg_run); seg_acc = seg_run = nullptr; cap_segs = 0;
        G16_CUDA(cudaMalloc(&seg_acc, sizeof(XYZZ<F>) * nsegs));
        G16_CUDA(cudaMalloc(&seg_run, sizeof(XYZZ<F>) * nsegs));
        cap_segs = nsegs;
    }
    if (ntiles > cap_tiles) {
        cudaFree(tile_sums); tile_sums = nullptr; cap_tiles = 0;
        G16_CUDA(cudaMalloc(&tile_sums, 4 * ntiles));
        cap_tiles = ntiles;
    }
    return G16_OK;
}

template <class F>
int MsmRunner<F>::run(const MsmBases<F>& bases, const Fr* d_scalars, size_t stride, const uint32_t* d_map,
                      int montgomery, size_t batch, Affine<F>* d_out, cudaStream_t st, const Fr* d_scalars1,
                      size_t stride1) {
    launches = 0;
    if (batch == 0) return G16_OK;
    if (!bases.table) {
        set_error("MsmRunner::run: bases not loaded");
        return G16_E_ARG;
    }
    G16_TRY(reserve(bases, batch));
    const MsmConfig cfg = bases.cfg;
    const uint32_t n = (uint32_t)bases.n;
    const size_t nbuckets = batch * cfg.nb;
    const uint32_t ntiles = cdiv(nbuckets, SCAN_TILE);
    dim3 dgrid(cdiv(n, MSM_DIGIT_THREADS), (unsigned)batch);

    G16_CUDA(cudaMemsetAsync(counts, 0, 4 * nbuckets, st));
    k_msm_digits<0><<<dgrid, MSM_DIGIT_THREADS, 0, st>>>(d_scalars, stride, d_scalars1, stride1, d_map, n, montgomery, cfg,
                                                         counts, nullptr);
    k_scan_tile_sums<<<ntiles, SCAN_THREADS, 0, st>>>(counts, nbuckets, tile_sums);
    k_scan_tiles<<<1, SCAN_THREADS, 0, st>>>(tile_sums, ntiles);
    k_scan_apply<<<ntiles, SCAN_THREADS, 0, st>>>(counts, nbuckets, tile_sums, starts, counts);
    k_msm_digits<1><<<dgrid, MSM_DIGIT_THREADS, 0, st>>>(d_scalars, stride, d_scalars1, stride1, d_map, n, montgomery, cfg,
                                                         counts, entries);
    // after the scatter, counts[k] (the cursor) is the END of bucket k
    // schedule the buckets by decreasing size
    G16_CUDA(cudaMemsetAsync(size_hist, 0, 4 * MSM_SIZE_BINS, st));
    const unsigned ogrid = cdiv(nbuckets, MSM_ORDER_THREADS * MSM_ORDER_ITEMS);
    k_msm_size_hist<<<ogrid, MSM_ORDER_THREADS, 0, st>>>(starts, counts, (uint32_t)nbuckets, size_hist);
    k_msm_size_bins<<<1, MSM_SIZE_BINS, 0, st>>>(size_hist, size_hist + MSM_SIZE_BINS);
    k_msm_order<<<ogrid, MSM_ORDER_THREADS, 0, st>>>(starts, counts, (uint32_t)nbuckets, size_hist + MSM_SIZE_BINS, order);
    if (prof) prof->begin(sizeof(F) == sizeof(Fp) ? PROF_MSM_ACC_G1 : PROF_MSM_ACC_G2, (double)batch * n, st);
    k_msm_accumulate<F><<<cdiv(nbuckets, 128), 128, 0, st>>>(bases.table, entries, starts, counts, order, buckets,
                                                             (uint32_t)nbuckets);
    if (prof) prof->end(st);
    const uint32_t nseg = cfg.nb / cfg.seg;
    const size_t nseg_total = batch * nseg;
    k_msm_reduce1<F><<<cdiv(nseg_total, 128), 128, 0, st>>>(buckets, cfg.nb, cfg.seg, (uint32_t)nseg_total, seg_acc,
                                                            seg_run);
    k_msm_reduce2<F><<<(unsigned)batch, MSM_R2_THREADS, 0, st>>>(seg_acc, seg_run, nseg, cfg.seg, d_out);
    launches = 11;
    G16_CUDA(cudaGetLastError());
    return G16_OK;
}

template class MsmBases<Fp>;
template class MsmBases<Fp2>;
template class MsmRunner<Fp>;
template class MsmRunner<Fp2>;

}  // namespace g16

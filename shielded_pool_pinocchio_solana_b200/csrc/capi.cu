// capi.cu -- the extern "C" surface declared in include/g16b200.h (context, standalone MSM /
// NTT entry points, integer-pipe microbenchmark).  Circuit-level entry points live in prove.cu.
#include "capi.cuh"

#include <string.h>

#include <chrono>
#include <functional>
#include <thread>
#include <vector>

namespace g16 {

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
const char* get_error() { return g_err.c_str(); }

void trace(const char* tag, long a) {
    static const bool on = getenv("G16_TRACE") && atoi(getenv("G16_TRACE")) != 0;
    if (!on) return;
    static const auto t0 = std::chrono::steady_clock::now();
    long us = (long)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count();
    fprintf(stderr, "[g16 %9ld us t%04x] %s %ld\n", us, (unsigned)(std::hash<std::thread::id>()(std::this_thread::get_id()) & 0xffff), tag, a);
}

// ---- byte-order helpers (gnark wire formats <-> little-endian limbs) ---------------------------
void be32_to_limbs(const uint8_t* be, uint32_t* limbs) {
    for (int i = 0; i < 8; i++) {
        const uint8_t* p = be + 28 - 4 * i;
        limbs[i] = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
    }
}
void limbs_to_be32(const uint32_t* limbs, uint8_t* be) {
    for (int i = 0; i < 8; i++) {
        uint8_t* p = be + 28 - 4 * i;
        p[0] = limbs[i] >> 24; p[1] = limbs[i] >> 16; p[2] = limbs[i] >> 8; p[3] = limbs[i];
    }
}
// gnark raw G1: X||Y.  gnark-crypto's RawBytes writes infinity as all zeros (for bn254 the 0b01 flag
// is the 32-byte COMPRESSED infinity); 0x40 is still accepted on input.
void g1_from_be(const uint8_t* be, G1Affine* out) {
    if ((be[0] & 0xc0) == 0x40) { *out = G1Affine::inf(); return; }
    uint8_t tmp[32];
    memcpy(tmp, be, 32);
    tmp[0] &= 0x3f;
    be32_to_limbs(tmp, out->x.v);
    be32_to_limbs(be + 32, out->y.v);
}
void g1_to_be(const G1Affine& p, uint8_t* be) {
    if (p.is_inf()) { memset(be, 0, 64); return; }
    limbs_to_be32(p.x.v, be);
    limbs_to_be32(p.y.v, be + 32);
}
// gnark raw G2: X.A1||X.A0||Y.A1||Y.A0
void g2_from_be(const uint8_t* be, G2Affine* out) {
    if ((be[0] & 0xc0) == 0x40) { *out = G2Affine::inf(); return; }
    uint8_t tmp[32];
    memcpy(tmp, be, 32);
    tmp[0] &= 0x3f;
    be32_to_limbs(tmp, out->x.c1.v);
    be32_to_limbs(be + 32, out->x.c0.v);
    be32_to_limbs(be + 64, out->y.c1.v);
    be32_to_limbs(be + 96, out->y.c0.v);
}
void g2_to_be(const G2Affine& p, uint8_t* be) {
    if (p.is_inf()) { memset(be, 0, 128); return; }
    limbs_to_be32(p.x.c1.v, be);
    limbs_to_be32(p.x.c0.v, be + 32);
    limbs_to_be32(p.y.c1.v, be + 64);
    limbs_to_be32(p.y.c0.v, be + 96);
}

int DeviceBuf::ensure(size_t bytes) {
    if (bytes <= cap) return G16_OK;
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
    G16_CUDA(cudaMalloc(&ptr, bytes));
    cap = bytes;
    return G16_OK;
}
DeviceBuf::~DeviceBuf() {
    if (ptr) cudaFree(ptr);
}

static __global__ void k_fr_to_mont_pub(Fr* v, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = v[i].to_mont();
}

// ---- integer-pipe microbenchmark ---------------------------------------------------------------
// 8 independent dependency chains per thread, 2048 resident threads per SM: measures the issue
// rate of the instruction, not its latency.
template <int KIND>
__global__ void __launch_bounds__(256) k_imad_peak(uint32_t* out, uint32_t seed, int iters) {
    uint32_t a = seed + threadIdx.x, b = seed * 3 + blockIdx.x;
    uint32_t x[16];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = a + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int rep = 0; rep < 4; rep++) {
            if (KIND == 0) {
#pragma unroll
                for (int i = 0; i < 16; i++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
            } else {
#pragma unroll
                for (int i = 0; i < 16; i += 4)
                    asm volatile(
                        "mad.lo.cc.u32 %0, %4, %5, %0; madc.hi.cc.u32 %1, %4, %5, %1;\n\t"
                        "madc.lo.cc.u32 %2, %4, %6, %2; madc.hi.u32 %3, %4, %6, %3;"
                        : "+r"(x[i]), "+r"(x[i + 1]), "+r"(x[i + 2]), "+r"(x[i + 3])
                        : "r"(a), "r"(b), "r"(seed));
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s ^= x[i];
    if (s == 0x12345678u) out[0] = s;  // keep the chains alive
}

// FP64 FMA issue rate (16 independent chains per thread), for the record: DESIGN.md discusses why the
// field multiplier stays on the integer pipe.
__global__ void __launch_bounds__(256) k_dfma_peak(double* out, double seed, int iters) {
    double a = seed + threadIdx.x, b = seed * 3.0 + blockIdx.x;
    double x[16];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = a + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int rep = 0; rep < 4; rep++) {
#pragma unroll
            for (int i = 0; i < 16; i++) x[i] = fma(x[i], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += x[i];
    if (s == 0.12345) out[0] = s;
}

}  // namespace g16

using namespace g16;

namespace {
// out[0] = parts[0] + ... + parts[n-1]   (affine, Montgomery)
template <class F>
__global__ void k_sum_parts(const Affine<F>* __restrict__ parts, int n, Affine<F>* __restrict__ out) {
    if (threadIdx.x || blockIdx.x) return;
    XYZZ<F> acc = XYZZ<F>::inf();
    for (int i = 0; i < n; i++) acc.madd(parts[i]);
    out[0] = acc.to_affine();
}

// One large MSM as MSM_PARTS point slices on as many streams: the sort of a slice runs under the accumulation of
// another one and only the last slice's bucket-reduction tail stays exposed.
template <class F, class Runner>
int msm_dev_sliced(g16_ctx* ctx, Runner* runners, const MsmBases<F>& b, const Fr* d_scalars, int montgomery, Affine<F>* d_out) {
    constexpr int P = g16_ctx::MSM_PARTS;
    cudaStream_t st = ctx->stream;
    if (!ctx->part_fork) {
        G16_CUDA(cudaEventCreateWithFlags(&ctx->part_fork, cudaEventDisableTiming));
        for (int p = 0; p < P; p++) {
            G16_CUDA(cudaStreamCreateWithFlags(&ctx->part_stream[p], cudaStreamNonBlocking));
            G16_CUDA(cudaEventCreateWithFlags(&ctx->part_done[p], cudaEventDisableTiming));
        }
    }
    G16_TRY(ctx->part_out.ensure(sizeof(Affine<F>) * P));
    Affine<F>* parts = (Affine<F>*)ctx->part_out.ptr;
    G16_CUDA(cudaEventRecord(ctx->part_fork, st));
    int launches = 0;
    for (int p = 0; p < P; p++) {
        const size_t lo = b.n * p / P, hi = b.n * (p + 1) / P;
        G16_CUDA(cudaStreamWaitEvent(ctx->part_stream[p], ctx->part_fork, 0));
        runners[p].prof = &ctx->prof;
        G16_TRY(runners[p].run(b, d_scalars, b.n, nullptr, montgomery, 1, parts + p, ctx->part_stream[p], nullptr, 0, lo, hi - lo));
        launches += runners[p].launches;
        G16_CUDA(cudaEventRecord(ctx->part_done[p], ctx->part_stream[p]));
        G16_CUDA(cudaStreamWaitEvent(st, ctx->part_done[p], 0));
    }
    k_sum_parts<F><<<1, 32, 0, st>>>(parts, P, d_out);
    G16_CUDA(cudaGetLastError());
    ctx->last_launches = launches + 1;
    return G16_OK;
}
}  // namespace


extern "C" {

const char* g16_last_error(void) { return get_error(); }

int g16_init(const int* device_ids, int n_devices, g16_ctx** out) {
    if (!out || n_devices != 1 || !device_ids) {
        set_error("g16_init: exactly one device per context (one process per GPU)");
        return G16_E_ARG;
    }
    // A circuit handle drives ~20 streams (one per MSM, the side streams, the witness stage).  With the default of 8
    // hardware connections they alias: the witness stage of group k+1 then waits behind whole kernel sequences of
    // group k that merely share its queue (measured: 38 ms for a 1 ms copy).  Read when the CUDA context is created,
    // so it only takes effect if this is the first CUDA user of the process (the Python package sets it at import).
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error(std::string("g16_init: no CUDA device: ") + cudaGetErrorString(e) +
                  " (this library has no CPU fallback)");
        return G16_E_CUDA;
    }
    if (device_ids[0] < 0 || device_ids[0] >= count) {
        set_error("g16_init: device id out of range");
        return G16_E_ARG;
    }
    G16_CUDA(cudaSetDevice(device_ids[0]));
    cudaDeviceProp prop;
    G16_CUDA(cudaGetDeviceProperties(&prop, device_ids[0]));
    if (prop.major < 10) {
        set_error("g16_init: kernels are built for sm_100a only; found sm_" + std::to_string(prop.major) +
                  std::to_string(prop.minor));
        return G16_E_CUDA;
    }
    g16_ctx* c = new g16_ctx();
    c->device = device_ids[0];
    c->sm_count = prop.multiProcessorCount;
    G16_CUDA(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    c->g1.prof = &c->prof;
    c->g2.prof = &c->prof;
    *out = c;
    return G16_OK;
}

void g16_shutdown(g16_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ctx->g1.release();
    ctx->g2.release();
    for (int p = 0; p < g16_ctx::MSM_PARTS; p++) {
        ctx->g1_part[p].release();
        ctx->g2_part[p].release();
        if (ctx->part_stream[p]) cudaStreamDestroy(ctx->part_stream[p]);
        if (ctx->part_done[p]) cudaEventDestroy(ctx->part_done[p]);
    }
    if (ctx->part_fork) cudaEventDestroy(ctx->part_fork);
    ctx->ntt.release();
    comm_release(ctx);
    cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

int g16_set_stream(g16_ctx* ctx, void* cuda_stream) {
    if (!ctx) return G16_E_ARG;
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return G16_OK;
}

int g16_sync(g16_ctx* ctx) {
    if (!ctx) return G16_E_ARG;
    G16_CUDA(cudaSetDevice(ctx->device));
    G16_LOCK(ctx);
    G16_TRY(g16_join(ctx));
    G16_CUDA(cudaStreamSynchronize(ctx->stream));
    return G16_OK;
}

int g16_last_launches(g16_ctx* ctx) { return ctx ? ctx->last_launches : 0; }

int g16_profile_enable(g16_ctx* ctx, int enable) {
    if (!ctx) return G16_E_ARG;
    ctx->prof.enabled = enable != 0;
    return G16_OK;
}
int g16_profile_read(g16_ctx* ctx, double ms[8], double launches[8], double units[8]) {
    if (!ctx || !ms || !launches || !units) return G16_E_ARG;
    G16_CUDA(cudaSetDevice(ctx->device));
    G16_LOCK(ctx);
    ctx->prof.timeline_dump();
    ctx->prof.read(ms, launches, units);
    return G16_OK;
}

int g16_fr_to_device(g16_ctx* ctx, const uint8_t* values_be, size_t count, void* d_out) {
    if (!ctx || !values_be || !d_out) return G16_E_ARG;
    G16_CUDA(cudaSetDevice(ctx->device));
    G16_LOCK(ctx);
    std::vector<Fr> tmp(count);
    for (size_t i = 0; i < count; i++) be32_to_limbs(values_be + 32 * i, tmp[i].v);
    G16_CUDA(cudaMemcpyAsync(d_out, tmp.data(), sizeof(Fr) * count, cudaMemcpyHostToDevice, ctx->stream));
    k_fr_to_mont_pub<<<cdiv(count, 256), 256, 0, ctx->stream>>>((Fr*)d_out, count);
    G16_CUDA(cudaStreamSynchronize(ctx->stream));
    return G16_OK;
}

int g16_measure_imad_peak(g16_ctx* ctx, int kind, double* instr_per_s) {
    if (!ctx || !instr_per_s) return G16_E_ARG;
    G16_CUDA(cudaSetDevice(ctx->device));
    G16_LOCK(ctx);
    G16_TRY(ctx->scratch.ensure(64));
    const int iters = 4096;
    const int blocks = ctx->sm_count * 8;
    cudaEvent_t e0, e1;
    G16_CUDA(cudaEventCreate(&e0));
    G16_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        G16_CUDA(cudaEventRecord(e0, ctx->stream));
        if (kind == 0) k_imad_peak<0><<<blocks, 256, 0, ctx->stream>>>((uint32_t*)ctx->scratch.ptr, 7u + rep, iters);
        else if (kind == 1) k_imad_peak<1><<<blocks, 256, 0, ctx->stream>>>((uint32_t*)ctx->scratch.ptr, 7u + rep, iters);
        else k_dfma_peak<<<blocks, 256, 0, ctx->stream>>>((double*)ctx->scratch.ptr, 1.0000001 + rep, iters);
        G16_CUDA(cudaEventRecord(e1, ctx->stream));
        G16_CUDA(cudaEventSynchronize(e1));
        float ms;
        G16_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    // kind 0: 64 IMAD per inner iteration; kind 1: 16 chains of 4 = 64 mad.lo/hi halves = 32 IMAD.WIDE
    // kind 2: 64 DFMA per inner iteration
    double per_thread = (double)iters * (kind == 1 ? 32.0 : 64.0);
    *instr_per_s = per_thread * 256.0 * blocks / (best * 1e-3);
    return G16_OK;
}

// ---- bases ---------------------------------------------------------------------------------------
static int bases_load(g16_ctx* ctx, const uint8_t* points_be, size_t n, int window, size_t batch_hint, int g2,
                      g16_bases** out) {
    if (!ctx || !points_be || !out || n == 0) {
        set_error("g16_bases_load: bad arguments");
        return G16_E_ARG;
    }
    G16_CUDA(cudaSetDevice(ctx->device));
    G16_LOCK(ctx);
    if (batch_hint == 0) batch_hint = 1;
    int c = window ? window : msm_pick_window(n, batch_hint);
    g16_bases* b = new g16_bases();
    b->g2 = g2;
    int rc;
    if (!g2) {
        std::vector<G1Affine> pts(n);
        for (size_t i = 0; i < n; i++) g1_from_be(points_be + 64 * i, &pts[i]);
        rc = b->b1.load(pts.data(), n, c, /*canonical=*/1, ctx->stream);
    } else {
        std::vector<G2Affine> pts(n);
        for (size_t i = 0; i < n; i++) g2_from_be(points_be + 128 * i, &pts[i]);
        rc = b->b2.load(pts.data(), n, c, 1, ctx->stream);
    }
    if (rc != G16_OK) {
        delete b;
        return rc;
    }
    *out = b;
    return G16_OK;
}

int g16_bases_load_g1(g16_ctx* ctx, const uint8_t* points_be, size_t n, int window, size_t batch_hint,
                      g16_bases** out) {
    return bases_load(ctx, points_be, n, window, batch_hint, 0, out);
}
int g16_bases_load_g2(g16_ctx* ctx, const uint8_t* points_be, size_t n, int window, size_t batch_hint,
                      g16_bases** out) {
    return bases_load(ctx, points_be, n, window, batch_hint, 1, out);
}
void g16_bases_free(g16_bases* b) { delete b; }
int g16_bases_window(const g16_bases* b) { return b ? (b->g2 ? b->b2.cfg.c : b->b1.cfg.c) : 0; }

int g16_msm_dev(g16_ctx* ctx, const g16_bases* bases, const void* d_scalars, int montgomery, size_t batch,
                void* d_out) {
    if (!ctx || !bases || !d_scalars || !d_out) {
        set_error("g16_msm_dev: bad arguments");
        return G16_E_ARG;
    }
    G16_CUDA(cudaSetDevice(ctx->device));
    G16_LOCK(ctx);
    // Slicing multiplies the bucket sets to reduce by MSM_PARTS: measured a gain (3-12 %) for windows up to 17 bits
    // (2^16 .. 2^20 points) and a loss from 20 bits on (2^21 and above), where the single stream stays.
    // G16_MSM_SLICE_MIN overrides the lower size bound and lifts the window bound (0 = never slice).
    const char* env = getenv("G16_MSM_SLICE_MIN");
    const long slice_min = env ? atol(env) : (1l << 16);
    const size_t npts = bases->g2 ? bases->b2.n : bases->b1.n;
    const int c = bases->g2 ? bases->b2.cfg.c : bases->b1.cfg.c;
    if (batch == 1 && slice_min > 0 && npts >= (size_t)slice_min && (env || c <= 17) && !ctx->prof.enabled) {
        if (!bases->g2) return msm_dev_sliced<Fp>(ctx, ctx->g1_part, bases->b1, (const Fr*)d_scalars, montgomery, (G1Affine*)d_out);
        return msm_dev_sliced<Fp2>(ctx, ctx->g2_part, bases->b2, (const Fr*)d_scalars, montgomery, (G2Affine*)d_out);
    }
    int rc;
    if (!bases->g2) {
        rc = ctx->g1.run(bases->b1, (const Fr*)d_scalars, bases->b1.n, nullptr, montgomery, batch, (G1Affine*)d_out,
                         ctx->stream);
        ctx->last_launches = ctx->g1.launches;
    } else {
        rc = ctx->g2.run(bases->b2, (const Fr*)d_scalars, bases->b2.n, nullptr, montgomery, batch, (G2Affine*)d_out,
                         ctx->stream);
        ctx->last_launches = ctx->g2.launches;
    }
    return rc;
}

static int msm_host(g16_ctx* ctx, const g16_bases* bases, const uint8_t* scalars_be, size_t batch, uint8_t* out_be,
                    int g2) {
    if (!ctx || !bases || !scalars_be || !out_be || bases->g2 != g2) {
        set_error("g16_msm: bad arguments (or G1/G2 bases mismatch)");
        return G16_E_ARG;
    }
    G16_CUDA(cudaSetDevice(ctx->device));
    G16_LOCK(ctx);
    size_t n = g2 ? bases->b2.n : bases->b1.n;
    size_t ptsz = g2 ? sizeof(G2Affine) : sizeof(G1Affine);
    std::vector<Fr> sc(batch * n);
    for (size_t i = 0; i < batch * n; i++) be32_to_limbs(scalars_be + 32 * i, sc[i].v);
    G16_TRY(ctx->scalars.ensure(sizeof(Fr) * batch * n));
    G16_TRY(ctx->results.ensure(ptsz * batch));
    G16_CUDA(cudaMemcpyAsync(ctx->scalars.ptr, sc.data(), sizeof(Fr) * batch * n, cudaMemcpyHostToDevice, ctx->stream));
    G16_TRY(g16_msm_dev(ctx, bases, ctx->scalars.ptr, /*montgomery=*/0, batch, ctx->results.ptr));
    size_t nfp = batch * ptsz / sizeof(Fp);
    k_fp_from_mont<<<cdiv(nfp, 256), 256, 0, ctx->stream>>>((Fp*)ctx->results.ptr, nfp);
    ctx->last_launches += 1;
    std::vector<uint8_t> host(ptsz * batch);
    G16_CUDA(cudaMemcpyAsync(host.data(), ctx->results.ptr, ptsz * batch, cudaMemcpyDeviceToHost, ctx->stream));
    G16_CUDA(cudaStreamSynchronize(ctx->stream));
    for (size_t b = 0; b < batch; b++) {
        if (!g2) g1_to_be(*reinterpret_cast<G1Affine*>(host.data() + ptsz * b), out_be + 64 * b);
        else g2_to_be(*reinterpret_cast<G2Affine*>(host.data() + ptsz * b), out_be + 128 * b);
    }
    return G16_OK;
}

int g16_msm_g1(g16_ctx* ctx, const g16_bases* bases, const uint8_t* scalars_be, size_t batch, uint8_t* out_be) {
    return msm_host(ctx, bases, scalars_be, batch, out_be, 0);
}
int g16_msm_g2(g16_ctx* ctx, const g16_bases* bases, const uint8_t* scalars_be, size_t batch, uint8_t* out_be) {
    return msm_host(ctx, bases, scalars_be, batch, out_be, 1);
}

// ---- NTT / quotient ---------------------------------------------------------------------------------
static __global__ void k_fr_to_mont(Fr* v, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = v[i].to_mont();
}
static __global__ void k_fr_from_mont(Fr* v, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = v[i].from_mont();
}

int g16_ntt_dev(g16_ctx* ctx, void* d_values, unsigned logn, size_t batch, int inverse, int coset) {
    if (!ctx || !d_values) return G16_E_ARG;
    G16_CUDA(cudaSetDevice(ctx->device));
    G16_LOCK(ctx);
    const NttDomain* d;
    G16_TRY(ctx->ntt.domain(logn, ctx->stream, &d));
    ctx->ntt.launches = 0;
    int rc;
    if (!inverse)
        rc = ctx->ntt.run((Fr*)d_values, logn, batch, NTT_DIF, false, coset ? d->coset_nat : nullptr, nullptr,
                          ctx->stream);
    else
        rc = ctx->ntt.run((Fr*)d_values, logn, batch, NTT_DIT, true, nullptr, coset ? d->cosetinv_nat : d->ninv_const,
                          ctx->stream);
    ctx->last_launches = ctx->ntt.launches;
    return rc;
}

// host buffers (big-endian canonical) -> device Montgomery limbs in ctx->scalars
static int upload_fr(g16_ctx* ctx, const uint8_t* be, size_t count) {
    std::vector<Fr> tmp(count);
    for (size_t i = 0; i < count; i++) be32_to_limbs(be + 32 * i, tmp[i].v);
    G16_TRY(ctx->scalars.ensure(sizeof(Fr) * count));
    G16_CUDA(cudaMemcpyAsync(ctx->scalars.ptr, tmp.data(), sizeof(Fr) * count, cudaMemcpyHostToDevice, ctx->stream));
    k_fr_to_mont<<<cdiv(count, 256), 256, 0, ctx->stream>>>((Fr*)ctx->scalars.ptr, count);
    G16_CUDA(cudaStreamSynchronize(ctx->stream));  // tmp goes out of scope
    return G16_OK;
}
static int download_fr(g16_ctx* ctx, Fr* d_src, size_t vec_len, size_t vec_stride, size_t nvec, uint8_t* be) {
    std::vector<Fr> tmp(vec_len * nvec);
    for (size_t v = 0; v < nvec; v++) {
        k_fr_from_mont<<<cdiv(vec_len, 256), 256, 0, ctx->stream>>>(d_src + v * vec_stride, vec_len);
        G16_CUDA(cudaMemcpyAsync(tmp.data() + v * vec_len, d_src + v * vec_stride, sizeof(Fr) * vec_len,
                                 cudaMemcpyDeviceToHost, ctx->stream));
    }
    G16_CUDA(cudaStreamSynchronize(ctx->stream));
    for (size_t i = 0; i < tmp.size(); i++) limbs_to_be32(tmp[i].v, be + 32 * i);
    return G16_OK;
}

int g16_ntt(g16_ctx* ctx, uint8_t* values_be, unsigned logn, size_t batch, int inverse, int coset) {
    if (!ctx || !values_be || logn < 1 || logn > 28) {
        set_error("g16_ntt: bad arguments");
        return G16_E_ARG;
    }
    G16_CUDA(cudaSetDevice(ctx->device));
    G16_LOCK(ctx);
    size_t n = (size_t)1 << logn;
    G16_TRY(upload_fr(ctx, values_be, n * batch));
    G16_TRY(g16_ntt_dev(ctx, ctx->scalars.ptr, logn, batch, inverse, coset));
    return download_fr(ctx, (Fr*)ctx->scalars.ptr, n * batch, n * batch, 1, values_be);
}

int g16_compute_h_dev(g16_ctx* ctx, void* d_abc, unsigned logn, size_t nproofs) {
    if (!ctx || !d_abc) return G16_E_ARG;
    G16_CUDA(cudaSetDevice(ctx->device));
    G16_LOCK(ctx);
    ctx->ntt.launches = 0;
    int rc = ctx->ntt.compute_h((Fr*)d_abc, logn, nproofs, ctx->stream);
    ctx->last_launches = ctx->ntt.launches;
    return rc;
}

int g16_compute_h(g16_ctx* ctx, const uint8_t* abc_be, unsigned logn, size_t nproofs, uint8_t* h_be) {
    if (!ctx || !abc_be || !h_be || logn < 1 || logn > 28) {
        set_error("g16_compute_h: bad arguments");
        return G16_E_ARG;
    }
    G16_CUDA(cudaSetDevice(ctx->device));
    G16_LOCK(ctx);
    size_t n = (size_t)1 << logn;
    G16_TRY(upload_fr(ctx, abc_be, 3 * n * nproofs));
    G16_TRY(g16_compute_h_dev(ctx, ctx->scalars.ptr, logn, nproofs));
    return download_fr(ctx, (Fr*)ctx->scalars.ptr, n, 3 * n, nproofs, h_be);
}

}  // extern "C"

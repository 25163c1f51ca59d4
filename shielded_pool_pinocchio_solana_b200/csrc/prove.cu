// prove.cu -- circuit handle + batched Groth16 proving pipeline behind g16_circuit_load /
// g16_prove* (include/g16b200.h).  See prove.cuh for what it replaces.
//
// Per batch of B proofs of one circuit (everything enqueued on the context stream):
//   host   : solve phase 1 (threads)                 -> committed values
//   device : commitment MSM  (B x 490 points)        -> D2H 64 B per proof
//   host   : hash_to_field challenge, solve phase 2  -> full wire vectors (Montgomery), H2D
//   device : R1CS SpMV  a,b,c = A.w, B.w, C.w        (k_r1cs_spmv)
//            quotient H                              (ntt.cu: 7 transforms, batched)
//            MSMs  A | B1 | B2 (G2) | K+Z | PoK      (msm.cu; alpha/beta/delta terms ride along as
//                                                     extra bases with scalars 1, r, s, -rs)
//            finalize: Krs += s*Ar + r*Bs1, out of Montgomery form
//   host   : gnark raw serialisation (388-byte proof, public witness)
#include "prove.cuh"

#include <string.h>

#include <algorithm>
#include <atomic>
#include <fstream>
#include <future>
#include <thread>

#include "hostutil.hpp"

namespace g16 {

// ------------------------------------------------------------------------------------------------
// proving-key file (gnark ProvingKey.WriteRawTo / WriteTo; layout per SURVEY.md 9.2, UNVERIFIED
// against a real gnark .pk -- none exists in the reference tree)
// ------------------------------------------------------------------------------------------------
namespace {
struct PkReader {
    const uint8_t* p;
    size_t len, off = 0;
    bool ok = true;
    bool need(size_t n) {
        if (off + n > len) ok = false;
        return ok;
    }
    uint32_t u32() {
        if (!need(4)) return 0;
        uint32_t v = ((uint32_t)p[off] << 24) | ((uint32_t)p[off + 1] << 16) | ((uint32_t)p[off + 2] << 8) | p[off + 3];
        off += 4;
        return v;
    }
    uint64_t u64() {
        uint64_t hi = u32();
        return (hi << 32) | u32();
    }
    bool g1(G1Affine* o) {
        if (!need(64)) return false;
        if ((p[off] & 0xc0) == 0x80 || (p[off] & 0xc0) == 0xc0) {  // compressed encodings
            ok = false;
            return false;
        }
        g1_from_be(p + off, o);
        off += 64;
        return true;
    }
    bool g2(G2Affine* o) {
        if (!need(128)) return false;
        if ((p[off] & 0xc0) == 0x80 || (p[off] & 0xc0) == 0xc0) {
            ok = false;
            return false;
        }
        g2_from_be(p + off, o);
        off += 128;
        return true;
    }
    bool g1s(std::vector<G1Affine>* v) {
        uint32_t n = u32();
        if (!ok || (size_t)n * 64 > len - off) return ok = false;
        v->resize(n);
        for (auto& x : *v)
            if (!g1(&x)) return false;
        return true;
    }
    bool g2s(std::vector<G2Affine>* v) {
        uint32_t n = u32();
        if (!ok || (size_t)n * 128 > len - off) return ok = false;
        v->resize(n);
        for (auto& x : *v)
            if (!g2(&x)) return false;
        return true;
    }
    bool bools(std::vector<uint8_t>* v) {
        uint32_t n = u32();
        if (!ok || n > len - off) return ok = false;
        v->assign(p + off, p + off + n);
        off += n;
        return true;
    }
};
}  // namespace

int parse_pk(const uint8_t* buf, size_t len, ProvingKeyHost* out) {
    // The fft.Domain header is 8 + 5*32 bytes, optionally followed by a 1-byte `withPrecompute`
    // flag (newer gnark-crypto).  Try both and keep the one that consumes the file exactly.
    for (int flag_byte = 1; flag_byte >= 0; flag_byte--) {
        PkReader r{buf, len};
        ProvingKeyHost pk;
        pk.domain = r.u64();
        r.need(5 * 32 + flag_byte);
        if (!r.ok) continue;
        r.off += 5 * 32 + flag_byte;
        r.g1(&pk.alpha1); r.g1(&pk.beta1); r.g1(&pk.delta1);
        r.g1s(&pk.A); r.g1s(&pk.B1); r.g1s(&pk.Z); r.g1s(&pk.K);
        r.g2(&pk.beta2); r.g2(&pk.delta2); r.g2s(&pk.B2);
        if (!r.ok) continue;
        uint64_t nw = r.u64(), ninf_a = r.u64(), ninf_b = r.u64();
        r.bools(&pk.infinity_a);
        r.bools(&pk.infinity_b);
        if (!r.ok || pk.infinity_a.size() != nw || pk.infinity_b.size() != nw) continue;
        uint64_t ca = 0, cb = 0;
        for (auto v : pk.infinity_a) ca += v != 0;
        for (auto v : pk.infinity_b) cb += v != 0;
        if (ca != ninf_a || cb != ninf_b || pk.A.size() != nw - ninf_a || pk.B1.size() != nw - ninf_b ||
            pk.B2.size() != pk.B1.size())
            continue;
        uint32_t nkeys = r.u32();
        if (!r.ok || nkeys > 64) continue;
        pk.commitment_keys.resize(nkeys);
        for (auto& k : pk.commitment_keys) {
            r.g1s(&k.basis);
            r.g1s(&k.basis_exp_sigma);
            if (r.ok && k.basis.size() != k.basis_exp_sigma.size()) r.ok = false;
        }
        if (!r.ok || r.off != len) continue;
        if (pk.domain == 0 || (pk.domain & (pk.domain - 1))) continue;
        *out = std::move(pk);
        return G16_OK;
    }
    set_error("parse_pk: not a raw (uncompressed) gnark groth16 proving key for BN254");
    return G16_E_PARSE;
}

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
// One thread per (proof, matrix, row): abc[proof][matrix][row] = sum coeff * w[wire].
// Rows >= nb_constraints are the zero padding of the FFT domain.
__global__ void __launch_bounds__(256)
k_r1cs_spmv(const uint32_t* __restrict__ rp0, const uint32_t* __restrict__ cid0, const uint32_t* __restrict__ wid0,
            const uint32_t* __restrict__ rp1, const uint32_t* __restrict__ cid1, const uint32_t* __restrict__ wid1,
            const uint32_t* __restrict__ rp2, const uint32_t* __restrict__ cid2, const uint32_t* __restrict__ wid2,
            const Fr* __restrict__ coeffs, const Fr* __restrict__ wires, size_t wstride, Fr* __restrict__ abc,
            uint32_t nrows, uint32_t n, int unit_ids) {
    uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t m = blockIdx.y, proof = blockIdx.z;
    if (row >= n) return;
    Fr acc = Fr::zero();
    if (row < nrows) {
        const uint32_t* rp = m == 0 ? rp0 : (m == 1 ? rp1 : rp2);
        const uint32_t* cid = m == 0 ? cid0 : (m == 1 ? cid1 : cid2);
        const uint32_t* wid = m == 0 ? wid0 : (m == 1 ? wid1 : wid2);
        const Fr* w = wires + (size_t)proof * wstride;
        for (uint32_t k = rp[row], e = rp[row + 1]; k < e; k++) {
            uint32_t c = cid[k];
            Fr x = w[wid[k]];
            // gnark's coefficient table starts 0, 1, 2, -1, -2 (checked at load -> unit_ids): skip
            // the multiplication for the +-1 entries that make up most real circuits
            if (unit_ids && c == 1) acc = acc + x;
            else if (unit_ids && c == 3) acc = acc - x;
            else if (!(unit_ids && c == 0)) acc = acc + coeffs[c] * x;
        }
    }
    abc[((size_t)proof * 3 + m) * n + row] = acc;
}

// 32 big-endian bytes (32-byte aligned) -> canonical little-endian limbs, reduced below r
__device__ __forceinline__ Fr fr_load_be(const uint4* __restrict__ p) {
    const uint4 hi = p[0], lo = p[1];
    Fr f;
    f.v[7] = __byte_perm(hi.x, 0, 0x0123); f.v[6] = __byte_perm(hi.y, 0, 0x0123);
    f.v[5] = __byte_perm(hi.z, 0, 0x0123); f.v[4] = __byte_perm(hi.w, 0, 0x0123);
    f.v[3] = __byte_perm(lo.x, 0, 0x0123); f.v[2] = __byte_perm(lo.y, 0, 0x0123);
    f.v[1] = __byte_perm(lo.z, 0, 0x0123); f.v[0] = __byte_perm(lo.w, 0, 0x0123);
    uint32_t m[8], d[8];
    Fr::modulus(m);
    for (int it = 0; it < 6; it++) {   // 256-bit inputs: at most 5 subtractions
        if (ff_sub8(d, f.v, m)) break;
#pragma unroll
        for (int l = 0; l < 8; l++) f.v[l] = d[l];
    }
    return f;
}

// Full wire vectors as the caller holds them (big-endian, nw per proof) -> Montgomery wire vectors with
// stride wstride; thread nw of every proof fills the X_* slots from rnd (r | s | unused), as k_assign does.
__global__ void __launch_bounds__(256)
k_wires_from_be(const uint4* __restrict__ wires_be, const uint4* __restrict__ rnd_be, Fr* __restrict__ wires,
                size_t wstride, uint32_t nw) {
    const uint32_t b = blockIdx.y;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    Fr* w = wires + (size_t)b * wstride;
    if (i < nw) {
        w[i] = fr_load_be(wires_be + ((size_t)b * nw + i) * 2).to_mont();
    } else if (i == nw) {
        const Fr r = fr_load_be(rnd_be + (size_t)b * 6).to_mont(), s = fr_load_be(rnd_be + (size_t)b * 6 + 2).to_mont();
        w[nw + X_ONE] = Fr::one();
        w[nw + X_R] = r;
        w[nw + X_S] = s;
        w[nw + X_NEG_RS] = (r * s).neg();
        for (int k = X_NEG_RS + 1; k < X_COUNT; k++) w[nw + k] = Fr::zero();
    }
}

__global__ void k_set_one(Fr* wires, size_t wstride, size_t slot, uint32_t n) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < n) wires[(size_t)b * wstride + slot] = Fr::one();
}

// [k]P for a 256-bit canonical scalar, affine base (double-and-add, MSB first)
__device__ G1XYZZ g1_scalar_mul(const G1Affine& p, const Fr& k) {
    G1XYZZ acc = G1XYZZ::inf();
    bool started = false;
    for (int bit = 253; bit >= 0; bit--) {
        if (started) acc = acc.dbl();
        if ((k.v[bit >> 5] >> (bit & 31)) & 1u) {
            acc.madd(p);
            started = true;
        }
    }
    return acc;
}

// The two 254-bit scalar multiplications of the Krs assembly, s*Ar and r*Bs1: ~3.5 K dependent modmuls each.
// One THREAD per ladder (2B threads): the step is issue-bound, and a warp with one working lane costs as many issue
// slots as a full one -- a CTA per proof with one ladder per warp was 1.5 % of a step's instructions for 128 ladders.
// Lanes diverge on the scalar bits (a warp runs the addition at every bit), which lengthens a ladder by a third; it
// runs on the side stream right behind the A and B1 MSMs, under the quotient and the K|Z MSM of the main stream.
__global__ void __launch_bounds__(32)
k_scalar_muls(const G1Affine* __restrict__ ar, const G1Affine* __restrict__ bs1, const Fr* __restrict__ wires,
              size_t wstride, size_t nw, uint32_t stride_b, uint32_t B, G1XYZZ* __restrict__ parts) {
    const uint32_t idx = blockIdx.x * 32 + threadIdx.x;
    if (idx >= 2 * B) return;
    const uint32_t b = idx >> 1, which = idx & 1;
    const Fr* x = wires + (size_t)b * wstride + nw;
    const Fr k = (which == 0 ? x[X_S] : x[X_R]).from_mont();
    parts[(size_t)which * stride_b + b] = g1_scalar_mul(which == 0 ? ar[b] : bs1[b], k);
}

// One 32-thread CTA per proof: lane 0 assembles Krs = s*Ar + r*Bs1 + (K|Z sum); lanes 0..3 convert the four
// outputs out of Montgomery form.
__global__ void __launch_bounds__(32)
k_finalize(const G1Affine* __restrict__ ar, const G1XYZZ* __restrict__ parts, uint32_t stride_b,
           const G2Affine* __restrict__ bs2, const G1Affine* __restrict__ kz, const G1Affine* __restrict__ pok,
           ProofPoints* __restrict__ out) {
    uint32_t b = blockIdx.x, t = threadIdx.x;
    auto unmont1 = [](G1Affine p) {
        p.x = p.x.from_mont();
        p.y = p.y.from_mont();
        return p;
    };
    if (t == 0) {
        G1XYZZ acc = parts[b];
        acc.add(parts[(size_t)stride_b + b]);
        acc.madd(kz[b]);
        out[b].krs = unmont1(acc.to_affine());
    } else if (t == 1) {
        out[b].ar = unmont1(ar[b]);
    } else if (t == 2) {
        G2Affine q = bs2[b];
        q.x.c0 = q.x.c0.from_mont(); q.x.c1 = q.x.c1.from_mont();
        q.y.c0 = q.y.c0.from_mont(); q.y.c1 = q.y.c1.from_mont();
        out[b].bs = q;
    } else if (t == 3) {
        out[b].pok = unmont1(pok[b]);
    }
}

}  // namespace g16

using namespace g16;

g16_circuit::~g16_circuit() {
    if (ctx) cudaSetDevice(ctx->device);
    for (int m = 0; m < 3; m++) {
        cudaFree(d_rowptr[m]); cudaFree(d_cid[m]); cudaFree(d_wid[m]);
    }
    cudaFree(d_coeffs);
    cudaFree(d_mapA); cudaFree(d_mapB); cudaFree(d_mapKZ); cudaFree(d_mapPok);
    if (ctx) cudaDeviceSynchronize();   // deferred chunks may still be in flight on the private streams
    for (int p = 0; p < 2; p++) {
        cudaFree(d_tmp_g1[p]); cudaFree(d_tmp_g2[p]); cudaFree(d_parts[p]);
        if (ev_fin[p]) cudaEventDestroy(ev_fin[p]);
    }
    if (fin_stream) cudaStreamDestroy(fin_stream);
    g1_kz.release();
    if (ctx) ctx->circuits.erase(std::remove(ctx->circuits.begin(), ctx->circuits.end(), this), ctx->circuits.end());
    cudaFree(d_map_commit);
    for (auto& sl : slots) {
        if (sl.h_stage) cudaFreeHost(sl.h_stage);
        if (sl.h_wires) cudaFreeHost(sl.h_wires);
        if (sl.ready) cudaEventDestroy(sl.ready);
    }
    for (int i = 0; i < 2; i++) {
        if (h_pts[i]) cudaFreeHost(h_pts[i]);
        if (ev_done[i]) cudaEventDestroy(ev_done[i]);
    }
    g1_aux.release();
    for (auto& r : g1_side) r.release();
    g2_side.release();
    if (aux_stream) cudaStreamDestroy(aux_stream);
    if (acc_stream) cudaStreamDestroy(acc_stream);
    for (int i = 0; i < N_SIDE; i++) {
        if (side[i]) cudaStreamDestroy(side[i]);
        if (ev_join[i]) cudaEventDestroy(ev_join[i]);
    }
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_h) cudaEventDestroy(ev_h);
}

namespace {

template <class T>
int upload_vec(const std::vector<T>& v, T** d, cudaStream_t st) {
    G16_CUDA(cudaMalloc(d, sizeof(T) * (v.size() ? v.size() : 1)));
    if (!v.empty()) G16_CUDA(cudaMemcpyAsync(*d, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice, st));
    return G16_OK;
}

int random_fr(HFr* out) {
    // one descriptor per thread, kept open: a 256-proof group draws 768 scalars
    static thread_local std::ifstream ur("/dev/urandom", std::ios::binary);
    if (!ur) {
        ur.clear();
        ur.close();
        ur.open("/dev/urandom", std::ios::binary);
    }
    for (int tries = 0; tries < 64; tries++) {
        uint8_t b[32];
        ur.read((char*)b, 32);
        if (!ur) break;
        b[0] &= 0x3f;
        uint64_t l[4];
        for (int i = 0; i < 4; i++) {
            uint64_t v = 0;
            for (int k = 0; k < 8; k++) v = (v << 8) | b[(3 - i) * 8 + k];
            l[i] = v;
        }
        if (!HFr::geq_mod(l)) {
            *out = HFr::from_be(b);
            return G16_OK;
        }
    }
    set_error("could not read /dev/urandom");
    return G16_E_INTERNAL;
}

size_t default_threads() {
    if (const char* e = getenv("G16_HOST_THREADS")) {
        long v = atol(e);
        if (v > 0) return (size_t)v;
    }
    unsigned hc = std::thread::hardware_concurrency();
    return hc ? hc : 4;
}

template <class Fn>
void parallel_for(size_t n, Fn fn) {
    size_t nt = std::min(default_threads(), n);
    if (nt <= 1) {
        for (size_t i = 0; i < n; i++) fn(i);
        return;
    }
    std::atomic<size_t> next{0};
    std::vector<std::thread> th;
    for (size_t t = 0; t < nt; t++)
        th.emplace_back([&] {
            for (size_t i; (i = next.fetch_add(1)) < n;) fn(i);
        });
    for (auto& x : th) x.join();
}

// Buffer set for the next chunk: the context stream waits until the chunk that last used it has been assembled.
// Call BEFORE staging that chunk's wires into d_wdev[p] / a pipeline slot.
int chunk_begin(g16_circuit* c, int* parity) {
    const int p = c->parity;
    c->parity ^= 1;
    G16_CUDA(cudaStreamWaitEvent(c->ctx->stream, c->ev_fin[p], 0));
    *parity = p;
    return G16_OK;
}

// The caller has enqueued whatever consumes d_out[p] on chunk_stream(c): mark the set reusable behind it and, unless
// joins are deferred, make the context stream wait (results are then ordered on the context stream as usual).
cudaStream_t chunk_stream(g16_circuit* c) { return c->ctx->prof.enabled ? c->ctx->stream : c->fin_stream; }
int chunk_end(g16_circuit* c, int p, bool defer) {
    G16_CUDA(cudaEventRecord(c->ev_fin[p], chunk_stream(c)));
    if (!defer) G16_CUDA(cudaStreamWaitEvent(c->ctx->stream, c->ev_fin[p], 0));
    return G16_OK;
}

// Device part of the pipeline: W = B wire vectors (stride wstride, X_* slots filled) in HBM; buffer set p
// (chunk_begin).  The assembled proof points land in d_out[p] on chunk_stream(c).
int prove_device(g16_circuit* c, size_t B, const Fr* W, int p) {
    g16_ctx* ctx = c->ctx;
    cudaStream_t st = ctx->stream;
    Fr* abc = (Fr*)c->d_abc[p].ptr;
    int launches = 0;
    if (!ctx->prof.enabled) {   // side streams start once the wires are in place
        G16_CUDA(cudaEventRecord(c->ev_fork, st));
        for (int i = 0; i < g16_circuit::SIDE_SM; i++) G16_CUDA(cudaStreamWaitEvent(c->side[i], c->ev_fork, 0));
    }
    ctx->prof.mark("chunk", "begin", st);
    dim3 grid(cdiv(c->n, 256), 3, (unsigned)B);
    k_r1cs_spmv<<<grid, 256, 0, st>>>(c->d_rowptr[0], c->d_cid[0], c->d_wid[0], c->d_rowptr[1], c->d_cid[1],
                                      c->d_wid[1], c->d_rowptr[2], c->d_cid[2], c->d_wid[2], c->d_coeffs, W,
                                      c->wstride, abc, c->circ.nb_constraints, (uint32_t)c->n, c->unit_ids);
    launches++;
    ctx->ntt.launches = 0;
    ctx->prof.mark("spmv", "done", st);
    G16_TRY(ctx->ntt.compute_h(abc, c->logn, B, st));
    ctx->prof.mark("quotient", "done", st);
    launches += ctx->ntt.launches;
    G1Affine* rA = c->d_tmp_g1[p];
    G1Affine* rB1 = c->d_tmp_g1[p] + c->max_batch;
    G1Affine* rKZ = c->d_tmp_g1[p] + 2 * c->max_batch;
    G1Affine* rPok = c->d_tmp_g1[p] + 3 * c->max_batch;
    G2Affine* rB2 = c->d_tmp_g2[p];
    G1XYZZ* parts = c->d_parts[p];
    cudaStream_t sFin = chunk_stream(c);
    // per-kernel profiling wants serial, un-overlapped launches
    const bool overlap = !ctx->prof.enabled;
    auto on = [&](int i) { return overlap ? c->side[i] : st; };
    cudaStream_t sA = on(g16_circuit::SIDE_A), sB1 = on(g16_circuit::SIDE_B1), sPok = on(g16_circuit::SIDE_POK),
                 sB2 = on(g16_circuit::SIDE_B2), sSm = on(g16_circuit::SIDE_SM);
    // measured: a shared low-priority accumulate stream serialises the small, latency-bound accumulations (A, PoK)
    // and costs 1.5 %; it stays available for timeline experiments (G16_ACC_STREAM=1)
    const bool low_acc = overlap && getenv("G16_ACC_STREAM") && atoi(getenv("G16_ACC_STREAM")) != 0;
    for (auto& r : c->g1_side) {
        r.prof = &ctx->prof;
        r.acc_stream = low_acc ? c->acc_stream : nullptr;
    }
    c->g2_side.prof = &ctx->prof;
    c->g2_side.acc_stream = low_acc ? c->acc_stream : nullptr;
    c->g1_side[0].label = "A";
    c->g1_side[1].label = "B1";
    c->g1_side[2].label = "PoK";
    c->g2_side.label = "B2";
    c->g1_kz.prof = &ctx->prof;
    c->g1_kz.label = "KZ";
    G16_TRY(c->g2_side.run(c->bB2, W, c->wstride, c->d_mapB, 1, B, rB2, sB2));   // the longest side chain first
    launches += c->g2_side.launches;
    G16_TRY(c->g1_side[0].run(c->bA, W, c->wstride, c->d_mapA, 1, B, rA, sA));
    launches += c->g1_side[0].launches;
    G16_TRY(c->g1_side[1].run(c->bB1, W, c->wstride, c->d_mapB, 1, B, rB1, sB1));
    launches += c->g1_side[1].launches;
    const bool split = ctx->world > 1;   // partial sums until the all-gather: the scalar multiplications wait for it
    if (!split) {
        if (overlap) {
            G16_CUDA(cudaEventRecord(c->ev_join[g16_circuit::SIDE_A], sA));
            G16_CUDA(cudaEventRecord(c->ev_join[g16_circuit::SIDE_B1], sB1));
            G16_CUDA(cudaStreamWaitEvent(sSm, c->ev_join[g16_circuit::SIDE_A], 0));
            G16_CUDA(cudaStreamWaitEvent(sSm, c->ev_join[g16_circuit::SIDE_B1], 0));
        }
        k_scalar_muls<<<cdiv(2 * B, 32), 32, 0, sSm>>>(rA, rB1, W, c->wstride, c->nw, (uint32_t)c->max_batch, (uint32_t)B, parts);
        ctx->prof.mark("scalar_muls", "done", sSm);
        launches++;
    }
    if (c->has_commitment) {
        G16_TRY(c->g1_side[2].run(c->bPok, W, c->wstride, c->d_mapPok, 1, B, rPok, sPok));
        launches += c->g1_side[2].launches;
    } else {
        G16_CUDA(cudaMemsetAsync(rPok, 0, sizeof(G1Affine) * B, sPok));
    }
    cudaStream_t sKZ = on(g16_circuit::SIDE_KZ);
    if (overlap) {   // the quotient is ready: K|Z leaves the (low-priority) context stream too
        G16_CUDA(cudaEventRecord(c->ev_h, st));
        G16_CUDA(cudaStreamWaitEvent(sKZ, c->ev_h, 0));
    }
    c->g1_kz.acc_stream = low_acc ? c->acc_stream : nullptr;
    G16_TRY(c->g1_kz.run(c->bKZ, W, c->wstride, c->d_mapKZ, 1, B, rKZ, sKZ, abc, 3 * c->n));
    launches += c->g1_kz.launches;
    if (overlap) {
        // A and B1 are joined through SIDE_SM (which waited for both) unless the circuit is split across ranks
        const int first = split ? g16_circuit::SIDE_A : g16_circuit::SIDE_POK;
        if (split) {
            G16_CUDA(cudaEventRecord(c->ev_join[g16_circuit::SIDE_A], sA));
            G16_CUDA(cudaEventRecord(c->ev_join[g16_circuit::SIDE_B1], sB1));
        }
        G16_CUDA(cudaEventRecord(c->ev_join[g16_circuit::SIDE_POK], sPok));
        G16_CUDA(cudaEventRecord(c->ev_join[g16_circuit::SIDE_B2], sB2));
        if (!split) G16_CUDA(cudaEventRecord(c->ev_join[g16_circuit::SIDE_SM], sSm));
        G16_CUDA(cudaEventRecord(c->ev_join[g16_circuit::SIDE_KZ], sKZ));
        for (int i = first; i < g16_circuit::N_SIDE; i++)
            if (!(split && i == g16_circuit::SIDE_SM)) G16_CUDA(cudaStreamWaitEvent(sFin, c->ev_join[i], 0));
    }
    if (ctx->world > 1) {   // partial sums of this rank's point ranges -> sums over all ranks (tiny NCCL all-gather)
        G16_TRY(comm_sum_points<Fp>(ctx, c->d_tmp_g1[p], 4 * c->max_batch, sFin));
        G16_TRY(comm_sum_points<Fp2>(ctx, rB2, c->max_batch, sFin));
        k_scalar_muls<<<cdiv(2 * B, 32), 32, 0, sFin>>>(rA, rB1, W, c->wstride, c->nw, (uint32_t)c->max_batch, (uint32_t)B, parts);
        launches += 3;
    }
    k_finalize<<<(unsigned)B, 32, 0, sFin>>>(rA, parts, (uint32_t)c->max_batch, rB2, rKZ, rPok, (ProofPoints*)c->d_out[p].ptr);
    ctx->prof.mark("finalize", "done", sFin);
    launches++;
    G16_CUDA(cudaGetLastError());
    c->last_launches = launches;
    ctx->last_launches = launches;
    return G16_OK;
}

// the X_* slots behind a wire vector: scalars of the alpha/beta/delta bases that ride along in the MSMs
void fill_extra_slots(HFr* x, const HFr& r, const HFr& s) {
    x[X_ONE] = HFr::one();
    x[X_R] = r;
    x[X_S] = s;
    x[X_NEG_RS] = (r * s).neg();
    for (int k = X_NEG_RS + 1; k < X_COUNT; k++) x[k] = HFr::zero();
}
void fill_extras(HFr* w, size_t nw, const HFr& r, const HFr& s) { fill_extra_slots(w + nw, r, s); }

void write_proof_bytes(const ProofPoints& pp, const G1Affine* commitment, uint8_t* out) {
    g1_to_be(pp.ar, out);
    g2_to_be(pp.bs, out + 64);
    g1_to_be(pp.krs, out + 192);
    out[256] = 0; out[257] = 0; out[258] = 0; out[259] = commitment ? 1 : 0;
    size_t off = 260;
    if (commitment) {
        g1_to_be(*commitment, out + off);
        off += 64;
    }
    g1_to_be(pp.pok, out + off);
}

}  // namespace

namespace g16 {
int witness_to_assignment(const Circuit& c, const uint8_t* gz, size_t gz_len, std::vector<uint8_t>* assignment_be);
}

extern "C" {

int g16_prove_assignment(g16_circuit* c, const uint8_t* assignment_be, size_t n_values, const uint8_t rnd[96],
                         uint8_t* proof, size_t* proof_len, uint8_t* pw, size_t* pw_len);

int g16_circuit_load(g16_ctx* ctx, const uint8_t* ccs, size_t ccs_len, const uint8_t* pk, size_t pk_len,
                     const char* acir_json, g16_circuit** out) {
    (void)acir_json;
    if (!ctx || !ccs || !pk || !out) {
        set_error("g16_circuit_load: bad arguments");
        return G16_E_ARG;
    }
    G16_CUDA(cudaSetDevice(ctx->device));
    G16_LOCK(ctx);
    std::unique_ptr<g16_circuit> c(new g16_circuit());
    c->ctx = ctx;
    G16_TRY(parse_ccs(ccs, ccs_len, &c->circ));
    ProvingKeyHost pkh;
    G16_TRY(parse_pk(pk, pk_len, &pkh));
    const Circuit& circ = c->circ;
    c->logn = circ.log_domain();
    c->n = (size_t)1 << c->logn;
    c->nw = circ.nb_wires();
    c->wstride = c->nw + X_COUNT;
    if (pkh.domain != c->n || pkh.infinity_a.size() != c->nw) {
        set_error("g16_circuit_load: proving key does not belong to this constraint system (domain / wire count)");
        return G16_E_ARG;
    }
    if (circ.commitments.size() > 1 || pkh.commitment_keys.size() != circ.commitments.size()) {
        set_error("g16_circuit_load: circuits with more than one BSB22 commitment are not supported yet");
        return G16_E_ARG;
    }
    if (pkh.Z.size() != c->n - 1 && pkh.Z.size() != c->n) {
        set_error("g16_circuit_load: pk.G1.Z has an unexpected length");
        return G16_E_PARSE;
    }
    cudaStream_t st = ctx->stream;
    // ---- scalar maps -------------------------------------------------------------------------
    const uint32_t X0 = (uint32_t)c->nw;
    std::vector<uint32_t> mapA, mapB, mapKZ, mapPok;
    for (uint32_t i = 0; i < c->nw; i++) {
        if (!pkh.infinity_a[i]) mapA.push_back(i);
        if (!pkh.infinity_b[i]) mapB.push_back(i);
    }
    std::vector<uint8_t> skip(c->nw, 0);
    for (auto& info : circ.commitments) {
        skip[info.commitment_index] = 1;
        for (uint32_t w : info.private_committed) skip[w] = 1;
    }
    for (uint32_t i = circ.nb_public; i < c->nw; i++)
        if (!skip[i]) mapKZ.push_back(i);
    if (mapKZ.size() != pkh.K.size()) {
        set_error("g16_circuit_load: pk.G1.K length does not match the private non-committed wires");
        return G16_E_ARG;
    }
    c->nA = mapA.size(); c->nB = mapB.size(); c->nK = pkh.K.size(); c->nZ = c->n - 1;
    // alpha/beta/delta ride along as extra bases
    std::vector<G1Affine> basesA = pkh.A, basesB1 = pkh.B1, basesKZ = pkh.K;
    std::vector<G2Affine> basesB2 = pkh.B2;
    basesA.push_back(pkh.alpha1); mapA.push_back(X0 + X_ONE);
    basesA.push_back(pkh.delta1); mapA.push_back(X0 + X_R);
    basesB1.push_back(pkh.beta1); basesB2.push_back(pkh.beta2); mapB.push_back(X0 + X_ONE);
    basesB1.push_back(pkh.delta1); basesB2.push_back(pkh.delta2); mapB.push_back(X0 + X_S);
    for (size_t j = 0; j < c->nZ; j++) {
        basesKZ.push_back(pkh.Z[j]);
        mapKZ.push_back(0x80000000u | (uint32_t)j);   // h lives in the first vector of the abc triple
    }
    basesKZ.push_back(pkh.delta1); mapKZ.push_back(X0 + X_NEG_RS);
    c->has_commitment = !circ.commitments.empty();
    if (c->has_commitment) {
        c->committed_wires = circ.commitments[0].private_committed;
        c->n_committed = c->committed_wires.size();
        if (pkh.commitment_keys[0].basis.size() != c->n_committed) {
            set_error("g16_circuit_load: commitment key size does not match PrivateCommitted");
            return G16_E_ARG;
        }
        mapPok = c->committed_wires;
    }
    c->n_committed_total = c->n_committed;
    // ---- single large proof across GPUs: this rank keeps one contiguous index range
    // of every MSM point set; partial sums are all-gathered before the final assembly (comm.cu)
    c->world = ctx->world;
    if (ctx->world > 1) {
        auto cut = [&](auto& pts, std::vector<uint32_t>* map) -> bool {
            size_t lo, hi;
            shard_range(pts.size(), ctx->rank, ctx->world, &lo, &hi);
            if (hi <= lo) return false;
            if (map) *map = std::vector<uint32_t>(map->begin() + lo, map->begin() + hi);
            pts = std::decay_t<decltype(pts)>(pts.begin() + lo, pts.begin() + hi);
            return true;
        };
        bool ok = cut(basesA, &mapA) && cut(basesB1, nullptr) && cut(basesB2, &mapB) && cut(basesKZ, &mapKZ);
        if (ok && c->has_commitment) {
            auto& key = pkh.commitment_keys[0];
            ok = cut(key.basis, &c->committed_wires) && cut(key.basis_exp_sigma, &mapPok);
            c->n_committed = c->committed_wires.size();
        }
        if (!ok) {
            set_error("g16_circuit_load: an MSM point set is smaller than the number of ranks");
            return G16_E_ARG;
        }
    }
    // ---- batch size: bounded by scratch memory (entries dominate: 4 B per base, window and proof)
    size_t max_batch = 64;
    if (c->n >= (1u << 18)) max_batch = 1;
    else if (c->n >= (1u << 16)) max_batch = 8;
    c->max_batch = max_batch;
    // the solver is latency-bound (one CTA per proof walks ~10^3 levels): give it 4 proving batches at a
    // time so its latency hides behind the proving of the previous group
    // Stage A solves a GROUP of chunks at once (the solver is latency-bound: 512 proofs take little longer than 64) while
    // the previous group is proved.  8 chunks per group keep its latency under load (25-55 ms) far from the group's
    // proving time (140-210 ms); the first group of a call is 4 chunks, since nothing can be proved before it is solved.
    const size_t group_chunks = getenv("G16_SOLVE_GROUP") ? std::max(1l, atol(getenv("G16_SOLVE_GROUP"))) : 8;
    const size_t first_chunks = getenv("G16_SOLVE_FIRST") ? std::max(1l, atol(getenv("G16_SOLVE_FIRST"))) : 4;
    const size_t solve_batch = max_batch >= 8 ? group_chunks * max_batch : max_batch;
    c->solve_batch = solve_batch;
    c->first_group = max_batch >= 8 ? std::min(first_chunks, group_chunks) * max_batch : max_batch;
    // ---- bases -------------------------------------------------------------------------------
    auto win = [&](size_t npts) { return msm_pick_window(npts, max_batch); };
    G16_TRY(c->bA.load(basesA.data(), basesA.size(), win(basesA.size()), 1, st));
    G16_TRY(c->bB1.load(basesB1.data(), basesB1.size(), win(basesB1.size()), 1, st));
    G16_TRY(c->bB2.load(basesB2.data(), basesB2.size(), win(basesB2.size()), 1, st));
    G16_TRY(c->bKZ.load(basesKZ.data(), basesKZ.size(), win(basesKZ.size()), 1, st));
    if (c->has_commitment) {
        auto& key = pkh.commitment_keys[0];
        G16_TRY(c->bCommit.load(key.basis.data(), key.basis.size(), win(key.basis.size()), 1, st));
        G16_TRY(c->bPok.load(key.basis_exp_sigma.data(), key.basis_exp_sigma.size(), win(key.basis.size()), 1, st));
    }
    G16_TRY(upload_vec(mapA, &c->d_mapA, st));
    G16_TRY(upload_vec(mapB, &c->d_mapB, st));
    G16_TRY(upload_vec(mapKZ, &c->d_mapKZ, st));
    G16_TRY(upload_vec(mapPok, &c->d_mapPok, st));
    // ---- R1CS ----------------------------------------------------------------------------------
    const Circuit::Csr* M[3] = {&circ.A, &circ.B, &circ.C};
    for (int m = 0; m < 3; m++) {
        G16_TRY(upload_vec(M[m]->rowptr, &c->d_rowptr[m], st));
        G16_TRY(upload_vec(M[m]->coeff, &c->d_cid[m], st));
        G16_TRY(upload_vec(M[m]->wire, &c->d_wid[m], st));
    }
    static_assert(sizeof(HFr) == sizeof(Fr), "host and device Fr must share a layout");
    c->unit_ids = circ.coeffs.size() >= 4 && circ.coeffs[0].is_zero() && circ.coeffs[1] == HFr::one() &&
                  circ.coeffs[3] == HFr::one().neg();
    G16_CUDA(cudaMalloc(&c->d_coeffs, sizeof(Fr) * circ.coeffs.size()));
    G16_CUDA(cudaMemcpyAsync(c->d_coeffs, circ.coeffs.data(), sizeof(Fr) * circ.coeffs.size(), cudaMemcpyHostToDevice, st));
    // ---- scratch -------------------------------------------------------------------------------
    for (int p = 0; p < 2; p++) {
        G16_TRY(c->d_abc[p].ensure(sizeof(Fr) * 3 * c->n * max_batch));
        G16_TRY(c->d_out[p].ensure(sizeof(ProofPoints) * max_batch));
        G16_TRY(c->d_wdev[p].ensure(sizeof(Fr) * c->wstride * max_batch));
        G16_CUDA(cudaEventCreateWithFlags(&c->ev_fin[p], cudaEventDisableTiming));
    }
    {
        int lo = 0, hi = 0;
        G16_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        G16_CUDA(cudaStreamCreateWithPriority(&c->aux_stream, cudaStreamNonBlocking, hi));
    }
    {
        int lo = 0, hi = 0;
        G16_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        G16_CUDA(cudaStreamCreateWithPriority(&c->acc_stream, cudaStreamNonBlocking, lo));
    }
    for (int i = 0; i < g16_circuit::N_SIDE; i++) {
        int lo = 0, hi = 0;
        G16_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));   // numerically lower = more urgent
        G16_CUDA(cudaStreamCreateWithPriority(&c->side[i], cudaStreamNonBlocking, hi < lo ? hi + 1 : hi));
        G16_CUDA(cudaEventCreateWithFlags(&c->ev_join[i], cudaEventDisableTiming));
    }
    G16_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    G16_CUDA(cudaEventCreateWithFlags(&c->ev_h, cudaEventDisableTiming));
    {
        int lo = 0, hi = 0;
        G16_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        G16_CUDA(cudaStreamCreateWithPriority(&c->fin_stream, cudaStreamNonBlocking, hi < lo ? hi + 1 : hi));
    }
    for (int i = 0; i < 2; i++) {
        G16_CUDA(cudaMallocHost((void**)&c->h_pts[i], sizeof(ProofPoints) * max_batch));
        G16_CUDA(cudaEventCreateWithFlags(&c->ev_done[i], cudaEventDisableTiming));
    }
    for (auto& sl : c->slots) {
        G16_TRY(sl.d_wires.ensure(sizeof(Fr) * c->wstride * solve_batch));
        G16_TRY(sl.d_commit_vals.ensure(sizeof(Fr) * (c->n_committed ? c->n_committed : 1) * solve_batch));
        G16_TRY(sl.d_commit_out.ensure(sizeof(G1Affine) * solve_batch));
        G16_CUDA(cudaMallocHost(&sl.h_wires, sizeof(Fr) * c->wstride * solve_batch));
        G16_CUDA(cudaEventCreateWithFlags(&sl.ready, cudaEventDisableTiming));
        const size_t nin = circ.nb_public - 1 + circ.nb_secret;
        G16_TRY(sl.d_asg_be.ensure(32 * nin * solve_batch));
        G16_TRY(sl.d_rnd_be.ensure(96 * solve_batch));
        G16_TRY(sl.d_err.ensure(4 * solve_batch));
        G16_TRY(sl.d_chal.ensure(sizeof(Fr) * solve_batch));
        G16_CUDA(cudaMallocHost(&sl.h_stage, (32 * nin + 96 + sizeof(Fr) + 4) * solve_batch));
    }
    G16_TRY(upload_vec(c->committed_wires, &c->d_map_commit, st));
    if (!getenv("G16_HOST_SOLVER")) {
        G16_TRY(c->plan.build(circ, st, &c->host_solver_reason));
    } else {
        c->host_solver_reason = "G16_HOST_SOLVER is set";
    }
    for (int p = 0; p < 2; p++) {
        G16_CUDA(cudaMalloc(&c->d_tmp_g1[p], sizeof(G1Affine) * 4 * max_batch));
        G16_CUDA(cudaMalloc(&c->d_tmp_g2[p], sizeof(G2Affine) * max_batch));
        G16_CUDA(cudaMalloc(&c->d_parts[p], sizeof(G1XYZZ) * 2 * max_batch));
        G16_CUDA(cudaMemsetAsync(c->d_tmp_g1[p], 0, sizeof(G1Affine) * 4 * max_batch, st));   // unused slots = infinity
        G16_CUDA(cudaMemsetAsync(c->d_tmp_g2[p], 0, sizeof(G2Affine) * max_batch, st));
    }
    const NttDomain* dom;
    G16_TRY(ctx->ntt.domain(c->logn, st, &dom));
    G16_CUDA(cudaStreamSynchronize(st));   // key tables and maps are in place
    if (c->plan.valid) {
        // size what stage A grows on first use (a cudaFree in the middle of a pipelined call drains the device):
        // the commitment MSM's scratch for a full group (zero wires: no bucket entries) and the hint outputs
        g16_circuit::Slot& sl = c->slots[0];
        if (c->plan.commit_level != (uint32_t)-1 && c->n_committed) {
            G16_CUDA(cudaMemsetAsync(sl.d_wires.ptr, 0, sizeof(Fr) * c->wstride * solve_batch, c->aux_stream));
            G16_TRY(c->g1_aux.run(c->bCommit, (const Fr*)sl.d_wires.ptr, c->wstride, c->d_map_commit, 1, solve_batch,
                                  (G1Affine*)sl.d_commit_out.ptr, c->aux_stream));
        }
        for (auto& s2 : c->slots)
            if (!c->plan.host_wires.empty()) G16_TRY(s2.d_pre.ensure(sizeof(Fr) * c->plan.host_wires.size() * solve_batch));
        G16_CUDA(cudaStreamSynchronize(c->aux_stream));
    }
    ctx->circuits.push_back(c.get());
    *out = c.release();
    return G16_OK;
}

int g16_set_deferred_join(g16_ctx* ctx, int on) {
    if (!ctx) return G16_E_ARG;
    G16_LOCK(ctx);
    ctx->deferred_join = on != 0;
    return G16_OK;
}

int g16_join(g16_ctx* ctx) {
    if (!ctx) return G16_E_ARG;
    G16_CUDA(cudaSetDevice(ctx->device));
    G16_LOCK(ctx);
    for (g16_circuit* c : ctx->circuits)
        for (int p = 0; p < 2; p++) G16_CUDA(cudaStreamWaitEvent(ctx->stream, c->ev_fin[p], 0));
    return G16_OK;
}

void g16_circuit_free(g16_circuit* c) {
    if (!c) return;
    g16_ctx* ctx = c->ctx;
    if (ctx) {
        G16_LOCK(ctx);
        delete c;
    } else {
        delete c;
    }
}

const char* g16_circuit_solver(const g16_circuit* c) {
    if (!c) return "";
    return c->plan.valid ? "gpu" : c->host_solver_reason.c_str();
}

int g16_circuit_info(const g16_circuit* c, uint64_t what[16]) {
    if (!c || !what) return G16_E_ARG;
    memset(what, 0, 16 * sizeof(uint64_t));
    what[0] = c->circ.nb_constraints;
    what[1] = c->nw;
    what[2] = c->circ.nb_public;
    what[3] = c->circ.nb_secret;
    what[4] = c->n;
    what[5] = c->circ.commitments.size();
    what[6] = c->nA; what[7] = c->nB; what[8] = c->nK; what[9] = c->nZ; what[10] = c->n_committed_total;
    what[11] = c->max_batch;
    what[12] = c->bA.cfg.c; what[13] = c->bB1.cfg.c; what[14] = c->bKZ.cfg.c; what[15] = c->bB2.cfg.c;
    return G16_OK;
}

int g16_prove_wires_dev(g16_circuit* c, size_t n, const void* d_wires, const uint8_t* rnd, void* d_proof_points) {
    if (!c || !d_wires || !d_proof_points || n == 0 || n > c->max_batch) {
        set_error("g16_prove_wires_dev: bad arguments (n must be 1..max_batch)");
        return G16_E_ARG;
    }
    G16_CUDA(cudaSetDevice(c->ctx->device));
    G16_LOCK(c->ctx);
    cudaStream_t st = c->ctx->stream;
    // blinding scalars: injected (tests, benchmarks) or drawn from the OS CSPRNG -- never silently zero
    std::vector<HFr> extras(X_COUNT * n);
    for (size_t b = 0; b < n; b++) {
        HFr r, s;
        if (rnd) {
            r = HFr::from_be(rnd + 96 * b);
            s = HFr::from_be(rnd + 96 * b + 32);
        } else {
            G16_TRY(random_fr(&r));
            G16_TRY(random_fr(&s));
        }
        fill_extra_slots(extras.data() + X_COUNT * b, r, s);
    }
    // wires arrive without the X_* slots: lay them out with stride wstride, then drop the extras in
    int p;
    G16_TRY(chunk_begin(c, &p));
    Fr* W = (Fr*)c->d_wdev[p].ptr;
    G16_CUDA(cudaMemcpy2DAsync(W, sizeof(Fr) * c->wstride, d_wires, sizeof(Fr) * c->nw, sizeof(Fr) * c->nw, n,
                               cudaMemcpyDeviceToDevice, st));
    // pageable source: the runtime stages it before returning, so `extras` may go out of scope
    G16_CUDA(cudaMemcpy2DAsync(W + c->nw, sizeof(Fr) * c->wstride, extras.data(), sizeof(Fr) * X_COUNT,
                               sizeof(Fr) * X_COUNT, n, cudaMemcpyHostToDevice, st));
    G16_TRY(prove_device(c, n, W, p));
    G16_CUDA(cudaMemcpyAsync(d_proof_points, c->d_out[p].ptr, sizeof(ProofPoints) * n, cudaMemcpyDeviceToDevice, chunk_stream(c)));
    return chunk_end(c, p, c->ctx->deferred_join);
}

// Shared host path: `wires` holds n full wire vectors (Montgomery) with stride wstride and the X_*
// slots filled; `commitments` (may be null) the commitment points in canonical form.
static int prove_from_host_wires(g16_circuit* c, size_t n, const HFr* wires, const G1Affine* commitments,
                                 uint8_t* proofs) {
    cudaStream_t st = c->ctx->stream;
    size_t bytes = sizeof(Fr) * c->wstride * n;
    g16_circuit::Slot& sl = c->slots[0];
    if ((const void*)wires != sl.h_wires) memcpy(sl.h_wires, wires, bytes);
    int p;
    G16_TRY(chunk_begin(c, &p));
    G16_CUDA(cudaStreamSynchronize(c->fin_stream));   // synchronous path: the slot is reused right away
    G16_CUDA(cudaMemcpyAsync(sl.d_wires.ptr, sl.h_wires, bytes, cudaMemcpyHostToDevice, st));
    G16_TRY(prove_device(c, n, (const Fr*)sl.d_wires.ptr, p));
    std::vector<ProofPoints> pts(n);
    G16_CUDA(cudaMemcpyAsync(pts.data(), c->d_out[p].ptr, sizeof(ProofPoints) * n, cudaMemcpyDeviceToHost, chunk_stream(c)));
    G16_TRY(chunk_end(c, p, false));
    G16_CUDA(cudaStreamSynchronize(st));
    const size_t plen = c->has_commitment ? 388 : 324;
    for (size_t b = 0; b < n; b++)
        write_proof_bytes(pts[b], commitments ? &commitments[b] : nullptr, proofs + plen * b);
    return G16_OK;
}

// Synchronous variant: host conversion, one chunk at a time on the context stream.  Used when the circuit is
// split across ranks (every call is then a collective and must enqueue its NCCL work in program order).
static int prove_wires_sync(g16_circuit* c, size_t n, const uint8_t* wires_be, const uint8_t* rnd, uint8_t* proofs) {
    const size_t plen = c->has_commitment ? 388 : 324;
    for (size_t done = 0; done < n;) {
        size_t B = std::min(c->max_batch, n - done);
        std::vector<HFr> w(c->wstride * B);
        std::vector<G1Affine> commits(B, G1Affine::inf());
        std::vector<int> rcs(B, G16_OK);
        trace("wires: convert begin", (long)done);
        parallel_for(B, [&](size_t b) {
            const uint8_t* src = wires_be + (done + b) * c->nw * 32;
            HFr* dst = w.data() + b * c->wstride;
            for (size_t i = 0; i < c->nw; i++) dst[i] = HFr::from_be(src + 32 * i);
            HFr r, s;
            if (rnd) {
                r = HFr::from_be(rnd + 96 * (done + b));
                s = HFr::from_be(rnd + 96 * (done + b) + 32);
            } else if (random_fr(&r) != G16_OK || random_fr(&s) != G16_OK) {
                rcs[b] = G16_E_INTERNAL;   // never fall back to r = s = 0 (a non-zero-knowledge proof)
                return;
            }
            fill_extras(dst, c->nw, r, s);
        });
        trace("wires: converted");
        for (int rc : rcs)
            if (rc != G16_OK) {
                set_error("could not read /dev/urandom");
                return rc;
            }
        if (c->has_commitment) {
            // commitment point = MSM of the committed wire values over the commitment basis
            std::vector<HFr> cv(c->n_committed * B);
            for (size_t b = 0; b < B; b++)
                for (size_t k = 0; k < c->n_committed; k++) cv[b * c->n_committed + k] = w[b * c->wstride + c->committed_wires[k]];
            cudaStream_t st = c->ctx->stream;
            g16_circuit::Slot& sl = c->slots[0];
            G16_CUDA(cudaMemcpyAsync(sl.d_commit_vals.ptr, cv.data(), sizeof(Fr) * cv.size(), cudaMemcpyHostToDevice, st));
            G16_TRY(c->ctx->g1.run(c->bCommit, (const Fr*)sl.d_commit_vals.ptr, c->n_committed, nullptr, 1, B,
                                   (G1Affine*)sl.d_commit_out.ptr, st));
            G16_TRY(comm_sum_points<Fp>(c->ctx, (G1Affine*)sl.d_commit_out.ptr, B, st));
            size_t nfp = B * 2;
            k_fp_from_mont<<<cdiv(nfp, 256), 256, 0, st>>>((Fp*)sl.d_commit_out.ptr, nfp);
            G16_CUDA(cudaMemcpyAsync(commits.data(), sl.d_commit_out.ptr, sizeof(G1Affine) * B, cudaMemcpyDeviceToHost, st));
            G16_CUDA(cudaStreamSynchronize(st));
        }
        trace("wires: commitments done");
        G16_TRY(prove_from_host_wires(c, B, w.data(), c->has_commitment ? commits.data() : nullptr, proofs + plen * done));
        trace("wires: proved");
        done += B;
    }
    return G16_OK;
}

// ---- stage A: host solve of one chunk into a pipeline slot (runs on a worker thread) ---------------
// On success the slot holds: h_wires (Montgomery, X_* slots filled), commits, and an H2D copy of the
// wires into d_wires has been enqueued on aux_stream with `ready` recorded behind it.
struct StageResult {
    int rc = G16_OK;
    std::string err;
};

static StageResult stage_solve(g16_circuit* c, int slot_id, size_t B, const uint8_t* assignments_be, const uint8_t* rnd,
                               size_t first_index, bool upload) {
    StageResult res;
    auto failm = [&](int rc, const std::string& m) {
        res.rc = rc;
        res.err = m;
        return res;
    };
    auto cuda_fail = [&](cudaError_t e, const char* what) {
        return failm(G16_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
    };
    cudaError_t ce = cudaSetDevice(c->ctx->device);
    if (ce != cudaSuccess) return cuda_fail(ce, "cudaSetDevice");
    const Circuit& circ = c->circ;
    const size_t nin = circ.nb_public - 1 + circ.nb_secret;
    g16_circuit::Slot& sl = c->slots[slot_id];
    cudaStream_t st = c->aux_stream;
    HFr* W = (HFr*)sl.h_wires;
    std::vector<SolveState> states(B);
    std::vector<HFr> rs(3 * B);
    std::vector<int> rcs(B, G16_OK);
    for (size_t b = 0; b < B; b++)
        for (int k = 0; k < 3; k++) {
            if (rnd) rs[3 * b + k] = HFr::from_be(rnd + 96 * b + 32 * k);
            else if (random_fr(&rs[3 * b + k]) != G16_OK) return failm(G16_E_INTERNAL, "could not read /dev/urandom");
        }
    parallel_for(B, [&](size_t b) {
        std::vector<HFr> asg(nin);
        const uint8_t* src = assignments_be + b * nin * 32;
        for (size_t i = 0; i < nin; i++) asg[i] = HFr::from_be(src + 32 * i);
        solve_begin(circ, asg.data(), &states[b]);
        rcs[b] = solve_run(circ, &states[b], &rs[3 * b + 2]);
    });
    sl.commits.assign(B, G1Affine::inf());
    bool any_commit = false;
    for (size_t b = 0; b < B; b++) {
        if (rcs[b] == SOLVE_NEED_COMMITMENT) any_commit = true;
        else if (rcs[b] != SOLVE_DONE) return failm(rcs[b], "proof " + std::to_string(first_index + b) + ": " + states[b].error);
    }
    if (any_commit) {
        std::vector<HFr> cv(c->n_committed * B, HFr::zero());
        for (size_t b = 0; b < B; b++) {
            if (rcs[b] != SOLVE_NEED_COMMITMENT || states[b].committed.size() != c->n_committed)
                return failm(G16_E_INTERNAL, "proof " + std::to_string(first_index + b) + ": inconsistent commitment hint");
            memcpy(&cv[b * c->n_committed], states[b].committed.data(), sizeof(HFr) * c->n_committed);
        }
        if ((ce = cudaMemcpyAsync(sl.d_commit_vals.ptr, cv.data(), sizeof(Fr) * cv.size(), cudaMemcpyHostToDevice, st)) != cudaSuccess)
            return cuda_fail(ce, "commit H2D");
        int rc = c->g1_aux.run(c->bCommit, (const Fr*)sl.d_commit_vals.ptr, c->n_committed, nullptr, 1, B,
                               (G1Affine*)sl.d_commit_out.ptr, st);
        if (rc != G16_OK) return failm(rc, get_error());
        k_fp_from_mont<<<cdiv(B * 2, 256), 256, 0, st>>>((Fp*)sl.d_commit_out.ptr, B * 2);
        if ((ce = cudaMemcpyAsync(sl.commits.data(), sl.d_commit_out.ptr, sizeof(G1Affine) * B, cudaMemcpyDeviceToHost, st)) != cudaSuccess)
            return cuda_fail(ce, "commit D2H");
        if ((ce = cudaStreamSynchronize(st)) != cudaSuccess) return cuda_fail(ce, "commitment MSM");
        parallel_for(B, [&](size_t b) {
            std::vector<uint8_t> msg(64 + 32 * states[b].hashed.size());
            g1_to_be(sl.commits[b], msg.data());
            for (size_t k = 0; k < states[b].hashed.size(); k++) states[b].hashed[k].to_be(msg.data() + 64 + 32 * k);
            solve_provide_challenge(&states[b], hash_to_fr(msg.data(), msg.size(), "bsb22-commitment"));
            rcs[b] = solve_run(circ, &states[b], &rs[3 * b + 2]);
        });
        for (size_t b = 0; b < B; b++)
            if (rcs[b] != SOLVE_DONE)
                return failm(rcs[b] == SOLVE_NEED_COMMITMENT ? G16_E_HINT : rcs[b],
                             "proof " + std::to_string(first_index + b) + ": " +
                                 (rcs[b] == SOLVE_NEED_COMMITMENT ? std::string("more than one commitment") : states[b].error));
    }
    parallel_for(B, [&](size_t b) {
        memcpy(&W[b * c->wstride], states[b].w.data(), sizeof(HFr) * c->nw);
        fill_extras(&W[b * c->wstride], c->nw, rs[3 * b], rs[3 * b + 1]);
    });
    if (upload) {
        if ((ce = cudaMemcpyAsync(sl.d_wires.ptr, sl.h_wires, sizeof(Fr) * c->wstride * B, cudaMemcpyHostToDevice, st)) != cudaSuccess)
            return cuda_fail(ce, "wires H2D");
        if ((ce = cudaEventRecord(sl.ready, st)) != cudaSuccess) return cuda_fail(ce, "event record");
    }
    return res;
}

// ---- stage A on the GPU: assignments -> full wire vectors in slot.d_wires (gpusolver.cu) ------------
static StageResult stage_solve_gpu(g16_circuit* c, int slot_id, size_t B, const uint8_t* assignments_be,
                                   const uint8_t* rnd, size_t first_index, bool tolerate = false) {
    StageResult res;
    auto failm = [&](int rc, const std::string& m) {
        res.rc = rc;
        res.err = m;
        return res;
    };
#define G16_STAGE_CUDA(expr)                                                                  \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) return failm(G16_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)
    G16_STAGE_CUDA(cudaSetDevice(c->ctx->device));
    const Circuit& circ = c->circ;
    const size_t nin = circ.nb_public - 1 + circ.nb_secret;
    g16_circuit::Slot& sl = c->slots[slot_id];
    cudaStream_t st = c->aux_stream;
    uint8_t* h_asg = (uint8_t*)sl.h_stage;
    uint8_t* h_rnd = h_asg + 32 * nin * c->solve_batch;
    HFr* h_chal = (HFr*)(h_rnd + 96 * c->solve_batch);
    uint32_t* h_err = (uint32_t*)((uint8_t*)h_chal + sizeof(HFr) * c->solve_batch);
    trace("solve_gpu: begin", (long)first_index);
    memcpy(h_asg, assignments_be, 32 * nin * B);
    if (rnd) memcpy(h_rnd, rnd, 96 * B);
    else
        for (size_t k = 0; k < 3 * B; k++) {
            HFr v;
            if (random_fr(&v) != G16_OK) return failm(G16_E_INTERNAL, "could not read /dev/urandom");
            v.to_be(h_rnd + 32 * k);
        }
    trace("solve_gpu: staged");
    // G16_TRACE_SYNC=1: wait for the stream after every step so that the trace shows where the device time goes
    static const bool trace_sync = getenv("G16_TRACE_SYNC") && atoi(getenv("G16_TRACE_SYNC")) != 0;
    auto tsync = [&](const char* tag) {
        if (!trace_sync) return;
        cudaStreamSynchronize(st);
        trace(tag);
    };
    Fr* W = (Fr*)sl.d_wires.ptr;
    uint32_t* d_err = (uint32_t*)sl.d_err.ptr;
    G16_STAGE_CUDA(cudaMemcpyAsync(sl.d_asg_be.ptr, h_asg, 32 * nin * B, cudaMemcpyHostToDevice, st));
    G16_STAGE_CUDA(cudaMemcpyAsync(sl.d_rnd_be.ptr, h_rnd, 96 * B, cudaMemcpyHostToDevice, st));
    G16_STAGE_CUDA(cudaMemsetAsync(d_err, 0xff, 4 * B, st));
    tsync("  sync: inputs copied");
    int rc = c->plan.assign((const uint8_t*)sl.d_asg_be.ptr, (const uint8_t*)sl.d_rnd_be.ptr, (uint32_t)nin, W, c->wstride, c->nw, B, st);
    if (rc != G16_OK) return failm(rc, get_error());
    tsync("  sync: k_assign");
    if (!c->plan.host_hints.empty()) {
        // integer hints that hang off the inputs (withdraw circuit: the Grumpkin scalar split and its emulated product
        // check): evaluated on the host per proof, scattered into the device wire vectors before level 0
        const size_t npre = c->plan.host_wires.size();
        std::vector<HFr> pre(npre * B);
        std::vector<int> rcs(B, G16_OK);
        std::vector<std::string> errs(B);
        const size_t workers = std::min<size_t>(default_threads(), B);
        trace("solve_gpu: inputs enqueued");
        parallel_for(workers, [&](size_t t) {
            SolveState stt;
            std::vector<HFr> asg(nin, HFr::zero());
            solve_begin(circ, asg.data(), &stt);
            for (size_t b = t; b < B; b += workers) {
                // only the inputs these hints read are converted; their outputs are forgotten between proofs
                const uint8_t* src = assignments_be + b * nin * 32;
                for (uint32_t wv : c->plan.host_inputs) stt.w[wv] = HFr::from_be(src + 32 * (size_t)(wv - 1));
                for (uint32_t wv : c->plan.host_wires) stt.known[wv] = 0;
                stt.error.clear();
                stt.tolerate = tolerate;
                for (uint32_t ins : c->plan.host_hints)
                    if ((rcs[b] = solve_run_hint(circ, &stt, ins)) != G16_OK) {
                        errs[b] = stt.error;
                        break;
                    }
                for (size_t k = 0; k < npre; k++) pre[b * npre + k] = stt.w[c->plan.host_wires[k]];
            }
        });
        for (size_t b = 0; b < B; b++)
            if (rcs[b] != G16_OK && !tolerate) return failm(rcs[b], "proof " + std::to_string(first_index + b) + ": " + errs[b]);
        trace("solve_gpu: host hints evaluated");
        if (sl.d_pre.ensure(sizeof(Fr) * npre * B) != G16_OK) return failm(G16_E_CUDA, get_error());
        G16_STAGE_CUDA(cudaMemcpyAsync(sl.d_pre.ptr, pre.data(), sizeof(Fr) * npre * B, cudaMemcpyHostToDevice, st));
        G16_STAGE_CUDA(cudaStreamSynchronize(st));   // `pre` is pageable and dies with this scope
        trace("solve_gpu: hint outputs on the device");
        if ((rc = c->plan.scatter_host_wires(W, c->wstride, (const Fr*)sl.d_pre.ptr, B, st)) != G16_OK) return failm(rc, get_error());
    }
    const bool commit = c->plan.commit_level != (uint32_t)-1;
    const uint32_t split = commit ? c->plan.commit_level + 1 : c->plan.nlevels;
    // All proofs of the group are solved at once.  Solving them `sub` at a time (G16_SOLVE_SUB; fewer resident solver
    // CTAs taking registers from the accumulate kernels next to them) was measured: the latency-bound solver then
    // runs four times as long beside the prover and becomes the bottleneck of the pipeline (19.1 -> 22.2 ms per audit
    // chunk, 29.5 -> 43.2 per withdraw chunk at sub = 64).
    static const size_t sub_env = getenv("G16_SOLVE_SUB") ? (size_t)atol(getenv("G16_SOLVE_SUB")) : 0;
    const size_t sub = sub_env ? sub_env : B;
    auto run_levels = [&](uint32_t lv0, uint32_t lv1) {
        for (size_t off = 0; off < B; off += sub) {
            int r = c->plan.run(c->d_coeffs, c->unit_ids, W + off * c->wstride, c->wstride, c->nw, std::min(sub, B - off), lv0, lv1,
                                d_err + off, st);
            if (r != G16_OK) return r;
        }
        return (int)G16_OK;
    };
    if ((rc = run_levels(0, split)) != G16_OK) return failm(rc, get_error());
    tsync("  sync: levels before the commitment");
    sl.commits.assign(B, G1Affine::inf());
    if (commit) {
        rc = c->g1_aux.run(c->bCommit, W, c->wstride, c->d_map_commit, 1, B, (G1Affine*)sl.d_commit_out.ptr, st);
        if (rc != G16_OK) return failm(rc, get_error());
        k_fp_from_mont<<<cdiv(B * 2, 256), 256, 0, st>>>((Fp*)sl.d_commit_out.ptr, B * 2);
        tsync("  sync: commitment MSM");
        trace("solve_gpu: phase1 enqueued");
        G16_STAGE_CUDA(cudaMemcpyAsync(sl.commits.data(), sl.d_commit_out.ptr, sizeof(G1Affine) * B, cudaMemcpyDeviceToHost, st));
        G16_STAGE_CUDA(cudaStreamSynchronize(st));
        trace("solve_gpu: phase1 done");
        const auto& hashed = c->plan.commit_hashed;
        std::vector<uint8_t> msg(64 + 32 * hashed.size());
        for (size_t b = 0; b < B; b++) {
            g1_to_be(sl.commits[b], msg.data());
            for (size_t h = 0; h < hashed.size(); h++) {   // committed public wires, from the caller's assignment
                HFr v = HFr::zero();
                for (const auto& term : hashed[h]) {
                    const HFr x = term.second == CCS_CONST_WIRE || term.second == 0
                                      ? HFr::one()
                                      : HFr::from_be(assignments_be + (b * nin + term.second - 1) * 32);
                    v = v + circ.coeffs[term.first] * x;
                }
                v.to_be(msg.data() + 64 + 32 * h);
            }
            h_chal[b] = hash_to_fr(msg.data(), msg.size(), "bsb22-commitment");
        }
        trace("solve_gpu: challenges hashed");
        G16_STAGE_CUDA(cudaMemcpyAsync(sl.d_chal.ptr, h_chal, sizeof(Fr) * B, cudaMemcpyHostToDevice, st));
        if ((rc = c->plan.set_wire(W, c->wstride, c->plan.commit_wire, (const Fr*)sl.d_chal.ptr, B, st)) != G16_OK) return failm(rc, get_error());
        if ((rc = run_levels(split, c->plan.nlevels)) != G16_OK) return failm(rc, get_error());
    }
    G16_STAGE_CUDA(cudaMemcpyAsync(h_err, d_err, 4 * B, cudaMemcpyDeviceToHost, st));
    G16_STAGE_CUDA(cudaEventRecord(sl.ready, st));
    G16_STAGE_CUDA(cudaStreamSynchronize(st));
    trace("solve_gpu: done");
    for (size_t b = 0; b < B && !tolerate; b++)
        if (h_err[b] != 0xffffffffu) {
            if (h_err[b] & 0x80000000u) return failm(G16_E_HINT, "proof " + std::to_string(first_index + b) + ": solver hint failed on the device");
            return failm(G16_E_UNSAT, "proof " + std::to_string(first_index + b) + ": constraint #" + std::to_string(h_err[b] - 1) + " is not satisfied");
        }
#undef G16_STAGE_CUDA
    return res;
}

// ---- stage A for callers that already hold full wire vectors: big-endian wires -> slot.d_wires ----------
static StageResult stage_wires(g16_circuit* c, int slot_id, size_t G, const uint8_t* wires_be, const uint8_t* rnd,
                               size_t first_index) {
    StageResult res;
    auto failm = [&](int rc, const std::string& m) {
        res.rc = rc;
        res.err = m;
        return res;
    };
#define G16_STAGE_CUDA(expr)                                                                  \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) return failm(G16_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)
    G16_STAGE_CUDA(cudaSetDevice(c->ctx->device));
    const size_t nin = c->circ.nb_public - 1 + c->circ.nb_secret;
    g16_circuit::Slot& sl = c->slots[slot_id];
    cudaStream_t st = c->aux_stream;
    trace("wires: stage begin", (long)first_index);
    const size_t bytes = 32 * c->nw * G;
    if (sl.d_wires_be.ensure(bytes) != G16_OK) return failm(G16_E_CUDA, get_error());
    // pageable caller memory -> pinned staging (threads), then one asynchronous copy
    uint8_t* h_be = (uint8_t*)sl.h_wires;   // pinned, wstride * solve_batch * 32 bytes >= bytes
    const size_t pieces = std::min<size_t>(G, 64), per = (bytes + pieces - 1) / pieces;
    parallel_for(pieces, [&](size_t k) {
        const size_t lo = k * per, hi = std::min(bytes, lo + per);
        if (lo < hi) memcpy(h_be + lo, wires_be + lo, hi - lo);
    });
    uint8_t* h_rnd = (uint8_t*)sl.h_stage + 32 * nin * c->solve_batch;
    if (rnd) {
        memcpy(h_rnd, rnd, 96 * G);
    } else {
        for (size_t k = 0; k < 3 * G; k++) {   // never fall back to r = s = 0 (a non-zero-knowledge proof)
            HFr v;
            if (random_fr(&v) != G16_OK) return failm(G16_E_INTERNAL, "could not read /dev/urandom");
            v.to_be(h_rnd + 32 * k);
        }
    }
    trace("wires: staged");
    G16_STAGE_CUDA(cudaMemcpyAsync(sl.d_wires_be.ptr, h_be, bytes, cudaMemcpyHostToDevice, st));
    G16_STAGE_CUDA(cudaMemcpyAsync(sl.d_rnd_be.ptr, h_rnd, 96 * G, cudaMemcpyHostToDevice, st));
    Fr* W = (Fr*)sl.d_wires.ptr;
    k_wires_from_be<<<dim3(cdiv(c->nw + 1, 256), (unsigned)G), 256, 0, st>>>((const uint4*)sl.d_wires_be.ptr,
                                                                            (const uint4*)sl.d_rnd_be.ptr, W, c->wstride,
                                                                            (uint32_t)c->nw);
    G16_STAGE_CUDA(cudaGetLastError());
    sl.commits.assign(G, G1Affine::inf());
    if (c->has_commitment) {   // commitment point = MSM of the committed wire values over the commitment basis
        int rc = c->g1_aux.run(c->bCommit, W, c->wstride, c->d_map_commit, 1, G, (G1Affine*)sl.d_commit_out.ptr, st);
        if (rc != G16_OK) return failm(rc, get_error());
        k_fp_from_mont<<<cdiv(G * 2, 256), 256, 0, st>>>((Fp*)sl.d_commit_out.ptr, G * 2);
        G16_STAGE_CUDA(cudaMemcpyAsync(sl.commits.data(), sl.d_commit_out.ptr, sizeof(G1Affine) * G, cudaMemcpyDeviceToHost, st));
    }
    G16_STAGE_CUDA(cudaEventRecord(sl.ready, st));
    G16_STAGE_CUDA(cudaStreamSynchronize(st));
    trace("wires: stage done");
#undef G16_STAGE_CUDA
    return res;
}

// ---- the two-stage pipeline shared by g16_prove_batch and g16_prove_wires ---------------------------------
// `launch(g)` starts stage A of group g (solve_batch proofs -> slot g&1: device wires, commitments, `ready`
// event); stage B proves the group in chunks of max_batch on the context stream.  Results return through two
// pinned buffers: chunk k-1 is serialised on the host while the device proves chunk k.  `after(first, B)` runs
// once the proofs [first, first+B) are written.
}  // extern "C" (templates need C++ linkage)
template <class Launch, class After>
static int run_pipeline(g16_circuit* c, size_t n, Launch launch, uint8_t* proofs, After after) {
    cudaStream_t st = c->ctx->stream;
    const size_t plen = c->has_commitment ? 388 : 324;
    // group schedule: the first group is small (nothing can be proved before it is solved), the others fill the slot
    const size_t SB = c->solve_batch, SB0 = std::min(SB, c->first_group);
    std::vector<size_t> gfirst(1, 0);
    for (size_t at = 0; at < n;) {
        at += std::min(gfirst.size() == 1 ? SB0 : SB, n - at);
        gfirst.push_back(at);
    }
    const size_t ngroups = gfirst.size() - 1;
    auto group_size = [&](size_t g) { return gfirst[g + 1] - gfirst[g]; };
    struct Pending {
        bool live = false;
        size_t first = 0, B = 0, off = 0;
        int slot = 0, ring = 0;
    } pend;
    auto drain = [&]() -> int {
        if (!pend.live) return G16_OK;
        pend.live = false;
        if (cudaEventSynchronize(c->ev_done[pend.ring]) != cudaSuccess) {
            set_error(std::string("device pipeline: ") + cudaGetErrorString(cudaGetLastError()));
            return G16_E_CUDA;
        }
        trace("pipeline: chunk done", (long)pend.first);
        const ProofPoints* pts = c->h_pts[pend.ring];
        const g16_circuit::Slot& sl = c->slots[pend.slot];
        for (size_t b = 0; b < pend.B; b++)
            write_proof_bytes(pts[b], c->has_commitment ? &sl.commits[pend.off + b] : nullptr, proofs + plen * (pend.first + b));
        after(pend.first, pend.B);
        return G16_OK;
    };
    std::future<StageResult> fut = launch((size_t)0, gfirst[0], group_size(0));
    int total_launches = 0, rc = G16_OK;
    size_t chunk_no = 0;
    for (size_t g = 0; g < ngroups && rc == G16_OK; g++) {
        StageResult sr = fut.get();
        if (sr.rc != G16_OK) {
            set_error(sr.err);
            rc = sr.rc;
            break;
        }
        // the slot stage g+1 writes was last read by the chunk that may still be in flight
        if ((rc = drain()) != G16_OK) break;
        if (g + 1 < ngroups) fut = launch(g + 1, gfirst[g + 1], group_size(g + 1));   // stage A of group g+1 runs while the device proves group g
        g16_circuit::Slot& sl = c->slots[g & 1];
        const size_t G = group_size(g);
        if (cudaStreamWaitEvent(st, sl.ready, 0) != cudaSuccess) {
            rc = G16_E_CUDA;
            set_error("cudaStreamWaitEvent failed");
        }
        for (size_t off = 0; off < G && rc == G16_OK; off += c->max_batch) {
            const size_t B = std::min(c->max_batch, G - off), first = gfirst[g] + off;
            const int ring = (int)(chunk_no++ & 1);
            trace("pipeline: chunk begin", (long)first);
            int p = 0;
            if ((rc = chunk_begin(c, &p)) != G16_OK) break;
            if ((rc = prove_device(c, B, (const Fr*)sl.d_wires.ptr + off * c->wstride, p)) != G16_OK) break;
            total_launches += c->last_launches;
            // consecutive chunks are independent on the device (two buffer sets): no join on the context stream
            if (cudaMemcpyAsync(c->h_pts[ring], c->d_out[p].ptr, sizeof(ProofPoints) * B, cudaMemcpyDeviceToHost, chunk_stream(c)) != cudaSuccess ||
                cudaEventRecord(c->ev_done[ring], chunk_stream(c)) != cudaSuccess || chunk_end(c, p, true) != G16_OK) {
                rc = G16_E_CUDA;
                set_error(std::string("device pipeline: ") + cudaGetErrorString(cudaGetLastError()));
                break;
            }
            trace("pipeline: chunk enqueued");
            Pending cur;
            cur.live = true; cur.first = first; cur.B = B; cur.off = off; cur.slot = (int)(g & 1); cur.ring = ring;
            if ((rc = drain()) != G16_OK) break;   // the previous chunk: its bytes are written while this one runs
            pend = cur;
        }
        if (rc != G16_OK && g + 1 < ngroups && fut.valid()) fut.wait();   // never leave the worker running on freed state
    }
    if (rc == G16_OK) rc = drain();
    else if (fut.valid()) fut.wait();
    if (rc != G16_OK) {
        cudaStreamSynchronize(st);
        return rc;
    }
    c->last_launches = total_launches;
    c->ctx->last_launches = total_launches;
    return G16_OK;
}

extern "C" {

int g16_prove_wires(g16_circuit* c, size_t n, const uint8_t* wires_be, const uint8_t* rnd, uint8_t* proofs) {
    if (!c || !wires_be || !proofs || n == 0) {
        set_error("g16_prove_wires: bad arguments");
        return G16_E_ARG;
    }
    G16_CUDA(cudaSetDevice(c->ctx->device));
    G16_LOCK(c->ctx);
    if (c->world > 1) return prove_wires_sync(c, n, wires_be, rnd, proofs);
    auto launch = [&](size_t g, size_t first, size_t G) {
        return std::async(std::launch::async, stage_wires, c, (int)(g & 1), G, wires_be + first * c->nw * 32,
                          rnd ? rnd + 96 * first : nullptr, first);
    };
    return run_pipeline(c, n, launch, proofs, [](size_t, size_t) {});
}

int g16_prove_batch(g16_circuit* c, size_t n, const uint8_t* assignments_be, size_t n_values, const uint8_t* rnd,
                    uint8_t* proofs, uint8_t* pws, size_t pw_stride) {
    if (!c || !assignments_be || !proofs || n == 0) {
        set_error("g16_prove_batch: bad arguments");
        return G16_E_ARG;
    }
    const Circuit& circ = c->circ;
    const size_t nin = circ.nb_public - 1 + circ.nb_secret;
    const size_t npub = circ.nb_public - 1;
    if (n_values != nin) {
        set_error("g16_prove_batch: assignment has " + std::to_string(n_values) + " values, circuit wants " + std::to_string(nin));
        return G16_E_ARG;
    }
    if (pws && pw_stride < 12 + 32 * npub) {
        set_error("g16_prove_batch: pw_stride too small");
        return G16_E_ARG;
    }
    G16_CUDA(cudaSetDevice(c->ctx->device));
    G16_LOCK(c->ctx);
    if (c->world > 1) {
        set_error("g16_prove_batch: this circuit is split across ranks (g16_comm_init): call g16_prove_wires / "
                  "g16_prove_wires_dev collectively on every rank");
        return G16_E_ARG;
    }
    // groups of solve_batch proofs are solved together (stage A); each group is proved in chunks of
    // max_batch (stage B) while the next group is being solved
    const bool solve_overlap = !(getenv("G16_SOLVE_OVERLAP") && atoi(getenv("G16_SOLVE_OVERLAP")) == 0);
    auto launch = [&](size_t g, size_t first, size_t G) {
        if (c->plan.valid)   // G16_SOLVE_OVERLAP=0 runs the device solver between groups instead of beside them
            return std::async(solve_overlap ? std::launch::async : std::launch::deferred, stage_solve_gpu, c, (int)(g & 1), G,
                              assignments_be + first * nin * 32, rnd ? rnd + 96 * first : nullptr, first, false);
        return std::async(std::launch::async, stage_solve, c, (int)(g & 1), G, assignments_be + first * nin * 32,
                          rnd ? rnd + 96 * first : nullptr, first, true);
    };
    return run_pipeline(c, n, launch, proofs, [&](size_t first, size_t B) {
        if (!pws) return;
        parallel_for(B, [&](size_t b) {
            uint8_t* o = pws + pw_stride * (first + b);
            const uint32_t hdr[3] = {(uint32_t)npub, 0, (uint32_t)npub};
            for (int q = 0; q < 3; q++) {
                o[4 * q] = hdr[q] >> 24; o[4 * q + 1] = hdr[q] >> 16; o[4 * q + 2] = hdr[q] >> 8; o[4 * q + 3] = hdr[q];
            }
            for (size_t i = 0; i < npub; i++)   // public wires = the first npub assignment values, reduced mod r
                HFr::from_be(assignments_be + ((first + b) * nin + i) * 32).to_be(o + 12 + 32 * i);
        });
    });
}

int g16_prove(g16_circuit* c, const uint8_t* witness_gz, size_t witness_len, const uint8_t rnd[96], uint8_t* proof,
              size_t* proof_len, uint8_t* pw, size_t* pw_len) {
    if (!c || !witness_gz) {
        set_error("g16_prove: bad arguments");
        return G16_E_ARG;
    }
    std::vector<uint8_t> asg;
    G16_TRY(witness_to_assignment(c->circ, witness_gz, witness_len, &asg));
    return g16_prove_assignment(c, asg.data(), asg.size() / 32, rnd, proof, proof_len, pw, pw_len);
}

int g16_complete_assignment(const uint8_t* ccs, size_t ccs_len, const uint32_t* known_wires, const uint8_t* known_values_be,
                            size_t n_known, uint8_t* assignment_be, size_t* n_values) {
    if (!ccs || !n_values || (n_known && (!known_wires || !known_values_be))) {
        set_error("g16_complete_assignment: bad arguments");
        return G16_E_ARG;
    }
    Circuit circ;
    G16_TRY(parse_ccs(ccs, ccs_len, &circ));
    const size_t nin = circ.nb_public - 1 + circ.nb_secret;
    if (!assignment_be) {
        *n_values = nin;
        return G16_OK;
    }
    if (*n_values < nin) {
        set_error("g16_complete_assignment: output buffer too small");
        return G16_E_ARG;
    }
    std::vector<std::pair<uint32_t, HFr>> known(n_known);
    for (size_t i = 0; i < n_known; i++) known[i] = {known_wires[i], HFr::from_be(known_values_be + 32 * i)};
    std::vector<HFr> asg;
    std::string err;
    int rc = complete_assignment(circ, known, &asg, &err);
    if (rc != G16_OK) {
        set_error("g16_complete_assignment: " + err);
        return rc;
    }
    for (size_t i = 0; i < nin; i++) asg[i].to_be(assignment_be + 32 * i);
    *n_values = nin;
    return G16_OK;
}

int g16_witness_to_assignment(const uint8_t* ccs, size_t ccs_len, const uint8_t* witness_gz, size_t witness_len,
                              uint8_t* assignment_be, size_t* n_values) {
    if (!ccs || !witness_gz || !n_values) {
        set_error("g16_witness_to_assignment: bad arguments");
        return G16_E_ARG;
    }
    Circuit circ;
    G16_TRY(parse_ccs(ccs, ccs_len, &circ));
    std::vector<uint8_t> asg;
    G16_TRY(witness_to_assignment(circ, witness_gz, witness_len, &asg));
    const size_t n = asg.size() / 32;
    if (assignment_be) {
        if (*n_values < n) {
            set_error("g16_witness_to_assignment: output buffer too small");
            return G16_E_ARG;
        }
        memcpy(assignment_be, asg.data(), asg.size());
    }
    *n_values = n;
    return G16_OK;
}

// Full wire vectors only (two-phase solve with the commitment MSM on the GPU), no proof.
int g16_witness_batch(g16_circuit* c, size_t n, const uint8_t* assignments_be, size_t n_values, const uint8_t* rnd,
                      uint8_t* wires_be) {
    if (!c || !assignments_be || !wires_be || n == 0) {
        set_error("g16_witness_batch: bad arguments");
        return G16_E_ARG;
    }
    const size_t nin = c->circ.nb_public - 1 + c->circ.nb_secret;
    if (n_values != nin) {
        set_error("g16_witness_batch: assignment size mismatch");
        return G16_E_ARG;
    }
    G16_LOCK(c->ctx);
    for (size_t done = 0; done < n;) {
        size_t B = std::min(c->solve_batch, n - done);
        StageResult sr = stage_solve(c, 0, B, assignments_be + done * nin * 32, rnd ? rnd + 96 * done : nullptr, done, false);
        if (sr.rc != G16_OK) {
            set_error(sr.err);
            return sr.rc;
        }
        const HFr* W = (const HFr*)c->slots[0].h_wires;
        parallel_for(B, [&](size_t b) {
            for (size_t i = 0; i < c->nw; i++) W[b * c->wstride + i].to_be(wires_be + ((done + b) * c->nw + i) * 32);
        });
        done += B;
    }
    return G16_OK;
}

// The same through the DEVICE solver (the one g16_prove_batch uses); fails with G16_E_HINT when the circuit has to be
// solved on the host.  G16_SOLVER_DIAG=1: unsatisfied rows do not fail the call (tests).
int g16_witness_batch_dev(g16_circuit* c, size_t n, const uint8_t* assignments_be, size_t n_values, const uint8_t* rnd,
                          uint8_t* wires_be) {
    if (!c || !assignments_be || !wires_be || n == 0) {
        set_error("g16_witness_batch_dev: bad arguments");
        return G16_E_ARG;
    }
    const size_t nin = c->circ.nb_public - 1 + c->circ.nb_secret;
    if (n_values != nin) {
        set_error("g16_witness_batch_dev: assignment size mismatch");
        return G16_E_ARG;
    }
    if (!c->plan.valid) {
        set_error("g16_witness_batch_dev: this circuit is solved on the host (" + c->host_solver_reason + ")");
        return G16_E_HINT;
    }
    G16_CUDA(cudaSetDevice(c->ctx->device));
    G16_LOCK(c->ctx);
    const bool tolerate = getenv("G16_SOLVER_DIAG") && atoi(getenv("G16_SOLVER_DIAG")) != 0;
    std::vector<HFr> w;
    for (size_t done = 0; done < n;) {
        const size_t B = std::min(c->solve_batch, n - done);
        StageResult sr = stage_solve_gpu(c, 0, B, assignments_be + done * nin * 32, rnd ? rnd + 96 * done : nullptr, done, tolerate);
        if (sr.rc != G16_OK) {
            set_error(sr.err);
            return sr.rc;
        }
        w.resize(c->wstride * B);
        G16_CUDA(cudaMemcpy(w.data(), c->slots[0].d_wires.ptr, sizeof(Fr) * c->wstride * B, cudaMemcpyDeviceToHost));
        parallel_for(B, [&](size_t b) {
            for (size_t i = 0; i < c->nw; i++) w[b * c->wstride + i].to_be(wires_be + ((done + b) * c->nw + i) * 32);
        });
        done += B;
    }
    return G16_OK;
}

int g16_prove_assignment(g16_circuit* c, const uint8_t* assignment_be, size_t n_values, const uint8_t rnd[96],
                         uint8_t* proof, size_t* proof_len, uint8_t* pw, size_t* pw_len) {
    if (!c || !proof || !proof_len) {
        set_error("g16_prove_assignment: bad arguments");
        return G16_E_ARG;
    }
    const size_t plen = c->has_commitment ? 388 : 324;
    const size_t wlen = 12 + 32 * (c->circ.nb_public - 1);
    if (*proof_len < plen || (pw && (!pw_len || *pw_len < wlen))) {
        set_error("g16_prove_assignment: output buffer too small");
        return G16_E_ARG;
    }
    G16_TRY(g16_prove_batch(c, 1, assignment_be, n_values, rnd, proof, pw, wlen));
    *proof_len = plen;
    if (pw_len) *pw_len = wlen;
    return G16_OK;
}

// Solver-only entry point (no GPU): extends an assignment to the full wire vector.  The BSB22
// challenges (one per commitment) are supplied by the caller; `committed_be` (optional) receives
// the committed values the challenge must be derived from.
int g16_solve_assignment(const uint8_t* ccs, size_t ccs_len, const uint8_t* assignment_be, size_t n_values,
                         const uint8_t* blinder_be, const uint8_t* challenges_be, size_t n_challenges,
                         uint8_t* wires_be, size_t wires_cap, uint8_t* committed_be, size_t committed_cap) {
    if (!ccs || !assignment_be || !wires_be) {
        set_error("g16_solve_assignment: bad arguments");
        return G16_E_ARG;
    }
    Circuit circ;
    G16_TRY(parse_ccs(ccs, ccs_len, &circ));
    const size_t nin = circ.nb_public - 1 + circ.nb_secret;
    if (n_values != nin || wires_cap < (size_t)circ.nb_wires() * 32) {
        set_error("g16_solve_assignment: size mismatch (assignment " + std::to_string(n_values) + " vs " + std::to_string(nin) + ")");
        return G16_E_ARG;
    }
    std::vector<HFr> asg(nin);
    for (size_t i = 0; i < nin; i++) asg[i] = HFr::from_be(assignment_be + 32 * i);
    HFr blinder = blinder_be ? HFr::from_be(blinder_be) : HFr::zero();
    SolveState stt;
    stt.tolerate = getenv("G16_SOLVER_DIAG") && atoi(getenv("G16_SOLVER_DIAG")) != 0;
    solve_begin(circ, asg.data(), &stt);
    size_t used = 0;
    for (;;) {
        int rc = solve_run(circ, &stt, blinder_be ? &blinder : nullptr);
        if (rc == SOLVE_DONE) break;
        if (rc == SOLVE_NEED_COMMITMENT) {
            if (committed_be && used == 0) {
                if (committed_cap < stt.committed.size() * 32) {
                    set_error("g16_solve_assignment: committed buffer too small");
                    return G16_E_ARG;
                }
                for (size_t k = 0; k < stt.committed.size(); k++) stt.committed[k].to_be(committed_be + 32 * k);
            }
            if (used >= n_challenges || !challenges_be) {
                set_error("g16_solve_assignment: circuit needs a commitment challenge that was not supplied");
                return G16_E_HINT;
            }
            solve_provide_challenge(&stt, HFr::from_be(challenges_be + 32 * used));
            used++;
            continue;
        }
        set_error(stt.error);
        return rc;
    }
    for (size_t i = 0; i < stt.w.size(); i++) stt.w[i].to_be(wires_be + 32 * i);
    return G16_OK;
}

}  // extern "C"

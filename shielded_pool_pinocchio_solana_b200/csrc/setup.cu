// setup.cu -- Groth16 trusted setup on the GPU (the `sunspot setup` step,
// /root/reference/noir_circuit/prove_linux.sh:73-79, scripts/generate_audit.py:671; gnark
// backend/groth16/bn254/setup.go -- SURVEY.md 8(f) rank 2, 9.3).
//
// The reference's proving keys are missing blobs (/root/reference/.MISSING_LARGE_BLOBS), so a
// key has to be regenerated before anything can be proved.  The toxic waste is derived from a
// caller-supplied seed (deterministic: tests compare the key byte-for-byte with the oracle's);
// a production ceremony would feed fresh randomness and discard it.
//   host  : Lagrange basis at tau, per-wire A_i(tau), B_i(tau), C_i(tau), all point scalars
//   device: every key point is a fixed-base multiple of the G1 / G2 generator (one thread each)
//   host  : gnark raw serialisation of ProvingKey / VerifyingKey
#include <string.h>

#include "capi.cuh"
#include "ccs.hpp"
#include "hostutil.hpp"

namespace g16 {

template <class F>
__global__ void __launch_bounds__(128) k_fixed_base_mul(Affine<F> gen, const Fr* __restrict__ scalars, uint32_t n,
                                                        Affine<F>* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr k = scalars[i];  // canonical
    XYZZ<F> acc = XYZZ<F>::inf();
    for (int w = 7; w >= 0; w--) {
        uint32_t limb = k.v[w];
        for (int b = 31; b >= 0; b--) {
            acc = acc.dbl();
            if ((limb >> b) & 1u) acc.madd(gen);
        }
    }
    Affine<F> a = acc.to_affine();
    Fp* c = reinterpret_cast<Fp*>(&a);
#pragma unroll
    for (int j = 0; j < (int)(sizeof(Affine<F>) / sizeof(Fp)); j++) c[j] = c[j].from_mont();
    out[i] = a;
}

namespace {

HFr toxic(const char* name, const uint8_t* seed, size_t seed_len) {
    for (int ctr = 0;; ctr++) {
        uint8_t d[64];
        for (int half = 0; half < 2; half++) {
            std::string pre = std::string(half ? "g16b200/setup2/" : "g16b200/setup/") + name + "/";
            Sha256 s;
            s.update((const uint8_t*)pre.data(), pre.size());
            s.update(seed, seed_len);
            std::string suf = "/" + std::to_string(ctr);
            s.update((const uint8_t*)suf.data(), suf.size());
            s.finish(d + 32 * half);
        }
        HFr v = HFr::from_be(d).to_mont() + HFr::from_be(d + 32);  // (hi * 2^256 + lo) mod r
        if (!v.is_zero()) return v;
    }
}

void put_u32(std::vector<uint8_t>& o, uint32_t v) {
    o.push_back(v >> 24); o.push_back(v >> 16); o.push_back(v >> 8); o.push_back(v);
}
void put_u64(std::vector<uint8_t>& o, uint64_t v) {
    put_u32(o, (uint32_t)(v >> 32));
    put_u32(o, (uint32_t)v);
}
void put_fr(std::vector<uint8_t>& o, const HFr& x) {
    uint8_t b[32];
    x.to_be(b);
    o.insert(o.end(), b, b + 32);
}
void put_g1(std::vector<uint8_t>& o, const G1Affine& p) {
    uint8_t b[64];
    g1_to_be(p, b);
    o.insert(o.end(), b, b + 64);
}
void put_g2(std::vector<uint8_t>& o, const G2Affine& p) {
    uint8_t b[128];
    g2_to_be(p, b);
    o.insert(o.end(), b, b + 128);
}

// batch inversion (Montgomery's trick); zeros stay zero
void batch_inverse(std::vector<HFr>& v) {
    std::vector<HFr> pref(v.size());
    HFr acc = HFr::one();
    for (size_t i = 0; i < v.size(); i++) {
        pref[i] = acc;
        if (!v[i].is_zero()) acc = acc * v[i];
    }
    HFr inv = acc.inverse();
    for (size_t i = v.size(); i-- > 0;) {
        if (v[i].is_zero()) continue;
        HFr t = inv * pref[i];
        inv = inv * v[i];
        v[i] = t;
    }
}

}  // namespace
}  // namespace g16

using namespace g16;

extern "C" int g16_setup(g16_ctx* ctx, const uint8_t* ccs, size_t ccs_len, const uint8_t* seed, size_t seed_len,
                         uint8_t* pk_out, size_t* pk_len, uint8_t* vk_out, size_t* vk_len) {
    if (!ctx || !ccs || !seed || !pk_len || !vk_len) {
        set_error("g16_setup: bad arguments");
        return G16_E_ARG;
    }
    G16_CUDA(cudaSetDevice(ctx->device));
    G16_LOCK(ctx);
    Circuit c;
    G16_TRY(parse_ccs(ccs, ccs_len, &c));
    if (c.commitments.size() > 1) {
        set_error("g16_setup: more than one commitment is not supported yet");
        return G16_E_ARG;
    }
    const unsigned logn = c.log_domain();
    const size_t n = (size_t)1 << logn, nw = c.nb_wires(), npub = c.nb_public;
    HFr tau = toxic("tau", seed, seed_len), alpha = toxic("alpha", seed, seed_len), beta = toxic("beta", seed, seed_len),
        gamma = toxic("gamma", seed, seed_len), delta = toxic("delta", seed, seed_len),
        sigma = toxic("sigma", seed, seed_len), g2k = toxic("g2k", seed, seed_len);
    HFr ginv = gamma.inverse(), dinv = delta.inverse();
    // domain generator
    HFr w;
    {
        static const uint64_t ROOT[4] = {0x9bd61b6e725b19f0ull, 0x402d111e41112ed4ull, 0x00e0a7eb8ef62abcull, 0x2a3c09f0a58a7e85ull};
        w = HFr{{ROOT[0], ROOT[1], ROOT[2], ROOT[3]}}.to_mont();
        for (unsigned i = logn; i < 28; i++) w = w.sqr();
    }
    // Lagrange basis at tau over the rows: L_k = (tau^n - 1) w^k / (n (tau - w^k))
    HFr tn = tau;
    for (unsigned i = 0; i < logn; i++) tn = tn.sqr();
    HFr zt = tn - HFr::one();
    HFr ninv = HFr::from_u64(n).inverse();
    std::vector<HFr> den(c.nb_constraints), wk(c.nb_constraints);
    {
        HFr cur = HFr::one();
        for (uint32_t k = 0; k < c.nb_constraints; k++) {
            wk[k] = cur;
            den[k] = tau - cur;
            cur = cur * w;
        }
    }
    batch_inverse(den);
    std::vector<HFr> lag(c.nb_constraints);
    HFr ztn = zt * ninv;
    for (uint32_t k = 0; k < c.nb_constraints; k++) lag[k] = ztn * wk[k] * den[k];
    std::vector<HFr> A(nw, HFr::zero()), Bv(nw, HFr::zero()), Cv(nw, HFr::zero());
    const Circuit::Csr* M[3] = {&c.A, &c.B, &c.C};
    std::vector<HFr>* dst[3] = {&A, &Bv, &Cv};
    for (int m = 0; m < 3; m++)
        for (uint32_t row = 0; row < c.nb_constraints; row++)
            for (uint32_t k = M[m]->rowptr[row]; k < M[m]->rowptr[row + 1]; k++) {
                HFr& d = (*dst[m])[M[m]->wire[k]];
                d = d + c.coeffs[M[m]->coeff[k]] * lag[row];
            }
    // ---- scalar lists -------------------------------------------------------------------------
    std::vector<uint8_t> inf_a(nw), inf_b(nw), skip(nw, 0);
    std::vector<uint32_t> committed;
    uint32_t commit_wire = 0;
    if (!c.commitments.empty()) {
        committed = c.commitments[0].private_committed;
        commit_wire = c.commitments[0].commitment_index;
        skip[commit_wire] = 1;
        for (uint32_t x : committed) skip[x] = 1;
    }
    std::vector<HFr> s1;   // G1 scalars, in the order: alpha beta delta | A | B | Z | K | vkK | basis | basisExpSigma
    std::vector<HFr> s2;   // G2 scalars: beta delta gamma | B | G GSigmaNeg
    s1.push_back(alpha); s1.push_back(beta); s1.push_back(delta);
    s2.push_back(beta); s2.push_back(delta); s2.push_back(gamma);
    size_t nA = 0, nB = 0;
    for (size_t i = 0; i < nw; i++) {
        inf_a[i] = A[i].is_zero();
        if (!inf_a[i]) { s1.push_back(A[i]); nA++; }
    }
    for (size_t i = 0; i < nw; i++) {
        inf_b[i] = Bv[i].is_zero();
        if (!inf_b[i]) { s1.push_back(Bv[i]); s2.push_back(Bv[i]); nB++; }
    }
    {
        HFr z = zt * dinv;
        std::vector<HFr> znat(n);
        for (size_t i = 0; i < n; i++) { znat[i] = z; z = z * tau; }
        for (size_t p = 0; p + 1 < n; p++) {   // bit-reversed order, n-1 entries
            size_t r = 0;
            for (unsigned b = 0; b < logn; b++) r |= ((p >> b) & 1) << (logn - 1 - b);
            s1.push_back(znat[r]);
        }
    }
    std::vector<HFr> t(nw);
    for (size_t i = 0; i < nw; i++) t[i] = beta * A[i] + alpha * Bv[i] + Cv[i];
    size_t nK = 0;
    for (size_t i = npub; i < nw; i++)
        if (!skip[i]) { s1.push_back(t[i] * dinv); nK++; }
    size_t nVk = 0;
    for (size_t i = 0; i < npub; i++) { s1.push_back(t[i] * ginv); nVk++; }
    if (!c.commitments.empty()) { s1.push_back(t[commit_wire] * ginv); nVk++; }
    for (uint32_t x : committed) s1.push_back(t[x] * ginv);
    for (uint32_t x : committed) s1.push_back(t[x] * ginv * sigma);
    if (!c.commitments.empty()) { s2.push_back(g2k); s2.push_back((sigma * g2k).neg()); }
    // ---- size query: everything below only fills bytes whose count is already known ---------------
    {
        const size_t ncom = c.commitments.size();
        const size_t need_pk = 8 + 5 * 32 + 1 + 3 * 64 + (4 + nA * 64) + (4 + nB * 64) + (4 + (n - 1) * 64) + (4 + nK * 64) +
                               2 * 128 + (4 + nB * 128) + 3 * 8 + 2 * (4 + nw) + 4 + (ncom ? 2 * (4 + committed.size() * 64) : 0);
        size_t need_vk = 2 * 64 + 2 * 128 + 64 + 128 + 4 + nVk * 64 + 4 + 4;
        for (auto& info : c.commitments) need_vk += 4 + 8 * info.public_and_commitment_committed.size();
        if (ncom) need_vk += 2 * 128;
        if (!pk_out && !vk_out) {
            *pk_len = need_pk;
            *vk_len = need_vk;
            return G16_OK;
        }
    }
    // ---- device: fixed-base multiplications ------------------------------------------------------
    cudaStream_t st = ctx->stream;
    std::vector<HFr> c1(s1.size()), c2(s2.size());
    for (size_t i = 0; i < s1.size(); i++) c1[i] = s1[i].from_mont();
    for (size_t i = 0; i < s2.size(); i++) c2[i] = s2[i].from_mont();
    DeviceBuf d_s1, d_s2, d_p1, d_p2;
    G16_TRY(d_s1.ensure(sizeof(Fr) * c1.size())); G16_TRY(d_p1.ensure(sizeof(G1Affine) * c1.size()));
    G16_TRY(d_s2.ensure(sizeof(Fr) * c2.size())); G16_TRY(d_p2.ensure(sizeof(G2Affine) * c2.size()));
    G16_CUDA(cudaMemcpyAsync(d_s1.ptr, c1.data(), sizeof(Fr) * c1.size(), cudaMemcpyHostToDevice, st));
    G16_CUDA(cudaMemcpyAsync(d_s2.ptr, c2.data(), sizeof(Fr) * c2.size(), cudaMemcpyHostToDevice, st));
    G1Affine g1gen;
    g1gen.x = Fp::one();
    g1gen.y = Fp::one().dbl();
    static const uint32_t X0[8] = {0xd992f6ed, 0x46debd5c, 0xf75edadd, 0x674322d4, 0x5e5c4479, 0x426a0066, 0x121f1e76, 0x1800deef};
    static const uint32_t X1[8] = {0xaef312c2, 0x97e485b7, 0x35a9e712, 0xf1aa4933, 0x31fb5d25, 0x7260bfb7, 0x920d483a, 0x198e9393};
    static const uint32_t Y0[8] = {0x66fa7daa, 0x4ce6cc01, 0x0c43d37b, 0xe3d1e769, 0x8dcb408f, 0x4aab7180, 0xdb8c6deb, 0x12c85ea5};
    static const uint32_t Y1[8] = {0xd122975b, 0x55acdadc, 0x70b38ef3, 0xbc4b3133, 0x690c3395, 0xec9e99ad, 0x585ff075, 0x090689d0};
    auto mk = [](const uint32_t l[8]) {
        Fp f;
        for (int i = 0; i < 8; i++) f.v[i] = l[i];
        return f.to_mont();
    };
    G2Affine g2gen;
    g2gen.x.c0 = mk(X0); g2gen.x.c1 = mk(X1); g2gen.y.c0 = mk(Y0); g2gen.y.c1 = mk(Y1);
    k_fixed_base_mul<Fp><<<cdiv(c1.size(), 128), 128, 0, st>>>(g1gen, (const Fr*)d_s1.ptr, (uint32_t)c1.size(), (G1Affine*)d_p1.ptr);
    k_fixed_base_mul<Fp2><<<cdiv(c2.size(), 128), 128, 0, st>>>(g2gen, (const Fr*)d_s2.ptr, (uint32_t)c2.size(), (G2Affine*)d_p2.ptr);
    G16_CUDA(cudaGetLastError());
    std::vector<G1Affine> p1(c1.size());
    std::vector<G2Affine> p2(c2.size());
    G16_CUDA(cudaMemcpyAsync(p1.data(), d_p1.ptr, sizeof(G1Affine) * p1.size(), cudaMemcpyDeviceToHost, st));
    G16_CUDA(cudaMemcpyAsync(p2.data(), d_p2.ptr, sizeof(G2Affine) * p2.size(), cudaMemcpyDeviceToHost, st));
    G16_CUDA(cudaStreamSynchronize(st));
    ctx->last_launches = 2;
    // ---- serialise (gnark raw layout) --------------------------------
    std::vector<uint8_t> pk, vk;
    put_u64(pk, n);
    put_fr(pk, ninv); put_fr(pk, w); put_fr(pk, w.inverse()); put_fr(pk, HFr::from_u64(5)); put_fr(pk, HFr::from_u64(5).inverse());
    pk.push_back(0);
    size_t o = 0;
    const G1Affine &pa = p1[0], &pb = p1[1], &pd = p1[2];
    o = 3;
    put_g1(pk, pa); put_g1(pk, pb); put_g1(pk, pd);
    auto slice1 = [&](std::vector<uint8_t>& dstv, size_t cnt) {
        put_u32(dstv, (uint32_t)cnt);
        for (size_t i = 0; i < cnt; i++) put_g1(dstv, p1[o + i]);
        o += cnt;
    };
    slice1(pk, nA);
    size_t oB = o;
    slice1(pk, nB);
    slice1(pk, n - 1);
    slice1(pk, nK);
    size_t oVk = o;
    o += nVk;
    put_g2(pk, p2[0]); put_g2(pk, p2[1]);
    put_u32(pk, (uint32_t)nB);
    for (size_t i = 0; i < nB; i++) put_g2(pk, p2[3 + i]);
    (void)oB;
    put_u64(pk, nw);
    uint64_t ia = 0, ib = 0;
    for (auto v : inf_a) ia += v;
    for (auto v : inf_b) ib += v;
    put_u64(pk, ia); put_u64(pk, ib);
    put_u32(pk, (uint32_t)nw); pk.insert(pk.end(), inf_a.begin(), inf_a.end());
    put_u32(pk, (uint32_t)nw); pk.insert(pk.end(), inf_b.begin(), inf_b.end());
    put_u32(pk, (uint32_t)c.commitments.size());
    if (!c.commitments.empty()) {
        slice1(pk, committed.size());
        slice1(pk, committed.size());
    }
    // vk: alpha1 beta1 beta2 gamma2 delta1 delta2 | K | PublicAndCommitmentCommitted | Pedersen vk
    put_g1(vk, pa); put_g1(vk, pb); put_g2(vk, p2[0]); put_g2(vk, p2[2]); put_g1(vk, pd); put_g2(vk, p2[1]);
    put_u32(vk, (uint32_t)nVk);
    for (size_t i = 0; i < nVk; i++) put_g1(vk, p1[oVk + i]);
    put_u32(vk, (uint32_t)c.commitments.size());
    for (auto& info : c.commitments) {
        put_u32(vk, (uint32_t)info.public_and_commitment_committed.size());
        for (uint32_t x : info.public_and_commitment_committed) put_u64(vk, x);
    }
    put_u32(vk, (uint32_t)c.commitments.size());
    if (!c.commitments.empty()) { put_g2(vk, p2[3 + nB]); put_g2(vk, p2[4 + nB]); }
    bool fits = pk_out && vk_out && *pk_len >= pk.size() && *vk_len >= vk.size();
    size_t need_pk = pk.size(), need_vk = vk.size();
    if (fits) {
        memcpy(pk_out, pk.data(), pk.size());
        memcpy(vk_out, vk.data(), vk.size());
    }
    *pk_len = need_pk;
    *vk_len = need_vk;
    if (!fits && (pk_out || vk_out)) {
        set_error("g16_setup: output buffers too small (sizes returned in *pk_len / *vk_len)");
        return G16_E_ARG;
    }
    return G16_OK;
}

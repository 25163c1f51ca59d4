// verify.cpp -- host-side Groth16 verifier for BN254 with gnark's BSB22 commitment extension:
// g16_verify() and the `g16prove verify <vk> <proof> <pw>` command.
//
// Stands in for `sunspot verify $VK $PROOF $PW` (/root/reference/noir_circuit/prove_linux.sh:87,
// audit_circuit/prove_audit.sh:99) = gnark groth16.Verify (backend/groth16/bn254/verify.go,
// third-party), whose equations SURVEY.md 9.4 records:
//     chal   = hash_to_field("bsb22-commitment", Commitment || committed public wires)
//     vk_x   = K_0 + sum pub_i K_i + chal * K_commit + Commitment
//     e(Ar, Bs) = e(alpha, beta) e(vk_x, gamma) e(Krs, delta)
//     e(Commitment, GSigmaNeg) e(PoK, G) = 1
// Not on the hot path: plain C++ on the host-compiled field code of ff.cuh.  Fp12 is the polynomial
// ring Fp[w]/(w^12 - 18 w^6 + 82) (Fp2's u = w^6 - 9), the Miller loop is the textbook optimal-ate loop
// on the twisted point embedded in Fp12, the final exponentiation is (p^6-1)(p^2+1) by Frobenius maps
// times a plain power by (p^4 - p^2 + 1)/r.
#include <string.h>

#include <mutex>
#include <vector>

#include "capi.cuh"
#include "hostutil.hpp"

namespace g16 {
namespace {

Fp fp_small(unsigned k) {   // k as a Montgomery field element
    Fp acc = Fp::zero(), base = Fp::one();
    for (; k; k >>= 1) {
        if (k & 1) acc = acc + base;
        base = base.dbl();
    }
    return acc;
}
Fp mul_small(const Fp& x, unsigned k) {
    Fp acc = Fp::zero(), base = x;
    for (; k; k >>= 1) {
        if (k & 1) acc = acc + base;
        base = base.dbl();
    }
    return acc;
}

struct Fq12 {
    Fp c[12];
    static Fq12 zero() {
        Fq12 r;
        for (auto& x : r.c) x = Fp::zero();
        return r;
    }
    static Fq12 one() {
        Fq12 r = zero();
        r.c[0] = Fp::one();
        return r;
    }
    static Fq12 scalar(const Fp& x) {
        Fq12 r = zero();
        r.c[0] = x;
        return r;
    }
    bool operator==(const Fq12& o) const {
        for (int i = 0; i < 12; i++)
            if (c[i] != o.c[i]) return false;
        return true;
    }
    bool is_zero() const {
        for (auto& x : c)
            if (!x.is_zero()) return false;
        return true;
    }
    Fq12 operator+(const Fq12& o) const {
        Fq12 r;
        for (int i = 0; i < 12; i++) r.c[i] = c[i] + o.c[i];
        return r;
    }
    Fq12 operator-(const Fq12& o) const {
        Fq12 r;
        for (int i = 0; i < 12; i++) r.c[i] = c[i] - o.c[i];
        return r;
    }
    Fq12 neg() const {
        Fq12 r;
        for (int i = 0; i < 12; i++) r.c[i] = c[i].neg();
        return r;
    }
    Fq12 small(unsigned k) const {
        Fq12 r;
        for (int i = 0; i < 12; i++) r.c[i] = mul_small(c[i], k);
        return r;
    }
    Fq12 operator*(const Fq12& o) const {
        Fp b[23];
        for (auto& x : b) x = Fp::zero();
        for (int i = 0; i < 12; i++) {
            if (c[i].is_zero()) continue;
            for (int j = 0; j < 12; j++)
                if (!o.c[j].is_zero()) b[i + j] = b[i + j] + c[i] * o.c[j];
        }
        for (int e = 22; e >= 12; e--) {   // w^12 = 18 w^6 - 82
            if (b[e].is_zero()) continue;
            b[e - 6] = b[e - 6] + mul_small(b[e], 18);
            b[e - 12] = b[e - 12] - mul_small(b[e], 82);
        }
        Fq12 r;
        for (int i = 0; i < 12; i++) r.c[i] = b[i];
        return r;
    }
    Fq12 mul_by_w() const {
        Fq12 r;
        r.c[0] = mul_small(c[11], 82).neg();
        for (int i = 1; i < 12; i++) r.c[i] = c[i - 1];
        r.c[6] = r.c[6] + mul_small(c[11], 18);
        return r;
    }
    // g with this * g = 1: the 12x12 linear system whose column j holds the coefficients of this * w^j
    Fq12 inverse() const {
        Fp m[12][13];
        Fq12 col = *this;
        for (int j = 0; j < 12; j++) {
            for (int i = 0; i < 12; i++) m[i][j] = col.c[i];
            col = col.mul_by_w();
        }
        for (int i = 0; i < 12; i++) m[i][12] = i == 0 ? Fp::one() : Fp::zero();
        for (int col_i = 0; col_i < 12; col_i++) {
            int piv = -1;
            for (int r = col_i; r < 12; r++)
                if (!m[r][col_i].is_zero()) {
                    piv = r;
                    break;
                }
            if (piv < 0) return zero();   // not invertible (only for 0)
            if (piv != col_i)
                for (int k = 0; k < 13; k++) {
                    Fp t = m[piv][k];
                    m[piv][k] = m[col_i][k];
                    m[col_i][k] = t;
                }
            Fp inv = m[col_i][col_i].inverse();
            for (int k = col_i; k < 13; k++) m[col_i][k] = m[col_i][k] * inv;
            for (int r = 0; r < 12; r++) {
                if (r == col_i || m[r][col_i].is_zero()) continue;
                Fp f = m[r][col_i];
                for (int k = col_i; k < 13; k++) m[r][k] = m[r][k] - f * m[col_i][k];
            }
        }
        Fq12 g;
        for (int i = 0; i < 12; i++) g.c[i] = m[i][12];
        return g;
    }
    Fq12 pow(const uint32_t* e, int nlimbs) const {
        Fq12 acc = one();
        bool started = false;
        for (int bit = nlimbs * 32 - 1; bit >= 0; bit--) {
            if (started) acc = acc * acc;
            if ((e[bit >> 5] >> (bit & 31)) & 1u) {
                acc = started ? acc * *this : *this;
                started = true;
            }
        }
        return acc;
    }
};

// (p^4 - p^2 + 1) / r
const uint32_t HARD_EXP[24] = {0xccdf42b1u, 0xe81bb482u, 0xf49c36d4u, 0x5abf5cc4u, 0x1da014fdu, 0xf1154e7eu,
                               0x87cdbacfu, 0xdcc7b44cu, 0x954bcf8au, 0xaaa441e3u, 0xd5095f23u, 0x6b887d56u,
                               0xf3fd90c6u, 0x79581e16u, 0xd189227du, 0x3b1b1355u, 0x61876f6bu, 0x4e529a58u,
                               0xd5b12278u, 0x6c0eb522u, 0x83177fafu, 0x331ec151u, 0x0b0759adu, 0x01baaa71u};
const uint64_t ATE_LOOP_LOW64 = 0x9d797039be763ba8ull;   // 6x + 2 = 2^64 + this
const uint32_t FR_MODULUS[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u,
                                0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};

// FW[i] = (w^i)^p : the Frobenius map is f = sum c_i w^i  ->  sum c_i FW[i]   (c_i in Fp)
struct Frob {
    Fq12 fw[12];
    Frob() {
        uint32_t pm1[8];
        Fp::modulus(pm1);
        pm1[0] -= 1;   // p is odd
        Fq12 w = Fq12::zero();
        w.c[1] = Fp::one();
        Fq12 wp = w.pow(pm1, 8) * w;   // w^p
        fw[0] = Fq12::one();
        for (int i = 1; i < 12; i++) fw[i] = fw[i - 1] * wp;
    }
    Fq12 operator()(const Fq12& f) const {
        Fq12 r = Fq12::zero();
        for (int i = 0; i < 12; i++) {
            if (f.c[i].is_zero()) continue;
            for (int k = 0; k < 12; k++) r.c[k] = r.c[k] + f.c[i] * fw[i].c[k];
        }
        return r;
    }
};
const Frob& frob() {
    static Frob f;   // C++11: thread-safe one-time initialisation
    return f;
}

struct Pt12 {
    Fq12 x, y;
    bool inf = false;
};

Pt12 pt_double(const Pt12& p) {
    Fq12 m = (p.x * p.x).small(3) * p.y.small(2).inverse();
    Pt12 r;
    r.x = m * m - p.x.small(2);
    r.y = m * (p.x - r.x) - p.y;
    return r;
}
Pt12 pt_add(const Pt12& a, const Pt12& b) {
    if (a.inf) return b;
    if (b.inf) return a;
    if (a.x == b.x) {
        if (a.y == b.y) return pt_double(a);
        Pt12 r;
        r.inf = true;
        return r;
    }
    Fq12 m = (b.y - a.y) * (b.x - a.x).inverse();
    Pt12 r;
    r.x = m * m - a.x - b.x;
    r.y = m * (a.x - r.x) - a.y;
    return r;
}
Fq12 linefunc(const Pt12& a, const Pt12& b, const Pt12& t) {
    if (!(a.x == b.x)) {
        Fq12 m = (b.y - a.y) * (b.x - a.x).inverse();
        return m * (t.x - a.x) - (t.y - a.y);
    }
    if (a.y == b.y) {
        Fq12 m = (a.x * a.x).small(3) * a.y.small(2).inverse();
        return m * (t.x - a.x) - (t.y - a.y);
    }
    return t.x - a.x;
}

// G2 point over Fp2 -> point of y^2 = x^3 + 3 over Fp12
Pt12 twist(const G2Affine& q) {
    Fq12 nx = Fq12::zero(), ny = Fq12::zero();
    nx.c[0] = q.x.c0 - mul_small(q.x.c1, 9);
    nx.c[6] = q.x.c1;
    ny.c[0] = q.y.c0 - mul_small(q.y.c1, 9);
    ny.c[6] = q.y.c1;
    Fq12 w2 = Fq12::zero(), w3 = Fq12::zero();
    w2.c[2] = Fp::one();
    w3.c[3] = Fp::one();
    Pt12 r;
    r.x = nx * w2;
    r.y = ny * w3;
    return r;
}

Fq12 miller_loop(const G2Affine& q, const G1Affine& p) {
    if (q.is_inf() || p.is_inf()) return Fq12::one();
    Pt12 Q = twist(q), Pt;
    Pt.x = Fq12::scalar(p.x);
    Pt.y = Fq12::scalar(p.y);
    Pt12 Rr = Q;
    Fq12 f = Fq12::one();
    for (int i = 63; i >= 0; i--) {
        f = f * f * linefunc(Rr, Rr, Pt);
        Rr = pt_double(Rr);
        if ((ATE_LOOP_LOW64 >> i) & 1ull) {
            f = f * linefunc(Rr, Q, Pt);
            Rr = pt_add(Rr, Q);
        }
    }
    const Frob& F = frob();
    Pt12 Q1, nQ2;
    Q1.x = F(Q.x);
    Q1.y = F(Q.y);
    nQ2.x = F(Q1.x);
    nQ2.y = F(Q1.y).neg();
    f = f * linefunc(Rr, Q1, Pt);
    Rr = pt_add(Rr, Q1);
    f = f * linefunc(Rr, nQ2, Pt);
    return f;
}

Fq12 final_exponentiation(const Fq12& f) {
    const Frob& F = frob();
    Fq12 f6 = f;
    for (int i = 0; i < 6; i++) f6 = F(f6);
    Fq12 a = f6 * f.inverse();          // f^(p^6 - 1)
    Fq12 b = F(F(a)) * a;               // ^(p^2 + 1)
    return b.pow(HARD_EXP, 24);
}

bool pairing_product_is_one(const std::vector<std::pair<G1Affine, G2Affine>>& pairs) {
    Fq12 f = Fq12::one();
    for (auto& pr : pairs) f = f * miller_loop(pr.second, pr.first);
    return final_exponentiation(f) == Fq12::one();
}

// ---- curve helpers (Montgomery coordinates) ------------------------------------------------------
G1Affine g1_mont(const G1Affine& p) {
    if (p.is_inf()) return p;
    return {p.x.to_mont(), p.y.to_mont()};
}
G2Affine g2_mont(const G2Affine& p) {
    if (p.is_inf()) return p;
    return {{p.x.c0.to_mont(), p.x.c1.to_mont()}, {p.y.c0.to_mont(), p.y.c1.to_mont()}};
}
bool g1_on_curve(const G1Affine& p) {
    if (p.is_inf()) return true;
    return p.y.sqr() == p.x.sqr() * p.x + fp_small(3);
}
bool g2_on_curve(const G2Affine& p) {
    if (p.is_inf()) return true;
    Fp2 b = Fp2{fp_small(3), Fp::zero()} * Fp2{fp_small(9), Fp::one()}.inverse();   // 3 / (9 + u)
    return p.y.sqr() == p.x.sqr() * p.x + b;
}
template <class F>
XYZZ<F> scalar_mul(const Affine<F>& p, const uint32_t* k /* 8 canonical limbs */) {
    XYZZ<F> acc = XYZZ<F>::inf();
    for (int bit = 255; bit >= 0; bit--) {
        acc = acc.dbl();
        if ((k[bit >> 5] >> (bit & 31)) & 1u) acc.madd(p);
    }
    return acc;
}
bool g2_in_subgroup(const G2Affine& p) { return scalar_mul(p, FR_MODULUS).is_inf(); }

bool fr_limbs_from_be(const uint8_t* be, uint32_t* out) {   // false when the value is not < r
    be32_to_limbs(be, out);
    uint32_t d[8];
    return ff_sub8(d, out, FR_MODULUS) != 0;
}

struct Reader {
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
    // raw points only; compressed encodings (top bits 10 / 11, or the 32-byte 01 infinity) are rejected
    bool g1(G1Affine* o) {
        if (!need(64) || (p[off] & 0xc0)) return ok = false;
        g1_from_be(p + off, o);
        off += 64;
        *o = g1_mont(*o);
        return true;
    }
    bool g2(G2Affine* o) {
        if (!need(128) || (p[off] & 0xc0)) return ok = false;
        g2_from_be(p + off, o);
        off += 128;
        *o = g2_mont(*o);
        return true;
    }
};

}  // namespace
}  // namespace g16

using namespace g16;

extern "C" int g16_verify(const uint8_t* vk, size_t vk_len, const uint8_t* proof, size_t proof_len, const uint8_t* pw,
                          size_t pw_len, int* ok) {
    if (!vk || !proof || !pw || !ok) {
        set_error("g16_verify: bad arguments");
        return G16_E_ARG;
    }
    *ok = 0;
    // ---- verifying key (VerifyingKey.WriteRawTo; layout pinned on the reference's committed .vk files)
    Reader r{vk, vk_len};
    G1Affine alpha1, beta1, delta1;
    G2Affine beta2, gamma2, delta2;
    r.g1(&alpha1); r.g1(&beta1); r.g2(&beta2); r.g2(&gamma2); r.g1(&delta1); r.g2(&delta2);
    uint32_t nk = r.u32();
    if (!r.ok || (size_t)nk * 64 > vk_len) {
        set_error("g16_verify: malformed verifying key");
        return G16_E_PARSE;
    }
    std::vector<G1Affine> K(nk);
    for (auto& k : K) r.g1(&k);
    uint32_t nsets = r.u32();
    if (!r.ok || nsets > 64) {
        set_error("g16_verify: malformed verifying key");
        return G16_E_PARSE;
    }
    std::vector<std::vector<uint64_t>> pacc(nsets);
    for (auto& lst : pacc) {
        uint32_t ln = r.u32();
        if (!r.ok || !r.need((size_t)ln * 8)) break;
        for (uint32_t i = 0; i < ln; i++) {
            uint64_t hi = r.u32();
            lst.push_back((hi << 32) | r.u32());
        }
    }
    uint32_t nkeys = r.u32();
    if (!r.ok || nkeys > 64) {
        set_error("g16_verify: malformed verifying key");
        return G16_E_PARSE;
    }
    std::vector<std::pair<G2Affine, G2Affine>> ckeys(nkeys);   // (G, GSigmaNeg)
    for (auto& k : ckeys) {
        r.g2(&k.first);
        r.g2(&k.second);
    }
    if (!r.ok || r.off != vk_len || nsets != nkeys) {
        set_error("g16_verify: malformed verifying key");
        return G16_E_PARSE;
    }
    // ---- public witness (witness.WriteTo, public part) --------------------------------------------------
    Reader w{pw, pw_len};
    uint32_t npub = w.u32(), nsec = w.u32(), nvec = w.u32();
    if (!w.ok || nsec != 0 || nvec != npub || pw_len != 12 + 32 * (size_t)npub) {
        set_error("g16_verify: malformed public witness");
        return G16_E_PARSE;
    }
    // ---- proof (Proof.WriteRawTo) -----------------------------------------------------------------------------
    Reader pr{proof, proof_len};
    G1Affine ar, krs, pok;
    G2Affine bs;
    pr.g1(&ar); pr.g2(&bs); pr.g1(&krs);
    uint32_t ncom = pr.u32();
    if (!pr.ok || ncom > 64 || proof_len != 324 + 64 * (size_t)ncom) {
        set_error("g16_verify: malformed proof");
        return G16_E_PARSE;
    }
    std::vector<G1Affine> commitments(ncom);
    for (auto& c : commitments) pr.g1(&c);
    pr.g1(&pok);
    if (!pr.ok) {
        set_error("g16_verify: malformed proof");
        return G16_E_PARSE;
    }
    // A well-formed but wrong proof is "not ok", not an error (sunspot verify exits non-zero either way).
    if (ncom != nkeys || nk != 1 + npub + ncom) return G16_OK;
    std::vector<std::vector<uint32_t>> pub(npub, std::vector<uint32_t>(8));
    for (uint32_t i = 0; i < npub; i++)
        if (!fr_limbs_from_be(pw + 12 + 32 * i, pub[i].data())) return G16_OK;   // gnark rejects values >= r
    for (const G1Affine* p : {&ar, &krs, &pok})
        if (!g1_on_curve(*p)) return G16_OK;
    for (auto& c : commitments)
        if (!g1_on_curve(c)) return G16_OK;
    if (!g2_on_curve(bs) || !g2_in_subgroup(bs)) return G16_OK;
    // vk_x = K_0 + sum pub_i K_{1+i} + sum_j (chal_j K_{1+npub+j} + C_j)
    G1XYZZ vkx = G1XYZZ::from_affine(K[0]);
    for (uint32_t i = 0; i < npub; i++) vkx.add(scalar_mul(K[1 + i], pub[i].data()));
    for (uint32_t j = 0; j < ncom; j++) {
        std::vector<uint8_t> msg(64 + 32 * pacc[j].size());
        G1Affine cj = commitments[j];
        if (!cj.is_inf()) cj = {cj.x.from_mont(), cj.y.from_mont()};
        g1_to_be(cj, msg.data());
        for (size_t k = 0; k < pacc[j].size(); k++) {
            // indices into the full public vector [1, pub...]
            uint64_t idx = pacc[j][k];
            if (idx > npub) return G16_OK;
            if (idx == 0) {
                memset(msg.data() + 64 + 32 * k, 0, 32);
                msg[64 + 32 * k + 31] = 1;
            } else {
                memcpy(msg.data() + 64 + 32 * k, pw + 12 + 32 * (idx - 1), 32);
            }
        }
        HFr chal = hash_to_fr(msg.data(), msg.size(), "bsb22-commitment");
        uint64_t c64[4];
        chal.canonical(c64);
        uint32_t c32[8];
        for (int i = 0; i < 4; i++) {
            c32[2 * i] = (uint32_t)c64[i];
            c32[2 * i + 1] = (uint32_t)(c64[i] >> 32);
        }
        vkx.add(scalar_mul(K[1 + npub + j], c32));
        vkx.madd(commitments[j]);
    }
    if (ncom) {
        // folded Pedersen proof of knowledge: e(sum rho^i C_i, GSigmaNeg) e(PoK, G) = 1, rho from DST "G16-BSB22"
        G1XYZZ folded = G1XYZZ::inf();
        if (ncom == 1) {
            folded = G1XYZZ::from_affine(commitments[0]);
        } else {
            std::vector<uint8_t> ser(64 * ncom);
            for (uint32_t j = 0; j < ncom; j++) {
                G1Affine cj = commitments[j];
                if (!cj.is_inf()) cj = {cj.x.from_mont(), cj.y.from_mont()};
                g1_to_be(cj, ser.data() + 64 * j);
            }
            HFr rho = hash_to_fr(ser.data(), ser.size(), "G16-BSB22"), coef = HFr::one();
            for (uint32_t j = 0; j < ncom; j++) {
                uint64_t c64[4];
                coef.canonical(c64);
                uint32_t c32[8];
                for (int i = 0; i < 4; i++) {
                    c32[2 * i] = (uint32_t)c64[i];
                    c32[2 * i + 1] = (uint32_t)(c64[i] >> 32);
                }
                folded.add(scalar_mul(commitments[j], c32));
                coef = coef * rho;
            }
        }
        if (!pairing_product_is_one({{folded.to_affine(), ckeys[0].second}, {pok, ckeys[0].first}})) return G16_OK;
    }
    G1Affine vkx_a = vkx.to_affine();
    bool good = pairing_product_is_one({{ar, bs},
                                        {alpha1.neg(), beta2},
                                        {vkx_a.neg(), gamma2},
                                        {krs.neg(), delta2}});
    *ok = good ? 1 : 0;
    return G16_OK;
}

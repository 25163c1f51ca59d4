// synth_ccs.cpp -- native generator of large synthetic constraint systems in gnark's v0.14 `.ccs`
// container (the format of /root/reference/noir_circuit/target/shielded_pool_verifier.ccs, SURVEY.md
// 8c-fmt).  BASELINE.json configs[3] asks for a "synthetic R1CS of 2^22 constraints": the Python writer
// in synth.py (used for the 26 K-row audit stand-in) would need minutes and gigabytes for four million
// rows, so the same container is produced here in a second.
//
// Shape: wire 0 = 1, n_public public and n_secret secret inputs; row k defines a
// fresh internal wire   (w_p) * (sum_{j<5} b_j w_qj) = w_new   with p, q drawn from the wires of earlier
// levels, so any input assignment extends to a satisfying witness.  No hints, no commitment.
#include <string.h>

#include <string>
#include <vector>

#include "common.cuh"
#include "hostfr.hpp"

namespace g16 {
namespace {

struct SplitMix {
    uint64_t s;
    uint64_t next() {
        uint64_t z = (s += 0x9e3779b97f4a7c15ull);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
        return z ^ (z >> 31);
    }
    uint32_t below(uint32_t n) { return (uint32_t)(((next() >> 32) * (uint64_t)n) >> 32); }
};

void put_u64le(std::vector<uint8_t>& o, uint64_t v) {
    for (int i = 0; i < 8; i++) o.push_back((uint8_t)(v >> (8 * i)));
}
void put_u32le(std::vector<uint8_t>& o, uint32_t v) {
    for (int i = 0; i < 4; i++) o.push_back((uint8_t)(v >> (8 * i)));
}

// C32 / C64 stream holding only the varbyte section (monotone values): SURVEY.md 8c-fmt
template <class W>
void varbyte_stream(std::vector<uint8_t>& o, const std::vector<uint64_t>& values) {
    std::vector<uint8_t> raw;
    uint64_t prev = 0;
    for (uint64_t v : values) {
        uint64_t d = v - prev;
        prev = v;
        for (;;) {
            uint8_t b = d & 0x7f;
            d >>= 7;
            if (d) raw.push_back(b | 0x80);
            else {
                raw.push_back(b);
                break;
            }
        }
    }
    const size_t wb = sizeof(W);
    while (raw.size() % wb) raw.push_back(0x80);
    std::vector<W> data(raw.size() / wb);
    for (size_t i = 0; i < data.size(); i++) {   // bytes big-endian within each word
        W w = 0;
        for (size_t k = 0; k < wb; k++) w = (W)((w << 8) | raw[i * wb + k]);
        data[i] = w;
    }
    std::vector<W> tail;
    if (wb == 4) {
        tail.push_back((W)values.size());
        tail.push_back((W)(data.size() + 2));
    } else {
        tail.push_back((W)((uint64_t)values.size() | ((uint64_t)(data.size() + 1) << 32)));
    }
    tail.insert(tail.end(), data.begin(), data.end());
    const size_t nwords = tail.size() + 1;
    put_u64le(o, nwords);
    for (W w : tail) {
        if (wb == 4) put_u32le(o, (uint32_t)w);
        else put_u64le(o, (uint64_t)w);
    }
    if (wb == 4) put_u32le(o, (uint32_t)tail.size());
    else put_u64le(o, (uint64_t)tail.size());
}

// bit-packed C32 stream of a constant value (the blueprint ids: all GenericR1C = 1).  Blocks of 128 values =
// 4 sub-blocks of 32 deltas; only the very first delta (0 -> value) is non-zero, every other sub-block has
// bit width 0 and carries no words.
void constant_stream32(std::vector<uint8_t>& o, size_t count, uint32_t value) {
    const size_t padded = (count + 127) / 128 * 128;
    std::vector<uint32_t> body;
    for (size_t blk = 0; blk < padded; blk += 128) {
        const uint32_t w0 = (blk == 0 && value) ? 32 - (uint32_t)__builtin_clz(value) : 0;
        body.push_back(w0 << 24);                       // descriptors of the 4 sub-blocks, MSB first
        for (uint32_t k = 0; k < w0; k++) body.push_back(k == 0 ? value : 0);   // 32 deltas x w0 bits, LSB first
    }
    std::vector<uint32_t> words;
    words.push_back((uint32_t)padded);
    words.push_back((uint32_t)(body.size() + 3));
    words.push_back(0);                                 // initial value
    words.insert(words.end(), body.begin(), body.end());
    words.push_back(0);                                 // empty varbyte tail
    put_u64le(o, words.size());
    for (uint32_t w : words) put_u32le(o, w);
}

struct Cbor {
    std::vector<uint8_t> o;
    void head(int major, uint64_t v) {
        if (v < 24) o.push_back((uint8_t)(major << 5 | v));
        else if (v < 256) { o.push_back((uint8_t)(major << 5 | 24)); o.push_back((uint8_t)v); }
        else if (v < 65536) { o.push_back((uint8_t)(major << 5 | 25)); o.push_back((uint8_t)(v >> 8)); o.push_back((uint8_t)v); }
        else if (v < (1ull << 32)) { o.push_back((uint8_t)(major << 5 | 26)); for (int i = 3; i >= 0; i--) o.push_back((uint8_t)(v >> (8 * i))); }
        else { o.push_back((uint8_t)(major << 5 | 27)); for (int i = 7; i >= 0; i--) o.push_back((uint8_t)(v >> (8 * i))); }
    }
    void uint(uint64_t v) { head(0, v); }
    void text(const std::string& s) { head(3, s.size()); o.insert(o.end(), s.begin(), s.end()); }
    void array(uint64_t n) { head(4, n); }
    void map(uint64_t n) { head(5, n); }
    void tag(uint64_t t) { head(6, t); }
    void null() { o.push_back(0xf6); }
};

}  // namespace

int synth_ccs(uint64_t n_constraints, uint32_t n_public, uint32_t n_secret, uint64_t seed, std::vector<uint8_t>* out) {
    if (n_constraints == 0 || n_constraints > (1ull << 26) || n_public < 1 || n_secret < 1) {
        set_error("g16_synth_ccs: bad sizes");
        return G16_E_ARG;
    }
    SplitMix rng{seed};
    const uint32_t n_coeff_rand = 256;
    std::vector<HFr> coeffs;
    coeffs.push_back(HFr::zero());
    coeffs.push_back(HFr::one());
    coeffs.push_back(HFr::from_u64(2));
    coeffs.push_back(HFr::one().neg());
    coeffs.push_back(HFr::from_u64(2).neg());
    for (uint32_t i = 0; i < n_coeff_rand; i++) {
        HFr x{{rng.next(), rng.next(), rng.next(), rng.next() >> 3}};   // < 2^253 < r
        if (x.is_zero()) x = HFr::one();
        coeffs.push_back(x);   // the value is used as the Montgomery representative directly
    }
    const uint32_t first_internal = 1 + n_public + n_secret;
    const uint64_t LEVEL = 65536;
    const uint64_t nlevels = (n_constraints + LEVEL - 1) / LEVEL;
    // ---- levels ----------------------------------------------------------------------------------
    std::vector<uint8_t> lv;
    put_u64le(lv, nlevels);
    {
        std::vector<uint64_t> idx;
        for (uint64_t l = 0; l < nlevels; l++) {
            idx.clear();
            for (uint64_t k = l * LEVEL; k < std::min(n_constraints, (l + 1) * LEVEL); k++) idx.push_back(k);
            varbyte_stream<uint32_t>(lv, idx);
        }
    }
    // ---- calldata + instruction streams --------------------------------------------------------------
    std::vector<uint8_t> cd;
    std::vector<uint64_t> coff(n_constraints), woff(n_constraints), start(n_constraints);
    const uint32_t words_per_row = 4 + 2 * (1 + 5 + 1);
    put_u64le(cd, n_constraints * words_per_row);
    auto uvar = [&](uint32_t v) {
        for (;;) {
            uint8_t b = v & 0x7f;
            v >>= 7;
            if (v) cd.push_back(b | 0x80);
            else {
                cd.push_back(b);
                break;
            }
        }
    };
    for (uint64_t k = 0; k < n_constraints; k++) {
        const uint32_t known = first_internal + (uint32_t)(k / LEVEL * LEVEL);   // wires of earlier levels + inputs
        const uint32_t fresh = first_internal + (uint32_t)k;
        coff[k] = k;
        woff[k] = fresh;
        start[k] = k * words_per_row;
        uvar(words_per_row); uvar(1); uvar(5); uvar(1);
        uvar(1); uvar(rng.below(known));
        for (int j = 0; j < 5; j++) {
            uvar(5 + rng.below(n_coeff_rand));
            uvar(rng.below(known));
        }
        uvar(1); uvar(fresh);
    }
    std::vector<uint8_t> ins;
    constant_stream32(ins, n_constraints, 1);
    varbyte_stream<uint32_t>(ins, coff);
    varbyte_stream<uint32_t>(ins, woff);
    varbyte_stream<uint64_t>(ins, start);
    // ---- body ---------------------------------------------------------------------------------------------
    Cbor b;
    b.map(15);
    b.text("Type"); b.uint(1);
    b.text("Public"); b.array(1 + n_public); b.text("1");
    for (uint32_t i = 0; i < n_public; i++) b.text("pub_" + std::to_string(i));
    b.text("Secret"); b.array(n_secret);
    for (uint32_t i = 0; i < n_secret; i++) b.text("__witness_" + std::to_string(n_public + i));
    b.text("NbInternalVariables"); b.uint(n_constraints);
    b.text("NbConstraints"); b.uint(n_constraints);
    b.text("ScalarField"); b.text("30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001");
    b.text("GnarkVersion"); b.text("0.14.0");
    b.text("Blueprints"); b.array(2); b.tag(5309735); b.map(0); b.tag(5309736); b.map(0);
    b.text("CommitmentInfo"); b.tag(5309737); b.array(0);
    b.text("MHintsDependencies"); b.map(0);
    b.text("GkrInfo"); b.null();
    b.text("Logs"); b.array(0);
    b.text("DebugInfo"); b.array(0);
    b.text("MDebug"); b.map(0);
    b.text("SymbolTable"); b.null();
    // ---- coefficient table ------------------------------------------------------------------------------
    std::vector<uint8_t> co;
    put_u64le(co, coeffs.size());
    for (auto& x : coeffs)
        for (int i = 0; i < 4; i++) put_u64le(co, x.l[i]);
    // ---- container ----------------------------------------------------------------------------------------
    std::vector<uint8_t>& o = *out;
    o.clear();
    const uint64_t payload = 32 + lv.size() + ins.size() + cd.size() + b.o.size() + co.size();
    o.reserve(32 + payload);
    put_u64le(o, payload); put_u64le(o, 0); put_u64le(o, 14); put_u64le(o, 0);
    put_u64le(o, lv.size()); put_u64le(o, ins.size()); put_u64le(o, cd.size()); put_u64le(o, b.o.size());
    o.insert(o.end(), lv.begin(), lv.end());
    o.insert(o.end(), ins.begin(), ins.end());
    o.insert(o.end(), cd.begin(), cd.end());
    o.insert(o.end(), b.o.begin(), b.o.end());
    o.insert(o.end(), co.begin(), co.end());
    return G16_OK;
}

}  // namespace g16

// Two-call pattern: out == NULL generates, caches and returns the size in *out_len; the second call
// (same parameters) copies the cached container out.
extern "C" int g16_synth_ccs(uint64_t n_constraints, uint32_t n_public, uint32_t n_secret, uint64_t seed, uint8_t* out,
                             size_t* out_len) {
    static thread_local std::vector<uint8_t> cache;
    static thread_local uint64_t key[4] = {0, 0, 0, 0};
    if (!out_len) return G16_E_ARG;
    const uint64_t k[4] = {n_constraints, n_public, n_secret, seed};
    if (cache.empty() || memcmp(k, key, sizeof k) != 0) {
        int rc = g16::synth_ccs(n_constraints, n_public, n_secret, seed, &cache);
        if (rc != G16_OK) return rc;
        memcpy(key, k, sizeof k);
    }
    if (!out) {
        *out_len = cache.size();
        return G16_OK;
    }
    if (*out_len < cache.size()) {
        g16::set_error("g16_synth_ccs: output buffer too small");
        return G16_E_ARG;
    }
    memcpy(out, cache.data(), cache.size());
    *out_len = cache.size();
    cache.clear();
    cache.shrink_to_fit();
    return G16_OK;
}

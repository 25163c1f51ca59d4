// hostutil.hpp -- SHA-256, RFC 9380 expand_message_xmd and gnark-crypto's fr hash-to-field, host
// side.  Used for the BSB22 commitment challenge (DST "bsb22-commitment") that gnark's prover
// computes in the middle of the witness solve.
#pragma once
#include <stdint.h>
#include <string.h>

#include <string>
#include <vector>

#include "hostfr.hpp"

namespace g16 {

struct Sha256 {
    uint32_t h[8];
    uint8_t buf[64];
    uint64_t len = 0;
    size_t fill = 0;
    Sha256() {
        static const uint32_t iv[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a,
                                       0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
        memcpy(h, iv, sizeof h);
    }
    static uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
    void block(const uint8_t* p) {
        static const uint32_t K[64] = {
            0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
            0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
            0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
            0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
            0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
            0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
            0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
            0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
        uint32_t w[64];
        for (int i = 0; i < 16; i++)
            w[i] = ((uint32_t)p[4 * i] << 24) | ((uint32_t)p[4 * i + 1] << 16) | ((uint32_t)p[4 * i + 2] << 8) | p[4 * i + 3];
        for (int i = 16; i < 64; i++) {
            uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
            uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
            w[i] = w[i - 16] + s0 + w[i - 7] + s1;
        }
        uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
        for (int i = 0; i < 64; i++) {
            uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25);
            uint32_t ch = (e & f) ^ (~e & g);
            uint32_t t1 = hh + S1 + ch + K[i] + w[i];
            uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22);
            uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
            uint32_t t2 = S0 + mj;
            hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        }
        h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
    }
    void update(const uint8_t* p, size_t n) {
        len += n;
        while (n) {
            size_t take = 64 - fill < n ? 64 - fill : n;
            memcpy(buf + fill, p, take);
            fill += take; p += take; n -= take;
            if (fill == 64) { block(buf); fill = 0; }
        }
    }
    void finish(uint8_t out[32]) {
        uint64_t bits = len * 8;
        uint8_t pad = 0x80;
        update(&pad, 1);
        uint8_t z = 0;
        while (fill != 56) update(&z, 1);
        uint8_t lb[8];
        for (int i = 0; i < 8; i++) lb[i] = (uint8_t)(bits >> (56 - 8 * i));
        update(lb, 8);
        for (int i = 0; i < 8; i++) {
            out[4 * i] = h[i] >> 24; out[4 * i + 1] = h[i] >> 16; out[4 * i + 2] = h[i] >> 8; out[4 * i + 3] = h[i];
        }
    }
};

inline void expand_message_xmd(const uint8_t* msg, size_t msg_len, const std::string& dst, uint8_t* out, size_t out_len) {
    size_t ell = (out_len + 31) / 32;
    std::vector<uint8_t> dst_prime(dst.begin(), dst.end());
    dst_prime.push_back((uint8_t)dst.size());
    uint8_t b0[32], bi[32];
    {
        Sha256 s;
        uint8_t zpad[64] = {0};
        s.update(zpad, 64);
        s.update(msg, msg_len);
        uint8_t l2[3] = {(uint8_t)(out_len >> 8), (uint8_t)out_len, 0};
        s.update(l2, 3);
        s.update(dst_prime.data(), dst_prime.size());
        s.finish(b0);
    }
    {
        Sha256 s;
        s.update(b0, 32);
        uint8_t one = 1;
        s.update(&one, 1);
        s.update(dst_prime.data(), dst_prime.size());
        s.finish(bi);
    }
    size_t off = 0;
    for (size_t i = 1; i <= ell; i++) {
        size_t take = out_len - off < 32 ? out_len - off : 32;
        memcpy(out + off, bi, take);
        off += take;
        if (i == ell) break;
        uint8_t x[32];
        for (int k = 0; k < 32; k++) x[k] = b0[k] ^ bi[k];
        Sha256 s;
        s.update(x, 32);
        uint8_t idx = (uint8_t)(i + 1);
        s.update(&idx, 1);
        s.update(dst_prime.data(), dst_prime.size());
        s.finish(bi);
    }
}

// gnark-crypto fr.Hash(msg, dst, 1)[0]: 48 uniform bytes, big-endian integer, reduced mod r
inline HFr hash_to_fr(const uint8_t* msg, size_t msg_len, const std::string& dst) {
    uint8_t raw[48];
    expand_message_xmd(msg, msg_len, dst, raw, 48);
    // value = hi(16 bytes) * 2^256 + lo(32 bytes)
    uint8_t hi_be[32] = {0};
    memcpy(hi_be + 16, raw, 16);
    HFr hi = HFr::from_be(hi_be);          // < 2^128, Montgomery
    HFr lo = HFr::from_be(raw + 16);       // reduced mod r, Montgomery
    // 2^256 mod r in Montgomery form is R2 * R^-1 ... simply: to_mont(x) = x * 2^256, so
    // hi * 2^256 (as a field element, Montgomery form) = to_mont(hi_mont)
    return hi.to_mont() + lo;
}

}  // namespace g16

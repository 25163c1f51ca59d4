// hostfr.hpp -- host-side BN254 Fr on 4x64-bit limbs (Montgomery, R = 2^256) for the witness
// solver.  Same in-memory bytes as the device `Fr` (8x32 little-endian limbs) and as the `.ccs`
// coefficient table, so solver output is copied to the GPU without conversion.
//
// Replaces the `fr.Element` arithmetic gnark's solver uses (constraint/bn254/solver.go, third-
// party).
#pragma once
#include <stdint.h>
#include <string.h>

namespace g16 {

struct HFr {
    uint64_t l[4];

    static constexpr uint64_t M[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull,
                                      0x30644e72e131a029ull};
    static constexpr uint64_t INV = 0xc2e1f593efffffffull;  // -r^{-1} mod 2^64
    static constexpr uint64_t ONE[4] = {0xac96341c4ffffffbull, 0x36fc76959f60cd29ull, 0x666ea36f7879462eull,
                                        0x0e0a77c19a07df2full};
    static constexpr uint64_t R2[4] = {0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull, 0x8c49833d53bb8085ull,
                                       0x0216d0b17f4e44a5ull};

    static HFr zero() { return HFr{{0, 0, 0, 0}}; }
    static HFr one() { return HFr{{ONE[0], ONE[1], ONE[2], ONE[3]}}; }
    bool is_zero() const { return (l[0] | l[1] | l[2] | l[3]) == 0; }
    bool operator==(const HFr& o) const { return l[0] == o.l[0] && l[1] == o.l[1] && l[2] == o.l[2] && l[3] == o.l[3]; }
    bool operator!=(const HFr& o) const { return !(*this == o); }

    static bool geq_mod(const uint64_t* a) {
        for (int i = 3; i >= 0; i--) {
            if (a[i] > M[i]) return true;
            if (a[i] < M[i]) return false;
        }
        return true;
    }
    static void sub_mod_inplace(uint64_t* a) {
        unsigned __int128 b = 0;
        for (int i = 0; i < 4; i++) {
            unsigned __int128 t = (unsigned __int128)a[i] - M[i] - (uint64_t)b;
            a[i] = (uint64_t)t;
            b = (t >> 64) & 1;
        }
    }
    friend HFr operator+(const HFr& a, const HFr& b) {
        HFr r;
        unsigned __int128 c = 0;
        for (int i = 0; i < 4; i++) {
            c += (unsigned __int128)a.l[i] + b.l[i];
            r.l[i] = (uint64_t)c;
            c >>= 64;
        }
        if (c || geq_mod(r.l)) sub_mod_inplace(r.l);
        return r;
    }
    friend HFr operator-(const HFr& a, const HFr& b) {
        HFr r;
        unsigned __int128 bw = 0;
        for (int i = 0; i < 4; i++) {
            unsigned __int128 t = (unsigned __int128)a.l[i] - b.l[i] - (uint64_t)bw;
            r.l[i] = (uint64_t)t;
            bw = (t >> 64) & 1;
        }
        if (bw) {
            unsigned __int128 c = 0;
            for (int i = 0; i < 4; i++) {
                c += (unsigned __int128)r.l[i] + M[i];
                r.l[i] = (uint64_t)c;
                c >>= 64;
            }
        }
        return r;
    }
    HFr neg() const { return is_zero() ? *this : zero() - *this; }

    // CIOS Montgomery product
    friend HFr operator*(const HFr& a, const HFr& b) {
        uint64_t t[6] = {0, 0, 0, 0, 0, 0};
        for (int i = 0; i < 4; i++) {
            unsigned __int128 c = 0;
            for (int j = 0; j < 4; j++) {
                c += (unsigned __int128)a.l[j] * b.l[i] + t[j];
                t[j] = (uint64_t)c;
                c >>= 64;
            }
            c += t[4];
            t[4] = (uint64_t)c;
            t[5] = (uint64_t)(c >> 64);
            uint64_t m = t[0] * INV;
            c = (unsigned __int128)m * M[0] + t[0];
            c >>= 64;
            for (int j = 1; j < 4; j++) {
                c += (unsigned __int128)m * M[j] + t[j];
                t[j - 1] = (uint64_t)c;
                c >>= 64;
            }
            c += t[4];
            t[3] = (uint64_t)c;
            t[4] = t[5] + (uint64_t)(c >> 64);
        }
        HFr r{{t[0], t[1], t[2], t[3]}};
        if (t[4] || geq_mod(r.l)) sub_mod_inplace(r.l);
        return r;
    }
    HFr sqr() const { return *this * *this; }

    HFr to_mont() const { return *this * HFr{{R2[0], R2[1], R2[2], R2[3]}}; }
    HFr from_mont() const { return *this * HFr{{1, 0, 0, 0}}; }

    // x^(r-2); 4-bit fixed window (0 -> 0)
    HFr inverse() const {
        HFr tab[16];
        tab[0] = one();
        for (int i = 1; i < 16; i++) tab[i] = tab[i - 1] * *this;
        uint64_t e[4] = {M[0] - 2, M[1], M[2], M[3]};
        HFr acc = one();
        for (int nib = 63; nib >= 0; nib--) {
            acc = acc.sqr().sqr().sqr().sqr();
            unsigned d = (unsigned)((e[nib >> 4] >> ((nib & 15) * 4)) & 15);
            if (d) acc = acc * tab[d];
        }
        return acc;
    }

    // canonical big-endian 32 bytes <-> Montgomery
    static HFr from_be(const uint8_t* be) {
        HFr x;
        for (int i = 0; i < 4; i++) {
            uint64_t v = 0;
            for (int k = 0; k < 8; k++) v = (v << 8) | be[(3 - i) * 8 + k];
            x.l[i] = v;
        }
        if (geq_mod(x.l)) {  // reduce values >= r (at most a few subtractions for 256-bit input)
            while (geq_mod(x.l)) sub_mod_inplace(x.l);
        }
        return x.to_mont();
    }
    void to_be(uint8_t* be) const {
        HFr c = from_mont();
        for (int i = 0; i < 4; i++)
            for (int k = 0; k < 8; k++) be[(3 - i) * 8 + k] = (uint8_t)(c.l[i] >> (56 - 8 * k));
    }
    static HFr from_u64(uint64_t v) { return HFr{{v, 0, 0, 0}}.to_mont(); }
    // canonical value as 4 limbs
    void canonical(uint64_t out[4]) const {
        HFr c = from_mont();
        memcpy(out, c.l, 32);
    }
};

}  // namespace g16

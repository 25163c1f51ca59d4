// hostbig.hpp -- a small signed big integer for the solver hints that work over the INTEGERS rather than in Fr
// (gnark std/math/emulated mulHint, sunspot sw-grumpkin decomposeScalar: SURVEY.md 9.6; restated in
// oracle/py/groth16.py `_hint_emulated_mul`, `glv_split_nonneg`).  Operands are a few hundred bits and the hints
// run a handful of times per proof, so clarity wins over speed: schoolbook product, shift-subtract division.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <vector>

#include "hostfr.hpp"

namespace g16 {

struct BigInt {
    std::vector<uint64_t> m;   // magnitude, little-endian, no leading zero limbs
    bool neg = false;

    BigInt() {}
    BigInt(uint64_t v) {
        if (v) m.push_back(v);
    }
    static BigInt from_fr(const HFr& x) {   // canonical representative in [0, r)
        uint64_t c[4];
        x.canonical(c);
        BigInt b;
        b.m.assign(c, c + 4);
        b.trim();
        return b;
    }
    static BigInt from_decimal(const char* s) {
        BigInt b;
        for (; *s; s++) b = b * BigInt(10) + BigInt((uint64_t)(*s - '0'));
        return b;
    }
    void trim() {
        while (!m.empty() && m.back() == 0) m.pop_back();
        if (m.empty()) neg = false;
    }
    bool is_zero() const { return m.empty(); }
    size_t bits() const {
        if (m.empty()) return 0;
        return 64 * (m.size() - 1) + (64 - (size_t)__builtin_clzll(m.back()));
    }
    bool bit(size_t i) const { return i / 64 < m.size() && ((m[i / 64] >> (i % 64)) & 1); }

    static int cmp_mag(const BigInt& a, const BigInt& b) {
        if (a.m.size() != b.m.size()) return a.m.size() < b.m.size() ? -1 : 1;
        for (size_t i = a.m.size(); i-- > 0;)
            if (a.m[i] != b.m[i]) return a.m[i] < b.m[i] ? -1 : 1;
        return 0;
    }
    static BigInt add_mag(const BigInt& a, const BigInt& b) {
        BigInt r;
        size_t n = std::max(a.m.size(), b.m.size());
        r.m.resize(n + 1);
        unsigned __int128 c = 0;
        for (size_t i = 0; i < n; i++) {
            c += (i < a.m.size() ? a.m[i] : 0);
            c += (i < b.m.size() ? b.m[i] : 0);
            r.m[i] = (uint64_t)c;
            c >>= 64;
        }
        r.m[n] = (uint64_t)c;
        r.trim();
        return r;
    }
    static BigInt sub_mag(const BigInt& a, const BigInt& b) {   // |a| >= |b|
        BigInt r;
        r.m.resize(a.m.size());
        uint64_t bw = 0;
        for (size_t i = 0; i < a.m.size(); i++) {
            unsigned __int128 t = (unsigned __int128)a.m[i] - (i < b.m.size() ? b.m[i] : 0) - bw;
            r.m[i] = (uint64_t)t;
            bw = (uint64_t)(t >> 64) & 1;
        }
        r.trim();
        return r;
    }
    friend BigInt operator+(const BigInt& a, const BigInt& b) {
        BigInt r;
        if (a.neg == b.neg) {
            r = add_mag(a, b);
            r.neg = a.neg;
        } else if (cmp_mag(a, b) >= 0) {
            r = sub_mag(a, b);
            r.neg = a.neg;
        } else {
            r = sub_mag(b, a);
            r.neg = b.neg;
        }
        r.trim();
        return r;
    }
    BigInt operator-() const {
        BigInt r = *this;
        if (!r.m.empty()) r.neg = !r.neg;
        return r;
    }
    friend BigInt operator-(const BigInt& a, const BigInt& b) { return a + (-b); }
    friend BigInt operator*(const BigInt& a, const BigInt& b) {
        BigInt r;
        if (a.m.empty() || b.m.empty()) return r;
        r.m.assign(a.m.size() + b.m.size(), 0);
        for (size_t i = 0; i < a.m.size(); i++) {
            unsigned __int128 c = 0;
            for (size_t j = 0; j < b.m.size(); j++) {
                c += (unsigned __int128)a.m[i] * b.m[j] + r.m[i + j];
                r.m[i + j] = (uint64_t)c;
                c >>= 64;
            }
            r.m[i + b.m.size()] = (uint64_t)c;
        }
        r.neg = a.neg != b.neg;
        r.trim();
        return r;
    }
    friend bool operator<(const BigInt& a, const BigInt& b) {
        if (a.neg != b.neg) return a.neg;
        int c = cmp_mag(a, b);
        return a.neg ? c > 0 : c < 0;
    }
    friend bool operator==(const BigInt& a, const BigInt& b) { return a.neg == b.neg && a.m == b.m; }

    BigInt shl(size_t k) const {
        BigInt r;
        if (m.empty()) return r;
        size_t w = k / 64, s = k % 64;
        r.m.assign(m.size() + w + 1, 0);
        for (size_t i = 0; i < m.size(); i++) {
            r.m[i + w] |= m[i] << s;
            if (s) r.m[i + w + 1] |= m[i] >> (64 - s);
        }
        r.neg = neg;
        r.trim();
        return r;
    }
    // magnitude shift (callers use it on exact multiples of 2^k or on non-negative values)
    BigInt shr_mag(size_t k) const {
        BigInt r;
        size_t w = k / 64, s = k % 64;
        if (w >= m.size()) return r;
        r.m.assign(m.size() - w, 0);
        for (size_t i = w; i < m.size(); i++) {
            r.m[i - w] = m[i] >> s;
            if (s && i + 1 < m.size()) r.m[i - w] |= m[i + 1] << (64 - s);
        }
        r.neg = neg;
        r.trim();
        return r;
    }
    // floor(x / 2^k), like Go's big.Int.Rsh on negative values
    BigInt shr_floor(size_t k) const {
        BigInt q = shr_mag(k);
        if (neg) {
            bool inexact = false;
            for (size_t i = 0; i < k && !inexact; i++) inexact = bit(i);
            if (inexact) q = q - BigInt(1);
        }
        return q;
    }
    BigInt low_bits(size_t k) const {   // |x| mod 2^k
        BigInt r;
        for (size_t i = 0; i < m.size() && 64 * i < k; i++) {
            uint64_t v = m[i];
            if (k - 64 * i < 64) v &= (1ull << (k - 64 * i)) - 1;
            r.m.push_back(v);
        }
        r.trim();
        return r;
    }
    // non-negative a, positive b: a = q*b + r
    static void divmod_mag(const BigInt& a, const BigInt& b, BigInt* q, BigInt* r) {
        *q = BigInt();
        *r = BigInt();
        q->m.assign(a.m.size(), 0);
        for (size_t i = a.bits(); i-- > 0;) {
            *r = r->shl(1);
            if (a.bit(i)) *r = *r + BigInt(1);
            if (cmp_mag(*r, b) >= 0) {
                *r = sub_mag(*r, b);
                q->m[i / 64] |= 1ull << (i % 64);
            }
        }
        q->trim();
    }
    // floor division (Python //) by a positive divisor
    static BigInt floordiv(const BigInt& a, const BigInt& b) {
        BigInt q, r, am = a;
        am.neg = false;
        divmod_mag(am, b, &q, &r);
        if (a.neg) {
            q = -q;
            if (!r.is_zero()) q = q - BigInt(1);
        }
        return q;
    }
    // x mod r as a Montgomery field element (negative values wrap)
    HFr to_fr() const {
        BigInt mod;
        mod.m.assign(HFr::M, HFr::M + 4);
        BigInt q, r, am = *this;
        am.neg = false;
        if (cmp_mag(am, mod) < 0) r = am;   // the usual case (limbs, carries): no division
        else divmod_mag(am, mod, &q, &r);
        uint64_t c[4] = {0, 0, 0, 0};
        for (size_t i = 0; i < r.m.size() && i < 4; i++) c[i] = r.m[i];
        HFr x = HFr{{c[0], c[1], c[2], c[3]}}.to_mont();
        return neg ? x.neg() : x;
    }
};

}  // namespace g16

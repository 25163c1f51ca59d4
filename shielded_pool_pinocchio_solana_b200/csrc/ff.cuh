// ff.cuh -- 254-bit prime-field arithmetic for BN254 (Fp and Fr) on 8x32-bit limbs.
//
// Replaces, on the device, what gnark-crypto's `ecc/bn254/fp` and `ecc/bn254/fr`
// (amd64 ADX/BMI2 assembly; third-party, not in /root/reference -- SURVEY.md 2,
// "Native / CUDA / collective inventory") do for the `sunspot prove` step invoked at
// /root/reference/client/proof.helper.ts:64.
//
// Representation: little-endian uint32 limbs, Montgomery form with R = 2^256 (the
// same radix gnark-crypto uses, so `.ccs` coefficient tables load without conversion),
// always fully reduced to [0, p).
//
// Multiplication is an operand-scanning Montgomery product whose partial products
// are split into two 256-bit accumulators ("lo-aligned" and "hi-aligned", offset by
// one limb) so that every 32x32->64 product lands on a 64-bit boundary of one of
// them.  ptxas then fuses each mad.lo.cc/madc.hi.cc pair into ONE `IMAD.WIDE.U32`
// with carry-in/out predicates: 8 rounds x (16 wide MADs + 1 IMAD) = 136 integer-pipe
// instructions per modmul -- the figure DESIGN.md / SURVEY.md 8(d) use as the
// algorithmic IMAD count.
//
// Every carry chain lives inside a single asm statement (the carry flag is never
// assumed to survive between statements).  The same functions compile for the host
// (portable 64-bit emulation of each chain) so the algorithms are unit-tested on CPU.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
// big curve routines used off the hot loop: real calls keep the non-hot kernels (and nvcc) small
#define FF_NOINLINE __host__ __device__ __noinline__
#define FF_HD __host__ __device__ __forceinline__
#define FF_D __device__ __forceinline__
#else
#define FF_NOINLINE
#define FF_HD inline
#define FF_D inline
#endif

namespace g16 {

struct alignas(16) u256 {
    uint32_t v[8];
};

// ---- modulus parameter packs ------------------------------------------------------
// Limbs are spelled out as immediates so ptxas folds them into the instruction stream
// (no constant-bank loads on the modulus operand of the reduction MADs).
struct FpParams {
    static constexpr uint32_t M0 = 0xd87cfd47u, M1 = 0x3c208c16u, M2 = 0x6871ca8du, M3 = 0x97816a91u,
                              M4 = 0x8181585du, M5 = 0xb85045b6u, M6 = 0xe131a029u, M7 = 0x30644e72u;
    static constexpr uint32_t INV = 0xe4866389u;  // -p^{-1} mod 2^32
    // R mod p  (Montgomery one)
    static constexpr uint32_t ONE_0 = 0xc58f0d9du, ONE_1 = 0xd35d438du, ONE_2 = 0xf5c70b3du, ONE_3 = 0x0a78eb28u, ONE_4 = 0x7879462cu, ONE_5 = 0x666ea36fu, ONE_6 = 0x9a07df2fu, ONE_7 = 0x0e0a77c1u;
    // R^2 mod p
    static constexpr uint32_t R2_0 = 0x538afa89u, R2_1 = 0xf32cfc5bu, R2_2 = 0xd44501fbu, R2_3 = 0xb5e71911u, R2_4 = 0x0a417ff6u, R2_5 = 0x47ab1effu, R2_6 = 0xcab8351fu, R2_7 = 0x06d89f71u;
};

struct FrParams {
    static constexpr uint32_t M0 = 0xf0000001u, M1 = 0x43e1f593u, M2 = 0x79b97091u, M3 = 0x2833e848u,
                              M4 = 0x8181585du, M5 = 0xb85045b6u, M6 = 0xe131a029u, M7 = 0x30644e72u;
    static constexpr uint32_t INV = 0xefffffffu;  // -r^{-1} mod 2^32
    static constexpr uint32_t ONE_0 = 0x4ffffffbu, ONE_1 = 0xac96341cu, ONE_2 = 0x9f60cd29u, ONE_3 = 0x36fc7695u, ONE_4 = 0x7879462eu, ONE_5 = 0x666ea36fu, ONE_6 = 0x9a07df2fu, ONE_7 = 0x0e0a77c1u;
    static constexpr uint32_t R2_0 = 0xae216da7u, R2_1 = 0x1bb8e645u, R2_2 = 0xe35c59e3u, R2_3 = 0x53fe3ab1u, R2_4 = 0x53bb8085u, R2_5 = 0x8c49833du, R2_6 = 0x7f4e44a5u, R2_7 = 0x0216d0b1u;
};

// ---- carry-chain primitives ---------------------------------------------------------
// r[0..7] = sum_k x_k * b * 2^(64k)   (four independent 32x32->64 products)
FF_HD void ff_mul4(uint32_t* r, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t b) {
#ifdef __CUDA_ARCH__
    asm("mul.lo.u32 %0, %8, %12; mul.hi.u32 %1, %8, %12;\n\t"
        "mul.lo.u32 %2, %9, %12; mul.hi.u32 %3, %9, %12;\n\t"
        "mul.lo.u32 %4, %10, %12; mul.hi.u32 %5, %10, %12;\n\t"
        "mul.lo.u32 %6, %11, %12; mul.hi.u32 %7, %11, %12;"
        : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7])
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(b));
#else
    uint32_t x[4] = {x0, x1, x2, x3};
    for (int k = 0; k < 4; k++) {
        uint64_t t = (uint64_t)x[k] * b;
        r[2 * k] = (uint32_t)t;
        r[2 * k + 1] = (uint32_t)(t >> 32);
    }
#endif
}

// r[0..7] += sum_k x_k * b * 2^(64k); returns the carry out of limb 7 (0 or 1)
FF_HD uint32_t ff_mad4(uint32_t* r, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t b) {
    uint32_t c;
#ifdef __CUDA_ARCH__
    asm("mad.lo.cc.u32 %0, %9, %13, %0; madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2; madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4; madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6; madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32 %8, 0, 0;"
        : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
          "=r"(c)
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(b));
#else
    uint32_t x[4] = {x0, x1, x2, x3};
    uint64_t carry = 0;
    for (int k = 0; k < 4; k++) {
        unsigned __int128 t = (unsigned __int128)x[k] * b;
        t += (uint64_t)r[2 * k] | ((uint64_t)r[2 * k + 1] << 32);
        t += carry;
        r[2 * k] = (uint32_t)t;
        r[2 * k + 1] = (uint32_t)(t >> 32);
        carry = (uint64_t)(t >> 64);
    }
    c = (uint32_t)carry;
#endif
    return c;
}

// The "shift" step of a round.  `lo` is the accumulator that is now aligned with
// limb 0, `y` the one whose limb 0 was just cleared by the reduction and must move
// down by two limbs to become the new hi-aligned accumulator:
//     lo[0] += y[1]                      (carry ripples into the new y)
//     y'[2k..2k+1] = x_k*b + y[2k+2..2k+3] + carry      k = 0..2
//     y'[6..7]     = x_3*b + carry
// No carry can leave y'[7] (see the bound in ff_mont_mul).
FF_HD void ff_mad4_shift(uint32_t* y, uint32_t& lo0, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3,
                         uint32_t b) {
#ifdef __CUDA_ARCH__
    asm("add.cc.u32 %8, %8, %1;\n\t"
        "madc.lo.cc.u32 %0, %9, %13, %2; madc.hi.cc.u32 %1, %9, %13, %3;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %4; madc.hi.cc.u32 %3, %10, %13, %5;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %6; madc.hi.cc.u32 %5, %11, %13, %7;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, 0; madc.hi.u32 %7, %12, %13, 0;"
        : "+r"(y[0]), "+r"(y[1]), "+r"(y[2]), "+r"(y[3]), "+r"(y[4]), "+r"(y[5]), "+r"(y[6]), "+r"(y[7]),
          "+r"(lo0)
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(b));
#else
    uint32_t x[4] = {x0, x1, x2, x3};
    uint64_t s = (uint64_t)lo0 + y[1];
    lo0 = (uint32_t)s;
    uint64_t carry = s >> 32;
    for (int k = 0; k < 4; k++) {
        unsigned __int128 t = (unsigned __int128)x[k] * b;
        if (k < 3) t += (uint64_t)y[2 * k + 2] | ((uint64_t)y[2 * k + 3] << 32);
        t += carry;
        y[2 * k] = (uint32_t)t;
        y[2 * k + 1] = (uint32_t)(t >> 32);
        carry = (uint64_t)(t >> 64);
    }
#endif
}

// r = a + b (8 limbs), returns carry
FF_HD uint32_t ff_add8(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    uint32_t c;
#ifdef __CUDA_ARCH__
    asm("add.cc.u32 %0, %9, %17; addc.cc.u32 %1, %10, %18; addc.cc.u32 %2, %11, %19; addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21; addc.cc.u32 %5, %14, %22; addc.cc.u32 %6, %15, %23; addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32 %8, 0, 0;"
        : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7]), "=&r"(c)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b[0]),
          "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
#else
    uint64_t carry = 0;
    for (int i = 0; i < 8; i++) {
        uint64_t t = (uint64_t)a[i] + b[i] + carry;
        r[i] = (uint32_t)t;
        carry = t >> 32;
    }
    c = (uint32_t)carry;
#endif
    return c;
}

// r = a - b (8 limbs), returns borrow (0 or 1)
FF_HD uint32_t ff_sub8(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    uint32_t bw;
#ifdef __CUDA_ARCH__
    asm("sub.cc.u32 %0, %9, %17; subc.cc.u32 %1, %10, %18; subc.cc.u32 %2, %11, %19; subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21; subc.cc.u32 %5, %14, %22; subc.cc.u32 %6, %15, %23; subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7]), "=&r"(bw)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b[0]),
          "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
    bw &= 1u;  // subc.u32 0-0-borrow = 0xffffffff when a borrow is pending
#else
    uint64_t borrow = 0;
    for (int i = 0; i < 8; i++) {
        uint64_t t = (uint64_t)a[i] - b[i] - borrow;
        r[i] = (uint32_t)t;
        borrow = (t >> 32) & 1u;
    }
    bw = (uint32_t)borrow;
#endif
    return bw;
}

// ---- the field ------------------------------------------------------------------------
template <class PM>
struct alignas(16) Field {
    uint32_t v[8];

    FF_HD static void modulus(uint32_t* m) {
        m[0] = PM::M0; m[1] = PM::M1; m[2] = PM::M2; m[3] = PM::M3;
        m[4] = PM::M4; m[5] = PM::M5; m[6] = PM::M6; m[7] = PM::M7;
    }
    FF_HD static Field zero() {
        Field r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = 0;
        return r;
    }
    FF_HD static Field one() {
        Field r;
        r.v[0] = PM::ONE_0; r.v[1] = PM::ONE_1; r.v[2] = PM::ONE_2; r.v[3] = PM::ONE_3;
        r.v[4] = PM::ONE_4; r.v[5] = PM::ONE_5; r.v[6] = PM::ONE_6; r.v[7] = PM::ONE_7;
        return r;
    }
    FF_HD static Field r2() {
        Field r;
        r.v[0] = PM::R2_0; r.v[1] = PM::R2_1; r.v[2] = PM::R2_2; r.v[3] = PM::R2_3;
        r.v[4] = PM::R2_4; r.v[5] = PM::R2_5; r.v[6] = PM::R2_6; r.v[7] = PM::R2_7;
        return r;
    }
    FF_HD bool is_zero() const {
        return (v[0] | v[1] | v[2] | v[3] | v[4] | v[5] | v[6] | v[7]) == 0;
    }
    FF_HD bool operator==(const Field& o) const {
        uint32_t d = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) d |= v[i] ^ o.v[i];
        return d == 0;
    }
    FF_HD bool operator!=(const Field& o) const { return !(*this == o); }

    // r = t - p if t >= p else t   (t < 2p, optional incoming carry bit `hi`)
    FF_HD static Field reduce_once(const uint32_t* t, uint32_t hi) {
        uint32_t m[8], d[8];
        modulus(m);
        uint32_t bw = ff_sub8(d, t, m);
        bool take = (hi != 0) || (bw == 0);
        Field r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = take ? d[i] : t[i];
        return r;
    }

    FF_HD friend Field operator+(const Field& a, const Field& b) {
        uint32_t t[8];
        uint32_t c = ff_add8(t, a.v, b.v);  // p < 2^254, so c is always 0; kept for generality
        return reduce_once(t, c);
    }
    FF_HD friend Field operator-(const Field& a, const Field& b) {
        uint32_t t[8], u[8], m[8];
        uint32_t bw = ff_sub8(t, a.v, b.v);
        modulus(m);
        ff_add8(u, t, m);
        Field r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = bw ? u[i] : t[i];
        return r;
    }
    FF_HD Field neg() const {
        if (is_zero()) return *this;
        uint32_t m[8];
        modulus(m);
        Field r;
        ff_sub8(r.v, m, v);
        return r;
    }
    FF_HD Field dbl() const { return *this + *this; }

    // Montgomery product a*b/R mod p.
    //
    // Invariant at the top of every round: T = lo + (hi << 32) < 2p.  During a round
    // T + a*b_i + m*p < 2^255 + 2^287 < 2^288, every term is non-negative, so the
    // hi-aligned accumulator (limbs 1..8 of T) never overflows and the carry leaving
    // the lo-aligned one is absorbed by its top limb.
    FF_HD friend Field operator*(const Field& a, const Field& b) {
        uint32_t E[8], O[8];
        // round 0 (accumulators start empty)
        ff_mul4(O, a.v[1], a.v[3], a.v[5], a.v[7], b.v[0]);
        ff_mul4(E, a.v[0], a.v[2], a.v[4], a.v[6], b.v[0]);
        round_reduce(E, O);
#pragma unroll
        for (int i = 1; i < 8; i++) {
            if (i & 1) {
                round_mul(O, E, a, b.v[i]);
                round_reduce(O, E);
            } else {
                round_mul(E, O, a, b.v[i]);
                round_reduce(E, O);
            }
        }
        // after 8 rounds E is lo-aligned again with one pending shift of O:
        // T = E + (O >> 32)  (O[0] == 0)
        uint32_t sh[8], t[8];
#pragma unroll
        for (int i = 0; i < 7; i++) sh[i] = O[i + 1];
        sh[7] = 0;
        ff_add8(t, E, sh);
        return reduce_once(t, 0);
    }
    FF_HD Field sqr() const { return (*this) * (*this); }

    // x (canonical) -> Montgomery, and back
    FF_HD Field to_mont() const { return (*this) * r2(); }
    FF_HD Field from_mont() const {
        Field o = zero();
        o.v[0] = 1;
        return (*this) * o;
    }

    FF_HD Field pow_u256(const uint32_t* e) const {
        Field acc = one();
        for (int i = 255; i >= 0; i--) {
            acc = acc.sqr();
            if ((e[i >> 5] >> (i & 31)) & 1u) acc = acc * (*this);
        }
        return acc;
    }
    // Fermat inverse (0 -> 0): ~380 dependent Montgomery products.  Kept as the cross-check of
    // inverse() in the host tests.
    FF_HD Field inverse_fermat() const {
        uint32_t e[8], two[8] = {2, 0, 0, 0, 0, 0, 0, 0}, m[8];
        modulus(m);
        ff_sub8(e, m, two);
        return pow_u256(e);
    }
    // x/2 mod p for x in [0, p)
    FF_HD Field halve() const {
        uint32_t t[8], m[8];
        modulus(m);
        const bool odd = v[0] & 1u;
        if (odd) ff_add8(t, v, m);   // < 2^255: no carry out
        Field r;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint32_t lo = odd ? t[i] : v[i];
            uint32_t hi = i < 7 ? (odd ? t[i + 1] : v[i + 1]) : 0u;
            r.v[i] = (lo >> 1) | (hi << 31);
        }
        return r;
    }
    // Inverse (0 -> 0) by the binary extended Euclidean algorithm on the Montgomery representative:
    // shifts, additions and subtractions on the ALU pipe only (~15 K simple instructions, no
    // multiplier), with x1 started at R^2 so that the result is already in Montgomery form:
    //   invariants  x1 * X = u * R^2,  x2 * X = v * R^2  (mod p)   =>  u = 1: x1 = R^2 / X = (a^-1) R.
    // Data-dependent loops: fine for one value on the host, divergent when the lanes of a warp invert different
    // values.  Kept as a second cross-check of inverse().
    FF_HD Field inverse_euclid() const {
        if (is_zero()) return *this;
        uint32_t u[8], w[8], t[8];
#pragma unroll
        for (int i = 0; i < 8; i++) u[i] = v[i];
        modulus(w);
        Field x1 = r2(), x2 = zero();
        auto is_one = [](const uint32_t* a) {
            return a[0] == 1u && (a[1] | a[2] | a[3] | a[4] | a[5] | a[6] | a[7]) == 0u;
        };
        auto shr1 = [](uint32_t* a) {
#pragma unroll
            for (int i = 0; i < 7; i++) a[i] = (a[i] >> 1) | (a[i + 1] << 31);
            a[7] >>= 1;
        };
        while (!is_one(u) && !is_one(w)) {
            while (!(u[0] & 1u)) {
                shr1(u);
                x1 = x1.halve();
            }
            while (!(w[0] & 1u)) {
                shr1(w);
                x2 = x2.halve();
            }
            if (ff_sub8(t, u, w) == 0) {   // u >= w
#pragma unroll
                for (int i = 0; i < 8; i++) u[i] = t[i];
                x1 = x1 - x2;
            } else {
                ff_sub8(t, w, u);
#pragma unroll
                for (int i = 0; i < 8; i++) w[i] = t[i];
                x2 = x2 - x1;
            }
        }
        return is_one(u) ? x1 : x2;
    }


    // Inverse (0 -> 0) with a fixed instruction sequence: the binary GCD on 64-bit approximations of Pornin
    // ("Optimized binary GCD for modular inversion", 2020).  17 rounds; each runs 30 halving/subtract steps on
    // (low 31 bits | top 33 bits) of a and b while tracking the 2x2 update matrix (f0 g0; f1 g1), |.| <= 2^30, then
    // applies it to the full values:  (a, b) <- (f0 a + g0 b, f1 a + g1 b) / 2^30  (exact) and
    // (u, v) <- (f0 u + g0 v, f1 u + g1 v) / 2^30 mod p  (one Montgomery word step).  Invariants a = u X / R^2,
    // b = v X / R^2 with u0 = R^2, v0 = 0; after 510 >= 2*254 - 1 steps b = gcd = 1 and v = R^2 / X = (a^-1) R.
    // No branch depends on the data, so the 32 lanes of a warp can invert 32 different values in lockstep (the
    // witness solver's division rows); ~3x fewer instructions than the shift/subtract loop of inverse_euclid().
    FF_HD Field inverse() const {
        uint32_t a[8], b[8], m[8];
        Field U = r2(), V = zero();
        modulus(m);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            a[i] = v[i];
            b[i] = m[i];
        }
        for (int round = 0; round < 17; round++) {
            // ---- approximations: n = bit length of max(a, b); keep bits [0, 31) and [n - 33, n)
            uint32_t hi_a = 0, mi_a = 0, lo_a = 0, hi_b = 0, mi_b = 0, lo_b = 0, top = 0;
            int j = 0;
#pragma unroll
            for (int i = 7; i >= 2; i--) {
                const bool here = top == 0 && (a[i] | b[i]) != 0;
                if (here) {
                    top = a[i] | b[i];
                    j = i;
                    hi_a = a[i], mi_a = a[i - 1], lo_a = a[i - 2];
                    hi_b = b[i], mi_b = b[i - 1], lo_b = b[i - 2];
                }
            }
            uint64_t xa, xb;
            if (j == 0) {   // n <= 64: exact
                xa = ((uint64_t)a[1] << 32) | a[0];
                xb = ((uint64_t)b[1] << 32) | b[0];
            } else {
#ifdef __CUDA_ARCH__
                const int s = __clz((int)top);
#else
                const int s = __builtin_clz(top);   // top != 0
#endif
                const uint64_t ta = ((((uint64_t)hi_a << 32) | mi_a) << s) | (s ? (uint64_t)(lo_a >> (32 - s)) : 0);
                const uint64_t tb = ((((uint64_t)hi_b << 32) | mi_b) << s) | (s ? (uint64_t)(lo_b >> (32 - s)) : 0);
                xa = ((ta >> 31) << 31) | (a[0] & 0x7fffffffu);
                xb = ((tb >> 31) << 31) | (b[0] & 0x7fffffffu);
            }
            // ---- 30 steps on the approximations
            int32_t f0 = 1, g0 = 0, f1 = 0, g1 = 1;
            for (int i = 0; i < 30; i++) {
                const bool odd = xa & 1u;
                const bool swap = odd && xa < xb;
                const uint64_t ta = swap ? xb : xa, tb = swap ? xa : xb;
                const int32_t tf0 = swap ? f1 : f0, tf1 = swap ? f0 : f1, tg0 = swap ? g1 : g0, tg1 = swap ? g0 : g1;
                xa = (ta - (odd ? tb : 0)) >> 1;
                xb = tb;
                f0 = tf0 - (odd ? tf1 : 0);
                g0 = tg0 - (odd ? tg1 : 0);
                f1 = tf1 * 2;
                g1 = tg1 * 2;
            }
            // ---- (a, b) <- (f0 a + g0 b, f1 a + g1 b) / 2^30, made non-negative (flip the matrix row with it)
            uint32_t na[9], nb[9];
            ff_lincomb9(na, a, f0, b, g0);
            ff_lincomb9(nb, a, f1, b, g1);
            if (na[8] & 0x80000000u) {
                ff_neg9(na);
                f0 = -f0, g0 = -g0;
            }
            if (nb[8] & 0x80000000u) {
                ff_neg9(nb);
                f1 = -f1, g1 = -g1;
            }
#pragma unroll
            for (int i = 0; i < 8; i++) {
                a[i] = (na[i] >> 30) | (na[i + 1] << 2);
                b[i] = (nb[i] >> 30) | (nb[i + 1] << 2);
            }
            // ---- (u, v) <- (f0 u + g0 v, f1 u + g1 v) / 2^30 mod p
            const Field nu = mont_step30(U, f0, V, g0), nv = mont_step30(U, f1, V, g1);
            U = nu;
            V = nv;
        }
        return V;
    }

   private:
    // out (9 limbs, two's complement) = f x + g y,  |f| + |g| <= 2^30 ... 2^31: every partial sum fits an int64
    FF_HD static void ff_lincomb9(uint32_t* out, const uint32_t* x, int32_t f, const uint32_t* y, int32_t g) {
        int64_t acc = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            acc += (int64_t)x[i] * f + (int64_t)y[i] * g;
            out[i] = (uint32_t)acc;
            acc >>= 32;
        }
        out[8] = (uint32_t)acc;
    }
    FF_HD static void ff_neg9(uint32_t* t) {
        uint32_t c = 1;
#pragma unroll
        for (int i = 0; i < 9; i++) {
            const uint32_t x = ~t[i] + c;
            c = (c && x == 0) ? 1u : 0u;
            t[i] = x;
        }
    }
    // (f x + g y) / 2^30 mod p for x, y in [0, p), |f| + |g| <= 2^30
    FF_HD static Field mont_step30(const Field& x, int32_t f, const Field& y, int32_t g) {
        uint32_t t[9], m[8];
        modulus(m);
        ff_lincomb9(t, x.v, f, y.v, g);   // |t| < 2^30 p
        // t += 2^30 p: non-negative, same residue
        uint64_t c = 0;
#pragma unroll
        for (int i = 0; i < 9; i++) {
            const uint32_t lo = i == 0 ? 0u : (m[i - 1] >> 2), hi = i < 8 ? (m[i] << 30) : 0u;
            c += (uint64_t)t[i] + (lo | hi);
            t[i] = (uint32_t)c;
            c >>= 32;
        }
        // t += k p with k = -t / p mod 2^30: the low 30 bits clear
        const uint32_t k = (t[0] * PM::INV) & 0x3fffffffu;
        c = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            c += (uint64_t)t[i] + (uint64_t)m[i] * k;
            t[i] = (uint32_t)c;
            c >>= 32;
        }
        t[8] += (uint32_t)c;
        Field r;   // t / 2^30 < 3p
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = (t[i] >> 30) | (t[i + 1] << 2);
        uint32_t d[8];
        for (int it = 0; it < 2; it++) {
            const bool below = ff_sub8(d, r.v, m) != 0;
#pragma unroll
            for (int i = 0; i < 8; i++) r.v[i] = below ? r.v[i] : d[i];
        }
        return r;
    }

    // lo += a_even*bi ; y (old lo-aligned, limb0 cleared) shifts into hi alignment += a_odd*bi
    FF_HD static void round_mul(uint32_t* lo, uint32_t* y, const Field& a, uint32_t bi) {
        ff_mad4_shift(y, lo[0], a.v[1], a.v[3], a.v[5], a.v[7], bi);
        uint32_t c = ff_mad4(lo, a.v[0], a.v[2], a.v[4], a.v[6], bi);
        y[7] += c;
    }
    // m = lo[0]*INV ; hi += p_odd*m ; lo += p_even*m  (clears lo[0])
    FF_HD static void round_reduce(uint32_t* lo, uint32_t* hi) {
        uint32_t m = lo[0] * PM::INV;
        ff_mad4(hi, PM::M1, PM::M3, PM::M5, PM::M7, m);
        uint32_t c = ff_mad4(lo, PM::M0, PM::M2, PM::M4, PM::M6, m);
        hi[7] += c;
    }
};

typedef Field<FpParams> Fp;
typedef Field<FrParams> Fr;

// Fp product as a real call on the device: an Fp2 point addition is 28 base-field products; inlined, the G2 bucket
// kernels are ~100 KB of straight-line SASS and stall on instruction fetch.
#ifdef __CUDA_ARCH__
static __device__ __noinline__ Fp fp_mul_call(Fp a, Fp b) { return a * b; }
#define FP2_MUL(a, b) fp_mul_call((a), (b))
#else
#define FP2_MUL(a, b) ((a) * (b))
#endif

// ---- Fp2 = Fp[u]/(u^2+1) ------------------------------------------------------------------
struct Fp2 {
    Fp c0, c1;
    FF_HD static Fp2 zero() { return {Fp::zero(), Fp::zero()}; }
    FF_HD static Fp2 one() { return {Fp::one(), Fp::zero()}; }
    FF_HD bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    FF_HD bool operator==(const Fp2& o) const { return c0 == o.c0 && c1 == o.c1; }
    FF_HD bool operator!=(const Fp2& o) const { return !(*this == o); }
    FF_HD friend Fp2 operator+(const Fp2& a, const Fp2& b) { return {a.c0 + b.c0, a.c1 + b.c1}; }
    FF_HD friend Fp2 operator-(const Fp2& a, const Fp2& b) { return {a.c0 - b.c0, a.c1 - b.c1}; }
    FF_HD Fp2 neg() const { return {c0.neg(), c1.neg()}; }
    FF_HD Fp2 dbl() const { return {c0.dbl(), c1.dbl()}; }
    // Karatsuba: 3 base multiplications
    FF_HD static Fp2 mul_inl(const Fp2& a, const Fp2& b) {
        Fp t0 = a.c0 * b.c0;
        Fp t1 = a.c1 * b.c1;
        Fp t2 = (a.c0 + a.c1) * (b.c0 + b.c1);
        return {t0 - t1, t2 - t0 - t1};
    }
    // (a0+a1)(a0-a1) + 2 a0 a1 u : 2 base multiplications
    FF_HD static Fp2 sqr_inl(const Fp2& a) {
        Fp t0 = (a.c0 + a.c1) * (a.c0 - a.c1);
        Fp t1 = a.c0 * a.c1;
        return {t0, t1.dbl()};
    }
#ifdef __CUDA_ARCH__
    // Real calls on the device, at the Fp2 level: an inlined Fp2 point addition is ~100 KB of straight-line SASS
    // (instruction-fetch bound); one call per Fp2 product moves 48 registers instead of 72 for three Fp calls.
    static __device__ __noinline__ Fp2 mul_call(Fp2 a, Fp2 b) { return mul_inl(a, b); }
    static __device__ __noinline__ Fp2 sqr_call(Fp2 a) { return sqr_inl(a); }
    __device__ __forceinline__ friend Fp2 operator*(const Fp2& a, const Fp2& b) { return mul_call(a, b); }
    __device__ __forceinline__ Fp2 sqr() const { return sqr_call(*this); }
#else
    friend Fp2 operator*(const Fp2& a, const Fp2& b) { return mul_inl(a, b); }
    Fp2 sqr() const { return sqr_inl(*this); }
#endif
    FF_HD Fp2 inverse() const {
        Fp d = (c0.sqr() + c1.sqr()).inverse();
        return {c0 * d, (c1 * d).neg()};
    }
};

}  // namespace g16

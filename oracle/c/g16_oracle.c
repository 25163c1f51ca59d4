/* g16_oracle.c -- CPU restatement of the hot path of `sunspot prove` (gnark v0.14 groth16.Prove on
 * BN254) in plain C + OpenMP.  TEST ORACLE / CPU BASELINE ONLY: nothing on the product path links
 * or loads this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may.
 *
 * What it restates (third-party Go, absent from /root/reference; invoked there at
 * client/proof.helper.ts:64, noir_circuit/prove_linux.sh:83 -- SURVEY.md 3.2, 9.7):
 *   - gnark-crypto fp/fr Montgomery arithmetic (here 4x64-bit CIOS with unsigned __int128)
 *   - G1Jac/G2Jac.MultiExp: signed-digit bucket method, extended-Jacobian buckets, one task per
 *     window, running-sum bucket reduction, windows combined high to low
 *   - fr/fft: in-place radix-2 DIF (natural -> bit-reversed) / DIT (bit-reversed -> natural)
 *   - computeH and the A / B1 / B2 / K+Z / PoK assembly of groth16.Prove
 * "gnark-algorithm CPU restatement -- not gnark": PARITY UNPINNED against gnark itself (no Go
 * toolchain, no .pk/.proof fixtures in the reference); pinned against oracle/py (big integers)
 * by tests/test_oracle_c.py.
 *
 * All buffers are gnark wire formats: Fr 32 B big-endian, G1 64 B, G2 128 B (X.A1|X.A0|Y.A1|Y.A0).
 */
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t l[4]; } fe;          /* Montgomery residue mod `m->p` */

typedef struct { uint64_t p[4]; uint64_t inv; fe one; fe r2; } modulus;

static const modulus FP = {
    {0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull},
    0x87d20782e4866389ull,
    {{0xd35d438dc58f0d9dull, 0x0a78eb28f5c70b3dull, 0x666ea36f7879462cull, 0x0e0a77c19a07df2full}},
    {{0xf32cfc5b538afa89ull, 0xb5e71911d44501fbull, 0x47ab1eff0a417ff6ull, 0x06d89f71cab8351full}}};
static const modulus FR = {
    {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull},
    0xc2e1f593efffffffull,
    {{0xac96341c4ffffffbull, 0x36fc76959f60cd29ull, 0x666ea36f7879462eull, 0x0e0a77c19a07df2full}},
    {{0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull, 0x8c49833d53bb8085ull, 0x0216d0b17f4e44a5ull}}};

static int fe_geq(const uint64_t* a, const uint64_t* b) {
    for (int i = 3; i >= 0; i--) { if (a[i] > b[i]) return 1; if (a[i] < b[i]) return 0; }
    return 1;
}
static void fe_sub_p(uint64_t* a, const uint64_t* p) {
    u128 bw = 0;
    for (int i = 0; i < 4; i++) { u128 t = (u128)a[i] - p[i] - (uint64_t)bw; a[i] = (uint64_t)t; bw = (t >> 64) & 1; }
}
static int fe_is_zero(const fe* a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
static int fe_eq(const fe* a, const fe* b) { return memcmp(a, b, sizeof(fe)) == 0; }

static void fe_add(fe* r, const fe* a, const fe* b, const modulus* m) {
    u128 c = 0;
    for (int i = 0; i < 4; i++) { c += (u128)a->l[i] + b->l[i]; r->l[i] = (uint64_t)c; c >>= 64; }
    if (c || fe_geq(r->l, m->p)) fe_sub_p(r->l, m->p);
}
static void fe_sub(fe* r, const fe* a, const fe* b, const modulus* m) {
    u128 bw = 0;
    for (int i = 0; i < 4; i++) { u128 t = (u128)a->l[i] - b->l[i] - (uint64_t)bw; r->l[i] = (uint64_t)t; bw = (t >> 64) & 1; }
    if (bw) { u128 c = 0; for (int i = 0; i < 4; i++) { c += (u128)r->l[i] + m->p[i]; r->l[i] = (uint64_t)c; c >>= 64; } }
}
static void fe_neg(fe* r, const fe* a, const modulus* m) {
    fe z = {{0, 0, 0, 0}};
    if (fe_is_zero(a)) *r = *a; else fe_sub(r, &z, a, m);
}
static void fe_mul(fe* r, const fe* a, const fe* b, const modulus* m) {
    uint64_t t[6] = {0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) { c += (u128)a->l[j] * b->l[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
        c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
        uint64_t k = t[0] * m->inv;
        c = ((u128)k * m->p[0] + t[0]) >> 64;
        for (int j = 1; j < 4; j++) { c += (u128)k * m->p[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
        c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
    }
    memcpy(r->l, t, 32);
    if (t[4] || fe_geq(r->l, m->p)) fe_sub_p(r->l, m->p);
}
static void fe_sqr(fe* r, const fe* a, const modulus* m) { fe_mul(r, a, a, m); }
static void fe_dbl(fe* r, const fe* a, const modulus* m) { fe_add(r, a, a, m); }
static void fe_inv(fe* r, const fe* a, const modulus* m) {          /* a^(p-2) */
    uint64_t e[4] = {m->p[0] - 2, m->p[1], m->p[2], m->p[3]};
    fe acc = m->one, base = *a;
    for (int i = 0; i < 256; i++) {
        if ((e[i >> 6] >> (i & 63)) & 1) fe_mul(&acc, &acc, &base, m);
        fe_sqr(&base, &base, m);
    }
    *r = acc;
}
static void fe_from_be(fe* r, const uint8_t* be, const modulus* m) {
    fe x;
    for (int i = 0; i < 4; i++) { uint64_t v = 0; for (int k = 0; k < 8; k++) v = (v << 8) | be[(3 - i) * 8 + k]; x.l[i] = v; }
    while (fe_geq(x.l, m->p)) fe_sub_p(x.l, m->p);
    fe_mul(r, &x, &m->r2, m);
}
static void fe_canon(uint64_t out[4], const fe* a, const modulus* m) {
    fe one = {{1, 0, 0, 0}}, c;
    fe_mul(&c, a, &one, m);
    memcpy(out, c.l, 32);
}
static void fe_to_be(uint8_t* be, const fe* a, const modulus* m) {
    uint64_t c[4];
    fe_canon(c, a, m);
    for (int i = 0; i < 4; i++) for (int k = 0; k < 8; k++) be[(3 - i) * 8 + k] = (uint8_t)(c[i] >> (56 - 8 * k));
}

/* ---- Fp2 = Fp[u]/(u^2+1) -------------------------------------------------------------------- */
typedef struct { fe a0, a1; } fe2;
static void f2_add(fe2* r, const fe2* a, const fe2* b) { fe_add(&r->a0, &a->a0, &b->a0, &FP); fe_add(&r->a1, &a->a1, &b->a1, &FP); }
static void f2_sub(fe2* r, const fe2* a, const fe2* b) { fe_sub(&r->a0, &a->a0, &b->a0, &FP); fe_sub(&r->a1, &a->a1, &b->a1, &FP); }
static void f2_neg(fe2* r, const fe2* a) { fe_neg(&r->a0, &a->a0, &FP); fe_neg(&r->a1, &a->a1, &FP); }
static void f2_mul(fe2* r, const fe2* a, const fe2* b) {           /* schoolbook: 4 products */
    fe t0, t1, t2, t3;
    fe_mul(&t0, &a->a0, &b->a0, &FP); fe_mul(&t1, &a->a1, &b->a1, &FP);
    fe_mul(&t2, &a->a0, &b->a1, &FP); fe_mul(&t3, &a->a1, &b->a0, &FP);
    fe_sub(&r->a0, &t0, &t1, &FP); fe_add(&r->a1, &t2, &t3, &FP);
}
static int f2_is_zero(const fe2* a) { return fe_is_zero(&a->a0) && fe_is_zero(&a->a1); }
static void f2_inv(fe2* r, const fe2* a) {
    fe t0, t1, d;
    fe_sqr(&t0, &a->a0, &FP); fe_sqr(&t1, &a->a1, &FP); fe_add(&d, &t0, &t1, &FP); fe_inv(&d, &d, &FP);
    fe_mul(&r->a0, &a->a0, &d, &FP); fe_mul(&t0, &a->a1, &d, &FP); fe_neg(&r->a1, &t0, &FP);
}

/* ---- group law, generated for both coordinate fields ------------------------------------------
 * Extended Jacobian (X, Y, ZZ, ZZZ), x = X/ZZ, y = Y/ZZZ; ZZ == 0 is infinity.  Affine infinity is
 * x = y = 0.  (gnark-crypto keeps its MSM buckets in the same coordinate system.) */
#define DEFINE_GROUP(G, F, ADD, SUB, MUL, NEG, ISZERO, INV, ONE_INIT)                                              \
    typedef struct { F x, y; } G##_aff;                                                                            \
    typedef struct { F X, Y, ZZ, ZZZ; } G##_ext;                                                                   \
    static void G##_set_inf(G##_ext* p) { memset(p, 0, sizeof *p); }                                               \
    static int G##_is_inf(const G##_ext* p) { return ISZERO(&p->ZZ); }                                             \
    static int G##_aff_is_inf(const G##_aff* p) { return ISZERO(&p->x) && ISZERO(&p->y); }                         \
    static void G##_dbl(G##_ext* r, const G##_ext* p) {                                                            \
        if (G##_is_inf(p) || ISZERO(&p->Y)) { G##_set_inf(r); return; }                                            \
        F U, V, W, S, M, t, X3, Y3;                                                                                \
        ADD(&U, &p->Y, &p->Y); MUL(&V, &U, &U); MUL(&W, &U, &V); MUL(&S, &p->X, &V);                               \
        MUL(&t, &p->X, &p->X); ADD(&M, &t, &t); ADD(&M, &M, &t);                                                   \
        MUL(&X3, &M, &M); SUB(&X3, &X3, &S); SUB(&X3, &X3, &S);                                                    \
        SUB(&t, &S, &X3); MUL(&Y3, &M, &t); MUL(&t, &W, &p->Y); SUB(&Y3, &Y3, &t);                                 \
        F zz, zzz; MUL(&zz, &V, &p->ZZ); MUL(&zzz, &W, &p->ZZZ);                                                   \
        r->X = X3; r->Y = Y3; r->ZZ = zz; r->ZZZ = zzz;                                                            \
    }                                                                                                              \
    static void G##_from_aff(G##_ext* r, const G##_aff* p) {                                                       \
        if (G##_aff_is_inf(p)) { G##_set_inf(r); return; }                                                         \
        F one = ONE_INIT; r->X = p->x; r->Y = p->y; r->ZZ = one; r->ZZZ = one;                                     \
    }                                                                                                              \
    static void G##_add(G##_ext* r, const G##_ext* a, const G##_ext* b) {                                          \
        if (G##_is_inf(b)) { *r = *a; return; }                                                                    \
        if (G##_is_inf(a)) { *r = *b; return; }                                                                    \
        F U1, U2, S1, S2, Pq, Rq, PP, PPP, Q, t, X3, Y3;                                                           \
        MUL(&U1, &a->X, &b->ZZ); MUL(&U2, &b->X, &a->ZZ); MUL(&S1, &a->Y, &b->ZZZ); MUL(&S2, &b->Y, &a->ZZZ);      \
        SUB(&Pq, &U2, &U1); SUB(&Rq, &S2, &S1);                                                                    \
        if (ISZERO(&Pq)) { if (ISZERO(&Rq)) G##_dbl(r, a); else G##_set_inf(r); return; }                          \
        MUL(&PP, &Pq, &Pq); MUL(&PPP, &Pq, &PP); MUL(&Q, &U1, &PP);                                                \
        MUL(&X3, &Rq, &Rq); SUB(&X3, &X3, &PPP); SUB(&X3, &X3, &Q); SUB(&X3, &X3, &Q);                             \
        SUB(&t, &Q, &X3); MUL(&Y3, &Rq, &t); MUL(&t, &S1, &PPP); SUB(&Y3, &Y3, &t);                                \
        F zz, zzz; MUL(&zz, &a->ZZ, &b->ZZ); MUL(&zz, &zz, &PP); MUL(&zzz, &a->ZZZ, &b->ZZZ); MUL(&zzz, &zzz, &PPP); \
        r->X = X3; r->Y = Y3; r->ZZ = zz; r->ZZZ = zzz;                                                            \
    }                                                                                                              \
    static void G##_madd(G##_ext* r, const G##_aff* b, int negate) {                                               \
        if (G##_aff_is_inf(b)) return;                                                                             \
        G##_ext e; G##_from_aff(&e, b);                                                                            \
        if (negate) NEG(&e.Y, &e.Y);                                                                               \
        G##_ext s; G##_add(&s, r, &e); *r = s;                                                                     \
    }                                                                                                              \
    static void G##_to_aff(G##_aff* r, const G##_ext* p) {                                                         \
        if (G##_is_inf(p)) { memset(r, 0, sizeof *r); return; }                                                    \
        F izz, izzz; INV(&izz, &p->ZZ); INV(&izzz, &p->ZZZ); MUL(&r->x, &p->X, &izz); MUL(&r->y, &p->Y, &izzz);    \
    }

static void fp_add_(fe* r, const fe* a, const fe* b) { fe_add(r, a, b, &FP); }
static void fp_sub_(fe* r, const fe* a, const fe* b) { fe_sub(r, a, b, &FP); }
static void fp_mul_(fe* r, const fe* a, const fe* b) { fe_mul(r, a, b, &FP); }
static void fp_neg_(fe* r, const fe* a) { fe_neg(r, a, &FP); }
static void fp_inv_(fe* r, const fe* a) { fe_inv(r, a, &FP); }
#define FP_ONE_INIT {{0xd35d438dc58f0d9dull, 0x0a78eb28f5c70b3dull, 0x666ea36f7879462cull, 0x0e0a77c19a07df2full}}
#define FP2_ONE_INIT {FP_ONE_INIT, {{0, 0, 0, 0}}}
DEFINE_GROUP(g1, fe, fp_add_, fp_sub_, fp_mul_, fp_neg_, fe_is_zero, fp_inv_, FP_ONE_INIT)
DEFINE_GROUP(g2, fe2, f2_add, f2_sub, f2_mul, f2_neg, f2_is_zero, f2_inv, FP2_ONE_INIT)

static void g1_aff_from_be(g1_aff* p, const uint8_t* be) {
    if ((be[0] & 0xc0) == 0x40) { memset(p, 0, sizeof *p); return; }
    fe_from_be(&p->x, be, &FP); fe_from_be(&p->y, be + 32, &FP);
}
static void g1_aff_to_be(uint8_t* be, const g1_aff* p) {
    if (g1_aff_is_inf(p)) { memset(be, 0, 64); return; }
    fe_to_be(be, &p->x, &FP); fe_to_be(be + 32, &p->y, &FP);
}
static void g2_aff_from_be(g2_aff* p, const uint8_t* be) {
    if ((be[0] & 0xc0) == 0x40) { memset(p, 0, sizeof *p); return; }
    fe_from_be(&p->x.a1, be, &FP); fe_from_be(&p->x.a0, be + 32, &FP);
    fe_from_be(&p->y.a1, be + 64, &FP); fe_from_be(&p->y.a0, be + 96, &FP);
}
static void g2_aff_to_be(uint8_t* be, const g2_aff* p) {
    if (g2_aff_is_inf(p)) { memset(be, 0, 128); return; }
    fe_to_be(be, &p->x.a1, &FP); fe_to_be(be + 32, &p->x.a0, &FP);
    fe_to_be(be + 64, &p->y.a1, &FP); fe_to_be(be + 96, &p->y.a0, &FP);
}

/* ---- bucket-method MSM (gnark-crypto multiexp.go structure) ----------------------------------- */
static int msm_window(size_t n) {          /* gnark picks c from a table by size; same shape */
    int c = 4;
    while (c < 16 && ((size_t)1 << (c + 3)) < n) c++;   /* c ~ log2(n) - 3, clamped to [4,16] */
    return c;
}
/* scalars: canonical 4x64; digits[w*n + i] signed */
static int32_t* msm_digits(const uint64_t (*sc)[4], size_t n, int c, int* nwin) {
    int W = (254 + c - 1) / c + 1;
    *nwin = W;
    int32_t* d = (int32_t*)malloc(sizeof(int32_t) * n * W);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        int carry = 0;
        for (int w = 0; w < W; w++) {
            int bit = w * c;
            int64_t v = 0;
            if (bit < 256) {
                int wi = bit >> 6, sh = bit & 63;
                uint64_t x = sc[i][wi] >> sh;
                if (sh && wi + 1 < 4) x |= sc[i][wi + 1] << (64 - sh);
                v = (int64_t)(x & (((uint64_t)1 << c) - 1));
            }
            v += carry;
            carry = 0;
            if (v > ((int64_t)1 << (c - 1))) { v -= (int64_t)1 << c; carry = 1; }
            d[(size_t)w * n + i] = (int32_t)v;
        }
    }
    return d;
}

#define DEFINE_MSM(G)                                                                                              \
    static void G##_msm(G##_ext* out, const G##_aff* pts, const uint64_t (*sc)[4], size_t n) {                     \
        G##_set_inf(out);                                                                                          \
        if (n == 0) return;                                                                                        \
        int c = msm_window(n), W;                                                                                  \
        int32_t* dig = msm_digits(sc, n, c, &W);                                                                   \
        size_t nb = (size_t)1 << (c - 1);                                                                          \
        G##_ext* win = (G##_ext*)malloc(sizeof(G##_ext) * W);                                                      \
        _Pragma("omp parallel for schedule(dynamic, 1)")                                                           \
        for (int w = 0; w < W; w++) {                                                                              \
            G##_ext* bk = (G##_ext*)calloc(nb, sizeof(G##_ext));                                                   \
            const int32_t* d = dig + (size_t)w * n;                                                                \
            for (size_t i = 0; i < n; i++) {                                                                       \
                int32_t v = d[i];                                                                                  \
                if (v > 0) G##_madd(&bk[v - 1], &pts[i], 0);                                                       \
                else if (v < 0) G##_madd(&bk[-v - 1], &pts[i], 1);                                                 \
            }                                                                                                      \
            G##_ext run, acc, t;                                                                                   \
            G##_set_inf(&run); G##_set_inf(&acc);                                                                  \
            for (size_t k = nb; k-- > 0;) { G##_add(&t, &run, &bk[k]); run = t; G##_add(&t, &acc, &run); acc = t; } \
            win[w] = acc;                                                                                          \
            free(bk);                                                                                              \
        }                                                                                                          \
        G##_ext res = win[W - 1], t;                                                                               \
        for (int w = W - 2; w >= 0; w--) {                                                                         \
            for (int k = 0; k < c; k++) { G##_dbl(&t, &res); res = t; }                                            \
            G##_add(&t, &res, &win[w]); res = t;                                                                   \
        }                                                                                                          \
        *out = res;                                                                                                \
        free(win); free(dig);                                                                                      \
    }
DEFINE_MSM(g1)
DEFINE_MSM(g2)

/* ---- radix-2 NTT over Fr (gnark-crypto fr/fft) -------------------------------------------------- */
static const uint64_t ROOT_2_28[4] = {0x9bd61b6e725b19f0ull, 0x402d111e41112ed4ull, 0x00e0a7eb8ef62abcull, 0x2a3c09f0a58a7e85ull};
static void fr_from_u64(fe* r, uint64_t v) { fe x = {{v, 0, 0, 0}}; fe_mul(r, &x, &FR.r2, &FR); }
static void fr_pow_u64(fe* r, const fe* a, uint64_t e) {
    fe acc = FR.one, b = *a;
    while (e) { if (e & 1) fe_mul(&acc, &acc, &b, &FR); fe_sqr(&b, &b, &FR); e >>= 1; }
    *r = acc;
}
static fe* twiddles(unsigned logn, int inverse) {
    size_t half = ((size_t)1 << logn) / 2;
    fe w, raw = {{ROOT_2_28[0], ROOT_2_28[1], ROOT_2_28[2], ROOT_2_28[3]}};
    fe_mul(&w, &raw, &FR.r2, &FR);
    for (unsigned i = logn; i < 28; i++) fe_sqr(&w, &w, &FR);
    if (inverse) fe_inv(&w, &w, &FR);
    fe* tw = (fe*)malloc(sizeof(fe) * (half ? half : 1));
    tw[0] = FR.one;
    for (size_t i = 1; i < half; i++) fe_mul(&tw[i], &tw[i - 1], &w, &FR);
    return tw;
}
/* DIF: natural -> bit-reversed */
static void ntt_dif(fe* a, unsigned logn, const fe* tw) {
    size_t n = (size_t)1 << logn;
    for (unsigned s = logn; s-- > 0;) {
        size_t h = (size_t)1 << s, step = n / (2 * h);
#pragma omp parallel for schedule(static)
        for (size_t b = 0; b < n / 2; b++) {
            size_t blk = b / h, j = b % h, i0 = blk * 2 * h + j, i1 = i0 + h;
            fe u = a[i0], v = a[i1], d;
            fe_add(&a[i0], &u, &v, &FR);
            fe_sub(&d, &u, &v, &FR);
            fe_mul(&a[i1], &d, &tw[j * step], &FR);
        }
    }
}
/* DIT: bit-reversed -> natural */
static void ntt_dit(fe* a, unsigned logn, const fe* tw) {
    size_t n = (size_t)1 << logn;
    for (unsigned s = 0; s < logn; s++) {
        size_t h = (size_t)1 << s, step = n / (2 * h);
#pragma omp parallel for schedule(static)
        for (size_t b = 0; b < n / 2; b++) {
            size_t blk = b / h, j = b % h, i0 = blk * 2 * h + j, i1 = i0 + h;
            fe u = a[i0], v;
            fe_mul(&v, &a[i1], &tw[j * step], &FR);
            fe_add(&a[i0], &u, &v, &FR);
            fe_sub(&a[i1], &u, &v, &FR);
        }
    }
}
static size_t bitrev(size_t x, unsigned bits) {
    size_t r = 0;
    for (unsigned i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}
/* computeH (gnark prove.go): a,b,c natural-order evaluations -> h bit-reversed coefficients in a */
static void compute_h(fe* a, fe* b, fe* c, unsigned logn) {
    size_t n = (size_t)1 << logn;
    fe* twi = twiddles(logn, 1);
    fe* twf = twiddles(logn, 0);
    fe g, ginv, ninv, gn, den, one = FR.one;
    fr_from_u64(&g, 5); fe_inv(&ginv, &g, &FR);
    fr_from_u64(&ninv, (uint64_t)n); fe_inv(&ninv, &ninv, &FR);
    fr_pow_u64(&gn, &g, (uint64_t)n); fe_sub(&den, &gn, &one, &FR); fe_inv(&den, &den, &FR);
    fe* cos = (fe*)malloc(sizeof(fe) * n);      /* g^j / n   */
    fe* cosi = (fe*)malloc(sizeof(fe) * n);     /* g^-j / n  */
    cos[0] = ninv; cosi[0] = ninv;
    for (size_t j = 1; j < n; j++) { fe_mul(&cos[j], &cos[j - 1], &g, &FR); fe_mul(&cosi[j], &cosi[j - 1], &ginv, &FR); }
    fe* v[3] = {a, b, c};
    for (int k = 0; k < 3; k++) {
        ntt_dif(v[k], logn, twi);                           /* FFTInverse, DIF: coefficients, bit-reversed */
#pragma omp parallel for schedule(static)
        for (size_t p = 0; p < n; p++) fe_mul(&v[k][p], &v[k][p], &cos[bitrev(p, logn)], &FR);
        ntt_dit(v[k], logn, twf);                           /* FFT on the coset, DIT: natural order */
    }
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        fe t;
        fe_mul(&t, &a[i], &b[i], &FR); fe_sub(&t, &t, &c[i], &FR); fe_mul(&a[i], &t, &den, &FR);
    }
    ntt_dif(a, logn, twi);
#pragma omp parallel for schedule(static)
    for (size_t p = 0; p < n; p++) fe_mul(&a[p], &a[p], &cosi[bitrev(p, logn)], &FR);
    free(cos); free(cosi); free(twi); free(twf);
}

/* ================================ exported C ABI (ctypes) ======================================= */
static void scalars_from_be(uint64_t (*out)[4], const uint8_t* be, size_t n) {
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        for (int k = 0; k < 4; k++) { uint64_t v = 0; for (int b = 0; b < 8; b++) v = (v << 8) | be[32 * i + (3 - k) * 8 + b]; out[i][k] = v; }
        while (fe_geq(out[i], FR.p)) fe_sub_p(out[i], FR.p);
    }
}

int oracle_set_threads(int t) { if (t > 0) omp_set_num_threads(t); return omp_get_max_threads(); }

int oracle_msm_g1(const uint8_t* points_be, const uint8_t* scalars_be, size_t n, uint8_t* out_be) {
    g1_aff* pts = (g1_aff*)malloc(sizeof(g1_aff) * (n ? n : 1));
    uint64_t (*sc)[4] = malloc(32 * (n ? n : 1));
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) g1_aff_from_be(&pts[i], points_be + 64 * i);
    scalars_from_be(sc, scalars_be, n);
    g1_ext r; g1_aff a;
    g1_msm(&r, pts, (const uint64_t (*)[4])sc, n);
    g1_to_aff(&a, &r); g1_aff_to_be(out_be, &a);
    free(pts); free(sc);
    return 0;
}
int oracle_msm_g2(const uint8_t* points_be, const uint8_t* scalars_be, size_t n, uint8_t* out_be) {
    g2_aff* pts = (g2_aff*)malloc(sizeof(g2_aff) * (n ? n : 1));
    uint64_t (*sc)[4] = malloc(32 * (n ? n : 1));
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) g2_aff_from_be(&pts[i], points_be + 128 * i);
    scalars_from_be(sc, scalars_be, n);
    g2_ext r; g2_aff a;
    g2_msm(&r, pts, (const uint64_t (*)[4])sc, n);
    g2_to_aff(&a, &r); g2_aff_to_be(out_be, &a);
    free(pts); free(sc);
    return 0;
}
/* same semantics as g16_ntt: forward = natural -> bit-reversed (coset: times g^j first);
 * inverse = bit-reversed -> natural, times 1/n (coset: times g^-j after) */
int oracle_ntt(uint8_t* values_be, unsigned logn, int inverse, int coset) {
    size_t n = (size_t)1 << logn;
    fe* a = (fe*)malloc(sizeof(fe) * n);
    for (size_t i = 0; i < n; i++) fe_from_be(&a[i], values_be + 32 * i, &FR);
    fe* tw = twiddles(logn, inverse);
    fe g, f;
    fr_from_u64(&g, 5);
    if (!inverse) {
        if (coset) { f = FR.one; for (size_t j = 0; j < n; j++) { fe_mul(&a[j], &a[j], &f, &FR); fe_mul(&f, &f, &g, &FR); } }
        ntt_dif(a, logn, tw);
    } else {
        ntt_dit(a, logn, tw);
        fe ninv; fr_from_u64(&ninv, (uint64_t)n); fe_inv(&ninv, &ninv, &FR);
        fe gi; fe_inv(&gi, &g, &FR);
        f = ninv;
        for (size_t j = 0; j < n; j++) { fe_mul(&a[j], &a[j], &f, &FR); if (coset) fe_mul(&f, &f, &gi, &FR); }
    }
    for (size_t i = 0; i < n; i++) fe_to_be(values_be + 32 * i, &a[i], &FR);
    free(a); free(tw);
    return 0;
}
int oracle_compute_h(const uint8_t* abc_be, unsigned logn, uint8_t* h_be) {
    size_t n = (size_t)1 << logn;
    fe* v = (fe*)malloc(sizeof(fe) * 3 * n);
    for (size_t i = 0; i < 3 * n; i++) fe_from_be(&v[i], abc_be + 32 * i, &FR);
    compute_h(v, v + n, v + 2 * n, logn);
    for (size_t i = 0; i < n; i++) fe_to_be(h_be + 32 * i, &v[i], &FR);
    free(v);
    return 0;
}

/* ---- whole prove from a full wire vector (solver excluded) ------------- */
typedef struct {
    uint32_t n_wires, n_rows, logn, n_public;
    /* R1CS as three CSR matrices; coefficients as Montgomery fe (converted once by oracle_circuit_new) */
    const uint32_t* rowptr[3];
    const uint32_t* wire[3];
    fe* coeff[3];
    g1_aff *A, *B1, *K, *Z, *pok_basis;
    g2_aff* B2;
    uint32_t nA, nB, nK, nZ, nPok;
    uint32_t *mapA, *mapB, *mapK, *mapPok;
    g1_aff alpha1, beta1, delta1;
    g2_aff beta2, delta2;
} oracle_circuit;

static g1_aff* load_g1(const uint8_t* be, size_t n) {
    g1_aff* p = (g1_aff*)malloc(sizeof(g1_aff) * (n ? n : 1));
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) g1_aff_from_be(&p[i], be + 64 * i);
    return p;
}
static uint32_t* dup_u32(const uint32_t* s, size_t n) {
    uint32_t* d = (uint32_t*)malloc(4 * (n ? n : 1));
    memcpy(d, s, 4 * n);
    return d;
}

oracle_circuit* oracle_circuit_new(uint32_t n_wires, uint32_t n_rows, uint32_t logn, uint32_t n_public,
                                   const uint32_t* rowptr_a, const uint32_t* wire_a, const uint8_t* coeff_a_be,
                                   const uint32_t* rowptr_b, const uint32_t* wire_b, const uint8_t* coeff_b_be,
                                   const uint32_t* rowptr_c, const uint32_t* wire_c, const uint8_t* coeff_c_be,
                                   const uint8_t* A_be, const uint32_t* mapA, uint32_t nA,
                                   const uint8_t* B1_be, const uint8_t* B2_be, const uint32_t* mapB, uint32_t nB,
                                   const uint8_t* K_be, const uint32_t* mapK, uint32_t nK,
                                   const uint8_t* Z_be, uint32_t nZ,
                                   const uint8_t* pok_be, const uint32_t* mapPok, uint32_t nPok,
                                   const uint8_t* alpha1, const uint8_t* beta1, const uint8_t* delta1,
                                   const uint8_t* beta2, const uint8_t* delta2) {
    oracle_circuit* c = (oracle_circuit*)calloc(1, sizeof *c);
    c->n_wires = n_wires; c->n_rows = n_rows; c->logn = logn; c->n_public = n_public;
    const uint32_t* rp[3] = {rowptr_a, rowptr_b, rowptr_c};
    const uint32_t* wi[3] = {wire_a, wire_b, wire_c};
    const uint8_t* co[3] = {coeff_a_be, coeff_b_be, coeff_c_be};
    for (int m = 0; m < 3; m++) {
        size_t nnz = rp[m][n_rows];
        c->rowptr[m] = dup_u32(rp[m], n_rows + 1);
        c->wire[m] = dup_u32(wi[m], nnz);
        c->coeff[m] = (fe*)malloc(sizeof(fe) * (nnz ? nnz : 1));
        for (size_t k = 0; k < nnz; k++) fe_from_be(&c->coeff[m][k], co[m] + 32 * k, &FR);
    }
    c->A = load_g1(A_be, nA); c->mapA = dup_u32(mapA, nA); c->nA = nA;
    c->B1 = load_g1(B1_be, nB); c->mapB = dup_u32(mapB, nB); c->nB = nB;
    c->B2 = (g2_aff*)malloc(sizeof(g2_aff) * (nB ? nB : 1));
    for (size_t i = 0; i < nB; i++) g2_aff_from_be(&c->B2[i], B2_be + 128 * i);
    c->K = load_g1(K_be, nK); c->mapK = dup_u32(mapK, nK); c->nK = nK;
    c->Z = load_g1(Z_be, nZ); c->nZ = nZ;
    c->pok_basis = load_g1(pok_be, nPok); c->mapPok = dup_u32(mapPok, nPok); c->nPok = nPok;
    g1_aff_from_be(&c->alpha1, alpha1); g1_aff_from_be(&c->beta1, beta1); g1_aff_from_be(&c->delta1, delta1);
    g2_aff_from_be(&c->beta2, beta2); g2_aff_from_be(&c->delta2, delta2);
    return c;
}
void oracle_circuit_free(oracle_circuit* c) {
    if (!c) return;
    for (int m = 0; m < 3; m++) { free((void*)c->rowptr[m]); free((void*)c->wire[m]); free(c->coeff[m]); }
    free(c->A); free(c->B1); free(c->B2); free(c->K); free(c->Z); free(c->pok_basis);
    free(c->mapA); free(c->mapB); free(c->mapK); free(c->mapPok);
    free(c);
}

static void g1_scalar_mul(g1_ext* r, const g1_aff* p, const uint64_t k[4]) {
    g1_ext acc, t; g1_set_inf(&acc);
    for (int bit = 255; bit >= 0; bit--) {
        g1_dbl(&t, &acc); acc = t;
        if ((k[bit >> 6] >> (bit & 63)) & 1) g1_madd(&acc, p, 0);
    }
    *r = acc;
}
static void g2_scalar_mul(g2_ext* r, const g2_aff* p, const uint64_t k[4]) {
    g2_ext acc, t; g2_set_inf(&acc);
    for (int bit = 255; bit >= 0; bit--) {
        g2_dbl(&t, &acc); acc = t;
        if ((k[bit >> 6] >> (bit & 63)) & 1) g2_madd(&acc, p, 0);
    }
    *r = acc;
}

/* wires_be: n_wires * 32; rs_be: r | s; out: Ar(64) | Bs(128) | Krs(64) | PoK(64) = 320 bytes */
int oracle_prove_from_wires(const oracle_circuit* c, const uint8_t* wires_be, const uint8_t* rs_be, uint8_t* out) {
    size_t n = (size_t)1 << c->logn, nw = c->n_wires;
    fe* w = (fe*)malloc(sizeof(fe) * nw);
    uint64_t (*wc)[4] = malloc(32 * nw);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < nw; i++) { fe_from_be(&w[i], wires_be + 32 * i, &FR); fe_canon(wc[i], &w[i], &FR); }
    /* a, b, c = A.w, B.w, C.w */
    fe* abc = (fe*)calloc(3 * n, sizeof(fe));
    for (int m = 0; m < 3; m++) {
#pragma omp parallel for schedule(static)
        for (size_t row = 0; row < c->n_rows; row++) {
            fe acc = {{0, 0, 0, 0}}, t;
            for (uint32_t k = c->rowptr[m][row]; k < c->rowptr[m][row + 1]; k++) {
                fe_mul(&t, &c->coeff[m][k], &w[c->wire[m][k]], &FR);
                fe_add(&acc, &acc, &t, &FR);
            }
            abc[(size_t)m * n + row] = acc;
        }
    }
    compute_h(abc, abc + n, abc + 2 * n, c->logn);
    uint64_t (*hc)[4] = malloc(32 * n);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) fe_canon(hc[i], &abc[i], &FR);
    /* gather the scalar vectors */
    uint64_t (*sA)[4] = malloc(32 * (c->nA + 1)), (*sB)[4] = malloc(32 * (c->nB + 1)), (*sK)[4] = malloc(32 * (c->nK + 1)),
             (*sP)[4] = malloc(32 * (c->nPok + 1));
    for (size_t i = 0; i < c->nA; i++) memcpy(sA[i], wc[c->mapA[i]], 32);
    for (size_t i = 0; i < c->nB; i++) memcpy(sB[i], wc[c->mapB[i]], 32);
    for (size_t i = 0; i < c->nK; i++) memcpy(sK[i], wc[c->mapK[i]], 32);
    for (size_t i = 0; i < c->nPok; i++) memcpy(sP[i], wc[c->mapPok[i]], 32);
    fe r, s, rs;
    uint64_t rc[4], scn[4], nrs[4];
    fe_from_be(&r, rs_be, &FR); fe_from_be(&s, rs_be + 32, &FR);
    fe_mul(&rs, &r, &s, &FR); fe_neg(&rs, &rs, &FR);
    fe_canon(rc, &r, &FR); fe_canon(scn, &s, &FR); fe_canon(nrs, &rs, &FR);
    g1_ext ar, bs1, krs, kz, pok, t1, t2;
    g2_ext bs2, u1, u2;
    g1_msm(&ar, c->A, (const uint64_t (*)[4])sA, c->nA);
    g1_msm(&bs1, c->B1, (const uint64_t (*)[4])sB, c->nB);
    g2_msm(&bs2, c->B2, (const uint64_t (*)[4])sB, c->nB);
    g1_msm(&krs, c->K, (const uint64_t (*)[4])sK, c->nK);
    g1_msm(&kz, c->Z, (const uint64_t (*)[4])hc, c->nZ);
    g1_msm(&pok, c->pok_basis, (const uint64_t (*)[4])sP, c->nPok);
    /* Ar = A + alpha + r*delta ; Bs1 = B + beta + s*delta */
    g1_madd(&ar, &c->alpha1, 0); g1_scalar_mul(&t1, &c->delta1, rc); g1_add(&t2, &ar, &t1); ar = t2;
    g1_madd(&bs1, &c->beta1, 0); g1_scalar_mul(&t1, &c->delta1, scn); g1_add(&t2, &bs1, &t1); bs1 = t2;
    g2_madd(&bs2, &c->beta2, 0); g2_scalar_mul(&u1, &c->delta2, scn); g2_add(&u2, &bs2, &u1); bs2 = u2;
    g1_aff ar_a, bs1_a, krs_a, pok_a;
    g2_aff bs2_a;
    g1_to_aff(&ar_a, &ar); g1_to_aff(&bs1_a, &bs1);
    /* Krs = K + Z.h - rs*delta + s*Ar + r*Bs1 */
    g1_add(&t1, &krs, &kz); krs = t1;
    g1_scalar_mul(&t1, &c->delta1, nrs); g1_add(&t2, &krs, &t1); krs = t2;
    g1_scalar_mul(&t1, &ar_a, scn); g1_add(&t2, &krs, &t1); krs = t2;
    g1_scalar_mul(&t1, &bs1_a, rc); g1_add(&t2, &krs, &t1); krs = t2;
    g1_to_aff(&krs_a, &krs); g2_to_aff(&bs2_a, &bs2); g1_to_aff(&pok_a, &pok);
    g1_aff_to_be(out, &ar_a); g2_aff_to_be(out + 64, &bs2_a); g1_aff_to_be(out + 192, &krs_a); g1_aff_to_be(out + 256, &pok_a);
    free(w); free(wc); free(abc); free(hc); free(sA); free(sB); free(sK); free(sP);
    return 0;
}

/* ---- fixed-base batch multiplication: out[i] = [k_i] * base -------------------------------------------------
 * The trusted setup of the oracle (oracle/py/groth16.py setup(): gnark backend/groth16/bn254/setup.go computes
 * every key point as a multiple of the generator) and of bench.py's reference arm, which must build its own
 * proving key without touching the product library.  8-bit windows: table[w][d] = d * 2^(8w) * base. */
#define DEFINE_FIXED_BASE(G)                                                                                       \
    int oracle_fixed_base_##G(const uint8_t* base_be, const uint8_t* scalars_be, size_t n, uint8_t* out_be,       \
                              size_t pt_bytes) {                                                                  \
        G##_aff base; G##_aff_from_be(&base, base_be);                                                            \
        G##_aff* table = (G##_aff*)malloc(sizeof(G##_aff) * 32 * 256);                                            \
        G##_ext cur; G##_from_aff(&cur, &base);                                                                   \
        for (int w = 0; w < 32; w++) {                                                                            \
            G##_ext acc; G##_set_inf(&acc);                                                                       \
            memset(&table[w * 256], 0, sizeof(G##_aff));                                                          \
            for (int d = 1; d < 256; d++) { G##_ext t; G##_add(&t, &acc, &cur); acc = t; G##_to_aff(&table[w * 256 + d], &acc); } \
            G##_ext t; G##_add(&t, &acc, &cur); cur = t;                                                          \
        }                                                                                                         \
        uint64_t (*sc)[4] = malloc(32 * (n ? n : 1));                                                             \
        scalars_from_be(sc, scalars_be, n);                                                                       \
        _Pragma("omp parallel for schedule(dynamic, 64)")                                                         \
        for (size_t i = 0; i < n; i++) {                                                                          \
            G##_ext acc; G##_set_inf(&acc);                                                                       \
            for (int w = 0; w < 32; w++) {                                                                        \
                unsigned d = (unsigned)((sc[i][w >> 3] >> ((w & 7) * 8)) & 0xff);                                 \
                if (d) G##_madd(&acc, &table[w * 256 + d], 0);                                                    \
            }                                                                                                     \
            G##_aff a; G##_to_aff(&a, &acc); G##_aff_to_be(out_be + pt_bytes * i, &a);                            \
        }                                                                                                         \
        free(sc); free(table);                                                                                    \
        return 0;                                                                                                 \
    }
DEFINE_FIXED_BASE(g1)
DEFINE_FIXED_BASE(g2)


// ec.cuh -- short-Weierstrass group law (a = 0) for BN254 G1 (over Fp) and G2 (over Fp2).
//
// Device-side replacement for gnark-crypto `ecc/bn254/{g1,g2}.go` as used by
// `G1Jac.MultiExp` / `G2Jac.MultiExp` inside gnark's groth16.Prove (third-party; the
// reference only shells out to it: /root/reference/client/proof.helper.ts:64,
// SURVEY.md 8a rows a7-a12).
//
// Bucket accumulators use extended-Jacobian "XYZZ" coordinates
//   (X, Y, ZZ, ZZZ):  x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2,  ZZ == 0  <=> infinity
// because the mixed addition (XYZZ += affine) costs 8M + 2S -- the cheapest complete-
// enough formula for the Pippenger inner loop.  Affine infinity is encoded x = y = 0
// (never on the curve since b != 0), the same convention gnark-crypto serialises.
#pragma once
#include "ff.cuh"

namespace g16 {

template <class F>
struct Affine {
    F x, y;
    FF_HD bool is_inf() const { return x.is_zero() && y.is_zero(); }
    FF_HD static Affine inf() { return {F::zero(), F::zero()}; }
    FF_HD Affine neg() const { return {x, y.neg()}; }
};

template <class F>
struct XYZZ {
    F X, Y, ZZ, ZZZ;

    FF_HD static XYZZ inf() { return {F::zero(), F::zero(), F::zero(), F::zero()}; }
    FF_HD bool is_inf() const { return ZZ.is_zero(); }
    FF_HD static XYZZ from_affine(const Affine<F>& p) {
        if (p.is_inf()) return inf();
        return {p.x, p.y, F::one(), F::one()};
    }
    FF_HD XYZZ neg() const { return {X, Y.neg(), ZZ, ZZZ}; }

    // 2 * affine point (mdbl-2008-s-1)
    FF_HD static XYZZ dbl_affine(const Affine<F>& p) {
        if (p.is_inf() || p.y.is_zero()) return inf();
        F U = p.y.dbl();
        F V = U.sqr();
        F W = U * V;
        F S = p.x * V;
        F xx = p.x.sqr();
        F M = xx.dbl() + xx;
        F X3 = M.sqr() - S.dbl();
        F Y3 = M * (S - X3) - W * p.y;
        return {X3, Y3, V, W};
    }

    // dbl-2008-s-1
    FF_NOINLINE XYZZ dbl() const {
        if (is_inf() || Y.is_zero()) return inf();
        F U = Y.dbl();
        F V = U.sqr();
        F W = U * V;
        F S = X * V;
        F xx = X.sqr();
        F M = xx.dbl() + xx;
        F X3 = M.sqr() - S.dbl();
        F Y3 = M * (S - X3) - W * Y;
        return {X3, Y3, V * ZZ, W * ZZZ};
    }

    // this += p   (madd-2008-s, 8M + 2S) with the exceptional cases handled
    FF_HD void madd(const Affine<F>& p) {
        if (p.is_inf()) return;
        if (is_inf()) {
            X = p.x; Y = p.y; ZZ = F::one(); ZZZ = F::one();
            return;
        }
        F U2 = p.x * ZZ;
        F S2 = p.y * ZZZ;
        F Pq = U2 - X;
        F Rq = S2 - Y;
        if (Pq.is_zero()) {
            if (Rq.is_zero()) *this = dbl_affine(p);
            else *this = inf();
            return;
        }
        F PP = Pq.sqr();
        F PPP = Pq * PP;
        F Q = X * PP;
        F X3 = Rq.sqr() - PPP - Q.dbl();
        F Y3 = Rq * (Q - X3) - Y * PPP;
        X = X3; Y = Y3;
        ZZ = ZZ * PP;
        ZZZ = ZZZ * PPP;
    }

    // this += q   (add-2008-s, 12M + 2S)
    FF_NOINLINE void add(const XYZZ& q) {
        if (q.is_inf()) return;
        if (is_inf()) { *this = q; return; }
        F U1 = X * q.ZZ;
        F U2 = q.X * ZZ;
        F S1 = Y * q.ZZZ;
        F S2 = q.Y * ZZZ;
        F Pq = U2 - U1;
        F Rq = S2 - S1;
        if (Pq.is_zero()) {
            if (Rq.is_zero()) *this = dbl();
            else *this = inf();
            return;
        }
        F PP = Pq.sqr();
        F PPP = Pq * PP;
        F Q = U1 * PP;
        F X3 = Rq.sqr() - PPP - Q.dbl();
        F Y3 = Rq * (Q - X3) - S1 * PPP;
        X = X3; Y = Y3;
        ZZ = ZZ * q.ZZ * PP;
        ZZZ = ZZZ * q.ZZZ * PPP;
    }

    // affine normalisation -- one inversion; off the hot path
    FF_HD Affine<F> to_affine() const {
        if (is_inf()) return Affine<F>::inf();
        // 1/ZZZ, then 1/ZZ = ZZZ^-2 * ZZ^2 ... simpler: two-for-one via ZZ*ZZZ
        F t = (ZZ * ZZZ).inverse();
        F izz = t * ZZZ;
        F izzz = t * ZZ;
        return {X * izz, Y * izzz};
    }
};

typedef Affine<Fp> G1Affine;
typedef Affine<Fp2> G2Affine;
typedef XYZZ<Fp> G1XYZZ;
typedef XYZZ<Fp2> G2XYZZ;

}  // namespace g16

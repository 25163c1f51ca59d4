// hostshim.cpp -- exposes the __host__ __device__ field / curve routines of ff.cuh and
// ec.cuh through a C ABI so that the CPU test-suite (tests/test_host_arith.py) can check
// the very same algorithms the kernels run against the big-integer oracle.  Test tooling:
// nothing on the product path links this file.
#include <string.h>

#include "ec.cuh"

using namespace g16;

template <class T>
static T ld(const uint32_t* p) {
    T t;
    memcpy(&t, p, sizeof(T));
    return t;
}
template <class T>
static void st(uint32_t* p, const T& t) {
    memcpy(p, &t, sizeof(T));
}

extern "C" {

// op: 0 mul, 1 add, 2 sub, 3 neg(a), 4 inverse(a), 5 to_mont(a), 6 from_mont(a), 7 sqr(a), 8 inverse_fermat(a), 9 halve(a), 10 inverse_euclid(a)
void shim_fp_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
    Fp x = ld<Fp>(a), y = ld<Fp>(b), r;
    switch (op) {
        case 0: r = x * y; break;
        case 1: r = x + y; break;
        case 2: r = x - y; break;
        case 3: r = x.neg(); break;
        case 4: r = x.inverse(); break;
        case 5: r = x.to_mont(); break;
        case 6: r = x.from_mont(); break;
        case 8: r = x.inverse_fermat(); break;
        case 9: r = x.halve(); break;
        case 10: r = x.inverse_euclid(); break;
        default: r = x.sqr(); break;
    }
    st(out, r);
}
void shim_fr_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
    Fr x = ld<Fr>(a), y = ld<Fr>(b), r;
    switch (op) {
        case 0: r = x * y; break;
        case 1: r = x + y; break;
        case 2: r = x - y; break;
        case 3: r = x.neg(); break;
        case 4: r = x.inverse(); break;
        case 5: r = x.to_mont(); break;
        case 6: r = x.from_mont(); break;
        case 8: r = x.inverse_fermat(); break;
        case 9: r = x.halve(); break;
        case 10: r = x.inverse_euclid(); break;
        default: r = x.sqr(); break;
    }
    st(out, r);
}
// op: 0 mul, 1 sqr(a), 2 inverse(a)
void shim_fp2_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
    Fp2 x = ld<Fp2>(a), y = ld<Fp2>(b), r;
    switch (op) {
        case 0: r = x * y; break;
        case 1: r = x.sqr(); break;
        default: r = x.inverse(); break;
    }
    st(out, r);
}

// acc (XYZZ, 32 words) op= operand; op: 0 madd(affine 16 words), 1 add(XYZZ), 2 dbl, 3 to_affine -> out 16 words
void shim_g1_op(int op, uint32_t* acc, const uint32_t* operand, uint32_t* out_affine) {
    G1XYZZ a = ld<G1XYZZ>(acc);
    switch (op) {
        case 0: a.madd(ld<G1Affine>(operand)); break;
        case 1: a.add(ld<G1XYZZ>(operand)); break;
        case 2: a = a.dbl(); break;
        default: st(out_affine, a.to_affine()); return;
    }
    st(acc, a);
}
void shim_g2_op(int op, uint32_t* acc, const uint32_t* operand, uint32_t* out_affine) {
    G2XYZZ a = ld<G2XYZZ>(acc);
    switch (op) {
        case 0: a.madd(ld<G2Affine>(operand)); break;
        case 1: a.add(ld<G2XYZZ>(operand)); break;
        case 2: a = a.dbl(); break;
        default: st(out_affine, a.to_affine()); return;
    }
    st(acc, a);
}
}

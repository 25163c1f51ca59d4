// ntt.cuh -- batched radix-2 number-theoretic transforms over BN254 Fr and the Groth16
// quotient polynomial H(x).
//
// Replaces gnark-crypto `fr/fft` (Domain.FFT / FFTInverse with DIF/DIT and OnCoset) and gnark
// `computeH` in backend/groth16/bn254/prove.go -- SURVEY.md 3.2 step 3, 8a row a6 (third-party
// Go reached through `sunspot prove`, /root/reference/client/proof.helper.ts:64).
//
// Layout: a vector is n = 2^logn Fr elements (8x32-bit limbs, Montgomery), `batch` vectors back
// to back.  A transform is a few PASSES; one pass runs k <= 8 consecutive butterfly stages on a
// tile of 2^k x C elements staged in shared memory (limb-major / SoA so that a warp touching 32
// consecutive elements hits 32 banks), and reads + writes every element exactly once:
//   * "strided" passes own the high stages: a tile is 2^k rows spaced 2^s_lo apart, each row
//     C contiguous elements (C*32 B contiguous per row: whole 32-B sectors, no partial lines);
//   * the "contiguous" pass owns stages s < k_c: a tile is C groups of 2^k_c adjacent elements.
//   DIF (natural -> bit-reversed) walks stages high -> low, DIT (bit-reversed -> natural) low ->
//   high, so no transposition or bit-reversal pass is ever needed; scaling tables (coset powers,
//   1/n) are folded into the first load or the last store of a transform.
// At n = 2^22 a transform moves 3 x 2 x 128 MiB while doing 46 M modular multiplications:
// it is bound by the integer pipe, not by HBM (SURVEY.md 7 "Roofline honesty for NTT").
#pragma once
#include <map>

#include "common.cuh"
#include "ff.cuh"

namespace g16 {

enum NttDir { NTT_DIF = 0, NTT_DIT = 1 };

struct NttDomain {
    unsigned logn = 0;
    size_t n = 0;
    Fr* tw_fwd = nullptr;      // w^k, k < n/2
    Fr* tw_inv = nullptr;      // w^-k
    Fr* coset_nat = nullptr;   // g^j                  indexed by natural position j
    Fr* cosetinv_nat = nullptr;  // g^-j / n           indexed by natural position j
    Fr* coset_br = nullptr;    // g^rev(p) / n         indexed by bit-reversed position p
    Fr* cosetinv_br = nullptr; // g^-rev(p) / n        indexed by bit-reversed position p
    Fr* ninv_const = nullptr;  // n copies of 1/n (post-scale of a plain inverse)
    Fr h_den;                  // (g^n - 1)^-1, Montgomery (host copy)
};

class NttEngine {
   public:
    ~NttEngine() { release(); }
    // builds (once) and returns the tables of the size-2^logn domain
    int domain(unsigned logn, cudaStream_t st, const NttDomain** out);
    // in-place transform of `batch` vectors.  pre/post: optional element-wise factors indexed by
    // the position of the element in the input (pre) / output (post) array.
    int run(Fr* d_data, unsigned logn, size_t batch, NttDir dir, bool inverse_twiddles, const Fr* pre,
            const Fr* post, cudaStream_t st);
    // same with an explicit distance (in elements) between consecutive vectors
    int run_strided(Fr* d_data, size_t vec_stride, unsigned logn, size_t batch, NttDir dir, bool inverse_twiddles,
                    const Fr* pre, const Fr* post, cudaStream_t st);
    // quotient: a,b,c = 3 consecutive vectors per proof (evaluations of A.w, B.w, C.w on the
    // domain, natural order).  On return the FIRST vector of each triple holds the coefficients
    // of H in bit-reversed order (gnark's order, matching pk.G1.Z).
    int compute_h(Fr* d_abc, unsigned logn, size_t nproofs, cudaStream_t st);
    void release();
    int launches = 0;

   private:
    std::map<unsigned, NttDomain> domains;
    bool attr_done = false;   // dynamic-shared-memory opt-in done on this engine's device
};

}  // namespace g16

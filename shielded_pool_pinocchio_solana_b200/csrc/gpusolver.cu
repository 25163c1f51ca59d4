// gpusolver.cu -- the R1CS witness solve (gnark `constraint/bn254/solver.go`, SURVEY.md 3.2 step 1,
// 8a row a3, 8(f) rank 1) on the GPU, for a whole batch of proofs at once.
//
// gnark walks the instruction levels with goroutines; on a 16-core host that costs 10-20 ms per
// audit-size proof and caps a box at a few hundred proofs/s -- less than ONE B200 proves.  Here the
// level structure is compiled once per circuit into a static plan (which wire each row defines is
// independent of the witness) and executed for a whole group of proofs, wires in HBM in the layout the
// prover reads.  Three kernels share the plan (GpuSolverPlan::run picks per level range):
//   k_solve_2p      wide levels: CTA per proof; the non-unit products of a (sub-)level by one thread each into
//                   shared memory, then one thread per row sums its operands and finishes the row (default)
//   k_solve_narrow  runs of thin levels (the division chain at the end of the withdraw circuit): thread per proof
//   k_solve_tpi     the previous wide kernel: four lanes per row (G16_SOLVER_TPI=1); k_solve_levels: round 1's
//                   thread per row (G16_SOLVER_CTA=1).  Both kept as cross-checks (tests/test_parity_at_size_gpu.py).
// Alone the solve is latency-bound (~10^3 dependent levels); beside the prover it also competes for issue slots,
// which is why the default kernels issue as few instructions as they can (pipeline section).
// The BSB22 commitment splits the plan in two phases (prove.cu runs the commitment MSM between).
// Circuits using a hint this file does not implement keep the host solver (solver.cpp).
#include "gpusolver.cuh"

#include <string.h>

#include <algorithm>

namespace g16 {

namespace {

constexpr uint32_t NO_WIRE = 0xFFFFFFFEu;
constexpr int SOLVE_THREADS = 128;

__device__ __forceinline__ Fr lin_eval(const uint32_t* __restrict__ terms, uint32_t n, const Fr* __restrict__ w,
                                       const Fr* __restrict__ coeffs, uint32_t skip, int unit_ids) {
    Fr acc = Fr::zero();
    for (uint32_t k = 0; k < n; k++) {
        uint32_t cid = terms[2 * k], wid = terms[2 * k + 1];
        if (wid == skip) continue;
        if (wid == CCS_CONST_WIRE) {
            acc = acc + coeffs[cid];
            continue;
        }
        Fr x = w[wid];
        if (unit_ids && cid == 1) acc = acc + x;
        else if (unit_ids && cid == 3) acc = acc - x;
        else if (!(unit_ids && cid == 0)) acc = acc + coeffs[cid] * x;
    }
    return acc;
}

__device__ __forceinline__ uint32_t canon_bits(const Fr& c, uint32_t lo, uint32_t nbits) {
    // bits [lo, lo+nbits) of a canonical 256-bit value, nbits <= 32
    if (lo >= 256) return 0;
    uint32_t wi = lo >> 5, sh = lo & 31;
    uint64_t v = c.v[wi];
    if (wi + 1 < 8) v |= (uint64_t)c.v[wi + 1] << 32;
    v >>= sh;
    return nbits >= 32 ? (uint32_t)v : (uint32_t)v & ((1u << nbits) - 1u);
}

__device__ __forceinline__ Fr fr_from_u32(uint32_t x) {
    Fr f = Fr::zero();
    f.v[0] = x;
    return f.to_mont();
}

__device__ void run_hint(uint32_t kind, const uint32_t* __restrict__ cd, Fr* __restrict__ w,
                         const Fr* __restrict__ coeffs, int unit_ids, const Fr& blinder, uint32_t* err) {
    const uint32_t nin = cd[2];
    // locate the output range: walk the input expressions
    uint32_t p = 3;
    for (uint32_t i = 0; i < nin; i++) p += 1 + 2 * cd[p];
    const uint32_t o0 = cd[p], o1 = cd[p + 1], nout = o1 - o0;
    auto input = [&](uint32_t idx) {
        uint32_t q = 3;
        for (uint32_t i = 0; i < idx; i++) q += 1 + 2 * cd[q];
        return lin_eval(cd + q + 1, cd[q], w, coeffs, NO_WIRE, unit_ids);
    };
    switch (kind) {
        case HINT_NBITS: {
            Fr c = input(0).from_mont();
            for (uint32_t k = 0; k < nout; k++) w[o0 + k] = canon_bits(c, k, 1) ? Fr::one() : Fr::zero();
            break;
        }
        case HINT_INVZERO:
            w[o0] = input(0).inverse();
            break;
        case HINT_DECOMPOSE: {
            Fr ls = input(1).from_mont(), c = input(2).from_mont();
            uint32_t limb = ls.v[0];
            if (limb == 0 || limb > 32 || (ls.v[1] | ls.v[2] | ls.v[3] | ls.v[4] | ls.v[5] | ls.v[6] | ls.v[7])) {
                atomicMin(err, 0x80000000u | 1u);
                break;
            }
            for (uint32_t k = 0; k < nout; k++) w[o0 + k] = fr_from_u32(canon_bits(c, k * limb, limb));
            break;
        }
        case HINT_COUNT: {
            // logderivarg.countHint with one column: out[j] = #queries equal to table row j
            Fr a = input(0).from_mont();
            uint32_t size = a.v[0];
            for (uint32_t k = 0; k < nout; k++) w[o0 + k] = Fr::zero();
            // table rows are inputs 2..2+size, queries follow; walk the expressions once
            uint32_t q = 3;
            for (uint32_t i = 0; i < 2; i++) q += 1 + 2 * cd[q];
            uint32_t table_q = q;
            for (uint32_t i = 0; i < size; i++) q += 1 + 2 * cd[q];
            Fr one = Fr::one();
            // range-check tables are the values 0..size-1 in order: a query v then counts for row v (no other row
            // holds v, so this is the first match the scan would find); any other table takes the scan
            bool identity = q - table_q == 3 * size;
            {
                Fr cur = Fr::zero();
                for (uint32_t j = 0; j < size && identity; j++) {
                    identity = cd[table_q + 3 * j] == 1 && lin_eval(cd + table_q + 3 * j + 1, 1, w, coeffs, NO_WIRE, unit_ids) == cur;
                    cur = cur + one;
                }
            }
            for (uint32_t qi = 2 + size; qi < nin; qi++) {
                Fr val = lin_eval(cd + q + 1, cd[q], w, coeffs, NO_WIRE, unit_ids);
                q += 1 + 2 * cd[q];
                if (identity) {
                    const Fr c = val.from_mont();
                    if (c.v[0] < size && !(c.v[1] | c.v[2] | c.v[3] | c.v[4] | c.v[5] | c.v[6] | c.v[7])) {
                        w[o0 + c.v[0]] = w[o0 + c.v[0]] + one;
                        continue;
                    }
                }
                uint32_t t = table_q;
                bool found = false;
                for (uint32_t j = 0; j < size; j++) {
                    Fr row = lin_eval(cd + t + 1, cd[t], w, coeffs, NO_WIRE, unit_ids);
                    t += 1 + 2 * cd[t];
                    if (row == val) {
                        w[o0 + j] = w[o0 + j] + one;
                        found = true;
                        break;
                    }
                }
                if (!found) atomicMin(err, 0x80000000u | 2u);
            }
            break;
        }
        case HINT_RANDOMIZE:
            for (uint32_t k = 0; k < nout; k++) w[o0 + k] = blinder;
            break;
        default:
            atomicMin(err, 0x80000000u | 3u);
    }
}

// One CTA per proof; levels [lvl_begin, lvl_end).
__global__ void __launch_bounds__(SOLVE_THREADS)
k_solve_levels(const uint32_t* __restrict__ lvl_off, const uint32_t* __restrict__ lvl_instr,
               const uint4* __restrict__ info, const uint32_t* __restrict__ instr_cd,
               const uint32_t* __restrict__ calldata, const Fr* __restrict__ coeffs,
               const Fr* __restrict__ coeff_invs, Fr* __restrict__ wires, size_t wstride, size_t blinder_slot,
               uint32_t lvl_begin, uint32_t lvl_end, int unit_ids, uint32_t* __restrict__ err) {
    const uint32_t b = blockIdx.x;
    Fr* w = wires + (size_t)b * wstride;
    uint32_t* e = err + b;
    for (uint32_t lv = lvl_begin; lv < lvl_end; lv++) {
        const uint32_t s = lvl_off[lv], t = lvl_off[lv + 1];
        for (uint32_t k = s + threadIdx.x; k < t; k += SOLVE_THREADS) {
            const uint32_t ins = lvl_instr[k];
            const uint4 inf = info[ins];
            const uint32_t* cd = calldata + instr_cd[ins];
            if (inf.x >= 16) {
                run_hint(inf.x - 16, cd, w, coeffs, unit_ids, w[blinder_slot], e);
                continue;
            }
            const uint32_t nl = cd[1], nr = cd[2], no = cd[3];
            const uint32_t skip = inf.x == 0 ? NO_WIRE : inf.y;
            Fr L = lin_eval(cd + 4, nl, w, coeffs, skip, unit_ids);
            Fr Rr = lin_eval(cd + 4 + 2 * nl, nr, w, coeffs, skip, unit_ids);
            Fr O = lin_eval(cd + 4 + 2 * (nl + nr), no, w, coeffs, skip, unit_ids);
            if (inf.x == 0) {
                if (L * Rr != O) atomicMin(e, inf.w + 1);           // constraint row (1-based)
            } else if (inf.x == 1) {
                w[inf.y] = (L * Rr - O) * coeff_invs[inf.z];
            } else if (inf.x == 2) {
                // zero divisor: gnark leaves the wire at 0 and only checks the row (0 == O)
                if (Rr.is_zero()) {
                    if (!O.is_zero()) atomicMin(e, inf.w + 1);
                    w[inf.y] = Fr::zero();
                } else w[inf.y] = (O * Rr.inverse() - L) * coeff_invs[inf.z];
            } else {
                if (L.is_zero()) {
                    if (!O.is_zero()) atomicMin(e, inf.w + 1);
                    w[inf.y] = Fr::zero();
                } else w[inf.y] = (O * L.inverse() - Rr) * coeff_invs[inf.z];
            }
        }
        __syncthreads();
    }
}

// ---- term-parallel variant (default) ------------------------------------------------------------------
// A level is a handful of rows (median ~20), a row ~10 terms, a term one modmul of ~600 cycles latency: with
// one thread per row the level costs the SUM of its terms' latencies.  Here SOLVE_TPI lanes share a row: each
// takes every SOLVE_TPI-th term (descriptors, wires and coefficients of up to four terms are in flight
// together), the partial sums of L, R and O meet through two xor-shuffles and the first lane finishes the row.
#ifndef SOLVE_TPI_LANES
#define SOLVE_TPI_LANES 4
#endif
#ifndef SOLVE_TPI_THREADS
#define SOLVE_TPI_THREADS 128
#endif
constexpr int SOLVE_TPI = SOLVE_TPI_LANES;
constexpr int SOLVE_TT = SOLVE_TPI_THREADS;             // threads per proof of the term-parallel kernel
constexpr int SOLVE_ROWS = SOLVE_TT / SOLVE_TPI;        // rows per pass of a CTA

__device__ __forceinline__ Fr shfl_xor_fr(uint32_t mask, const Fr& a, int d) {
    Fr r;
#pragma unroll
    for (int l = 0; l < 8; l++) r.v[l] = __shfl_xor_sync(mask, a.v[l], d);
    return r;
}

// rec[2k]   = (mode, defined wire, coefficient id of the defined wire's term, constraint row)
// rec[2k+1] = (calldata offset of the instruction, nL, nR, nO)          k = position in level order
#ifndef SOLVE_MIN_CTAS
#define SOLVE_MIN_CTAS 1
#endif
__global__ void __launch_bounds__(SOLVE_TT, SOLVE_MIN_CTAS)
k_solve_tpi(const uint32_t* __restrict__ lvl_off, const uint4* __restrict__ rec, const uint32_t* __restrict__ calldata,
            const Fr* __restrict__ coeffs, const Fr* __restrict__ coeff_invs, Fr* wires, size_t wstride,
            size_t blinder_slot, uint32_t lvl_begin, uint32_t lvl_end, int unit_ids, uint32_t* err) {
    const uint32_t b = blockIdx.x;
    const uint32_t grp = threadIdx.x / SOLVE_TPI, sub = threadIdx.x % SOLVE_TPI;
    const uint32_t gmask = ((1u << SOLVE_TPI) - 1u) << ((threadIdx.x & 31u) & ~(uint32_t)(SOLVE_TPI - 1));
    Fr* w = wires + (size_t)b * wstride;
    uint32_t* e = err + b;
    for (uint32_t lv = lvl_begin; lv < lvl_end; lv++) {
        const uint32_t s = lvl_off[lv], t = lvl_off[lv + 1];
        for (uint32_t k = s + grp; k < t; k += SOLVE_ROWS) {   // uniform over the SOLVE_TPI lanes of a row
            const uint4 inf = rec[2 * k], shp = rec[2 * k + 1];
            const uint32_t* cd = calldata + shp.x;
            if (inf.x >= 16) {
                if (sub == 0) run_hint(inf.x - 16, cd, w, coeffs, unit_ids, w[blinder_slot], e);
                continue;
            }
            const uint32_t nl = shp.y, nlr = shp.y + shp.z, nt = nlr + shp.w;
            const uint32_t skip = inf.x == 0 ? NO_WIRE : inf.y;
            const uint32_t* terms = cd + 4;
            Fr L = Fr::zero(), Rr = Fr::zero(), O = Fr::zero();
            for (uint32_t i0 = sub; i0 < nt; i0 += 4 * SOLVE_TPI) {
                uint32_t cid[4], wid[4];
                Fr x[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const uint32_t i = i0 + j * SOLVE_TPI;
                    cid[j] = i < nt ? terms[2 * i] : 0u;
                    wid[j] = i < nt ? terms[2 * i + 1] : skip;
                }
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    x[j] = Fr::zero();
                    if (i0 + j * SOLVE_TPI < nt && wid[j] != skip && wid[j] != CCS_CONST_WIRE) x[j] = w[wid[j]];
                }
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const uint32_t i = i0 + j * SOLVE_TPI;
                    if (i >= nt || wid[j] == skip) continue;
                    Fr v;
                    if (wid[j] == CCS_CONST_WIRE) v = coeffs[cid[j]];
                    else if (unit_ids && cid[j] == 1) v = x[j];
                    else if (unit_ids && cid[j] == 3) v = x[j].neg();
                    else if (unit_ids && cid[j] == 0) continue;
                    else v = coeffs[cid[j]] * x[j];
                    if (i < nl) L = L + v;
                    else if (i < nlr) Rr = Rr + v;
                    else O = O + v;
                }
            }
#pragma unroll
            for (int d = 1; d < SOLVE_TPI; d <<= 1) {
                if (nl > 1) L = L + shfl_xor_fr(gmask, L, d);    // a single L term sits in lane 0 of the row
                Rr = Rr + shfl_xor_fr(gmask, Rr, d);
                O = O + shfl_xor_fr(gmask, O, d);
            }
            if (sub != 0) continue;
            if (inf.x == 0) {
                if (L * Rr != O) atomicMin(e, inf.w + 1);           // constraint row (1-based)
            } else if (inf.x == 1) {
                // the defined wire almost always enters its row with coefficient +1 or -1: no division then
                const Fr v = L * Rr - O;
                w[inf.y] = (unit_ids && inf.z == 1) ? v : ((unit_ids && inf.z == 3) ? v.neg() : v * coeff_invs[inf.z]);
            } else if (inf.x == 2) {
                // zero divisor: gnark leaves the wire at 0 and only checks the row (0 == O)
                if (Rr.is_zero()) {
                    if (!O.is_zero()) atomicMin(e, inf.w + 1);
                    w[inf.y] = Fr::zero();
                } else w[inf.y] = (O * Rr.inverse() - L) * coeff_invs[inf.z];
            } else {
                if (L.is_zero()) {
                    if (!O.is_zero()) atomicMin(e, inf.w + 1);
                    w[inf.y] = Fr::zero();
                } else w[inf.y] = (O * L.inverse() - Rr) * coeff_invs[inf.z];
            }
        }
        __syncthreads();
    }
}

// ---- two-phase variant (default for wide levels) ------------------------------------------------------------
// Beside the prover the solve is not bound by its latency but by the instructions it issues (the step is issue-bound
// and k_solve_tpi executes ~5 M warp instructions per proof: its four lanes per row all run the multiply and the three
// side additions whenever one lane of the warp needs them).  Here the plan is compiled further on the host:
//   * every term whose coefficient is not 0 / +1 / -1 becomes a PRODUCT of its (sub-)level: phase A computes them with
//     one thread per product -- full warps of multiplies and nothing else -- into shared memory;
//   * a row keeps, per side, a list of (code, ref) operands: a product slot, +wire, -wire or a constant; zero terms and
//     the term of the wire the row defines are dropped at compile time.  Phase B: one thread per row adds the operands
//     of L, R and O and finishes the row.
// Levels with more products than the shared-memory buffer are split into sub-levels (rows of one level are
// independent).  Two barriers per sub-level.
constexpr int SOLVE2_THREADS = 128;
constexpr uint32_t SOLVE2_PRODUCTS = 1024;   // 32 KB of shared memory
enum : uint32_t { OPD_PRODUCT = 0, OPD_PLUS = 1, OPD_CONST = 2, OPD_MINUS = 3, OPD_INLINE = 4 };

__device__ __forceinline__ Fr side_sum(const uint2* __restrict__ ops, uint32_t n, const Fr* w, const Fr* __restrict__ coeffs,
                                       const Fr* prod, const uint2* __restrict__ inline_prods) {
    Fr acc = Fr::zero();
#pragma unroll 2
    for (uint32_t i = 0; i < n; i++) {
        const uint2 d = ops[i];
        if (d.x == OPD_INLINE) {   // a row with more products than the shared buffer holds multiplies in place
            const uint2 pr = inline_prods[d.y];
            acc = acc + coeffs[pr.x] * w[pr.y];
            continue;
        }
        const Fr* src = d.x == OPD_PRODUCT ? prod + d.y : (d.x == OPD_CONST ? coeffs + d.y : w + d.y);
        const Fr v = *src;
        acc = acc + (d.x == OPD_MINUS ? v.neg() : v);
    }
    return acc;
}

__global__ void __launch_bounds__(SOLVE2_THREADS, 4)
k_solve_2p(const uint32_t* __restrict__ sub_off, const uint32_t* __restrict__ prod_off, const uint2* __restrict__ prods,
           const uint4* __restrict__ rec, const uint2* __restrict__ ops, const uint2* __restrict__ inline_prods,
           const uint32_t* __restrict__ calldata, const Fr* __restrict__ coeffs, const Fr* __restrict__ coeff_invs, Fr* wires, size_t wstride, size_t blinder_slot,
           uint32_t sub_begin, uint32_t sub_end, int unit_ids, uint32_t* err) {
    __shared__ Fr prod[SOLVE2_PRODUCTS];
    const uint32_t b = blockIdx.x;
    Fr* w = wires + (size_t)b * wstride;
    uint32_t* e = err + b;
    for (uint32_t sl = sub_begin; sl < sub_end; sl++) {
        const uint32_t p0 = prod_off[sl], p1 = prod_off[sl + 1];
        for (uint32_t t = p0 + threadIdx.x; t < p1; t += SOLVE2_THREADS) {
            const uint2 d = prods[t];
            prod[t - p0] = coeffs[d.x] * w[d.y];
        }
        if (p1 > p0) __syncthreads();
        const uint32_t s = sub_off[sl], t1 = sub_off[sl + 1];
        for (uint32_t k = s + threadIdx.x; k < t1; k += SOLVE2_THREADS) {
            const uint4 inf = rec[2 * k], shp = rec[2 * k + 1];
            if (inf.x >= 16) {
                run_hint(inf.x - 16, calldata + shp.x, w, coeffs, unit_ids, w[blinder_slot], e);
                continue;
            }
            const uint2* o = ops + shp.x;
            const Fr L = side_sum(o, shp.y, w, coeffs, prod, inline_prods);
            const Fr Rr = side_sum(o + shp.y, shp.z, w, coeffs, prod, inline_prods);
            const Fr O = side_sum(o + shp.y + shp.z, shp.w, w, coeffs, prod, inline_prods);
            if (inf.x == 0) {
                if (L * Rr != O) atomicMin(e, inf.w + 1);           // constraint row (1-based)
            } else if (inf.x == 1) {
                const Fr v = L * Rr - O;
                w[inf.y] = inf.z == 1 ? v : (inf.z == 3 ? v.neg() : v * coeff_invs[inf.z >> 2]);
            } else {
                // division rows: zero divisor -> gnark leaves the wire at 0 and only checks the row (0 == O)
                const Fr den = inf.x == 2 ? Rr : L, oth = inf.x == 2 ? L : Rr;
                if (den.is_zero()) {
                    if (!O.is_zero()) atomicMin(e, inf.w + 1);
                    w[inf.y] = Fr::zero();
                } else {
                    const Fr v = O * den.inverse() - oth;
                    w[inf.y] = inf.z == 1 ? v : (inf.z == 3 ? v.neg() : v * coeff_invs[inf.z >> 2]);
                }
            }
        }
        __syncthreads();
    }
}

// ---- thread-per-proof variant for runs of thin levels -----------------------------------------------------
// The tail of the reference withdraw circuit is one dependent chain: ~650 levels of 1-6 short rows, every second
// or third of them a division (the affine point additions of the in-circuit scalar multiplication).  A CTA per
// proof would sit on an SM for the whole chain with one lane working, and (measured) takes the registers of one
// bucket-accumulation CTA of the concurrently running MSMs with it.  Here a THREAD owns a proof and walks the rows
// of [lvl_begin, lvl_end) in plan order (no barrier: program order is the dependency order), so a 256-proof group
// is 8 warps; the plan is the same for every lane and Fr::inverse() has no data-dependent branch, so the 32 proofs
// of a warp stay in lockstep.  Wire loads are one 32-byte sector per lane.
constexpr int NARROW_THREADS = 32;
__global__ void __launch_bounds__(NARROW_THREADS)
k_solve_narrow(const uint32_t* __restrict__ lvl_off, const uint4* __restrict__ rec, const uint32_t* __restrict__ calldata,
               const Fr* __restrict__ coeffs, const Fr* __restrict__ coeff_invs, Fr* wires, size_t wstride,
               size_t blinder_slot, uint32_t lvl_begin, uint32_t lvl_end, int unit_ids, uint32_t* err, uint32_t B) {
    const uint32_t b = blockIdx.x * NARROW_THREADS + threadIdx.x;
    if (b >= B) return;
    Fr* w = wires + (size_t)b * wstride;
    uint32_t* e = err + b;
    const uint32_t k_end = lvl_off[lvl_end];
    for (uint32_t k = lvl_off[lvl_begin]; k < k_end; k++) {
        const uint4 inf = rec[2 * k], shp = rec[2 * k + 1];
        const uint32_t* cd = calldata + shp.x;
        if (inf.x >= 16) {
            run_hint(inf.x - 16, cd, w, coeffs, unit_ids, w[blinder_slot], e);
            continue;
        }
        const uint32_t nl = shp.y, nlr = shp.y + shp.z, nt = nlr + shp.w;
        const uint32_t skip = inf.x == 0 ? NO_WIRE : inf.y;
        const uint32_t* terms = cd + 4;
        Fr L = Fr::zero(), Rr = Fr::zero(), O = Fr::zero();
        for (uint32_t i0 = 0; i0 < nt; i0 += 4) {
            uint32_t cid[4], wid[4];
            Fr x[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t i = i0 + j;
                cid[j] = i < nt ? terms[2 * i] : 0u;
                wid[j] = i < nt ? terms[2 * i + 1] : skip;
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                x[j] = Fr::zero();
                if (i0 + j < nt && wid[j] != skip && wid[j] != CCS_CONST_WIRE) x[j] = w[wid[j]];
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t i = i0 + j;
                if (i >= nt || wid[j] == skip) continue;
                Fr v;
                if (wid[j] == CCS_CONST_WIRE) v = coeffs[cid[j]];
                else if (unit_ids && cid[j] == 1) v = x[j];
                else if (unit_ids && cid[j] == 3) v = x[j].neg();
                else if (unit_ids && cid[j] == 0) continue;
                else v = coeffs[cid[j]] * x[j];
                if (i < nl) L = L + v;
                else if (i < nlr) Rr = Rr + v;
                else O = O + v;
            }
        }
        if (inf.x == 0) {
            if (L * Rr != O) atomicMin(e, inf.w + 1);
        } else if (inf.x == 1) {
            const Fr v = L * Rr - O;
            w[inf.y] = (unit_ids && inf.z == 1) ? v : ((unit_ids && inf.z == 3) ? v.neg() : v * coeff_invs[inf.z]);
        } else {
            // division rows: zero divisor -> gnark leaves the wire at 0 and only checks the row (0 == O)
            const Fr den = inf.x == 2 ? Rr : L, oth = inf.x == 2 ? L : Rr;
            if (den.is_zero()) {
                if (!O.is_zero()) atomicMin(e, inf.w + 1);
                w[inf.y] = Fr::zero();
            } else w[inf.y] = (O * den.inverse() - oth) * coeff_invs[inf.z];
        }
    }
}

// assignment (big-endian canonical) -> wires[b][1..nin] (Montgomery), wire 0 = 1, X_* slots from rnd
__global__ void __launch_bounds__(256)
k_assign(const uint8_t* __restrict__ asg_be, const uint8_t* __restrict__ rnd_be, uint32_t nin, Fr* __restrict__ wires,
         size_t wstride, size_t nw) {
    const uint32_t b = blockIdx.y;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    Fr* w = wires + (size_t)b * wstride;
    auto load_be = [](const uint8_t* p) {
        Fr f;
#pragma unroll
        for (int l = 0; l < 8; l++) {
            const uint8_t* q = p + 28 - 4 * l;
            f.v[l] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3];
        }
        // reduce values >= r (inputs are 256-bit): at most 5 subtractions
        uint32_t m[8], d[8];
        Fr::modulus(m);
        for (int it = 0; it < 6; it++) {
            if (ff_sub8(d, f.v, m)) break;
#pragma unroll
            for (int l = 0; l < 8; l++) f.v[l] = d[l];
        }
        return f.to_mont();
    };
    if (i < nin) w[1 + i] = load_be(asg_be + ((size_t)b * nin + i) * 32);
    if (i == 0) {
        w[0] = Fr::one();
        Fr r = load_be(rnd_be + (size_t)b * 96), s = load_be(rnd_be + (size_t)b * 96 + 32),
           bl = load_be(rnd_be + (size_t)b * 96 + 64);
        w[nw + X_ONE] = Fr::one();
        w[nw + X_R] = r;
        w[nw + X_S] = s;
        w[nw + X_NEG_RS] = (r * s).neg();
        w[nw + X_BLINDER] = bl;
        for (int k = X_BLINDER + 1; k < X_COUNT; k++) w[nw + k] = Fr::zero();
    }
}

__global__ void k_scatter_wires(Fr* wires, size_t wstride, const uint32_t* __restrict__ ids, const Fr* __restrict__ values,
                                uint32_t nids, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nids * n) return;
    uint32_t b = i / nids, k = i % nids;
    wires[(size_t)b * wstride + ids[k]] = values[i];
}

__global__ void k_set_wire(Fr* wires, size_t wstride, uint32_t wire, const Fr* __restrict__ values, uint32_t n) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < n) wires[(size_t)b * wstride + wire] = values[b];
}

template <class T>
int upload(const std::vector<T>& v, T** d, cudaStream_t st) {
    G16_CUDA(cudaMalloc(d, sizeof(T) * (v.size() ? v.size() : 1)));
    if (!v.empty()) G16_CUDA(cudaMemcpyAsync(*d, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice, st));
    return G16_OK;
}

}  // namespace

void GpuSolverPlan::release() {
    cudaFree(d_lvl_off); cudaFree(d_lvl_instr); cudaFree(d_info); cudaFree(d_instr_cd); cudaFree(d_calldata);
    cudaFree(d_coeff_invs);
    cudaFree(d_rec);
    d_rec = nullptr;
    cudaFree(d_sub_off); cudaFree(d_prod_off); cudaFree(d_prods); cudaFree(d_ops); cudaFree(d_rec2); cudaFree(d_iprods);
    d_sub_off = d_prod_off = nullptr;
    d_prods = d_ops = d_iprods = nullptr;
    d_rec2 = nullptr;
    lvl_to_sub.clear();
    cudaFree(d_host_wires);
    d_host_wires = nullptr;
    host_hints.clear();
    host_wires.clear();
    host_inputs.clear();
    segments.clear();
    commit_hashed.clear();
    d_lvl_off = d_lvl_instr = d_instr_cd = d_calldata = nullptr;
    d_info = nullptr;
    d_coeff_invs = nullptr;
    valid = false;
}

int GpuSolverPlan::build(const Circuit& c, cudaStream_t st, std::string* why_not) {
    release();
    const uint32_t nw = c.nb_wires();
    std::vector<uint8_t> known(nw, 0), host_known(nw, 0);
    for (uint32_t i = 0; i < c.nb_public + c.nb_secret; i++) known[i] = host_known[i] = 1;
    const size_t ninstr = c.blueprint.size();
    std::vector<uint4> info(ninstr, make_uint4(0, 0, 0, 0));
    std::vector<uint32_t> lvl_off(1, 0), lvl_instr, instr_cd(ninstr);
    for (size_t i = 0; i < ninstr; i++) {
        if (c.start_calldata[i] > 0xffffffffull) {
            *why_not = "calldata larger than 2^32 words";
            return G16_OK;
        }
        instr_cd[i] = (uint32_t)c.start_calldata[i];
    }
    commit_level = (uint32_t)-1;
    size_t commits_seen = 0;
    for (size_t lv = 0; lv < c.levels.size(); lv++) {
        std::vector<uint32_t> newly;
        for (uint32_t ins : c.levels[lv]) {
            const uint32_t* cd = c.calldata.data() + c.start_calldata[ins];
            if (c.blueprint[ins] == 1) {
                uint32_t cnt[3] = {cd[1], cd[2], cd[3]};
                uint32_t unk = NO_WIRE, unk_terms = 0, unk_side = 0, unk_cid = 0;
                size_t p = 4;
                bool multi = false;
                for (int side = 0; side < 3; side++)
                    for (uint32_t k = 0; k < cnt[side]; k++, p += 2) {
                        uint32_t wid = cd[p + 1];
                        if (wid == CCS_CONST_WIRE || known[wid]) continue;
                        if (unk != NO_WIRE && wid != unk) multi = true;
                        unk = wid;
                        unk_terms++;
                        unk_side = side;
                        unk_cid = cd[p];
                    }
                if (multi || unk_terms > 1) {
                    *why_not = "a row defines its wire through more than one term";
                    return G16_OK;
                }
                uint32_t mode = unk == NO_WIRE ? 0 : (unk_side == 2 ? 1 : (unk_side == 0 ? 2 : 3));
                info[ins] = make_uint4(mode, unk, unk_cid, c.constraint_offset[ins]);
                if (unk != NO_WIRE) newly.push_back(unk);
                lvl_instr.push_back(ins);
            } else {
                uint32_t hid = cd[1], nin = cd[2];
                auto kit = c.hint_kinds.find(hid);
                HintKind kind = kit == c.hint_kinds.end() ? HINT_UNKNOWN : kit->second;
                size_t p = 3;
                for (uint32_t i = 0; i < nin; i++) p += 1 + 2 * (size_t)cd[p];
                uint32_t o0 = cd[p], o1 = cd[p + 1];
                for (uint32_t wv = o0; wv < o1 && wv < nw; wv++) newly.push_back(wv);
                if (kind == HINT_COMMIT) {
                    if (commits_seen++ || c.commitments.empty()) {
                        *why_not = "more than one commitment";
                        return G16_OK;
                    }
                    // public wires committed to (api.Commit over public inputs): hashed into the challenge on the host
                    // next to the commitment point; their expressions must hang off the circuit's inputs
                    {
                        const size_t nh = c.commitments[0].public_and_commitment_committed.size();
                        if (nin < 1 + nh) {
                            *why_not = "commitment hint with fewer inputs than hashed wires";
                            return G16_OK;
                        }
                        commit_hashed.assign(nh, {});
                        size_t q = 3 + 1 + 2 * (size_t)cd[3];   // input 0 is the commitment's index
                        for (size_t h = 0; h < nh; h++) {
                            const uint32_t ln = cd[q++];
                            for (uint32_t t = 0; t < ln; t++, q += 2) {
                                if (cd[q + 1] != CCS_CONST_WIRE && cd[q + 1] >= c.nb_public + c.nb_secret) {
                                    *why_not = "commitment hashes a solved wire";
                                    return G16_OK;
                                }
                                commit_hashed[h].push_back({cd[q], cd[q + 1]});
                            }
                        }
                    }
                    commit_level = (uint32_t)lv;
                    commit_wire = o0;
                    continue;   // executed by prove.cu between the two phases
                }
                if (kind == HINT_EMULATED_MUL || kind == HINT_GRUMPKIN_SPLIT || kind == HINT_GRUMPKIN_LIMBS) {
                    // integer hints: host-evaluated up front when they hang off the inputs only
                    bool from_inputs = true;
                    size_t q = 3;
                    for (uint32_t i = 0; i < nin && from_inputs; i++) {
                        const uint32_t ln = cd[q++];
                        for (uint32_t t = 0; t < ln; t++, q += 2)
                            if (cd[q + 1] != CCS_CONST_WIRE && !host_known[cd[q + 1]]) from_inputs = false;
                    }
                    if (!from_inputs) {
                        auto nm = c.hint_names.find(hid);
                        *why_not = "integer hint fed by solved wires (host solver only): " + (nm == c.hint_names.end() ? std::to_string(hid) : nm->second);
                        return G16_OK;
                    }
                    host_hints.push_back(ins);
                    {
                        size_t q2 = 3;
                        for (uint32_t i = 0; i < nin; i++) {
                            const uint32_t ln = cd[q2++];
                            for (uint32_t t = 0; t < ln; t++, q2 += 2) {
                                const uint32_t wid = cd[q2 + 1];
                                if (wid != CCS_CONST_WIRE && wid != 0 && wid < c.nb_public + c.nb_secret &&
                                    std::find(host_inputs.begin(), host_inputs.end(), wid) == host_inputs.end())
                                    host_inputs.push_back(wid);
                            }
                        }
                    }
                    for (uint32_t wv = o0; wv < o1 && wv < nw; wv++) {
                        host_wires.push_back(wv);
                        host_known[wv] = 1;
                        known[wv] = 1;   // in place before level 0 runs
                    }
                    continue;
                }
                if (kind == HINT_UNKNOWN) {
                    auto nm = c.hint_names.find(hid);
                    *why_not = "hint without a device implementation: " + (nm == c.hint_names.end() ? std::to_string(hid) : nm->second);
                    return G16_OK;
                }
                if (kind == HINT_DECOMPOSE) {
                    // the device version extracts limbs of at most 32 bits (the host one 64): larger
                    // limb sizes, or a limb size that is not a compile-time constant, go to the host solver
                    bool ok = nin == 3;
                    if (ok) {
                        size_t q = 3 + 1 + 2 * (size_t)cd[3];
                        ok = cd[q] == 1 && cd[q + 2] == CCS_CONST_WIRE;
                        if (ok) {
                            uint64_t ls[4];
                            c.coeffs[cd[q + 1]].canonical(ls);
                            ok = !(ls[1] | ls[2] | ls[3]) && ls[0] >= 1 && ls[0] <= 32;
                        }
                    }
                    if (!ok) {
                        *why_not = "DecomposeHint with a limb size above 32 bits (host solver handles up to 64)";
                        return G16_OK;
                    }
                }
                if (kind == HINT_COUNT) {
                    // device version handles the one-column form only (what rangecheck emits)
                    // nbCols is input 1: a constant expression
                    size_t q = 3 + 1 + 2 * (size_t)cd[3];
                    bool ok = cd[q] == 1 && cd[q + 2] == CCS_CONST_WIRE;
                    if (ok) {
                        HFr one = HFr::one();
                        ok = c.coeffs[cd[q + 1]] == one;
                    }
                    if (!ok) {
                        *why_not = "countHint with more than one column";
                        return G16_OK;
                    }
                }
                info[ins] = make_uint4(16 + (uint32_t)kind, 0, 0, 0);
                lvl_instr.push_back(ins);
            }
        }
        for (uint32_t wv : newly) known[wv] = 1;
        lvl_off.push_back((uint32_t)lvl_instr.size());
    }
    for (uint32_t i = 0; i < nw; i++)
        if (!known[i]) {
            *why_not = "wire " + std::to_string(i) + " is never defined";
            return G16_OK;
        }
    nlevels = (uint32_t)c.levels.size();
    // runs of thin levels (few short rows each) go to the thread-per-proof kernel
    {
        static const long thin_rows = getenv("G16_SOLVER_THIN_ROWS") ? atol(getenv("G16_SOLVER_THIN_ROWS")) : 8;
        static const long thin_terms_env = getenv("G16_SOLVER_THIN_TERMS") ? atol(getenv("G16_SOLVER_THIN_TERMS")) : 256;
        const uint32_t thin_terms = (uint32_t)thin_terms_env, min_run = 16;
        std::vector<uint8_t> thin(nlevels, 0);
        for (uint32_t lv = 0; lv < nlevels; lv++) {
            const uint32_t s0 = lvl_off[lv], t0 = lvl_off[lv + 1];
            if ((long)(t0 - s0) > thin_rows) continue;
            uint32_t terms = 0;
            bool hint = false;
            for (uint32_t k = s0; k < t0; k++) {
                const uint32_t ins = lvl_instr[k];
                const uint32_t* cd = c.calldata.data() + c.start_calldata[ins];
                if (c.blueprint[ins] == 1) terms += cd[1] + cd[2] + cd[3];
                else hint = true;
            }
            thin[lv] = !hint && terms <= thin_terms;
        }
        segments.clear();
        for (uint32_t lv = 0; lv < nlevels;) {
            uint32_t e = lv;
            while (e < nlevels && thin[e] == thin[lv]) e++;
            // a thread walks the rows of a run one after the other (~3 us each) where the wide kernel pays ~5 us per
            // level: worth its small footprint only while the levels average a few rows
            const bool narrow = thin[lv] && e - lv >= min_run && lvl_off[e] - lvl_off[lv] <= 3 * (e - lv);
            if (!segments.empty() && segments.back().narrow == narrow) segments.back().end = e;
            else segments.push_back({lv, e, narrow});
            lv = e;
        }
        // a short thin run between two wide ones stays wide: merge neighbours of equal kind
        std::vector<Segment> merged;
        for (const Segment& sg : segments) {
            if (!merged.empty() && merged.back().narrow == sg.narrow) merged.back().end = sg.end;
            else merged.push_back(sg);
        }
        segments.swap(merged);
    }
    // ---- two-phase plan (k_solve_2p): products per sub-level, operand lists per row
    {
        const HFr one = HFr::one(), minus_one = HFr::one().neg();
        std::vector<uint32_t> sub_off(1, 0), prod_off(1, 0);
        std::vector<uint2> prods, ops, iprods;
        std::vector<uint4> rec2;
        lvl_to_sub.assign(nlevels + 1, 0);
        for (uint32_t lv = 0; lv < nlevels; lv++) {
            lvl_to_sub[lv] = (uint32_t)sub_off.size() - 1;
            uint32_t in_sub = 0;   // products of the open sub-level
            auto close_sub = [&]() {
                sub_off.push_back((uint32_t)(rec2.size() / 2));
                prod_off.push_back((uint32_t)prods.size());
                in_sub = 0;
            };
            for (uint32_t k = lvl_off[lv]; k < lvl_off[lv + 1]; k++) {
                const uint32_t ins = lvl_instr[k];
                const uint32_t* cd = c.calldata.data() + c.start_calldata[ins];
                if (c.blueprint[ins] != 1) {
                    rec2.push_back(info[ins]);
                    rec2.push_back(make_uint4(instr_cd[ins], 0, 0, 0));
                    continue;
                }
                const uint32_t cnt[3] = {cd[1], cd[2], cd[3]};
                const uint32_t unk = info[ins].x == 0 ? NO_WIRE : info[ins].y;
                uint32_t row_products = 0;
                for (uint32_t t = 0; t < cnt[0] + cnt[1] + cnt[2]; t++) {
                    const uint32_t cid = cd[4 + 2 * t], wid = cd[5 + 2 * t];
                    if (wid != unk && wid != CCS_CONST_WIRE && !c.coeffs[cid].is_zero() && !(c.coeffs[cid] == one) && !(c.coeffs[cid] == minus_one))
                        row_products++;
                }
                const bool inline_row = row_products > SOLVE2_PRODUCTS;   // multiplies in place, thread per row
                if (!inline_row && in_sub + row_products > SOLVE2_PRODUCTS) close_sub();
                uint32_t kept[3] = {0, 0, 0};
                const uint32_t first_op = (uint32_t)ops.size();
                size_t p = 4;
                for (int side = 0; side < 3; side++)
                    for (uint32_t t = 0; t < cnt[side]; t++, p += 2) {
                        const uint32_t cid = cd[p], wid = cd[p + 1];
                        if (wid == unk) continue;                       // the wire this row defines
                        const HFr& cf = c.coeffs[cid];
                        if (cf.is_zero()) continue;
                        if (wid == CCS_CONST_WIRE) ops.push_back(make_uint2(OPD_CONST, cid));
                        else if (cf == one) ops.push_back(make_uint2(OPD_PLUS, wid));
                        else if (cf == minus_one) ops.push_back(make_uint2(OPD_MINUS, wid));
                        else if (inline_row) {
                            ops.push_back(make_uint2(OPD_INLINE, (uint32_t)iprods.size()));
                            iprods.push_back(make_uint2(cid, wid));
                        } else {
                            ops.push_back(make_uint2(OPD_PRODUCT, in_sub++));
                            prods.push_back(make_uint2(cid, wid));
                        }
                        kept[side]++;
                    }
                // coefficient of the defined wire: 1 -> +1, 3 -> -1 (by VALUE), else 4*cid (never 1 or 3)
                uint4 inf = info[ins];
                if (unk != NO_WIRE) {
                    const HFr& uc = c.coeffs[inf.z];
                    inf.z = uc == one ? 1u : (uc == minus_one ? 3u : inf.z << 2);
                }
                rec2.push_back(inf);
                rec2.push_back(make_uint4(first_op, kept[0], kept[1], kept[2]));
            }
            close_sub();
        }
        lvl_to_sub[nlevels] = (uint32_t)sub_off.size() - 1;
        if (c.coeffs.size() >= (1u << 30)) {
            *why_not = "coefficient table too large";
            return G16_OK;
        }
        G16_TRY(upload(sub_off, &d_sub_off, st));
        G16_TRY(upload(prod_off, &d_prod_off, st));
        G16_TRY(upload(prods, &d_prods, st));
        G16_TRY(upload(iprods, &d_iprods, st));
        G16_TRY(upload(ops, &d_ops, st));
        G16_TRY(upload(rec2, &d_rec2, st));
    }
    std::vector<uint4> rec(2 * lvl_instr.size());
    for (size_t k = 0; k < lvl_instr.size(); k++) {
        const uint32_t ins = lvl_instr[k];
        const uint32_t* cd = c.calldata.data() + c.start_calldata[ins];
        rec[2 * k] = info[ins];
        rec[2 * k + 1] = c.blueprint[ins] == 1 ? make_uint4(instr_cd[ins], cd[1], cd[2], cd[3]) : make_uint4(instr_cd[ins], 0, 0, 0);
    }
    G16_TRY(upload(rec, &d_rec, st));
    G16_TRY(upload(host_wires, &d_host_wires, st));
    G16_TRY(upload(lvl_off, &d_lvl_off, st));
    G16_TRY(upload(lvl_instr, &d_lvl_instr, st));
    G16_TRY(upload(info, &d_info, st));
    G16_TRY(upload(instr_cd, &d_instr_cd, st));
    G16_TRY(upload(c.calldata, &d_calldata, st));
    G16_CUDA(cudaMalloc(&d_coeff_invs, sizeof(Fr) * c.coeff_invs.size()));
    G16_CUDA(cudaMemcpyAsync(d_coeff_invs, c.coeff_invs.data(), sizeof(Fr) * c.coeff_invs.size(), cudaMemcpyHostToDevice, st));
    G16_CUDA(cudaStreamSynchronize(st));
    valid = true;
    return G16_OK;
}

int GpuSolverPlan::assign(const uint8_t* d_asg_be, const uint8_t* d_rnd_be, uint32_t nin, Fr* d_wires, size_t wstride,
                          size_t nw, size_t B, cudaStream_t st) const {
    dim3 grid(cdiv(nin ? nin : 1, 256), (unsigned)B);
    k_assign<<<grid, 256, 0, st>>>(d_asg_be, d_rnd_be, nin, d_wires, wstride, nw);
    G16_CUDA(cudaGetLastError());
    return G16_OK;
}

int GpuSolverPlan::run(const Fr* d_coeffs, int unit_ids, Fr* d_wires, size_t wstride, size_t nw, size_t B,
                       uint32_t lvl_begin, uint32_t lvl_end, uint32_t* d_err, cudaStream_t st) const {
    if (lvl_begin >= lvl_end) return G16_OK;
    const bool cta_per_proof = getenv("G16_SOLVER_CTA") && atoi(getenv("G16_SOLVER_CTA")) != 0;   // round-1 kernel
    if (cta_per_proof)
        k_solve_levels<<<(unsigned)B, SOLVE_THREADS, 0, st>>>(d_lvl_off, d_lvl_instr, d_info, d_instr_cd, d_calldata, d_coeffs,
                                                              d_coeff_invs, d_wires, wstride, nw + X_BLINDER, lvl_begin,
                                                              lvl_end, unit_ids, d_err);
    else {
        const bool no_narrow = getenv("G16_SOLVER_NARROW") && atoi(getenv("G16_SOLVER_NARROW")) == 0;
        const bool tpi_kernel = getenv("G16_SOLVER_TPI") && atoi(getenv("G16_SOLVER_TPI")) != 0;   // lanes-per-row kernel
        static const bool trace_sync = getenv("G16_TRACE_SYNC") && atoi(getenv("G16_TRACE_SYNC")) != 0;
        for (const Segment& sg : segments) {
            const uint32_t lo = std::max(sg.begin, lvl_begin), hi = std::min(sg.end, lvl_end);
            if (lo >= hi) continue;
            if (trace_sync) {
                cudaStreamSynchronize(st);
                trace(sg.narrow ? "  sync: before narrow segment" : "  sync: before wide segment", (long)lo);
            }
            if (sg.narrow && !no_narrow)
                k_solve_narrow<<<cdiv(B, NARROW_THREADS), NARROW_THREADS, 0, st>>>(d_lvl_off, d_rec, d_calldata, d_coeffs,
                                                                                  d_coeff_invs, d_wires, wstride, nw + X_BLINDER,
                                                                                  lo, hi, unit_ids, d_err, (uint32_t)B);
            else if (!tpi_kernel)
                k_solve_2p<<<(unsigned)B, SOLVE2_THREADS, 0, st>>>(d_sub_off, d_prod_off, d_prods, d_rec2, d_ops, d_iprods, d_calldata, d_coeffs,
                                                                   d_coeff_invs, d_wires, wstride, nw + X_BLINDER, lvl_to_sub[lo],
                                                                   lvl_to_sub[hi], unit_ids, d_err);
            else
                k_solve_tpi<<<(unsigned)B, SOLVE_TT, 0, st>>>(d_lvl_off, d_rec, d_calldata, d_coeffs, d_coeff_invs, d_wires,
                                                              wstride, nw + X_BLINDER, lo, hi, unit_ids, d_err);
        }
    }
    G16_CUDA(cudaGetLastError());
    return G16_OK;
}

int GpuSolverPlan::scatter_host_wires(Fr* d_wires, size_t wstride, const Fr* d_values, size_t B, cudaStream_t st) const {
    const size_t total = host_wires.size() * B;
    if (!total) return G16_OK;
    k_scatter_wires<<<cdiv(total, 256), 256, 0, st>>>(d_wires, wstride, d_host_wires, d_values, (uint32_t)host_wires.size(), (uint32_t)B);
    G16_CUDA(cudaGetLastError());
    return G16_OK;
}

int GpuSolverPlan::set_wire(Fr* d_wires, size_t wstride, uint32_t wire, const Fr* d_values, size_t B, cudaStream_t st) const {
    k_set_wire<<<cdiv(B, 128), 128, 0, st>>>(d_wires, wstride, wire, d_values, (uint32_t)B);
    G16_CUDA(cudaGetLastError());
    return G16_OK;
}

}  // namespace g16

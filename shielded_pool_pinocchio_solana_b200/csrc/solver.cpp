// solver.cpp -- see solver.hpp.
#include "solver.hpp"
#include "hostbig.hpp"

#include <string.h>

#include <map>

#include "common.cuh"

namespace g16 {

namespace {

struct Lin {
    HFr sum;              // sum of the known terms
    int unknowns = 0;     // distinct unknown wires seen (0, 1, or >1)
    uint32_t unk_wire = 0;
    HFr unk_coeff;        // accumulated coefficient of that wire
    uint32_t unk_cid = 0; // coefficient id when the wire appears in exactly one term
    int unk_terms = 0;
};

inline void eval_terms(const Circuit& c, const SolveState& st, const uint32_t* terms, uint32_t n, Lin* out) {
    out->sum = HFr::zero();
    out->unknowns = 0;
    out->unk_coeff = HFr::zero();
    out->unk_terms = 0;
    static const HFr ONE = HFr::one();
    for (uint32_t k = 0; k < n; k++) {
        uint32_t cid = terms[2 * k], wid = terms[2 * k + 1];
        const HFr& coef = c.coeffs[cid];
        if (wid == CCS_CONST_WIRE) {
            out->sum = out->sum + coef;
            continue;
        }
        if (!st.known[wid]) {
            if (out->unknowns == 0 || out->unk_wire != wid) out->unknowns++;
            out->unk_wire = wid;
            out->unk_coeff = out->unk_coeff + coef;
            out->unk_cid = cid;
            out->unk_terms++;
            continue;
        }
        if (coef == ONE) out->sum = out->sum + st.w[wid];
        else out->sum = out->sum + coef * st.w[wid];
    }
}

inline uint64_t canon_bits(const uint64_t c[4], unsigned lo, unsigned nbits) {
    // bits [lo, lo+nbits) of a 256-bit little-endian integer, nbits <= 64
    if (lo >= 256) return 0;
    unsigned wi = lo >> 6, sh = lo & 63;
    uint64_t v = c[wi] >> sh;
    if (sh && wi + 1 < 4) v |= c[wi + 1] << (64 - sh);
    if (nbits < 64) v &= (((uint64_t)1) << nbits) - 1;
    return v;
}

int fail(SolveState* st, int code, const std::string& msg) {
    st->error = msg;
    return code;
}

int run_hint(const Circuit& c, SolveState* st, uint32_t instr, const HFr* blinder, bool* paused) {
    const uint32_t* cd = c.calldata.data() + c.start_calldata[instr];
    const uint32_t total = cd[0], hid = cd[1], nin = cd[2];
    size_t p = 3;
    std::vector<HFr> ins(nin);
    for (uint32_t i = 0; i < nin; i++) {
        if (p >= total) return fail(st, G16_E_PARSE, "hint calldata truncated");
        uint32_t ln = cd[p++];
        if (p + 2 * (size_t)ln > total) return fail(st, G16_E_PARSE, "hint calldata truncated");
        Lin l;
        eval_terms(c, *st, cd + p, ln, &l);
        if (l.unknowns) return fail(st, G16_E_UNSAT, "hint input uses unsolved wire " + std::to_string(l.unk_wire));
        ins[i] = l.sum;
        p += 2 * (size_t)ln;
    }
    if (p + 2 != total) return fail(st, G16_E_PARSE, "hint calldata length mismatch");
    uint32_t o0 = cd[p], o1 = cd[p + 1];
    if (o1 < o0 || o1 > st->w.size()) return fail(st, G16_E_PARSE, "hint output range out of bounds");
    const uint32_t nout = o1 - o0;
    auto kind_it = c.hint_kinds.find(hid);
    HintKind kind = kind_it == c.hint_kinds.end() ? HINT_UNKNOWN : kind_it->second;
    auto set_out = [&](uint32_t k, const HFr& v) {
        st->w[o0 + k] = v;
        st->known[o0 + k] = 1;
    };
    switch (kind) {
        case HINT_NBITS: {
            if (nin != 1) return fail(st, G16_E_HINT, "nBits: expected 1 input");
            uint64_t cn[4];
            ins[0].canonical(cn);
            for (uint32_t k = 0; k < nout; k++) set_out(k, HFr::from_u64(canon_bits(cn, k, 1)));
            return G16_OK;
        }
        case HINT_INVZERO:
            if (nin != 1 || nout != 1) return fail(st, G16_E_HINT, "InvZeroHint: expected 1 input, 1 output");
            set_out(0, ins[0].inverse());
            return G16_OK;
        case HINT_DECOMPOSE: {
            if (nin != 3) return fail(st, G16_E_HINT, "DecomposeHint: expected 3 inputs");
            uint64_t ls[4], cn[4];
            ins[1].canonical(ls);
            ins[2].canonical(cn);
            if (ls[1] | ls[2] | ls[3] || ls[0] == 0 || ls[0] > 64) return fail(st, G16_E_HINT, "DecomposeHint: bad limb size");
            for (uint32_t k = 0; k < nout; k++) set_out(k, HFr::from_u64(canon_bits(cn, k * (unsigned)ls[0], (unsigned)ls[0])));
            return G16_OK;
        }
        case HINT_COUNT: {
            if (nin < 2) return fail(st, G16_E_HINT, "countHint: too few inputs");
            uint64_t a[4], b[4];
            ins[0].canonical(a);
            ins[1].canonical(b);
            uint64_t size = a[0], cols = b[0];
            if (a[1] | a[2] | a[3] | b[1] | b[2] | b[3] || cols == 0 || 2 + size * cols > nin || size != nout ||
                (nin - 2 - size * cols) % cols)
                return fail(st, G16_E_HINT, "countHint: inconsistent sizes");
            std::map<std::string, uint32_t> index;
            auto key_of = [&](size_t first) {
                return std::string((const char*)&ins[first], sizeof(HFr) * cols);
            };
            for (uint64_t r = 0; r < size; r++) index.emplace(key_of(2 + r * cols), (uint32_t)r);
            std::vector<uint64_t> mult(size, 0);
            for (size_t q = 2 + size * cols; q < nin; q += cols) {
                auto it = index.find(key_of(q));
                if (it == index.end()) {
                    if (!st->tolerate) return fail(st, G16_E_HINT, "countHint: query not in table");
                    continue;   // diagnostic mode: count what is there (what the device solver does, too)
                }
                mult[it->second]++;
            }
            for (uint32_t k = 0; k < nout; k++) set_out(k, HFr::from_u64(mult[k]));
            return G16_OK;
        }
        case HINT_RANDOMIZE:
            if (!blinder) return fail(st, G16_E_HINT, "hints.Randomize: no blinder supplied");
            for (uint32_t k = 0; k < nout; k++) set_out(k, *blinder);
            return G16_OK;
        case HINT_COMMIT: {
            if (st->commitments_done >= c.commitments.size() || nout != 1)
                return fail(st, G16_E_HINT, "commitment hint without CommitmentInfo");
            const CommitmentInfo& info = c.commitments[st->commitments_done];
            size_t nh = info.public_and_commitment_committed.size(), np = info.private_committed.size();
            if (nin != 1 + nh + np) return fail(st, G16_E_HINT, "commitment hint: input count mismatch");
            st->hashed.assign(ins.begin() + 1, ins.begin() + 1 + nh);
            st->committed.assign(ins.begin() + 1 + nh, ins.end());
            st->challenge_wire = o0;
            *paused = true;
            return G16_OK;
        }
        case HINT_EMULATED_MUL: {
            // gnark std/math/emulated mulHint: quo | rem | carries with
            //   a(X) b(X) = rem(X) + quo(X) p(X) + (2^nbBits - X) carry(X)   over the integers
            // (restated in oracle/py/groth16.py `_hint_emulated_mul`)
            if (nin < 6) return fail(st, G16_E_HINT, "emulated.mulHint: too few inputs");
            uint64_t hdr[4];
            for (int k = 0; k < 4; k++) {
                uint64_t cn[4];
                ins[k].canonical(cn);
                if (cn[1] | cn[2] | cn[3] || cn[0] > 4096) return fail(st, G16_E_HINT, "emulated.mulHint: bad header");
                hdr[k] = cn[0];
            }
            const size_t nbits = hdr[0], nlimbs = hdr[1], na = hdr[2], nquo = hdr[3];
            if (nbits == 0 || nbits > 128 || nlimbs == 0 || 4 + nlimbs + na >= nin) return fail(st, G16_E_HINT, "emulated.mulHint: malformed inputs");
            const size_t nb = nin - 4 - nlimbs - na;
            const size_t ncarry = std::max(na + nb - 1, nquo + nlimbs - 1) - 1;
            if (nout < nquo + ncarry) return fail(st, G16_E_HINT, "emulated.mulHint: output count mismatch");
            const size_t nrem = nout - nquo - ncarry;
            if (nrem != 0 && nrem != nlimbs) return fail(st, G16_E_HINT, "emulated.mulHint: output count mismatch");
            std::vector<BigInt> pl(nlimbs), al(na), bl(nb);
            for (size_t i = 0; i < nlimbs; i++) pl[i] = BigInt::from_fr(ins[4 + i]);
            for (size_t i = 0; i < na; i++) al[i] = BigInt::from_fr(ins[4 + nlimbs + i]);
            for (size_t i = 0; i < nb; i++) bl[i] = BigInt::from_fr(ins[4 + nlimbs + na + i]);
            auto recompose = [&](const std::vector<BigInt>& l) {
                BigInt v;
                for (size_t i = l.size(); i-- > 0;) v = v.shl(nbits) + l[i];
                return v;
            };
            const BigInt p = recompose(pl), a = recompose(al), b = recompose(bl);
            if (p.is_zero()) return fail(st, G16_E_HINT, "emulated.mulHint: zero modulus");
            BigInt quo, rem;
            BigInt::divmod_mag(a * b, p, &quo, &rem);
            if (nrem == 0 && !rem.is_zero()) return fail(st, G16_E_HINT, "emulated.mulHint: product is not a multiple of the modulus");
            if (quo.bits() > nbits * nquo) return fail(st, G16_E_HINT, "emulated.mulHint: quotient does not fit");
            std::vector<BigInt> ql(nquo), rl(nrem);
            for (size_t i = 0; i < nquo; i++) ql[i] = quo.shr_mag(nbits * i).low_bits(nbits);
            for (size_t i = 0; i < nrem; i++) rl[i] = rem.shr_mag(nbits * i).low_bits(nbits);
            std::vector<BigInt> xp(na + nb - 1), yp(nquo + nlimbs - 1);
            for (size_t i = 0; i < na; i++)
                for (size_t j = 0; j < nb; j++) xp[i + j] = xp[i + j] + al[i] * bl[j];
            for (size_t i = 0; i < nlimbs; i++) {
                if (i < nrem) yp[i] = yp[i] + rl[i];
                for (size_t j = 0; j < nquo; j++) yp[i + j] = yp[i + j] + ql[j] * pl[i];
            }
            for (size_t i = 0; i < nquo; i++) set_out((uint32_t)i, ql[i].to_fr());
            for (size_t i = 0; i < nrem; i++) set_out((uint32_t)(nquo + i), rl[i].to_fr());
            BigInt carry;
            for (size_t i = 0; i < ncarry; i++) {
                if (i < xp.size()) carry = carry + xp[i];
                if (i < yp.size()) carry = carry - yp[i];
                carry = carry.shr_floor(nbits);
                set_out((uint32_t)(nquo + nrem + i), carry.to_fr());
            }
            return G16_OK;
        }
        case HINT_GRUMPKIN_LIMBS: {
            // the native scalar as nout 64-bit limbs (its representation in the emulated scalar field)
            if (nin != 1 || nout == 0 || nout > 4) return fail(st, G16_E_HINT, "sw-grumpkin.decompose: expected 1 input, <= 4 outputs");
            uint64_t cn[4];
            ins[0].canonical(cn);
            for (uint32_t k = nout; k < 4; k++)
                if (cn[k]) return fail(st, G16_E_HINT, "sw-grumpkin.decompose: scalar does not fit");
            for (uint32_t k = 0; k < nout; k++) set_out(k, HFr::from_u64(cn[k]));
            return G16_OK;
        }
        case HINT_GRUMPKIN_SPLIT: {
            // s -> (s1, s2), 0 <= s1, s2 < 2^127, s1 = s + LAMBDA s2 (mod q): what instructions 18-21 and 150 of the
            // withdraw circuit enforce.  The box can hold several such pairs and sunspot's own choice is not
            // observable offline: smallest max(s1, s2), then smallest s2 (UNPINNED; oracle/py `glv_split_nonneg`).
            static const BigInt Q = BigInt::from_decimal("21888242871839275222246405745257275088696311157297823662689037894645226208583");
            static const BigInt A1 = BigInt::from_decimal("9931322734385697762");
            static const BigInt B1 = BigInt::from_decimal("147946756881789319000765030803803410729");
            static const BigInt A2 = BigInt::from_decimal("147946756881789319010696353538189108491");
            static const BigInt B2 = -BigInt::from_decimal("9931322734385697762");
            uint64_t cn[4];
            if (nin < 9) return fail(st, G16_E_HINT, "sw-grumpkin.decomposeScalar: unexpected input framing");
            auto small = [&](size_t k, uint64_t* v) {
                ins[k].canonical(cn);
                *v = cn[0];
                return !(cn[1] | cn[2] | cn[3]);
            };
            uint64_t n_native, n_out, nlimbs, nbits;
            if (!small(0, &n_native) || !small(3, &n_out) || !small(7, &nlimbs) || !small(8, &nbits) || n_native != 1 || n_out != 2 ||
                nlimbs != 4 || nbits != 64 || nin != 9 + nlimbs || nout != n_out * nlimbs)
                return fail(st, G16_E_HINT, "sw-grumpkin.decomposeScalar: unexpected input framing");
            BigInt mod;
            for (size_t i = nlimbs; i-- > 0;) mod = mod.shl(nbits) + BigInt::from_fr(ins[9 + i]);
            if (!(mod == Q)) return fail(st, G16_E_HINT, "sw-grumpkin.decomposeScalar: not the Grumpkin scalar field");
            BigInt s, sq, sr;
            BigInt::divmod_mag(BigInt::from_fr(ins[6]), Q, &sq, &s);
            const BigInt det = A1 * B2 - B1 * A2;   // = -q
            const BigInt half = BigInt(1).shl(126), top = BigInt(1).shl(127);
            const BigInt tx = s - half, ty = -half;
            // floor((tx*B2 - ty*A2) / det) with det < 0: negate numerator and denominator
            const BigInt c1 = BigInt::floordiv(-(tx * B2 - ty * A2), -det);
            const BigInt c2 = BigInt::floordiv(-(ty * A1 - tx * B1), -det);
            bool have = false;
            BigInt bx, by, bkey;
            for (int d1 = -2; d1 < 4; d1++)
                for (int d2 = -2; d2 < 4; d2++) {
                    const BigInt k1 = c1 + (d1 < 0 ? -BigInt((uint64_t)-d1) : BigInt((uint64_t)d1));
                    const BigInt k2 = c2 + (d2 < 0 ? -BigInt((uint64_t)-d2) : BigInt((uint64_t)d2));
                    const BigInt x = s - k1 * A1 - k2 * A2, y = -(k1 * B1) - k2 * B2;
                    if (x.neg || y.neg || !(x < top) || !(y < top)) continue;
                    const BigInt key = x < y ? y : x;
                    if (!have || key < bkey || (key == bkey && y < by)) {
                        have = true;
                        bx = x; by = y; bkey = key;
                    }
                }
            if (!have) return fail(st, G16_E_HINT, "sw-grumpkin.decomposeScalar: no decomposition in range");
            for (uint32_t k = 0; k < 4; k++) {
                set_out(k, bx.shr_mag(64 * k).low_bits(64).to_fr());
                set_out(4 + k, by.shr_mag(64 * k).low_bits(64).to_fr());
            }
            return G16_OK;
        }
        default: {
            auto nm = c.hint_names.find(hid);
            return fail(st, G16_E_HINT, "solver hint not implemented: " + (nm == c.hint_names.end() ? std::to_string(hid) : nm->second));
        }
    }
}

int run_r1c(const Circuit& c, SolveState* st, uint32_t instr) {
    const uint32_t* cd = c.calldata.data() + c.start_calldata[instr];
    uint32_t nl = cd[1], nr = cd[2], no = cd[3];
    if (4 + 2 * ((uint64_t)nl + nr + no) != cd[0]) return fail(st, G16_E_PARSE, "R1C calldata length mismatch");
    Lin L, Rr, O;
    eval_terms(c, *st, cd + 4, nl, &L);
    eval_terms(c, *st, cd + 4 + 2 * nl, nr, &Rr);
    eval_terms(c, *st, cd + 4 + 2 * (nl + nr), no, &O);
    int unknowns = L.unknowns + Rr.unknowns + O.unknowns;
    const uint32_t row = c.constraint_offset[instr];
    if (unknowns == 0) {
        if (L.sum * Rr.sum != O.sum) {
            if (!st->tolerate) return fail(st, G16_E_UNSAT, "constraint #" + std::to_string(row) + " is not satisfied");
            st->failed_rows++;
        }
        return G16_OK;
    }
    // gnark's blueprint guarantees one unknown wire per row, appearing on one side
    bool same = (L.unknowns <= 1 && Rr.unknowns <= 1 && O.unknowns <= 1);
    uint32_t wid = L.unknowns ? L.unk_wire : (Rr.unknowns ? Rr.unk_wire : O.unk_wire);
    if (!same || (L.unknowns && L.unk_wire != wid) || (Rr.unknowns && Rr.unk_wire != wid) || (O.unknowns && O.unk_wire != wid) ||
        unknowns != 1)
        return fail(st, G16_E_UNSAT, "constraint #" + std::to_string(row) + " has more than one unsolved wire");
    // 1/coefficient: from the table precomputed at parse time unless the wire repeats in the row
    auto coeff_inv = [&](const Lin& l) { return l.unk_terms == 1 ? c.coeff_invs[l.unk_cid] : l.unk_coeff.inverse(); };
    HFr val;
    if (O.unknowns) {
        val = (L.sum * Rr.sum - O.sum) * coeff_inv(O);
    } else if (L.unknowns) {
        // gnark solveR1C: with a zero divisor the wire stays 0 and the row is only checked (a*b == c),
        // the DivUnchecked(0, 0) = 0 convention
        if (Rr.sum.is_zero()) {
            if (!O.sum.is_zero()) {
                if (!st->tolerate) return fail(st, G16_E_UNSAT, "constraint #" + std::to_string(row) + " is not satisfied");
                st->failed_rows++;
            }
            val = HFr::zero();
        } else {
            val = (O.sum * Rr.sum.inverse() - L.sum) * coeff_inv(L);
        }
    } else {
        if (L.sum.is_zero()) {
            if (!O.sum.is_zero()) {
                if (!st->tolerate) return fail(st, G16_E_UNSAT, "constraint #" + std::to_string(row) + " is not satisfied");
                st->failed_rows++;
            }
            val = HFr::zero();
        } else {
            val = (O.sum * L.sum.inverse() - Rr.sum) * coeff_inv(Rr);
        }
    }
    st->w[wid] = val;
    st->known[wid] = 1;
    return G16_OK;
}

}  // namespace

void solve_begin(const Circuit& c, const HFr* assignment, SolveState* st) {
    const uint32_t nw = c.nb_wires();
    st->w.assign(nw, HFr::zero());
    st->known.assign(nw, 0);
    st->w[0] = HFr::one();
    st->known[0] = 1;
    const uint32_t nin = c.nb_public - 1 + c.nb_secret;
    for (uint32_t i = 0; i < nin; i++) {
        st->w[1 + i] = assignment[i];
        st->known[1 + i] = 1;
    }
    st->level = st->pos = 0;
    st->commitments_done = 0;
    st->error.clear();
}

int solve_run(const Circuit& c, SolveState* st, const HFr* blinder) {
    for (; st->level < c.levels.size(); st->level++, st->pos = 0) {
        const auto& lvl = c.levels[st->level];
        while (st->pos < lvl.size()) {
            uint32_t instr = lvl[st->pos++];
            if (c.blueprint[instr] == 1) {
                int rc = run_r1c(c, st, instr);
                if (rc != G16_OK) return rc;
            } else {
                bool paused = false;
                int rc = run_hint(c, st, instr, blinder, &paused);
                if (rc != G16_OK) return rc;
                if (paused) return SOLVE_NEED_COMMITMENT;
            }
        }
    }
    for (size_t i = 0; i < st->known.size(); i++)
        if (!st->known[i]) return fail(st, G16_E_UNSAT, "wire " + std::to_string(i) + " was never solved");
    return SOLVE_DONE;
}

int solve_run_hint(const Circuit& c, SolveState* st, uint32_t instr) {
    bool paused = false;
    return run_hint(c, st, instr, nullptr, &paused);
}

void solve_provide_challenge(SolveState* st, const HFr& challenge) {
    st->w[st->challenge_wire] = challenge;
    st->known[st->challenge_wire] = 1;
    st->commitments_done++;
}

}  // namespace g16

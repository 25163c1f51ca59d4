// solver.cpp -- see solver.hpp.
#include "solver.hpp"
#include "hostbig.hpp"

#include <string.h>

#include <algorithm>
#include <functional>
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

// ---------------------------------------------------------------------------------------------------------------------
// witness completion (see solver.hpp); restated in oracle/py/witness_completion.py, which generated the golden
// assignments this is tested against (tests/test_withdraw_real.py)
// ---------------------------------------------------------------------------------------------------------------------
namespace g16 {
namespace {

typedef std::vector<std::pair<uint32_t, HFr>> Unk;   // unknown wire -> accumulated coefficient

void unk_add(Unk* u, uint32_t wire, const HFr& coef) {
    for (auto& e : *u)
        if (e.first == wire) {
            e.second = e.second + coef;
            return;
        }
    u->push_back({wire, coef});
}
void unk_prune(Unk* u) {
    Unk o;
    for (auto& e : *u)
        if (!e.second.is_zero()) o.push_back(e);
    u->swap(o);
}
const HFr* unk_find(const Unk& u, uint32_t wire) {
    for (auto& e : u)
        if (e.first == wire) return &e.second;
    return nullptr;
}

struct Completer {
    const Circuit& c;
    SolveState st;
    std::vector<uint8_t> done, queued;
    std::vector<uint32_t> use_ptr, use_idx;   // CSR: wire -> instructions that READ it
    std::vector<uint32_t> range_bits;         // 0 = unknown
    std::vector<uint32_t> queue;
    size_t qhead = 0;
    std::string error;
    size_t violated = 0;

    explicit Completer(const Circuit& circ) : c(circ) {}

    const uint32_t* cd(uint32_t ins) const { return c.calldata.data() + c.start_calldata[ins]; }

    template <class Fn>
    void for_each_input_wire(uint32_t ins, Fn fn) const {
        const uint32_t* d = cd(ins);
        if (c.blueprint[ins] == 1) {
            const uint32_t nt = d[1] + d[2] + d[3];
            for (uint32_t k = 0; k < nt; k++)
                if (d[5 + 2 * k] != CCS_CONST_WIRE) fn(d[5 + 2 * k]);
        } else {
            size_t p = 3;
            for (uint32_t i = 0; i < d[2]; i++) {
                const uint32_t ln = d[p++];
                for (uint32_t k = 0; k < ln; k++, p += 2)
                    if (d[p + 1] != CCS_CONST_WIRE) fn(d[p + 1]);
            }
        }
    }

    void lin(const uint32_t* terms, uint32_t n, HFr* sum, Unk* unk) const {
        *sum = HFr::zero();
        unk->clear();
        for (uint32_t k = 0; k < n; k++) {
            const uint32_t cid = terms[2 * k], wid = terms[2 * k + 1];
            const HFr& co = c.coeffs[cid];
            if (wid == CCS_CONST_WIRE) *sum = *sum + co;
            else if (!st.known[wid]) unk_add(unk, wid, co);
            else *sum = *sum + co * st.w[wid];
        }
        unk_prune(unk);
    }

    void set_wire(uint32_t x, const HFr& v) {
        if (st.known[x]) return;
        st.w[x] = v;
        st.known[x] = 1;
        for (uint32_t k = use_ptr[x]; k < use_ptr[x + 1]; k++) {
            const uint32_t ins = use_idx[k];
            if (!done[ins] && !queued[ins]) {
                queue.push_back(ins);
                queued[ins] = 1;
            }
        }
    }

    static BigInt canon(const HFr& x) { return BigInt::from_fr(x); }

    // radix rule on  sum coef_i u_i = K
    bool radix(const Unk& coefs, const HFr& K) {
        struct Item {
            size_t e;
            uint32_t wire;
        };
        static const BigInt RM = [] {
            BigInt m;
            m.m.assign(HFr::M, HFr::M + 4);
            return m;
        }();
        const BigInt half = RM.shr_mag(1);
        std::vector<Item> items;
        int sign = -1;
        for (auto& e : coefs) {
            BigInt co = canon(e.second);
            const bool neg = half < co;
            BigInt m = neg ? RM - co : co;
            if (m.is_zero()) return false;
            const size_t bits = m.bits();
            if (!(m == BigInt(1).shl(bits - 1))) return false;   // not a power of two
            if (sign < 0) sign = neg;
            else if (sign != (int)neg) return false;
            items.push_back({bits - 1, e.first});
        }
        std::sort(items.begin(), items.end(), [](const Item& a, const Item& b) { return a.e < b.e; });
        for (size_t i = 0; i + 1 < items.size(); i++)
            if (items[i].e == items[i + 1].e) return false;
        const BigInt Kp = canon(sign ? K.neg() : K);
        std::vector<BigInt> vals(items.size());
        BigInt recomposed;
        for (size_t i = 0; i < items.size(); i++) {
            const size_t gap = i + 1 < items.size() ? items[i + 1].e - items[i].e : 0;
            const size_t bw = range_bits[items[i].wire];
            size_t width = 0;   // 0 = everything above
            if (bw && gap) width = std::min(bw, gap);
            else if (bw) width = bw;
            else width = gap;
            BigInt v = Kp.shr_mag(items[i].e);
            if (width) v = v.low_bits(width);
            vals[i] = v;
            recomposed = recomposed + v.shl(items[i].e);
        }
        if (!(recomposed == Kp)) return false;
        for (size_t i = 0; i < items.size(); i++) set_wire(items[i].wire, vals[i].to_fr());
        return true;
    }

    // one instruction; returns false on a hard error (error set)
    bool step(uint32_t ins) {
        const uint32_t* d = cd(ins);
        if (c.blueprint[ins] != 1) {
            bool ready = true;
            for_each_input_wire(ins, [&](uint32_t x) { ready = ready && st.known[x]; });
            if (!ready) return true;
            auto kit = c.hint_kinds.find(d[1]);
            if (kit != c.hint_kinds.end() && kit->second == HINT_COMMIT) {   // the assignment does not depend on it
                done[ins] = 1;
                return true;
            }
            static const HFr zero_blinder = HFr::zero();
            bool paused = false;
            // run_hint marks its outputs known itself; re-announce them so that their readers are queued
            size_t p = 3;
            for (uint32_t i = 0; i < d[2]; i++) p += 1 + 2 * (size_t)d[p];
            const uint32_t o0 = d[p], o1 = d[p + 1];
            std::vector<uint8_t> was(o1 > o0 ? o1 - o0 : 0);
            for (uint32_t x = o0; x < o1; x++) was[x - o0] = st.known[x];
            if (run_hint(c, &st, ins, &zero_blinder, &paused) != G16_OK) {
                error = st.error;
                return false;
            }
            for (uint32_t x = o0; x < o1; x++)
                if (!was[x - o0]) {
                    st.known[x] = 0;
                    set_wire(x, st.w[x]);
                }
            done[ins] = 1;
            return true;
        }
        const uint32_t nl = d[1], nr = d[2], no = d[3];
        HFr a, b, cc;
        Unk ua, ub, uc;
        lin(d + 4, nl, &a, &ua);
        lin(d + 4 + 2 * nl, nr, &b, &ub);
        lin(d + 4 + 2 * (nl + nr), no, &cc, &uc);
        std::vector<uint32_t> unk;
        for (auto* u : {&ua, &ub, &uc})
            for (auto& e : *u)
                if (std::find(unk.begin(), unk.end(), e.first) == unk.end()) unk.push_back(e.first);
        if (unk.empty()) {
            done[ins] = 1;
            if (a * b != cc) {
                violated++;
                if (error.empty()) error = "constraint #" + std::to_string(c.constraint_offset[ins]) + " is not satisfied";
            }
            return true;
        }
        if (unk.size() == 1) {
            const uint32_t x = unk[0];
            const HFr *ca = unk_find(ua, x), *cb = unk_find(ub, x), *co = unk_find(uc, x);
            if (co && !ca && !cb) {
                set_wire(x, (a * b - cc) * co->inverse());
            } else if (ca && !cb && !co) {
                set_wire(x, b.is_zero() ? HFr::zero() : (cc * b.inverse() - a) * ca->inverse());
            } else if (cb && !ca && !co) {
                set_wire(x, a.is_zero() ? HFr::zero() : (cc * a.inverse() - b) * cb->inverse());
            } else if (ca && co && !cb) {
                const HFr den = *ca * b - *co;
                if (den.is_zero()) return true;
                set_wire(x, (cc - a * b) * den.inverse());
            } else if (cb && co && !ca) {
                const HFr den = *cb * a - *co;
                if (den.is_zero()) return true;
                set_wire(x, (cc - a * b) * den.inverse());
            } else {
                return true;   // quadratic in x: left to the linear-system pass or to other rows
            }
            done[ins] = 1;
            return true;
        }
        // several unknowns: radix rows
        if (ua.empty() && ub.empty()) {
            if (radix(uc, a * b - cc)) done[ins] = 1;
        } else if (ua.empty() && uc.empty()) {
            if (!a.is_zero() && radix(ub, cc * a.inverse() - b)) done[ins] = 1;
        } else if (ub.empty() && uc.empty()) {
            if (!b.is_zero() && radix(ua, cc * b.inverse() - a)) done[ins] = 1;
        }
        return true;
    }

    bool propagate() {
        while (qhead < queue.size()) {
            const uint32_t ins = queue[qhead++];
            queued[ins] = 0;
            if (done[ins]) continue;
            if (!step(ins)) return false;
        }
        queue.clear();
        qhead = 0;
        return true;
    }

    // small linear systems over the stalled rows; returns whether any wire was set
    bool linear_systems(bool fallback) {
        struct Eq {
            Unk co;
            HFr rhs;
        };
        std::vector<Eq> eqs;
        for (uint32_t ins = 0; ins < c.blueprint.size(); ins++) {
            if (done[ins] || c.blueprint[ins] != 1) continue;
            const uint32_t* d = cd(ins);
            const uint32_t nl = d[1], nr = d[2], no = d[3];
            HFr a, b, cc;
            Unk ua, ub, uc;
            lin(d + 4, nl, &a, &ua);
            lin(d + 4 + 2 * nl, nr, &b, &ub);
            lin(d + 4 + 2 * (nl + nr), no, &cc, &uc);
            if (!ua.empty() && !ub.empty()) continue;
            Eq e;
            for (auto& t : ua) unk_add(&e.co, t.first, t.second * b);
            for (auto& t : ub) unk_add(&e.co, t.first, t.second * a);
            for (auto& t : uc) unk_add(&e.co, t.first, t.second.neg());
            unk_prune(&e.co);
            e.rhs = cc - a * b;
            if (!e.co.empty() && e.co.size() <= 4) eqs.push_back(std::move(e));
        }
        if (eqs.empty()) return false;
        std::map<uint32_t, uint32_t> parent;
        std::function<uint32_t(uint32_t)> find = [&](uint32_t x) {
            auto it = parent.find(x);
            if (it == parent.end()) {
                parent[x] = x;
                return x;
            }
            if (it->second == x) return x;
            uint32_t r = find(it->second);
            parent[x] = r;
            return r;
        };
        for (auto& e : eqs)
            for (size_t k = 1; k < e.co.size(); k++) parent[find(e.co[k].first)] = find(e.co[0].first);
        std::map<uint32_t, std::vector<size_t>> comps;
        for (size_t i = 0; i < eqs.size(); i++) comps[find(eqs[i].co[0].first)].push_back(i);
        bool progress = false;
        for (auto& kv : comps) {
            std::vector<uint32_t> xs;
            for (size_t i : kv.second)
                for (auto& t : eqs[i].co)
                    if (std::find(xs.begin(), xs.end(), t.first) == xs.end()) xs.push_back(t.first);
            std::sort(xs.begin(), xs.end());
            if (xs.size() > 40) continue;
            const size_t nc = xs.size(), nrw = kv.second.size();
            std::vector<std::vector<HFr>> M(nrw, std::vector<HFr>(nc + 1, HFr::zero()));
            for (size_t r = 0; r < nrw; r++) {
                for (auto& t : eqs[kv.second[r]].co) {
                    const size_t col = std::lower_bound(xs.begin(), xs.end(), t.first) - xs.begin();
                    M[r][col] = t.second;
                }
                M[r][nc] = eqs[kv.second[r]].rhs;
            }
            std::vector<size_t> piv;
            size_t rr = 0;
            for (size_t col = 0; col < nc && rr < nrw; col++) {
                size_t p = rr;
                while (p < nrw && M[p][col].is_zero()) p++;
                if (p == nrw) continue;
                std::swap(M[rr], M[p]);
                const HFr iv = M[rr][col].inverse();
                for (auto& v : M[rr]) v = v * iv;
                for (size_t r = 0; r < nrw; r++)
                    if (r != rr && !M[r][col].is_zero()) {
                        const HFr f = M[r][col];
                        for (size_t k = 0; k <= nc; k++) M[r][k] = M[r][k] - f * M[rr][k];
                    }
                piv.push_back(col);
                rr++;
            }
            bool inconsistent = false;
            for (size_t r = rr; r < nrw; r++)
                if (!M[r][nc].is_zero()) inconsistent = true;
            if (inconsistent) continue;
            std::vector<uint8_t> is_piv(nc, 0);
            for (size_t col : piv) is_piv[col] = 1;
            bool set_any = false;
            for (size_t r = 0; r < piv.size(); r++) {
                bool determined = true;
                for (size_t col = 0; col < nc; col++)
                    if (!is_piv[col] && !M[r][col].is_zero()) determined = false;
                if ((determined || fallback) && !st.known[xs[piv[r]]]) {
                    set_wire(xs[piv[r]], M[r][nc]);
                    set_any = true;
                }
            }
            if (fallback) {
                for (size_t col = 0; col < nc; col++)
                    if (!is_piv[col] && !st.known[xs[col]]) {
                        set_wire(xs[col], HFr::zero());
                        set_any = true;
                    }
                if (set_any) return true;   // one component at a time in fallback mode
            }
            progress = progress || set_any;
        }
        return progress;
    }

    int run(const std::vector<std::pair<uint32_t, HFr>>& known_in, std::vector<HFr>* assignment) {
        const uint32_t nw = c.nb_wires();
        const size_t ninstr = c.blueprint.size();
        st.w.assign(nw, HFr::zero());
        st.known.assign(nw, 0);
        st.w[0] = HFr::one();
        st.known[0] = 1;
        done.assign(ninstr, 0);
        queued.assign(ninstr, 1);
        // wire -> reader instructions
        std::vector<uint32_t> cnt(nw + 1, 0);
        for (uint32_t ins = 0; ins < ninstr; ins++) for_each_input_wire(ins, [&](uint32_t x) { cnt[x + 1]++; });
        use_ptr.assign(nw + 1, 0);
        for (uint32_t x = 0; x < nw; x++) use_ptr[x + 1] = use_ptr[x] + cnt[x + 1];
        use_idx.assign(use_ptr[nw], 0);
        std::vector<uint32_t> cur(use_ptr.begin(), use_ptr.end() - 1);
        for (uint32_t ins = 0; ins < ninstr; ins++) for_each_input_wire(ins, [&](uint32_t x) { use_idx[cur[x]++] = ins; });
        // range information
        range_bits.assign(nw, 0);
        auto note = [&](uint32_t wire, uint64_t bits) {
            if (bits && bits < 400 && (!range_bits[wire] || bits < range_bits[wire])) range_bits[wire] = (uint32_t)bits;
        };
        for (uint32_t ins = 0; ins < ninstr; ins++) {
            const uint32_t* d = cd(ins);
            if (c.blueprint[ins] != 1) {
                auto kit = c.hint_kinds.find(d[1]);
                if (kit == c.hint_kinds.end()) continue;
                size_t p = 3;
                std::vector<size_t> starts;
                for (uint32_t i = 0; i < d[2]; i++) {
                    starts.push_back(p);
                    p += 1 + 2 * (size_t)d[p];
                }
                const uint32_t nout = d[p + 1] - d[p];
                auto single_wire = [&](size_t q, uint32_t* wire) {
                    if (d[q] != 1 || d[q + 2] == CCS_CONST_WIRE || !(c.coeffs[d[q + 1]] == HFr::one())) return false;
                    *wire = d[q + 2];
                    return true;
                };
                uint32_t wire;
                if (kit->second == HINT_DECOMPOSE && d[2] == 3 && d[starts[0]] == 1 && d[starts[0] + 2] == CCS_CONST_WIRE &&
                    single_wire(starts[2], &wire)) {
                    uint64_t cn[4];
                    c.coeffs[d[starts[0] + 1]].canonical(cn);
                    if (!(cn[1] | cn[2] | cn[3])) note(wire, cn[0]);
                } else if (kit->second == HINT_NBITS && d[2] == 1 && single_wire(starts[0], &wire)) {
                    note(wire, nout);
                }
            } else if (d[1] == 1 && d[2] == 2 && d[3] <= 1 && d[5] != CCS_CONST_WIRE && c.coeffs[d[4]] == HFr::one()) {
                // b * (1 - b) = 0
                const uint32_t bw = d[5];
                bool one_ok = false, minus_ok = false, zero_rhs = true;
                for (int k = 0; k < 2; k++) {
                    const uint32_t cid = d[6 + 2 * k], wid = d[7 + 2 * k];
                    if ((wid == 0 || wid == CCS_CONST_WIRE) && c.coeffs[cid] == HFr::one()) one_ok = true;
                    if (wid == bw && c.coeffs[cid] == HFr::one().neg()) minus_ok = true;
                }
                if (d[3] == 1 && !c.coeffs[d[10]].is_zero()) zero_rhs = false;
                if (one_ok && minus_ok && zero_rhs) note(bw, 1);
            }
        }
        for (auto& kv : known_in) {
            if (kv.first >= nw) {
                error = "known wire out of range";
                return G16_E_ARG;
            }
            st.w[kv.first] = kv.second;
            st.known[kv.first] = 1;
        }
        queue.resize(ninstr);
        for (uint32_t i = 0; i < ninstr; i++) queue[i] = i;
        if (!propagate()) return G16_E_HINT;
        bool fallback = false, fallback_tried = false, progress = true;
        const uint32_t nin = c.nb_public + c.nb_secret;
        auto inputs_complete = [&] {
            for (uint32_t i = 0; i < nin; i++)
                if (!st.known[i] && use_ptr[i] != use_ptr[i + 1]) return false;
            return true;
        };
        while (progress || !fallback_tried) {
            if (!progress) {
                fallback = true;
                fallback_tried = true;
            } else {
                fallback_tried = false;
                fallback = false;
            }
            progress = linear_systems(fallback);
            if (!propagate()) return G16_E_HINT;
            if (!progress && fallback_tried) break;
        }
        if (violated) return G16_E_UNSAT;
        if (!inputs_complete()) {
            uint32_t first = 0;
            while (first < nin && (st.known[first] || use_ptr[first] == use_ptr[first + 1])) first++;
            error = "the constraints do not determine input wire " + std::to_string(first);
            return G16_E_UNSAT;
        }
        assignment->assign(st.w.begin() + 1, st.w.begin() + nin);   // unread inputs stay 0
        return G16_OK;
    }
};

}  // namespace

int complete_assignment(const Circuit& c, const std::vector<std::pair<uint32_t, HFr>>& known, std::vector<HFr>* assignment,
                        std::string* err) {
    Completer k(c);
    int rc = k.run(known, assignment);
    if (rc != G16_OK && err) *err = k.error.empty() ? "witness completion failed" : k.error;
    return rc;
}

}  // namespace g16

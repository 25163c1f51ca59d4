// ccs.cpp -- parser for the gnark v0.14 `.ccs` container (see ccs.hpp).
#include "ccs.hpp"

#include <string.h>

#include <memory>
#include <stdexcept>

#include "common.cuh"

namespace g16 {

namespace {

struct ParseError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

struct Reader {
    const uint8_t* p;
    size_t len, off = 0;
    void need(size_t n) const {
        if (off + n > len) throw ParseError("truncated .ccs (need " + std::to_string(n) + " bytes at " + std::to_string(off) + ")");
    }
    uint64_t u64() {
        need(8);
        uint64_t v;
        memcpy(&v, p + off, 8);
        off += 8;
        return v;
    }
    uint32_t u32() {
        need(4);
        uint32_t v;
        memcpy(&v, p + off, 4);
        off += 4;
        return v;
    }
    uint64_t uvarint() {
        uint64_t v = 0;
        int shift = 0;
        for (;;) {
            need(1);
            uint8_t b = p[off++];
            v |= (uint64_t)(b & 0x7f) << shift;
            if (!(b & 0x80)) break;
            shift += 7;
            if (shift > 63) throw ParseError("uvarint overflow");
        }
        return v;
    }
};

// ---- intcomp streams (github.com/ronanh/intcomp as gnark uses it) --------------------------------
template <class W>
void decode_stream(const std::vector<W>& words, std::vector<W>* out) {
    constexpr int WB = sizeof(W) * 8;
    constexpr int GROUP = WB;  // 32 ints per sub-block for u32, 64 for u64
    out->clear();
    if (words.empty()) return;
    size_t T = (size_t)words.back();
    if (T + 1 > words.size()) throw ParseError("stream: bad trailer length");
    size_t body_end = words.size() - 1 - T;
    size_t p = 0;
    while (p < body_end) {
        uint64_t N, section;
        W init;
        size_t q;
        if (WB == 32) {
            if (p + 3 > body_end) throw ParseError("stream: short bit-packed header");
            N = words[p]; section = words[p + 1]; init = words[p + 2];
            q = p + 3;
        } else {
            if (p + 2 > body_end) throw ParseError("stream: short bit-packed header");
            N = (uint64_t)words[p] & 0xffffffffu; section = (uint64_t)words[p] >> 32; init = words[p + 1];
            q = p + 2;
        }
        size_t end = p + section;
        if (end > body_end || section == 0) throw ParseError("stream: bad section length");
        W cur = init;
        uint64_t produced = 0;
        while (produced < N) {
            if (q >= end) throw ParseError("stream: block overruns section");
            uint32_t header = (uint32_t)words[q++];
            for (int sb = 0; sb < 4; sb++) {
                uint32_t desc = (header >> (24 - 8 * sb)) & 0xff;
                unsigned width = desc & 0x7f;
                bool zz = desc >> 7;
                if (width > (unsigned)WB || q + width > end) throw ParseError("stream: bad bit width");
                // `width` words hold GROUP deltas of `width` bits each, LSB first
                for (int i = 0; i < GROUP; i++) {
                    W d = 0;
                    if (width) {
                        size_t bit = (size_t)i * width;
                        size_t wi = bit / WB, sh = bit % WB;
                        d = words[q + wi] >> sh;
                        if (sh + width > (size_t)WB) d |= words[q + wi + 1] << (WB - sh);
                        if (width < (unsigned)WB) d &= (((W)1) << width) - 1;
                    }
                    if (zz) d = (d >> 1) ^ (W)(0 - (d & 1));
                    cur = cur + d;
                    out->push_back(cur);
                }
                q += width;
                produced += GROUP;
            }
        }
        if (q != end) throw ParseError("stream: section length mismatch");
        p = end;
    }
    if (T) {
        const W* tail = words.data() + body_end;
        uint64_t count;
        size_t hdr;
        if (WB == 32) {
            if (T < 2) throw ParseError("stream: short varbyte header");
            count = tail[0];
            hdr = 2;
        } else {
            count = (uint64_t)tail[0] & 0xffffffffu;
            hdr = 1;
        }
        // bytes are taken big-endian inside each word
        std::vector<uint8_t> raw;
        raw.reserve((T - hdr) * sizeof(W));
        for (size_t i = hdr; i < T; i++)
            for (int b = (int)sizeof(W) - 1; b >= 0; b--) raw.push_back((uint8_t)(tail[i] >> (8 * b)));
        W cur = 0;
        size_t i = 0;
        for (uint64_t k = 0; k < count; k++) {
            W v = 0;
            int shift = 0;
            for (;;) {
                if (i >= raw.size()) throw ParseError("stream: varbyte section truncated");
                uint8_t b = raw[i++];
                v |= (W)(b & 0x7f) << shift;
                if (!(b & 0x80)) break;
                shift += 7;
            }
            cur = cur + v;
            out->push_back(cur);
        }
    }
}

template <class W>
void read_stream(Reader& r, std::vector<W>* out) {
    uint64_t n = r.u64();
    r.need(n * sizeof(W));
    std::vector<W> words(n);
    if (n) memcpy(words.data(), r.p + r.off, n * sizeof(W));
    r.off += n * sizeof(W);
    decode_stream<W>(words, out);
}

// ---- minimal CBOR (RFC 8949) tree -------------------------------------------------------------------
struct Cbor {
    enum Type { UINT, NINT, BYTES, TEXT, ARRAY, MAP, SIMPLE } type = SIMPLE;
    uint64_t u = 0;  // UINT value, NINT encoded value, SIMPLE value
    std::string s;   // BYTES / TEXT
    std::vector<Cbor> items;                       // ARRAY
    std::vector<std::pair<Cbor, Cbor>> entries;    // MAP
    const Cbor* get(const char* key) const {
        for (auto& e : entries)
            if (e.first.type == TEXT && e.first.s == key) return &e.second;
        return nullptr;
    }
};

Cbor cbor_read(Reader& r, int depth = 0) {
    if (depth > 64) throw ParseError("CBOR nesting too deep");
    r.need(1);
    uint8_t ib = r.p[r.off++];
    unsigned major = ib >> 5, info = ib & 31;
    uint64_t val = 0;
    bool indefinite = false;
    if (info < 24) val = info;
    else if (info <= 27) {
        int n = 1 << (info - 24);
        r.need(n);
        for (int i = 0; i < n; i++) val = (val << 8) | r.p[r.off++];
    } else if (info == 31) indefinite = true;
    else throw ParseError("CBOR: reserved additional info");
    Cbor c;
    switch (major) {
        case 0: c.type = Cbor::UINT; c.u = val; break;
        case 1: c.type = Cbor::NINT; c.u = val; break;
        case 2:
        case 3:
            c.type = major == 2 ? Cbor::BYTES : Cbor::TEXT;
            if (indefinite) {
                for (;;) {
                    r.need(1);
                    if (r.p[r.off] == 0xff) { r.off++; break; }
                    Cbor chunk = cbor_read(r, depth + 1);
                    c.s += chunk.s;
                }
            } else {
                r.need(val);
                c.s.assign((const char*)r.p + r.off, val);
                r.off += val;
            }
            break;
        case 4:
            c.type = Cbor::ARRAY;
            if (indefinite) {
                for (;;) {
                    r.need(1);
                    if (r.p[r.off] == 0xff) { r.off++; break; }
                    c.items.push_back(cbor_read(r, depth + 1));
                }
            } else {
                if (val > r.len) throw ParseError("CBOR: array too long");
                c.items.reserve(val);
                for (uint64_t i = 0; i < val; i++) c.items.push_back(cbor_read(r, depth + 1));
            }
            break;
        case 5:
            c.type = Cbor::MAP;
            if (indefinite) {
                for (;;) {
                    r.need(1);
                    if (r.p[r.off] == 0xff) { r.off++; break; }
                    Cbor k = cbor_read(r, depth + 1);
                    Cbor v = cbor_read(r, depth + 1);
                    c.entries.emplace_back(std::move(k), std::move(v));
                }
            } else {
                if (val > r.len) throw ParseError("CBOR: map too long");
                for (uint64_t i = 0; i < val; i++) {
                    Cbor k = cbor_read(r, depth + 1);
                    Cbor v = cbor_read(r, depth + 1);
                    c.entries.emplace_back(std::move(k), std::move(v));
                }
            }
            break;
        case 6:  // tag: transparent
            return cbor_read(r, depth + 1);
        default:
            c.type = Cbor::SIMPLE;
            c.u = val;  // 20 false, 21 true, 22 null; floats are skipped as raw bits
            break;
    }
    return c;
}

HintKind classify_hint(const std::string& name) {
    auto ends_with = [&](const char* suf) {
        size_t n = strlen(suf);
        return name.size() >= n && name.compare(name.size() - n, n, suf) == 0;
    };
    if (ends_with("std/math/bits.nBits")) return HINT_NBITS;
    if (ends_with("constraint/solver.InvZeroHint")) return HINT_INVZERO;
    if (ends_with("std/rangecheck.DecomposeHint")) return HINT_DECOMPOSE;
    if (ends_with("std/internal/logderivarg.countHint")) return HINT_COUNT;
    if (ends_with("internal/hints.Randomize")) return HINT_RANDOMIZE;
    if (ends_with("frontend/cs.Bsb22CommitmentComputePlaceholder")) return HINT_COMMIT;
    if (ends_with("std/math/emulated.mulHint")) return HINT_EMULATED_MUL;
    if (ends_with("sw-grumpkin.decomposeScalar")) return HINT_GRUMPKIN_SPLIT;
    if (ends_with("sw-grumpkin.decompose")) return HINT_GRUMPKIN_LIMBS;
    return HINT_UNKNOWN;
}

void build_csr(Circuit& c) {
    Circuit::Csr* M[3] = {&c.A, &c.B, &c.C};
    std::vector<uint64_t> row_instr(c.nb_constraints, (uint64_t)-1);
    for (size_t i = 0; i < c.blueprint.size(); i++)
        if (c.blueprint[i] == 1) {
            if (c.constraint_offset[i] >= c.nb_constraints) throw ParseError("constraint offset out of range");
            row_instr[c.constraint_offset[i]] = i;
        }
    for (int m = 0; m < 3; m++) {
        M[m]->rowptr.assign(1, 0);
        M[m]->coeff.clear();
        M[m]->wire.clear();
    }
    const uint32_t nw = c.nb_wires();
    for (uint32_t row = 0; row < c.nb_constraints; row++) {
        if (row_instr[row] == (uint64_t)-1) throw ParseError("constraint row " + std::to_string(row) + " has no instruction");
        size_t s = c.start_calldata[row_instr[row]];
        const uint32_t* cd = c.calldata.data() + s;
        uint32_t n = cd[0];
        uint32_t cnt[3] = {cd[1], cd[2], cd[3]};
        if (4 + 2 * ((uint64_t)cnt[0] + cnt[1] + cnt[2]) != n) throw ParseError("R1C calldata length mismatch");
        size_t p = 4;
        for (int m = 0; m < 3; m++) {
            for (uint32_t k = 0; k < cnt[m]; k++, p += 2) {
                uint32_t cid = cd[p], wid = cd[p + 1];
                if (wid == CCS_CONST_WIRE) wid = 0;
                if (cid >= c.coeffs.size() || wid >= nw) throw ParseError("R1C term out of range");
                M[m]->coeff.push_back(cid);
                M[m]->wire.push_back(wid);
            }
            M[m]->rowptr.push_back((uint32_t)M[m]->coeff.size());
        }
    }
}

}  // namespace

int parse_ccs(const uint8_t* buf, size_t len, Circuit* out) {
    try {
        Circuit c;
        Reader r{buf, len};
        uint64_t total = r.u64(), v0 = r.u64(), v1 = r.u64(), v2 = r.u64();
        if (total != len - 32 || v0 != 0 || v1 != 14 || v2 != 0)
            throw ParseError("not a gnark 0.14 .ccs (bad header)");
        uint64_t lv_len = r.u64(), ins_len = r.u64(), cd_len = r.u64(), body_len = r.u64();
        size_t end = r.off + lv_len;
        uint64_t nlev = r.u64();
        if (nlev > lv_len) throw ParseError("bad level count");
        c.levels.resize(nlev);
        size_t ninstr = 0;
        for (auto& l : c.levels) {
            read_stream<uint32_t>(r, &l);
            ninstr += l.size();
        }
        if (r.off != end) throw ParseError("levels section length mismatch");
        end = r.off + ins_len;
        read_stream<uint32_t>(r, &c.blueprint);
        read_stream<uint32_t>(r, &c.constraint_offset);
        read_stream<uint32_t>(r, &c.wire_offset);
        read_stream<uint64_t>(r, &c.start_calldata);
        if (r.off != end) throw ParseError("instructions section length mismatch");
        if (c.blueprint.size() < ninstr || c.constraint_offset.size() < ninstr || c.wire_offset.size() < ninstr ||
            c.start_calldata.size() < ninstr)
            throw ParseError("instruction streams shorter than the level lists");
        c.blueprint.resize(ninstr);
        c.constraint_offset.resize(ninstr);
        c.wire_offset.resize(ninstr);
        c.start_calldata.resize(ninstr);
        end = r.off + cd_len;
        uint64_t count = r.u64();
        if (count > cd_len) throw ParseError("bad calldata count");
        c.calldata.resize(count);
        for (auto& v : c.calldata) v = (uint32_t)r.uvarint();
        if (r.off != end) throw ParseError("calldata section length mismatch");
        for (size_t i = 0; i < ninstr; i++) {
            uint64_t s = c.start_calldata[i];
            if (s >= c.calldata.size() || s + c.calldata[s] > c.calldata.size() || c.calldata[s] < 3)
                throw ParseError("instruction calldata out of range");
        }
        // levels must partition the instructions
        {
            std::vector<uint8_t> seen(ninstr, 0);
            for (auto& l : c.levels)
                for (uint32_t i : l) {
                    if (i >= ninstr || seen[i]) throw ParseError("levels do not partition the instructions");
                    seen[i] = 1;
                }
        }
        // body
        r.need(body_len);
        Reader br{buf + r.off, (size_t)body_len};
        Cbor body = cbor_read(br);
        r.off += body_len;
        if (body.type != Cbor::MAP) throw ParseError("body is not a CBOR map");
        auto req = [&](const char* k) {
            const Cbor* v = body.get(k);
            if (!v) throw ParseError(std::string("body lacks ") + k);
            return v;
        };
        for (auto& it : req("Public")->items) c.public_names.push_back(it.s);
        for (auto& it : req("Secret")->items) c.secret_names.push_back(it.s);
        c.nb_public = (uint32_t)c.public_names.size();
        c.nb_secret = (uint32_t)c.secret_names.size();
        c.nb_internal = (uint32_t)req("NbInternalVariables")->u;
        c.nb_constraints = (uint32_t)req("NbConstraints")->u;
        if (const Cbor* ci = body.get("CommitmentInfo")) {
            for (auto& it : ci->items) {
                CommitmentInfo info;
                if (const Cbor* v = it.get("CommitmentIndex")) info.commitment_index = (uint32_t)v->u;
                if (const Cbor* v = it.get("PrivateCommitted"))
                    for (auto& w : v->items) info.private_committed.push_back((uint32_t)w.u);
                if (const Cbor* v = it.get("NbPublicCommitted")) info.nb_public_committed = (uint32_t)v->u;
                if (const Cbor* v = it.get("PublicAndCommitmentCommitted"))
                    for (auto& w : v->items) info.public_and_commitment_committed.push_back((uint32_t)w.u);
                c.commitments.push_back(std::move(info));
            }
        }
        if (const Cbor* mh = body.get("MHintsDependencies")) {
            for (auto& e : mh->entries) {
                c.hint_names[(uint32_t)e.first.u] = e.second.s;
                c.hint_kinds[(uint32_t)e.first.u] = classify_hint(e.second.s);
            }
        }
        // coefficient table
        uint64_t ncoef = r.u64();
        r.need(ncoef * 32);
        c.coeffs.resize(ncoef);
        for (auto& x : c.coeffs) {
            memcpy(x.l, r.p + r.off, 32);
            r.off += 32;
            if (HFr::geq_mod(x.l)) throw ParseError("coefficient not reduced");
        }
        if (r.off != len) throw ParseError("trailing bytes after the coefficient table");
        {   // batch inversion of the table (Montgomery's trick)
            c.coeff_invs = c.coeffs;
            std::vector<HFr> pref(ncoef);
            HFr acc = HFr::one();
            for (size_t i = 0; i < ncoef; i++) {
                pref[i] = acc;
                if (!c.coeffs[i].is_zero()) acc = acc * c.coeffs[i];
            }
            HFr inv = acc.inverse();
            for (size_t i = ncoef; i-- > 0;) {
                if (c.coeffs[i].is_zero()) continue;
                c.coeff_invs[i] = inv * pref[i];
                inv = inv * c.coeffs[i];
            }
        }
        size_t nrows = 0;
        for (auto b : c.blueprint) nrows += (b == 1);
        if (nrows != c.nb_constraints) throw ParseError("NbConstraints does not match the R1C instruction count");
        // every id the solvers (host and device) and the key loader will index with is range-checked
        // once, here: hint calldata structure, wire / coefficient ids, CommitmentInfo wire ids
        {
            const uint64_t nw = c.nb_wires();
            auto check_pair = [&](uint32_t cid, uint32_t wid) {
                if (cid >= ncoef) throw ParseError("coefficient id out of range");
                if (wid != 0xFFFFFFFFu && wid >= nw) throw ParseError("wire id out of range");
            };
            for (size_t i = 0; i < ninstr; i++) {
                const uint32_t* cd = c.calldata.data() + c.start_calldata[i];
                const uint64_t n = cd[0];
                if (c.blueprint[i] == 1) {
                    if (n < 4 || 4 + 2 * ((uint64_t)cd[1] + cd[2] + cd[3]) != n) throw ParseError("R1C calldata length mismatch");
                    for (uint64_t k = 4; k < n; k += 2) check_pair(cd[k], cd[k + 1]);
                } else if (c.blueprint[i] == 0) {
                    if (n < 5) throw ParseError("hint calldata too short");
                    uint64_t p = 3;
                    for (uint32_t in = 0; in < cd[2]; in++) {
                        if (p >= n) throw ParseError("hint calldata: input list overruns the instruction");
                        uint64_t len_i = cd[p];
                        if (p + 1 + 2 * len_i > n) throw ParseError("hint calldata: input expression overruns the instruction");
                        for (uint64_t k = 0; k < len_i; k++) check_pair(cd[p + 1 + 2 * k], cd[p + 2 + 2 * k]);
                        p += 1 + 2 * len_i;
                    }
                    if (p + 2 != n) throw ParseError("hint calldata length mismatch");
                    if (cd[p] > cd[p + 1] || cd[p + 1] > nw) throw ParseError("hint output range out of bounds");
                } else {
                    throw ParseError("unknown blueprint id " + std::to_string(c.blueprint[i]));
                }
            }
            for (auto& info : c.commitments) {
                if (info.commitment_index >= nw) throw ParseError("CommitmentInfo.CommitmentIndex out of range");
                for (uint32_t w : info.private_committed)
                    if (w >= nw) throw ParseError("CommitmentInfo.PrivateCommitted wire out of range");
                for (uint32_t w : info.public_and_commitment_committed)
                    if (w >= nw) throw ParseError("CommitmentInfo.PublicAndCommitmentCommitted wire out of range");
            }
        }
        build_csr(c);
        *out = std::move(c);
        return G16_OK;
    } catch (const std::exception& e) {
        set_error(std::string("parse_ccs: ") + e.what());
        return G16_E_PARSE;
    }
}

}  // namespace g16

// witness.cpp -- Noir witness container (`target/<name>.gz`, written by `nargo execute`) ->
// gnark assignment, i.e. the second argument of `sunspot prove`
// (/root/reference/client/proof.helper.ts:55-64).
//
// Format (UNVERIFIED: the reference commits no witness file, /root/reference/.gitignore:51-54):
// gzip( bincode WitnessStack ) with fixed-width little-endian integers,
//   u64 nItems { u32 functionIndex ; u64 nEntries { u32 witness ; FieldElement } }
// where FieldElement is encoded like the constants inside the committed ACIR bytecode
// (noir_circuit/target/shielded_pool_verifier.json, Noir 1.0.0-beta.18): u64 length = 32 followed
// by 32 big-endian bytes.  A 64-character hex string (older Noir) is accepted as well.
// Mapping to gnark variables follows the `.ccs` name lists: secret `__witness_K` <- ACIR witness
// K; public inputs <- the lowest ACIR witness indices that are not secret, in ascending order.
#include <string.h>
#include <zlib.h>

#include <map>
#include <set>

#include "ccs.hpp"
#include "common.cuh"

namespace g16 {

static int inflate_gzip(const uint8_t* in, size_t len, std::vector<uint8_t>* out) {
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (inflateInit2(&zs, 15 + 32) != Z_OK) {  // +32: auto-detect gzip / zlib headers
        set_error("witness: zlib init failed");
        return G16_E_INTERNAL;
    }
    zs.next_in = const_cast<uint8_t*>(in);
    zs.avail_in = (uInt)len;
    std::vector<uint8_t> buf(1 << 16);
    int rc;
    do {
        zs.next_out = buf.data();
        zs.avail_out = (uInt)buf.size();
        rc = inflate(&zs, Z_NO_FLUSH);
        if (rc != Z_OK && rc != Z_STREAM_END) {
            inflateEnd(&zs);
            set_error("witness: not a gzip stream");
            return G16_E_PARSE;
        }
        out->insert(out->end(), buf.data(), buf.data() + (buf.size() - zs.avail_out));
        if (out->size() > ((size_t)1 << 30)) {   // a 2^22-wire witness stack is ~200 MB; refuse zip bombs
            inflateEnd(&zs);
            set_error("witness: inflated stream exceeds 1 GiB");
            return G16_E_PARSE;
        }
    } while (rc != Z_STREAM_END);
    inflateEnd(&zs);
    return G16_OK;
}

static int hexval(uint8_t c) {
    if (c >= '0' && c <= '9') return c - '0';
    if (c >= 'a' && c <= 'f') return c - 'a' + 10;
    if (c >= 'A' && c <= 'F') return c - 'A' + 10;
    return -1;
}

typedef std::map<uint32_t, std::vector<uint8_t>> WitnessValues;   // ACIR witness index -> 32 big-endian bytes

static bool field_from_bytes(const uint8_t* p, size_t len, std::vector<uint8_t>* be) {
    be->assign(32, 0);
    if (len == 32) {
        memcpy(be->data(), p, 32);
        return true;
    }
    if (len == 64) {   // hex string
        for (int k = 0; k < 32; k++) {
            int hi = hexval(p[2 * k]), lo = hexval(p[2 * k + 1]);
            if (hi < 0 || lo < 0) return false;
            (*be)[k] = (uint8_t)(hi * 16 + lo);
        }
        return true;
    }
    return false;
}

// bincode WitnessStack { stack: Vec<StackItem { index: u32, witness: BTreeMap<u32, FieldElement> }> }; the values of
// the LAST stack item (= main) are kept
static bool parse_bincode_stack(const std::vector<uint8_t>& raw, size_t off, WitnessValues* values, std::string* why) {
    auto need = [&](size_t n) { return off + n <= raw.size(); };
    auto u64 = [&](uint64_t* v) {
        if (!need(8)) return false;
        memcpy(v, raw.data() + off, 8);
        off += 8;
        return true;
    };
    auto u32 = [&](uint32_t* v) {
        if (!need(4)) return false;
        memcpy(v, raw.data() + off, 4);
        off += 4;
        return true;
    };
    uint64_t nitems;
    if (!u64(&nitems) || nitems == 0 || nitems > 1024) {
        *why = "not a bincode WitnessStack";
        return false;
    }
    for (uint64_t it = 0; it < nitems; it++) {
        uint32_t fidx;
        uint64_t nent;
        if (!u32(&fidx) || !u64(&nent) || nent > raw.size()) {
            *why = "truncated stack item";
            return false;
        }
        values->clear();
        for (uint64_t e = 0; e < nent; e++) {
            uint32_t wi;
            uint64_t flen;
            if (!u32(&wi) || !u64(&flen) || !need(flen)) {
                *why = "truncated entry";
                return false;
            }
            std::vector<uint8_t> be;
            if (!field_from_bytes(raw.data() + off, flen, &be)) {
                *why = "unsupported field element encoding (length " + std::to_string(flen) + ")";
                return false;
            }
            off += flen;
            (*values)[wi] = std::move(be);
        }
    }
    if (off != raw.size()) {
        *why = "trailing bytes after the bincode WitnessStack";
        return false;
    }
    return true;
}

// ---- MessagePack (the serialisation newer nargo releases write behind a one-byte format marker) --------------------
// Field names and struct-as-map / struct-as-array ("compact") framing differ between releases, so the reader does not
// depend on them: it decodes the document generically and takes the LAST map whose keys are all non-negative integers
// and whose values all decode as field elements (32 raw bytes, 64 hex characters, or an array of 32 byte values) --
// the witness map of main().
struct MpNode {
    enum Kind { NIL, BOOL, INT, BYTES, ARRAY, MAP } kind = NIL;
    uint64_t u = 0;
    bool negative = false;
    const uint8_t* p = nullptr;   // BYTES (str or bin)
    size_t len = 0;
    std::vector<MpNode> items;    // ARRAY: elements; MAP: key, value, key, value, ...
};

static bool mp_parse(const uint8_t* d, size_t n, size_t* off, MpNode* out, int depth) {
    if (depth > 32 || *off >= n) return false;
    auto be = [&](int bytes, uint64_t* v) {
        if (*off + bytes > n) return false;
        *v = 0;
        for (int i = 0; i < bytes; i++) *v = (*v << 8) | d[(*off)++];
        return true;
    };
    auto bytes_node = [&](uint64_t len) {
        if (*off + len > n) return false;
        out->kind = MpNode::BYTES;
        out->p = d + *off;
        out->len = (size_t)len;
        *off += len;
        return true;
    };
    auto seq = [&](uint64_t count, bool map) {
        if (count > n) return false;
        out->kind = map ? MpNode::MAP : MpNode::ARRAY;
        out->items.resize((size_t)(map ? 2 * count : count));
        for (auto& it : out->items)
            if (!mp_parse(d, n, off, &it, depth + 1)) return false;
        return true;
    };
    const uint8_t t = d[(*off)++];
    uint64_t v;
    if (t <= 0x7f) { out->kind = MpNode::INT; out->u = t; return true; }
    if (t >= 0xe0) { out->kind = MpNode::INT; out->negative = true; out->u = (uint64_t)(int8_t)t; return true; }
    if ((t & 0xf0) == 0x80) return seq(t & 0x0f, true);
    if ((t & 0xf0) == 0x90) return seq(t & 0x0f, false);
    if ((t & 0xe0) == 0xa0) return bytes_node(t & 0x1f);
    switch (t) {
        case 0xc0: out->kind = MpNode::NIL; return true;
        case 0xc2: case 0xc3: out->kind = MpNode::BOOL; out->u = t == 0xc3; return true;
        case 0xc4: case 0xd9: return be(1, &v) && bytes_node(v);
        case 0xc5: case 0xda: return be(2, &v) && bytes_node(v);
        case 0xc6: case 0xdb: return be(4, &v) && bytes_node(v);
        case 0xcc: out->kind = MpNode::INT; return be(1, &out->u);
        case 0xcd: out->kind = MpNode::INT; return be(2, &out->u);
        case 0xce: out->kind = MpNode::INT; return be(4, &out->u);
        case 0xcf: out->kind = MpNode::INT; return be(8, &out->u);
        case 0xd0: out->kind = MpNode::INT; if (!be(1, &v)) return false; out->negative = (int8_t)v < 0; out->u = v; return true;
        case 0xd1: out->kind = MpNode::INT; if (!be(2, &v)) return false; out->negative = (int16_t)v < 0; out->u = v; return true;
        case 0xd2: out->kind = MpNode::INT; if (!be(4, &v)) return false; out->negative = (int32_t)v < 0; out->u = v; return true;
        case 0xd3: out->kind = MpNode::INT; if (!be(8, &v)) return false; out->negative = (int64_t)v < 0; out->u = v; return true;
        case 0xdc: return be(2, &v) && seq(v, false);
        case 0xdd: return be(4, &v) && seq(v, false);
        case 0xde: return be(2, &v) && seq(v, true);
        case 0xdf: return be(4, &v) && seq(v, true);
        default: return false;   // floats and ext types do not occur in a witness stack
    }
}

static bool mp_field(const MpNode& n, std::vector<uint8_t>* be) {
    if (n.kind == MpNode::BYTES) return field_from_bytes(n.p, n.len, be);
    if (n.kind == MpNode::ARRAY && n.items.size() == 32) {
        be->assign(32, 0);
        for (int k = 0; k < 32; k++) {
            if (n.items[k].kind != MpNode::INT || n.items[k].negative || n.items[k].u > 255) return false;
            (*be)[k] = (uint8_t)n.items[k].u;
        }
        return true;
    }
    return false;
}

static void mp_find_witness_maps(const MpNode& n, WitnessValues* last, bool* found) {
    if (n.kind == MpNode::MAP && !n.items.empty()) {
        WitnessValues cand;
        bool ok = true;
        for (size_t i = 0; ok && i < n.items.size(); i += 2) {
            const MpNode& k = n.items[i];
            std::vector<uint8_t> be;
            ok = k.kind == MpNode::INT && !k.negative && k.u <= 0xffffffffull && mp_field(n.items[i + 1], &be);
            if (ok) cand[(uint32_t)k.u] = std::move(be);
        }
        if (ok) {
            *last = std::move(cand);
            *found = true;
            return;
        }
    }
    if (n.kind == MpNode::ARRAY || n.kind == MpNode::MAP)
        for (auto& c : n.items) mp_find_witness_maps(c, last, found);
}

static bool parse_msgpack_stack(const std::vector<uint8_t>& raw, size_t off, WitnessValues* values, std::string* why) {
    MpNode root;
    size_t o = off;
    if (!mp_parse(raw.data(), raw.size(), &o, &root, 0) || o != raw.size()) {
        *why = "not a MessagePack document";
        return false;
    }
    bool found = false;
    mp_find_witness_maps(root, values, &found);
    if (!found) *why = "MessagePack document without a witness map";
    return found;
}

// -> assignment (nb_public-1 + nb_secret values, big-endian 32 B each)
int witness_to_assignment(const Circuit& c, const uint8_t* gz, size_t gz_len, std::vector<uint8_t>* assignment_be) {
    std::vector<uint8_t> raw;
    G16_TRY(inflate_gzip(gz, gz_len, &raw));
    // `nargo execute` has written three framings over time: plain bincode, a format byte (0/1 = bincode, 2/3 = msgpack /
    // msgpack-compact) followed by the payload.  None can be checked against a real file offline, so
    // every framing is tried and the first that parses completely wins.
    WitnessValues values;
    std::string why, first_why;
    bool ok = parse_bincode_stack(raw, 0, &values, &first_why);
    if (!ok && !raw.empty() && raw[0] <= 1) ok = parse_bincode_stack(raw, 1, &values, &why);
    if (!ok && !raw.empty() && (raw[0] == 2 || raw[0] == 3)) ok = parse_msgpack_stack(raw, 1, &values, &why);
    if (!ok) ok = parse_msgpack_stack(raw, 0, &values, &why);
    if (!ok) {
        set_error("witness: unsupported container (" + first_why + "; " + why + ")");
        return G16_E_PARSE;
    }
    std::set<uint32_t> secret_idx;
    std::vector<uint32_t> secret_order;
    for (auto& name : c.secret_names) {
        const char* pfx = "__witness_";
        if (name.compare(0, strlen(pfx), pfx) != 0) {
            set_error("witness: secret variable '" + name + "' is not an ACIR witness name");
            return G16_E_PARSE;
        }
        uint32_t k = (uint32_t)strtoul(name.c_str() + strlen(pfx), nullptr, 10);
        secret_idx.insert(k);
        secret_order.push_back(k);
    }
    std::vector<uint32_t> public_order;
    for (uint32_t k = 0; public_order.size() + 1 < c.nb_public; k++)
        if (!secret_idx.count(k)) public_order.push_back(k);
    assignment_be->clear();
    auto push = [&](uint32_t k) {
        auto itv = values.find(k);
        if (itv == values.end()) {
            set_error("witness: ACIR witness " + std::to_string(k) + " is missing");
            return false;
        }
        assignment_be->insert(assignment_be->end(), itv->second.begin(), itv->second.end());
        return true;
    };
    for (uint32_t k : public_order)
        if (!push(k)) return G16_E_PARSE;
    for (uint32_t k : secret_order)
        if (!push(k)) return G16_E_PARSE;
    return G16_OK;
}

}  // namespace g16

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

// -> assignment (nb_public-1 + nb_secret values, big-endian 32 B each)
int witness_to_assignment(const Circuit& c, const uint8_t* gz, size_t gz_len, std::vector<uint8_t>* assignment_be) {
    std::vector<uint8_t> raw;
    G16_TRY(inflate_gzip(gz, gz_len, &raw));
    size_t off = 0;
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
        set_error("witness: unsupported container (expected a bincode WitnessStack)");
        return G16_E_PARSE;
    }
    std::map<uint32_t, std::vector<uint8_t>> values;  // of the LAST stack item (= main)
    for (uint64_t it = 0; it < nitems; it++) {
        uint32_t fidx;
        uint64_t nent;
        if (!u32(&fidx) || !u64(&nent) || nent > raw.size()) {
            set_error("witness: truncated stack item");
            return G16_E_PARSE;
        }
        values.clear();
        for (uint64_t e = 0; e < nent; e++) {
            uint32_t wi;
            uint64_t flen;
            if (!u32(&wi) || !u64(&flen) || !need(flen)) {
                set_error("witness: truncated entry");
                return G16_E_PARSE;
            }
            std::vector<uint8_t> be(32, 0);
            if (flen == 32) {
                memcpy(be.data(), raw.data() + off, 32);
            } else if (flen == 64) {
                for (int k = 0; k < 32; k++) {
                    int hi = hexval(raw[off + 2 * k]), lo = hexval(raw[off + 2 * k + 1]);
                    if (hi < 0 || lo < 0) {
                        set_error("witness: bad hex field element");
                        return G16_E_PARSE;
                    }
                    be[k] = (uint8_t)(hi * 16 + lo);
                }
            } else {
                set_error("witness: unsupported field element encoding (length " + std::to_string(flen) + ")");
                return G16_E_PARSE;
            }
            off += flen;
            values[wi] = std::move(be);
        }
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

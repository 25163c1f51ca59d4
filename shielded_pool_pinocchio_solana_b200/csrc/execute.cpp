// execute.cpp -- `nargo execute` for circuits whose constraints determine their own witnesses (host only).
//
// The reference produces a witness with `nargo execute` (noir_circuit/prove_linux.sh:62, client/proof.helper.ts:55:
// Prover.toml + the compiled program -> target/<name>.gz) and hands it to `sunspot prove`.  For the withdraw circuit the
// R1CS pins every intermediate ACIR witness (solver.hpp: complete_assignment), so the same file can be produced from
// Prover.toml, the program's ABI (`abi.parameters` of target/<name>.json) and the .ccs alone:
//     ABI parameter k (arrays flattened) = ACIR witness k  ->  gnark wire (public: 1 + k; secret: by its `__witness_k` name)
//     -> complete_assignment -> WitnessStack { [ { index: 0, witness: { k: value } } ] } (bincode, gzip)
// g16_execute is that; `g16prove execute` / `g16prove prove <acir> <Prover.toml> ...` are its command-line forms.
#include <string.h>
#include <zlib.h>

#include <map>
#include <string>
#include <vector>

#include "ccs.hpp"
#include "common.cuh"
#include "solver.hpp"

namespace g16 {
namespace {

// ---- a JSON value, just enough for `abi.parameters` -------------------------------------------------------------------
struct Json {
    enum Kind { NUL, BOOL, NUM, STR, ARR, OBJ } kind = NUL;
    std::string s;     // STR, or the literal of a NUM
    std::vector<Json> a;
    std::vector<std::pair<std::string, Json>> o;
    const Json* get(const char* key) const {
        for (auto& kv : o)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
};

struct JsonParser {
    const char* p;
    const char* end;
    bool ok = true;
    void ws() {
        while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++;
    }
    bool str(std::string* out) {
        if (p >= end || *p != '"') return ok = false;
        p++;
        out->clear();
        while (p < end && *p != '"') {
            if (*p == '\\' && p + 1 < end) {
                p++;
                switch (*p) {
                    case 'n': out->push_back('\n'); break;
                    case 't': out->push_back('\t'); break;
                    case 'u': p += 4; out->push_back('?'); break;   // not needed for identifiers
                    default: out->push_back(*p);
                }
                p++;
            } else {
                out->push_back(*p++);
            }
        }
        if (p >= end) return ok = false;
        p++;
        return true;
    }
    bool value(Json* v, int depth) {
        if (depth > 64) return ok = false;
        ws();
        if (p >= end) return ok = false;
        if (*p == '{') {
            v->kind = Json::OBJ;
            p++;
            ws();
            if (p < end && *p == '}') { p++; return true; }
            for (;;) {
                ws();
                std::string k;
                if (!str(&k)) return false;
                ws();
                if (p >= end || *p != ':') return ok = false;
                p++;
                v->o.emplace_back(k, Json());
                if (!value(&v->o.back().second, depth + 1)) return false;
                ws();
                if (p < end && *p == ',') { p++; continue; }
                if (p < end && *p == '}') { p++; return true; }
                return ok = false;
            }
        }
        if (*p == '[') {
            v->kind = Json::ARR;
            p++;
            ws();
            if (p < end && *p == ']') { p++; return true; }
            for (;;) {
                v->a.emplace_back();
                if (!value(&v->a.back(), depth + 1)) return false;
                ws();
                if (p < end && *p == ',') { p++; continue; }
                if (p < end && *p == ']') { p++; return true; }
                return ok = false;
            }
        }
        if (*p == '"') {
            v->kind = Json::STR;
            return str(&v->s);
        }
        const char* q = p;
        while (p < end && *p != ',' && *p != '}' && *p != ']' && *p != ' ' && *p != '\n' && *p != '\r' && *p != '\t') p++;
        v->s.assign(q, p);
        if (v->s == "null") v->kind = Json::NUL;
        else if (v->s == "true" || v->s == "false") v->kind = Json::BOOL;
        else v->kind = Json::NUM;
        return !v->s.empty() || (ok = false);
    }
};

// number of ACIR witnesses an ABI type occupies (fields, integers, booleans: 1; arrays, tuples, structs: the sum)
bool abi_width(const Json& type, size_t* out) {
    const Json* kind = type.get("kind");
    if (!kind) return false;
    if (kind->s == "field" || kind->s == "integer" || kind->s == "boolean") {
        *out = 1;
        return true;
    }
    if (kind->s == "array") {
        const Json *len = type.get("length"), *inner = type.get("type");
        size_t w;
        if (!len || !inner || !abi_width(*inner, &w)) return false;
        *out = w * (size_t)strtoull(len->s.c_str(), nullptr, 10);
        return true;
    }
    if (kind->s == "tuple" || kind->s == "struct") {
        const Json* fields = type.get("fields");
        if (!fields) return false;
        size_t total = 0;
        for (auto& f : fields->a) {
            const Json* t = f.kind == Json::OBJ && f.get("type") ? f.get("type") : &f;
            size_t w;
            if (!abi_width(*t, &w)) return false;
            total += w;
        }
        *out = total;
        return true;
    }
    return false;   // strings etc.: not used by the reference's circuits
}

// ---- Prover.toml: `key = "0x.."`, `key = 123`, `key = [ ... ]` (possibly over several lines), `#` comments ------------
bool parse_scalar(const std::string& tok, HFr* out) {
    std::string t = tok;
    if (t.size() >= 2 && t.front() == '"' && t.back() == '"') t = t.substr(1, t.size() - 2);
    if (t.empty()) return false;
    uint8_t be[32] = {0};
    if (t.size() > 2 && t[0] == '0' && (t[1] == 'x' || t[1] == 'X')) {
        std::string h = t.substr(2);
        if (h.size() > 64) return false;
        h = std::string(64 - h.size(), '0') + h;
        for (int k = 0; k < 32; k++) {
            auto hv = [](char c) { return c >= '0' && c <= '9' ? c - '0' : (c >= 'a' && c <= 'f' ? c - 'a' + 10 : (c >= 'A' && c <= 'F' ? c - 'A' + 10 : -1)); };
            int hi = hv(h[2 * k]), lo = hv(h[2 * k + 1]);
            if (hi < 0 || lo < 0) return false;
            be[k] = (uint8_t)(hi * 16 + lo);
        }
    } else {
        // decimal, any length below 2^256
        for (char c : t) {
            if (c < '0' || c > '9') return false;
            unsigned carry = (unsigned)(c - '0');
            for (int k = 31; k >= 0; k--) {
                unsigned v = be[k] * 10u + carry;
                be[k] = (uint8_t)v;
                carry = v >> 8;
            }
            if (carry) return false;
        }
    }
    *out = HFr::from_be(be);
    return true;
}

bool parse_toml(const std::string& text, std::map<std::string, std::vector<HFr>>* out, std::string* why) {
    size_t i = 0;
    const size_t n = text.size();
    auto skip_ws = [&](bool newlines) {
        while (i < n) {
            if (text[i] == '#') {
                while (i < n && text[i] != '\n') i++;
            } else if (text[i] == ' ' || text[i] == '\t' || text[i] == '\r' || (newlines && text[i] == '\n')) {
                i++;
            } else {
                break;
            }
        }
    };
    auto token = [&](std::string* tok) {
        tok->clear();
        if (i < n && text[i] == '"') {
            size_t j = text.find('"', i + 1);
            if (j == std::string::npos) return false;
            *tok = text.substr(i, j - i + 1);
            i = j + 1;
            return true;
        }
        while (i < n && (isalnum((unsigned char)text[i]) || text[i] == '_')) tok->push_back(text[i++]);
        return !tok->empty();
    };
    for (;;) {
        skip_ws(true);
        if (i >= n) return true;
        if (text[i] == '[') {   // a [table] header: nested inputs are not used by the reference's circuits
            *why = "Prover.toml tables are not supported";
            return false;
        }
        std::string key, tok;
        while (i < n && (isalnum((unsigned char)text[i]) || text[i] == '_' || text[i] == '-')) key.push_back(text[i++]);
        skip_ws(false);
        if (key.empty() || i >= n || text[i] != '=') {
            *why = "Prover.toml: expected `key = value`";
            return false;
        }
        i++;
        skip_ws(false);
        std::vector<HFr> vals;
        if (i < n && text[i] == '[') {
            i++;
            for (;;) {
                skip_ws(true);
                if (i < n && text[i] == ']') { i++; break; }
                HFr v;
                if (!token(&tok) || !parse_scalar(tok, &v)) {
                    *why = "Prover.toml: bad array element for `" + key + "`";
                    return false;
                }
                vals.push_back(v);
                skip_ws(true);
                if (i < n && text[i] == ',') i++;
            }
        } else {
            HFr v;
            if (!token(&tok) || !parse_scalar(tok, &v)) {
                *why = "Prover.toml: bad value for `" + key + "`";
                return false;
            }
            vals.push_back(v);
        }
        (*out)[key] = std::move(vals);
    }
}

int gzip_bytes(const std::vector<uint8_t>& raw, std::vector<uint8_t>* out) {
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (deflateInit2(&zs, Z_DEFAULT_COMPRESSION, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) {
        set_error("execute: zlib init failed");
        return G16_E_INTERNAL;
    }
    out->resize(deflateBound(&zs, (uLong)raw.size()) + 64);
    zs.next_in = const_cast<uint8_t*>(raw.data());
    zs.avail_in = (uInt)raw.size();
    zs.next_out = out->data();
    zs.avail_out = (uInt)out->size();
    int rc = deflate(&zs, Z_FINISH);
    size_t produced = out->size() - zs.avail_out;
    deflateEnd(&zs);
    if (rc != Z_STREAM_END) {
        set_error("execute: deflate failed");
        return G16_E_INTERNAL;
    }
    out->resize(produced);
    return G16_OK;
}

}  // namespace
}  // namespace g16

using namespace g16;

extern "C" int g16_execute(const uint8_t* ccs, size_t ccs_len, const char* acir_json, size_t acir_len, const char* prover_toml,
                           size_t toml_len, uint8_t* witness_gz, size_t* witness_len) {
    if (!ccs || !acir_json || !prover_toml || !witness_len) {
        set_error("g16_execute: bad arguments");
        return G16_E_ARG;
    }
    Circuit circ;
    G16_TRY(parse_ccs(ccs, ccs_len, &circ));
    // ---- ABI -----------------------------------------------------------------------------------------------------------
    Json root;
    JsonParser jp{acir_json, acir_json + acir_len};
    if (!jp.value(&root, 0)) {
        set_error("g16_execute: the program file is not JSON");
        return G16_E_PARSE;
    }
    const Json* abi = root.get("abi");
    const Json* params = abi ? abi->get("parameters") : nullptr;
    if (!params || params->kind != Json::ARR) {
        set_error("g16_execute: no abi.parameters in the program file");
        return G16_E_PARSE;
    }
    // ---- ACIR witness -> gnark wire ------------------------------------------------------------------------------------
    std::map<uint32_t, uint32_t> wire_of;
    std::vector<uint32_t> acir_of_input(circ.nb_public - 1 + circ.nb_secret);
    for (uint32_t pos = 0; pos < circ.secret_names.size(); pos++) {
        const std::string& name = circ.secret_names[pos];
        const char* pfx = "__witness_";
        if (name.compare(0, strlen(pfx), pfx) != 0) {
            set_error("g16_execute: secret variable '" + name + "' is not an ACIR witness name");
            return G16_E_PARSE;
        }
        const uint32_t k = (uint32_t)strtoul(name.c_str() + strlen(pfx), nullptr, 10);
        wire_of[k] = circ.nb_public + pos;
        acir_of_input[circ.nb_public - 1 + pos] = k;
    }
    {   // the public inputs are the ACIR witnesses that are not secrets, in ascending order (witness.cpp does the same)
        uint32_t k = 0;
        for (uint32_t pub = 0; pub + 1 < circ.nb_public; k++)
            if (!wire_of.count(k)) {
                wire_of[k] = 1 + pub;
                acir_of_input[pub] = k;
                pub++;
            }
    }
    // ---- Prover.toml ---------------------------------------------------------------------------------------------------
    std::map<std::string, std::vector<HFr>> inputs;
    std::string why;
    if (!parse_toml(std::string(prover_toml, toml_len), &inputs, &why)) {
        set_error("g16_execute: " + why);
        return G16_E_PARSE;
    }
    std::vector<std::pair<uint32_t, HFr>> known;
    uint32_t k = 0;
    for (auto& prm : params->a) {
        const Json *name = prm.get("name"), *type = prm.get("type");
        size_t width;
        if (!name || !type || !abi_width(*type, &width)) {
            set_error("g16_execute: unsupported ABI parameter type");
            return G16_E_PARSE;
        }
        auto it = inputs.find(name->s);
        if (it != inputs.end()) {
            if (it->second.size() != width) {
                set_error("g16_execute: `" + name->s + "` has " + std::to_string(it->second.size()) + " values, the ABI wants " + std::to_string(width));
                return G16_E_ARG;
            }
            for (size_t i = 0; i < width; i++) {
                auto w = wire_of.find(k + (uint32_t)i);
                if (w != wire_of.end()) known.push_back({w->second, it->second[i]});
            }
        }
        k += (uint32_t)width;
    }
    // ---- complete, then write the witness stack -----------------------------------------------------------------------------
    std::vector<HFr> asg;
    int rc = complete_assignment(circ, known, &asg, &why);
    if (rc != G16_OK) {
        set_error("g16_execute: " + why);
        return rc;
    }
    std::map<uint32_t, const HFr*> by_acir;
    for (size_t i = 0; i < asg.size(); i++) by_acir[acir_of_input[i]] = &asg[i];
    std::vector<uint8_t> raw;
    auto put = [&](const void* p, size_t n) { raw.insert(raw.end(), (const uint8_t*)p, (const uint8_t*)p + n); };
    const uint64_t one = 1, nent = by_acir.size(), flen = 32;
    const uint32_t zero = 0;
    put(&one, 8);
    put(&zero, 4);
    put(&nent, 8);
    for (auto& kv : by_acir) {
        uint8_t be[32];
        kv.second->to_be(be);
        put(&kv.first, 4);
        put(&flen, 8);
        put(be, 32);
    }
    std::vector<uint8_t> gz;
    G16_TRY(gzip_bytes(raw, &gz));
    if (witness_gz) {
        if (*witness_len < gz.size()) {
            set_error("g16_execute: output buffer too small");
            *witness_len = gz.size();
            return G16_E_ARG;
        }
        memcpy(witness_gz, gz.data(), gz.size());
    }
    *witness_len = gz.size();
    return G16_OK;
}

// ccs.hpp -- gnark v0.14 R1CS (`.ccs`) container, host side.
//
// Reads what `sunspot prove` reads as its third argument
// (/root/reference/client/proof.helper.ts:61-64; the committed
// /root/reference/noir_circuit/target/shielded_pool_verifier.ccs is the format pin).  Layout and
// the intcomp stream compression are documented in SURVEY.md 8(c)-fmt; gnark's own reader is
// `constraint/marshal.go` (third-party).
#pragma once
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "hostfr.hpp"

namespace g16 {

constexpr uint32_t CCS_CONST_WIRE = 0xFFFFFFFFu;

struct CommitmentInfo {
    uint32_t commitment_index = 0;               // wire that receives the challenge
    std::vector<uint32_t> private_committed;     // wire ids
    uint32_t nb_public_committed = 0;
    std::vector<uint32_t> public_and_commitment_committed;
};

enum HintKind {
    HINT_UNKNOWN = 0,
    HINT_NBITS,
    HINT_INVZERO,
    HINT_DECOMPOSE,
    HINT_COUNT,
    HINT_RANDOMIZE,
    HINT_COMMIT,
    HINT_EMULATED_MUL,        // gnark std/math/emulated.mulHint
    HINT_GRUMPKIN_SPLIT,      // sunspot sw-grumpkin.decomposeScalar
    HINT_GRUMPKIN_LIMBS,      // sunspot sw-grumpkin.decompose
};

struct Circuit {
    // instruction stream
    std::vector<std::vector<uint32_t>> levels;
    std::vector<uint32_t> blueprint, constraint_offset, wire_offset;
    std::vector<uint64_t> start_calldata;
    std::vector<uint32_t> calldata;
    std::vector<HFr> coeffs;  // Montgomery
    std::vector<HFr> coeff_invs;  // 1/coeff (0 for 0), for the solver
    // sizes
    uint32_t nb_public = 0;   // including the ONE wire
    uint32_t nb_secret = 0;
    uint32_t nb_internal = 0;
    uint32_t nb_constraints = 0;
    uint32_t nb_wires() const { return nb_public + nb_secret + nb_internal; }
    unsigned log_domain() const {
        unsigned l = 0;
        while (((uint64_t)1 << l) < nb_constraints) l++;
        return l;
    }
    std::vector<std::string> public_names, secret_names;
    std::vector<CommitmentInfo> commitments;
    std::map<uint32_t, std::string> hint_names;
    std::map<uint32_t, HintKind> hint_kinds;

    // R1CS rows in constraint order, CSR over (coeff id, wire id); wire CONST -> wire 0 (ONE)
    struct Csr {
        std::vector<uint32_t> rowptr, coeff, wire;
    };
    Csr A, B, C;
};

// returns G16_OK or G16_E_PARSE (message via set_error)
int parse_ccs(const uint8_t* buf, size_t len, Circuit* out);

}  // namespace g16

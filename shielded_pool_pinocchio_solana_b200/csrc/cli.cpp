// cli.cpp -- `g16prove`: command-line stand-in for the sunspot binary on the proving path.
//
//   g16prove prove <acir.json> <witness.gz> <ccs> <pk>      same argv as `sunspot prove`
//        (/root/reference/client/proof.helper.ts:64, noir_circuit/prove_linux.sh:83); writes
//        <dir-of-ccs>/<name>.proof and <name>.pw, the files proof.helper.ts:68-69 and
//        client/generate-proof-hex.ts:18-27 read back.
//        <witness.gz> may be a Prover.toml instead (name ending in .toml): the witness is then rebuilt from it
//        (g16_execute), for circuits whose constraints determine their witnesses -- the withdraw circuit
//   g16prove execute <acir.json> <Prover.toml> <ccs> <witness.gz-out>   that step alone (`nargo execute`, prove_linux.sh:62;
//        host only, no GPU needed)
//   g16prove setup <ccs>                                    same argv as `sunspot setup` (prove_linux.sh:78):
//        writes <name>.pk and <name>.vk next to the .ccs; fresh OS entropy (aborts if it cannot be read)
//   g16prove setup <ccs> <pk-out> <vk-out> [seed]           explicit outputs / reproducible seed (tests)
//   g16prove verify <vk> <proof> <pw>                       same argv as `sunspot verify` (prove_linux.sh:87):
//        exit 0 iff the proof is accepted (host-only pairing check, no GPU needed)
// Exit code 0 on success; non-zero with the library's message on stderr otherwise (execSync throws
// on non-zero exit, which is the reference's only error channel).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <fstream>
#include <iterator>
#include <string>
#include <vector>

#include "../../include/g16b200.h"

static bool slurp(const char* path, std::vector<uint8_t>* out) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    out->assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
    return true;
}
static bool spill(const std::string& path, const uint8_t* p, size_t n) {
    std::ofstream f(path, std::ios::binary);
    f.write((const char*)p, (std::streamsize)n);
    return (bool)f;
}
static int die(const char* what) {
    fprintf(stderr, "g16prove: %s: %s\n", what, g16_last_error());
    return 1;
}

int main(int argc, char** argv) {
    if (argc < 2) {
        fprintf(stderr, "usage: g16prove prove <acir.json> <witness.gz | Prover.toml> <ccs> <pk>\n"
                        "       g16prove execute <acir.json> <Prover.toml> <ccs> <witness.gz-out>\n"
                        "       g16prove setup <ccs> [<pk-out> <vk-out> [seed]]\n"
                        "       g16prove verify <vk> <proof> <pw>\n");
        return 2;
    }
    std::string cmd = argv[1];
    if (cmd == "verify" && argc == 5) {   // host only: no CUDA context
        std::vector<uint8_t> vk, proof, pw;
        if (!slurp(argv[2], &vk) || !slurp(argv[3], &proof) || !slurp(argv[4], &pw)) {
            fprintf(stderr, "g16prove: cannot read input files\n");
            return 1;
        }
        int ok = 0;
        if (g16_verify(vk.data(), vk.size(), proof.data(), proof.size(), pw.data(), pw.size(), &ok) != G16_OK) return die("verify");
        if (!ok) {
            fprintf(stderr, "g16prove: proof rejected\n");
            return 1;
        }
        printf("proof accepted\n");
        return 0;
    }
    auto execute = [](const char* acir_path, const char* toml_path, const std::vector<uint8_t>& ccs, std::vector<uint8_t>* gz) {
        std::vector<uint8_t> acir, toml;
        if (!slurp(acir_path, &acir) || !slurp(toml_path, &toml)) {
            fprintf(stderr, "g16prove: cannot read %s / %s\n", acir_path, toml_path);
            return false;
        }
        size_t n = 0;
        if (g16_execute(ccs.data(), ccs.size(), (const char*)acir.data(), acir.size(), (const char*)toml.data(), toml.size(), nullptr, &n) != G16_OK) {
            die("execute");
            return false;
        }
        gz->resize(n);
        if (g16_execute(ccs.data(), ccs.size(), (const char*)acir.data(), acir.size(), (const char*)toml.data(), toml.size(), gz->data(), &n) != G16_OK) {
            die("execute");
            return false;
        }
        gz->resize(n);
        return true;
    };
    if (cmd == "execute" && argc == 6) {   // host only
        std::vector<uint8_t> ccs, gz;
        if (!slurp(argv[4], &ccs)) {
            fprintf(stderr, "g16prove: cannot read %s\n", argv[4]);
            return 1;
        }
        if (!execute(argv[2], argv[3], ccs, &gz)) return 1;
        if (!spill(argv[5], gz.data(), gz.size())) {
            fprintf(stderr, "g16prove: cannot write %s\n", argv[5]);
            return 1;
        }
        return 0;
    }
    int dev = getenv("G16_DEVICE") ? atoi(getenv("G16_DEVICE")) : 0;
    g16_ctx* ctx = nullptr;
    if (g16_init(&dev, 1, &ctx) != G16_OK) return die("init");
    if (cmd == "prove" && argc == 6) {
        std::vector<uint8_t> acir, gz, ccs, pk;
        const std::string wpath = argv[3];
        const bool from_toml = wpath.size() > 5 && wpath.compare(wpath.size() - 5, 5, ".toml") == 0;
        if ((!from_toml && !slurp(argv[3], &gz)) || !slurp(argv[4], &ccs) || !slurp(argv[5], &pk)) {
            fprintf(stderr, "g16prove: cannot read input files\n");
            return 1;
        }
        if (from_toml && !execute(argv[2], argv[3], ccs, &gz)) return 1;
        slurp(argv[2], &acir);  // optional: the .ccs name lists carry the witness mapping
        g16_circuit* c = nullptr;
        if (g16_circuit_load(ctx, ccs.data(), ccs.size(), pk.data(), pk.size(), nullptr, &c) != G16_OK) return die("load");
        uint8_t proof[G16_PROOF_LEN];
        std::vector<uint8_t> pw(12 + 32 * 4096);
        size_t pl = sizeof proof, wl = pw.size();
        if (g16_prove(c, gz.data(), gz.size(), nullptr, proof, &pl, pw.data(), &wl) != G16_OK) return die("prove");
        std::string base = argv[4];
        size_t dot = base.rfind('.');
        if (dot != std::string::npos && base.find('/', dot) == std::string::npos) base.resize(dot);
        if (!spill(base + ".proof", proof, pl) || !spill(base + ".pw", pw.data(), wl)) {
            fprintf(stderr, "g16prove: cannot write outputs next to %s\n", argv[4]);
            return 1;
        }
        g16_circuit_free(c);
    } else if (cmd == "setup" && (argc == 3 || argc == 5 || argc == 6)) {
        std::vector<uint8_t> ccs;
        if (!slurp(argv[2], &ccs)) {
            fprintf(stderr, "g16prove: cannot read %s\n", argv[2]);
            return 1;
        }
        std::string seed = argc == 6 ? argv[5] : "";
        if (seed.empty()) {  // fresh entropy unless a seed is given (tests / reproducible keys)
            std::ifstream ur("/dev/urandom", std::ios::binary);
            seed.resize(32);
            ur.read(&seed[0], 32);
            if (!ur || ur.gcount() != 32) {   // an all-zero seed would be publicly known toxic waste
                fprintf(stderr, "g16prove: cannot read 32 bytes of entropy from /dev/urandom\n");
                return 1;
            }
        }
        std::string pk_path, vk_path;
        if (argc == 3) {
            std::string base = argv[2];
            size_t dot = base.rfind('.');
            if (dot != std::string::npos && base.find('/', dot) == std::string::npos) base.resize(dot);
            pk_path = base + ".pk";
            vk_path = base + ".vk";
        } else {
            pk_path = argv[3];
            vk_path = argv[4];
        }
        size_t pl = 0, vl = 0;
        if (g16_setup(ctx, ccs.data(), ccs.size(), (const uint8_t*)seed.data(), seed.size(), nullptr, &pl, nullptr, &vl) != G16_OK)
            return die("setup");
        std::vector<uint8_t> pk(pl), vk(vl);
        if (g16_setup(ctx, ccs.data(), ccs.size(), (const uint8_t*)seed.data(), seed.size(), pk.data(), &pl, vk.data(), &vl) != G16_OK)
            return die("setup");
        if (!spill(pk_path, pk.data(), pl) || !spill(vk_path, vk.data(), vl)) {
            fprintf(stderr, "g16prove: cannot write key files\n");
            return 1;
        }
    } else {
        fprintf(stderr, "g16prove: bad command line\n");
        return 2;
    }
    g16_shutdown(ctx);
    return 0;
}

// gpusolver.cuh -- batched R1CS witness solver on the GPU (see gpusolver.cu).
#pragma once
#include <string>
#include <utility>
#include <vector>

#include "ccs.hpp"
#include "common.cuh"
#include "ff.cuh"

namespace g16 {

// slots appended to every proof's wire vector on the device
enum { X_ONE = 0, X_R = 1, X_S = 2, X_NEG_RS = 3, X_BLINDER = 4, X_COUNT = 8 };

struct GpuSolverPlan {
    bool valid = false;
    uint32_t nlevels = 0;
    uint32_t commit_level = (uint32_t)-1;   // level holding the BSB22 commitment hint, or -1
    uint32_t commit_wire = 0;
    // linear expressions over INPUT wires whose values are hashed into the challenge after the commitment point
    std::vector<std::vector<std::pair<uint32_t, uint32_t>>> commit_hashed;   // (coefficient id, wire)
    uint32_t *d_lvl_off = nullptr, *d_lvl_instr = nullptr, *d_instr_cd = nullptr, *d_calldata = nullptr;
    uint4* d_info = nullptr;
    uint4* d_rec = nullptr;                 // flattened plan: two uint4 per instruction, in level order
    // two-phase plan (k_solve_2p): levels split into sub-levels of bounded product count
    uint32_t *d_sub_off = nullptr, *d_prod_off = nullptr;   // sub-level -> row range / product range
    uint2 *d_prods = nullptr, *d_ops = nullptr;             // (coefficient id, wire) per product; (code, ref) per operand
    uint2* d_iprods = nullptr;                              // products of rows too long for the shared buffer (multiplied in place)
    uint4* d_rec2 = nullptr;                                // per row: (mode, wire, coefficient code, row), (first operand, nL, nR, nO)
    std::vector<uint32_t> lvl_to_sub;                       // level -> first sub-level (size nlevels + 1)
    Fr* d_coeff_invs = nullptr;
    // Hints that work over the integers (emulated.mulHint, the sw-grumpkin scalar split) have no device version; when
    // their inputs only depend on the circuit's inputs (and on each other) the host evaluates them per proof before
    // the device solve and their outputs are scattered into the wire vectors like extra inputs.
    std::vector<uint32_t> host_hints;   // instruction ids, in dependency order
    std::vector<uint32_t> host_wires;   // their output wires, concatenated
    std::vector<uint32_t> host_inputs;  // the circuit INPUT wires those hints read (only these are converted on the host)
    uint32_t* d_host_wires = nullptr;
    // level ranges by kernel: wide levels run a CTA per proof (k_solve_tpi), runs of thin levels a thread per proof
    struct Segment {
        uint32_t begin, end;
        bool narrow;
    };
    std::vector<Segment> segments;

    // Compiles the plan; leaves valid == false (and says why) when the circuit needs the host solver.
    int build(const Circuit& c, cudaStream_t st, std::string* why_not);
    // big-endian assignments / (r,s,blinder) triples already on the device -> wires
    int assign(const uint8_t* d_asg_be, const uint8_t* d_rnd_be, uint32_t nin, Fr* d_wires, size_t wstride, size_t nw,
               size_t B, cudaStream_t st) const;
    int run(const Fr* d_coeffs, int unit_ids, Fr* d_wires, size_t wstride, size_t nw, size_t B, uint32_t lvl_begin,
            uint32_t lvl_end, uint32_t* d_err, cudaStream_t st) const;
    int set_wire(Fr* d_wires, size_t wstride, uint32_t wire, const Fr* d_values, size_t B, cudaStream_t st) const;
    // d_values[b * host_wires.size() + i] -> wire host_wires[i] of proof b
    int scatter_host_wires(Fr* d_wires, size_t wstride, const Fr* d_values, size_t B, cudaStream_t st) const;
    void release();
    ~GpuSolverPlan() { release(); }
};

}  // namespace g16

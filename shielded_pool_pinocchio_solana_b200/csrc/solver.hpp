// solver.hpp -- R1CS witness solver (gnark `constraint/bn254/solver.go` semantics): walks the
// instruction levels of a parsed `.ccs`, solves each R1C row for its single unknown wire and runs
// the solver hints.  SURVEY.md 3.2 step 1, 8a rows a3 / a3-H.
//
// The BSB22 commitment hint needs a Pedersen MSM in the middle of the solve; that MSM runs on the
// GPU, so the solver is resumable: run() returns SOLVE_NEED_COMMITMENT with the committed values,
// the caller (prove.cu) computes the commitment for the whole batch in one launch, hashes it and
// calls provide_challenge() before run() again.
#pragma once
#include <string>
#include <utility>
#include <vector>

#include "ccs.hpp"

namespace g16 {

enum { SOLVE_DONE = 0, SOLVE_NEED_COMMITMENT = 100 };

struct SolveState {
    std::vector<HFr> w;          // all wires, Montgomery
    std::vector<uint8_t> known;
    size_t level = 0, pos = 0;   // resume point
    size_t commitments_done = 0;
    // filled when run() returns SOLVE_NEED_COMMITMENT
    std::vector<HFr> committed;  // private committed values, in PrivateCommitted order
    std::vector<HFr> hashed;     // public/commitment committed values
    uint32_t challenge_wire = 0;
    std::string error;           // filled on failure
    // diagnostic mode (tests: G16_SOLVER_DIAG=1 through g16_solve_assignment): unsatisfied rows are counted, not
    // fatal, and a failing table lookup yields zeros -- lets a test drive the hints of a circuit whose full
    // witness cannot be produced offline
    bool tolerate = false;
    size_t failed_rows = 0;
};

// assignment: nb_public-1 public values then nb_secret secret values (Montgomery)
void solve_begin(const Circuit& c, const HFr* assignment, SolveState* st);
// returns SOLVE_DONE, SOLVE_NEED_COMMITMENT, or a G16_E_* code (st->error holds the message)
int solve_run(const Circuit& c, SolveState* st, const HFr* blinder);
void solve_provide_challenge(SolveState* st, const HFr& challenge);
// one hint instruction on its own (the device solver's host-evaluated hints, gpusolver.cuh); G16_OK or a G16_E_* code
int solve_run_hint(const Circuit& c, SolveState* st, uint32_t instr);

// ---- witness completion ("ACVM-lite") ----------------------------------------------------------
// sunspot declares every ACIR witness a circuit's constraints read as a secret input, so an assignment normally comes
// from `nargo execute`.  When the constraints themselves determine those witnesses -- arithmetic gates, limb / bit
// decompositions, is-zero gadgets: the case of the reference's withdraw circuit -- this rebuilds the assignment from a
// subset of the inputs (the program's ABI inputs) by data-flow propagation over the R1CS:
//   1. a row with one unknown wire defines it; a hint runs once its inputs are known;
//   2. a linear row whose unknowns carry distinct powers of two is a radix decomposition: the known side is split,
//      widths from the range checks (rangecheck.DecomposeHint, bits.nBits, b * (1 - b) = 0 rows);
//   3. what stalls after that is solved as small linear systems (rows linear in their unknowns, per connected
//      component, Gaussian elimination over Fr); a variable that stays free is set to 0.
// `known`: (wire id, value) pairs.  On success `assignment` holds nb_public-1 + nb_secret values (Montgomery).
// Returns G16_OK, G16_E_UNSAT (a fully known row does not hold, or inputs stay undetermined) or G16_E_HINT.
int complete_assignment(const Circuit& c, const std::vector<std::pair<uint32_t, HFr>>& known, std::vector<HFr>* assignment,
                        std::string* err);

}  // namespace g16

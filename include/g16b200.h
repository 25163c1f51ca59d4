/* g16b200.h -- C ABI of libg16b200.so, a B200-native Groth16 (BN254) prover.
 *
 * Drop-in boundary for the ONE step the reference shells out for:
 *     execSync(`sunspot prove ${acir} ${witness.gz} ${ccs} ${pk}`)
 *         /root/reference/client/proof.helper.ts:58-66
 *     same argv in noir_circuit/prove_linux.sh:83, audit_circuit/prove_audit.sh:95,
 *     scripts/generate_audit.py:680, scripts/benchmark_all.py:683
 * and for the files that step leaves behind (`<name>.proof`, `<name>.pw`,
 * client/proof.helper.ts:68-69, client/generate-proof-hex.ts:18-27).
 *
 * Conventions
 *   - every function returns 0 (G16_OK) or a G16_E_* code; g16_last_error() gives the
 *     thread-local message for the last failure on the calling thread.
 *   - the caller owns every host buffer; the library owns device memory behind the opaque
 *     handles.  A proving key is uploaded (and expanded into window tables) ONCE per
 *     g16_circuit and reused for every proof.
 *   - wire formats are gnark's: Fr = 32 B big-endian canonical; G1 = X||Y (64 B raw);
 *     G2 = X.A1||X.A0||Y.A1||Y.A0 (128 B raw); infinity = all zeros (gnark-crypto RawBytes).
 *   - "dev" entry points take DEVICE pointers (e.g. torch tensors' data_ptr()) holding
 *     little-endian 8x32-bit limbs; they only enqueue work on the context stream.
 *   - there is no CPU fallback: every compute entry point fails with G16_E_CUDA when no
 *     sm_100-class device is usable.
 */
#ifndef G16B200_H
#define G16B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    G16_OK = 0,
    G16_E_ARG = 1,
    G16_E_PARSE = 2,  /* malformed .ccs / .pk / witness */
    G16_E_UNSAT = 3,  /* witness does not satisfy the R1CS (message names the row) */
    G16_E_CUDA = 4,
    G16_E_HINT = 5,   /* unknown / failing solver hint */
    G16_E_NOMEM = 6,
    G16_E_INTERNAL = 7
};

#define G16_PROOF_LEN 388 /* shielded_pool_program/src/instructions/withdraw.rs:13, submit_audit.rs:18 */

typedef struct g16_ctx g16_ctx;
typedef struct g16_circuit g16_circuit;
typedef struct g16_bases g16_bases;

/* ---- context ------------------------------------------------------------------------- */
/* One context drives one GPU (one process per GPU under torchrun).  n_devices must be 1. */
int g16_init(const int* device_ids, int n_devices, g16_ctx** out);
void g16_shutdown(g16_ctx* ctx);
const char* g16_last_error(void);
/* Run all subsequent work of `ctx` on the caller's CUDA stream (a cudaStream_t, e.g.
 * torch.cuda.current_stream().cuda_stream); NULL restores the context's own stream. */
int g16_set_stream(g16_ctx* ctx, void* cuda_stream);
int g16_sync(g16_ctx* ctx);
/* Device-side calls (g16_prove_wires_dev) normally end by making the context stream wait for their last kernel, so
 * that their outputs are ordered on that stream.  With deferred joins ON they return without that wait: consecutive
 * calls then overlap on the device (the latency-bound end of one under the start of the next).  Outputs are
 * complete once work enqueued after g16_join(ctx) -- which inserts the pending waits into the context stream --
 * runs, or after g16_sync.  Host-side calls (g16_prove*, g16_prove_batch, g16_prove_wires) are unaffected. */
int g16_set_deferred_join(g16_ctx* ctx, int on);
int g16_join(g16_ctx* ctx);
/* kernels launched by the last compute call on this context (bench.py `gpu_launches`) */
int g16_last_launches(g16_ctx* ctx);
/* Per-kernel device timing with CUDA events on the context stream (bench.py's roofline line).
 * Slots: 0 = G1 bucket accumulation, 1 = G2 bucket accumulation (ms, launches, points). */
int g16_profile_enable(g16_ctx* ctx, int enable);
int g16_profile_read(g16_ctx* ctx, double ms[8], double launches[8], double units[8]);
/* Convert `count` Fr values (32 B big-endian) to Montgomery limb form in a caller-owned device
 * buffer (count * 32 bytes), e.g. to stage inputs for the *_dev entry points. */
int g16_fr_to_device(g16_ctx* ctx, const uint8_t* values_be, size_t count, void* d_out);
/* Pipe microbenchmark: kind 0 = IMAD (mad.lo.u32), 2 = FP64 FMA, 1 = IMAD.WIDE.U32 with carry
 * (mad.lo.cc/madc.hi.cc pairs, the instruction the field multiplier is made of).
 * Writes instructions per second over the whole chip. */
int g16_measure_imad_peak(g16_ctx* ctx, int kind, double* instr_per_s);

/* ---- standalone kernels ("kernels exposed for the sweep") ------------------ */
/* Upload n bases and expand them into window tables.  window = 0 picks one for
 * (n, batch_hint).  Replaces the pk.G1.{A,B,K,Z} / pk.G2.B slices gnark keeps on the heap. */
int g16_bases_load_g1(g16_ctx* ctx, const uint8_t* points_be, size_t n, int window, size_t batch_hint,
                      g16_bases** out);
int g16_bases_load_g2(g16_ctx* ctx, const uint8_t* points_be, size_t n, int window, size_t batch_hint,
                      g16_bases** out);
void g16_bases_free(g16_bases* b);
int g16_bases_window(const g16_bases* b);
/* out[b] = sum_i scalars[b][i] * P_i  -- replaces G1Jac.MultiExp / G2Jac.MultiExp.
 * scalars_be: batch*n*32 B; out_be: batch*64 B (G1) / batch*128 B (G2). Host buffers. */
int g16_msm_g1(g16_ctx* ctx, const g16_bases* bases, const uint8_t* scalars_be, size_t batch, uint8_t* out_be);
int g16_msm_g2(g16_ctx* ctx, const g16_bases* bases, const uint8_t* scalars_be, size_t batch, uint8_t* out_be);
/* Device-resident variant: d_scalars = batch*n Fr (8 LE limbs, Montgomery iff montgomery!=0),
 * d_out = batch affine points in Montgomery limb form.  Asynchronous on the context stream. */
int g16_msm_dev(g16_ctx* ctx, const g16_bases* bases, const void* d_scalars, int montgomery, size_t batch,
                void* d_out);

/* Synthetic bases for the kernel sweep / throughput runs:
 * P_i = [k_i]G, k_i = 253-bit value built from SplitMix64(seed + 4i .. 4i+3), computed on the
 * device; written as gnark raw points to a host buffer (n*64 B, or n*128 B when g2 != 0). */
int g16_generate_points(g16_ctx* ctx, int g2, uint64_t seed, size_t n, uint8_t* out_be);

/* In-place Fr NTT of `batch` vectors of length n = 2^logn (replaces gnark-crypto fr/fft
 * Domain.FFT / FFTInverse).
 *   inverse = 0: natural-order input  -> bit-reversed output (DIF), optional coset shift g=5 first
 *   inverse = 1: bit-reversed input   -> natural-order output (DIT), 1/n and optional g^-j after
 * values_be: batch*n*32 B big-endian canonical, transformed in place (host buffer). */
int g16_ntt(g16_ctx* ctx, uint8_t* values_be, unsigned logn, size_t batch, int inverse, int coset);
/* device-resident variant: d_values = batch*n Fr in Montgomery limb form */
int g16_ntt_dev(g16_ctx* ctx, void* d_values, unsigned logn, size_t batch, int inverse, int coset);

/* Quotient polynomial of Groth16 (replaces gnark computeH, backend/groth16/bn254/prove.go;
 * SURVEY.md 8a row a6).  abc_be: per proof three vectors of n = 2^logn Fr (A.w, B.w, C.w on the
 * domain, natural order, zero padded).  h_be: per proof n Fr = coefficients of
 * H = (A*B - C)/(X^n - 1) in BIT-REVERSED order (entry n-1 is zero), the order pk.G1.Z uses. */
int g16_compute_h(g16_ctx* ctx, const uint8_t* abc_be, unsigned logn, size_t nproofs, uint8_t* h_be);
/* device variant: d_abc = nproofs * 3n Fr (Montgomery limbs); H overwrites the first vector of
 * every triple. */
int g16_compute_h_dev(g16_ctx* ctx, void* d_abc, unsigned logn, size_t nproofs);

/* ---- circuits and proofs ---------------------------------------------------------------- */
/* Parse a gnark v0.14 R1CS (.ccs) and proving key (.pk), upload and expand the key.
 * acir_json may be NULL: the .ccs `Public`/`Secret` name lists carry the witness mapping. */
int g16_circuit_load(g16_ctx* ctx, const uint8_t* ccs, size_t ccs_len, const uint8_t* pk, size_t pk_len,
                     const char* acir_json, g16_circuit** out);
void g16_circuit_free(g16_circuit* c);
/* "gpu" when witnesses of this circuit are solved by the batched device solver; otherwise the
 * reason the host solver is used (an unimplemented hint, G16_HOST_SOLVER set, ...). */
const char* g16_circuit_solver(const g16_circuit* c);
/* sizes: what[0]=nbConstraints [1]=nbWires [2]=nbPublic(incl. ONE) [3]=nbSecret [4]=domain size
 *        [5]=nbCommitments [6..10] = MSM sizes A,B,K,Z,commit [11]=max proofs per device batch
 *        [12..15] = window bits chosen for the A, B1, K+Z and B2 MSMs */
int g16_circuit_info(const g16_circuit* c, uint64_t what[16]);

/* Trusted setup on the GPU -- the `sunspot setup` step (noir_circuit/prove_linux.sh:73-79).  The
 * toxic waste is derived from `seed` (deterministic; use fresh entropy and discard it for a real
 * ceremony).  Writes gnark raw ProvingKey / VerifyingKey bytes.  Call with pk_out = vk_out = NULL
 * to obtain the sizes in *pk_len / *vk_len. */
int g16_setup(g16_ctx* ctx, const uint8_t* ccs, size_t ccs_len, const uint8_t* seed, size_t seed_len,
              uint8_t* pk_out, size_t* pk_len, uint8_t* vk_out, size_t* vk_len);

/* One proof from a Noir witness (`target/<name>.gz`, as written by `nargo execute`).
 * rnd = r || s || commitment blinder (3 x 32 B big-endian); NULL draws them from the OS CSPRNG.
 * proof: G16_PROOF_LEN bytes (gnark Proof.WriteRawTo); pw: 12 + 32*nPublic bytes (witness.WriteTo,
 * public part).  *_len: in = capacity, out = bytes written. */
int g16_prove(g16_circuit* c, const uint8_t* witness_gz, size_t witness_len, const uint8_t rnd[96],
              uint8_t* proof, size_t* proof_len, uint8_t* pw, size_t* pw_len);
/* Host only (no GPU): the witness-ingest step alone -- Noir witness file (gzip of a WitnessStack: plain bincode, or a
 * format byte followed by bincode / MessagePack, field elements as 32 raw bytes or 64 hex characters) -> the
 * public + secret assignment of the circuit in `.ccs` order, 32 B big-endian each.  Two-call pattern: out = NULL
 * returns the value count in *n_values. */
int g16_witness_to_assignment(const uint8_t* ccs, size_t ccs_len, const uint8_t* witness_gz, size_t witness_len,
                              uint8_t* assignment_be, size_t* n_values);
/* Host only: "ACVM-lite" (SURVEY.md 8f-3).  sunspot declares every ACIR witness a circuit reads as a secret input;
 * when the constraints themselves determine those witnesses (arithmetic gates, limb / bit decompositions, is-zero
 * gadgets -- the reference's withdraw circuit) this rebuilds the whole assignment from a subset of the input wires,
 * typically the program's ABI inputs (noir_circuit/src/main.nr:39-53; values as in client/prover-params.toml).
 * known_wires[i] = wire id (1.. = public inputs, then secret inputs, `.ccs` order), known_values_be = 32 B big-endian
 * each.  Public inputs may be left out: they are derived like any other wire.  Two-call pattern: assignment_be = NULL
 * returns the value count.  G16_E_UNSAT when the inputs contradict the constraints or do not determine the rest. */
int g16_complete_assignment(const uint8_t* ccs, size_t ccs_len, const uint32_t* known_wires, const uint8_t* known_values_be,
                            size_t n_known, uint8_t* assignment_be, size_t* n_values);
/* Host only: `nargo execute` for such circuits (noir_circuit/prove_linux.sh:62, client/proof.helper.ts:55).  Prover.toml
 * (scalars as "0x.." / decimal, arrays) + the compiled program's JSON (only `abi.parameters` is read) + the .ccs ->
 * the witness file `sunspot prove` / g16_prove read (gzip of a bincode WitnessStack).  Inputs the TOML leaves out
 * (e.g. the public ones) are derived; inputs it gives are checked.  Two-call pattern on witness_gz / *witness_len. */
int g16_execute(const uint8_t* ccs, size_t ccs_len, const char* acir_json, size_t acir_len, const char* prover_toml,
                size_t toml_len, uint8_t* witness_gz, size_t* witness_len);
/* Witness only: full wire vectors (n * nbWires * 32 B big-endian) for n assignments -- the R1CS
 * solve of gnark's Prove, with the BSB22 commitment MSM on the GPU.  rnd as in g16_prove_batch
 * (only the blinder is used). */
int g16_witness_batch(g16_circuit* c, size_t n, const uint8_t* assignments_be, size_t n_values, const uint8_t* rnd,
                      uint8_t* wires_be);
/* The same through the batched DEVICE solver (the one g16_prove_batch uses); G16_E_HINT when the circuit has to be
 * solved on the host (g16_circuit_solver() says why). */
int g16_witness_batch_dev(g16_circuit* c, size_t n, const uint8_t* assignments_be, size_t n_values, const uint8_t* rnd,
                          uint8_t* wires_be);
/* Same, from the public+secret assignment directly (32 B big-endian each, `.ccs` order:
 * Public[1..] then Secret[..]) -- skips the Noir container, runs the R1CS solver. */
int g16_prove_assignment(g16_circuit* c, const uint8_t* assignment_be, size_t n_values, const uint8_t rnd[96],
                         uint8_t* proof, size_t* proof_len, uint8_t* pw, size_t* pw_len);
/* n independent proofs of the same circuit in one device batch.  assignments_be = n *
 * n_values * 32 B; rnd = n*96 B or NULL; proofs = n*388 B; pws = n*pw_stride B. */
int g16_prove_batch(g16_circuit* c, size_t n, const uint8_t* assignments_be, size_t n_values, const uint8_t* rnd,
                    uint8_t* proofs, uint8_t* pws, size_t pw_stride);
/* Bypass the solver: full wire vectors (nbWires * 32 B big-endian each, wire 0 = 1); rnd as in
 * g16_prove_batch (NULL => CSPRNG).  The
 * commitment wire must already hold the BSB22 challenge consistent with `rnd`'s blinder when a
 * valid proof is wanted; for throughput runs any vector does identical work. */
int g16_prove_wires(g16_circuit* c, size_t n, const uint8_t* wires_be, const uint8_t* rnd, uint8_t* proofs);
/* Device-resident throughput path: d_wires = n * nbWires Fr in Montgomery limb form already in
 * HBM (n <= max_batch, g16_circuit_info what[11]).  rnd = n*96 B host bytes (r || s || unused) or
 * NULL => r, s from the OS CSPRNG (never zero blinding).  d_proof_points receives per proof
 * 320 bytes: Ar (G1) | Bs (G2) | Krs (G1) | PoK (G1), affine, canonical little-endian limbs.
 * Asynchronous on the context stream. */
int g16_prove_wires_dev(g16_circuit* c, size_t n, const void* d_wires, const uint8_t* rnd, void* d_proof_points);
/* Solver only (no GPU): extend an assignment to the full wire vector (nbWires * 32 B BE).
 * challenges_be: the BSB22 challenge of each commitment, supplied by the caller;
 * committed_be (optional): receives the committed values of the first commitment. */
int g16_solve_assignment(const uint8_t* ccs, size_t ccs_len, const uint8_t* assignment_be, size_t n_values,
                         const uint8_t* blinder_be, const uint8_t* challenges_be, size_t n_challenges,
                         uint8_t* wires_be, size_t wires_cap, uint8_t* committed_be, size_t committed_cap);

/* ---- one large proof across several GPUs (BASELINE.json configs[3]) ----------------
 * One process per GPU.  Rank 0 draws an id (g16_comm_unique_id) and hands it to the others (any channel:
 * torch.distributed broadcast, a file, MPI); every rank then calls g16_comm_init on its context.  A circuit
 * loaded on such a context keeps only this rank's contiguous index range of every MSM point set;
 * g16_prove_wires / g16_prove_wires_dev become COLLECTIVE calls (same wires on every rank): each rank
 * runs SpMV + H on its own copy and Pippenger on its slice, the partial G1 / G2 sums are combined with one
 * NCCL all-gather of a few hundred bytes and every rank returns the same proof bytes.  Batches of
 * independent proofs need none of this: they shard by proof with no collective. */
int g16_comm_unique_id(uint8_t id[128]);
int g16_comm_init(g16_ctx* ctx, const uint8_t id[128], int rank, int world);
/* Synthetic constraint system in gnark's .ccs container (row k: (w_p)(sum_5 b_j w_qj) = w_new), generated
 * natively because configs[3] needs 2^22 rows.  Two-call pattern: out = NULL returns the size. */
int g16_synth_ccs(uint64_t n_constraints, uint32_t n_public, uint32_t n_secret, uint64_t seed, uint8_t* out,
                  size_t* out_len);

/* ---- verification (host only, no GPU needed) ------------------------------------------------- */
/* `sunspot verify <vk> <proof> <pw>` (noir_circuit/prove_linux.sh:87, audit_circuit/prove_audit.sh:99):
 * gnark groth16.Verify incl. the BSB22 commitment and its Pedersen proof of knowledge.
 * vk = VerifyingKey.WriteRawTo bytes (e.g. noir_circuit/target/shielded_pool_verifier.vk), proof =
 * Proof.WriteRawTo (388 B), pw = public witness.  Returns G16_OK with *ok = 1 (accepted) or 0
 * (well-formed but rejected); G16_E_PARSE when an input cannot be decoded.  Thread-safe. */
int g16_verify(const uint8_t* vk, size_t vk_len, const uint8_t* proof, size_t proof_len, const uint8_t* pw,
               size_t pw_len, int* ok);

#ifdef __cplusplus
}
#endif
#endif /* G16B200_H */

/// <reference types="node" />
/** Opaque handle: parsed constraint system + window-expanded proving key resident on one GPU. */
export type Circuit = { readonly __g16circuit: unique symbol };
export interface ProofFiles { proof: Buffer; publicWitness: Buffer }
export function loadCircuit(ccs: Buffer, pk: Buffer, device?: number): Circuit;
export function proveSync(circuit: Circuit, witnessGz: Buffer): ProofFiles;
export function prove(circuit: Circuit, witnessGz: Buffer): Promise<ProofFiles>;
export function proveBatch(circuit: Circuit, assignments: Buffer, n: number):
  Promise<{ proofs: Buffer; publicWitnesses: Buffer; pwStride: number }>;
export function verify(vk: Buffer, proof: Buffer, publicWitness: Buffer): boolean;
/** `nargo execute` stand-in for circuits whose constraints determine their witnesses: Prover.toml -> witness file. */
export function execute(ccs: Buffer, programJson: Buffer, proverToml: Buffer): Buffer;
export function setup(ccs: Buffer, seed: Buffer, device?: number): { pk: Buffer; vk: Buffer };

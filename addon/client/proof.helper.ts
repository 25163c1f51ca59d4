// Drop-in for the reference's client/proof.helper.ts: same exported names, argument shapes and result
// ({ proof: Buffer, publicWitness: Buffer }), but the `sunspot prove` child process (reference line 64) is
// replaced by the in-process g16b200 prover.  `nargo execute` (reference line 55) stays upstream: it is the
// Noir toolchain's witness generator, not part of the proving step.
//
// UNTESTED in this repository (no node toolchain in the build image); kept as source for maintainers.
import { execSync } from "child_process";
import * as fs from "fs";
import * as path from "path";
import * as g16 from "g16b200-napi";

export interface ShieldedPoolInputs {
  root: string; nullifier: string; recipient: string; amount: number | string; wa_commitment: string;
  secret_key: string; owner_x: string; owner_y: string; randomness: string; index: number | string;
  siblings: string[];
}
export interface CircuitConfig { circuitDir: string; circuitName: string }

// keys Noir expects as bare numbers in Prover.toml; everything else is a quoted field element
const NUMERIC_KEYS = new Set(["amount", "index"]);

function proverToml(inputs: Record<string, unknown>): string {
  const line = (key: string, value: unknown): string => {
    if (Array.isArray(value)) return `${key} = [\n${value.map((v) => `  "${v}",\n`).join("")}]\n`;
    return NUMERIC_KEYS.has(key) ? `${key} = ${value}\n` : `${key} = "${value}"\n`;
  };
  return Object.entries(inputs).map(([k, v]) => line(k, v)).join("");
}

// one resident circuit (constraint system + proving key on the GPU) per target directory
const resident = new Map<string, g16.Circuit>();

function circuitFor(config: CircuitConfig): g16.Circuit {
  const target = path.join(config.circuitDir, "target");
  const key = path.join(target, config.circuitName);
  let circuit = resident.get(key);
  if (!circuit) {
    circuit = g16.loadCircuit(fs.readFileSync(`${key}.ccs`), fs.readFileSync(`${key}.pk`));
    resident.set(key, circuit);
  }
  return circuit;
}

function runNargo(config: CircuitConfig, inputs: ShieldedPoolInputs): Buffer {
  fs.writeFileSync(path.join(config.circuitDir, "Prover.toml"), proverToml(inputs as unknown as Record<string, unknown>));
  if (process.env.G16_NO_NARGO) {
    // withdraw circuit: its constraints determine every intermediate witness, so the library rebuilds the witness file
    // from Prover.toml itself (g16_execute) -- no Noir toolchain on the proving host
    const base = path.join(config.circuitDir, "target", config.circuitName);
    return g16.execute(fs.readFileSync(`${base}.ccs`), fs.readFileSync(`${base}.json`),
                       fs.readFileSync(path.join(config.circuitDir, "Prover.toml")));
  }
  execSync("nargo execute", { cwd: config.circuitDir });
  return fs.readFileSync(path.join(config.circuitDir, "target", `${config.circuitName}.gz`));
}

/** Synchronous, like the reference's generateProof. */
export function generateProof(config: CircuitConfig, inputs: ShieldedPoolInputs) {
  const witness = runNargo(config, inputs);
  const { proof, publicWitness } = g16.proveSync(circuitFor(config), witness);
  // callers such as client/generate-proof-hex.ts:18-27 read the files back: keep writing them
  const base = path.join(config.circuitDir, "target", config.circuitName);
  fs.writeFileSync(`${base}.proof`, proof);
  fs.writeFileSync(`${base}.pw`, publicWitness);
  return { proof, publicWitness };
}

/** Same, without blocking the event loop while the GPU works. */
export async function generateProofAsync(config: CircuitConfig, inputs: ShieldedPoolInputs) {
  const witness = runNargo(config, inputs);
  return g16.prove(circuitFor(config), witness);
}

/** `sunspot verify` stand-in (noir_circuit/prove_linux.sh:87). */
export function verifyProof(config: CircuitConfig, proof: Buffer, publicWitness: Buffer): boolean {
  const vk = fs.readFileSync(path.join(config.circuitDir, "target", `${config.circuitName}.vk`));
  return g16.verify(vk, proof, publicWitness);
}

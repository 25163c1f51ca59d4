"""Generates tests/golden/prove_small.json with the big-integer oracle (oracle/py/groth16.py).

    python tests/gen_golden_prove.py

A 300-constraint synthetic gnark-format circuit (one BSB22 commitment, all standard solver hints),
the oracle's trusted setup with known toxic waste, and for three assignments the proof / public-
witness bytes the oracle produces under fixed (r, s, blinder).  The CUDA prover must reproduce
those bytes exactly (tests/test_prove_gpu.py); the C++ solver must reproduce the wires
(tests/test_host_solver.py).
"""
import hashlib, json, os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'py'))
import importlib.util
spec = importlib.util.spec_from_file_location('synth', os.path.join(ROOT, 'shielded_pool_pinocchio_solana_b200', 'synth.py'))
synth = importlib.util.module_from_spec(spec); spec.loader.exec_module(synth)
import bn254 as B, ccs, groth16 as G, serialize as S


def rnd_for(i):
    """(r, s, blinder) = SHA-256("g16b200/rnd/" || i || k) mod r  """
    return [int.from_bytes(hashlib.sha256(b"g16b200/rnd/%d/%d" % (i, k)).digest(), 'big') % B.R for k in range(3)]


def main():
    sc = synth.build(300, n_public=2, n_secret=16, n_committed=12, seed=7)
    c = ccs.parse_ccs(sc.ccs)
    pk, vk, tx = G.setup(c, b'golden-small')
    out = {'ccs': sc.ccs.hex(), 'pk': G.write_pk(pk).hex(), 'vk': G.write_vk(vk).hex(),
           'nb_wires': c.nb_wires, 'n_committed': len(c.commitments[0]['PrivateCommitted']), 'cases': []}
    enc = lambda v: b''.join(S.fr_to_bytes(x) for x in v)
    for i in range(3):
        asg = sc.assignment(100 + i)
        r, s, bl = rnd_for(i)
        proof, pw, aux = G.prove(c, pk, asg, r, s, bl)
        assert G.verify(vk, proof, pw), 'oracle proof does not verify'
        assert G.closed_form_check(c, tx, aux['wires'], proof, r, s)
        w = aux['wires']
        out['cases'].append({'assignment': enc(asg).hex(), 'rnd': enc([r, s, bl]).hex(), 'proof': proof.hex(),
                             'pw': pw.hex(), 'wires': enc(w).hex(),
                             'challenge': S.fr_to_bytes(w[c.commitments[0]['CommitmentIndex']]).hex(),
                             'h': enc(aux['h']).hex()})
    path = os.path.join(ROOT, 'tests', 'golden', 'prove_small.json')
    json.dump(out, open(path, 'w'))
    print('wrote', path, os.path.getsize(path))

if __name__ == '__main__':
    main()

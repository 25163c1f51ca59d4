"""Generates tests/golden/msm_small.json with the big-integer oracle (oracle/py/bn254.py).

    python tests/gen_golden_msm.py

Deterministic (seeded).  The file holds gnark-raw big-endian bytes as hex so the same vectors
drive the C oracle test (CPU) and the CUDA parity test (GPU box, where /root/reference and slow
Python loops are not wanted).
"""
import json, os, random, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'oracle', 'py'))
import bn254 as B
import serialize as S

def scalar_mix(rng, n):
    """witness-like mix + adversarial values (SURVEY.md 7 'Skewed scalars', 8d config 3)"""
    special = [0, 1, 2, B.R - 1, B.R - 2, (1 << 253), (1 << 254) - 1 if (1 << 254) - 1 < B.R else B.R - 3,
               0xffff, 0x10000, 0x8000, 0x7fff, (1 << 128) - 1]
    out = []
    for i in range(n):
        t = rng.random()
        if i < len(special): out.append(special[i] % B.R)
        elif t < 0.2: out.append(0)
        elif t < 0.35: out.append(1)
        elif t < 0.5: out.append(rng.randrange(256))
        else: out.append(rng.randrange(B.R))
    return out

def main():
    rng = random.Random(0xB200)
    fb1, fb2 = B.g1_fixed_base(), B.g2_fixed_base()
    cases = {}
    # G1: 300 bases incl. two infinities and a repeated point (forces the P == acc doubling path)
    n1, batch1 = 300, 3
    pts = [fb1.mul(rng.randrange(1, B.R)) for _ in range(n1)]
    pts[7] = None; pts[123] = None
    pts[50] = pts[49]; pts[51] = B.g1_neg(pts[49])
    sc = [scalar_mix(rng, n1) for _ in range(batch1)]
    sc[1] = [5] * n1                      # all-equal scalars: every point in one bucket
    sc[2][49] = sc[2][50] = sc[2][51] = 77  # P + P + (-P)
    res = [B.g1_msm(pts, s) for s in sc]
    cases['g1'] = dict(n=n1, batch=batch1, points=b''.join(S.g1_to_bytes(p) for p in pts).hex(),
                       scalars=b''.join(S.fr_to_bytes(x) for s in sc for x in s).hex(),
                       results=b''.join(S.g1_to_bytes(p) for p in res).hex())
    n2, batch2 = 96, 2
    pts = [fb2.mul(rng.randrange(1, B.R)) for _ in range(n2)]
    pts[5] = None; pts[11] = pts[10]
    sc = [scalar_mix(rng, n2) for _ in range(batch2)]
    sc[1][10] = sc[1][11] = 9
    res = [B.g2_msm(pts, s) for s in sc]
    cases['g2'] = dict(n=n2, batch=batch2, points=b''.join(S.g2_to_bytes(p) for p in pts).hex(),
                       scalars=b''.join(S.fr_to_bytes(x) for s in sc for x in s).hex(),
                       results=b''.join(S.g2_to_bytes(p) for p in res).hex())
    out = os.path.join(os.path.dirname(__file__), 'golden', 'msm_small.json')
    with open(out, 'w') as f:
        json.dump(cases, f)
    print('wrote', out)

if __name__ == '__main__':
    main()

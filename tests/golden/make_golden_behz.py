"""Generates tests/golden/behz_golden.json from the pure-Python big-int model (oracle/pymodel.py).

PARITY UNPINNED (SEAL absent): these vectors pin the BEHZ multiply / relinearize restatement
(SURVEY.md Appendix A.6-A.7).  They are produced by the big-int model, NOT by the C oracle or the
CUDA path, so both are checked against an implementation they share no code with.
Run:  python tests/golden/make_golden_behz.py
"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from pymodel import Behz, Model  # noqa: E402

P55 = [0x7fffffff380001, 0x7ffffffef00001, 0x7ffffffeac0001, 0x7ffffffe700001]
CASES = [
    dict(n=16, q=P55[:2], t=1 << 14, sizes=(2, 2), seed=1),
    dict(n=16, q=P55[:2], t=1 << 14, sizes=(3, 2), seed=2),
    dict(n=32, q=P55[:4], t=1 << 14, sizes=(2, 2), seed=3),
    dict(n=16, q=P55[:1], t=65537, sizes=(4, 3), seed=4),
    dict(n=16, q=[0xffffffffffc0001, 0xfffffffff840001], t=(1 << 59) + 12345, sizes=(2, 2), seed=5),  # 60-bit primes: |B| = k + 1
]
out = []
for c in CASES:
    m = Model(c["n"], c["q"], c["t"])
    bz = Behz(m)
    rng = random.Random(c["seed"])
    a = [[rng.randrange(m.q) for _ in range(m.n)] for _ in range(c["sizes"][0])]
    b = [[rng.randrange(m.q) for _ in range(m.n)] for _ in range(c["sizes"][1])]
    prod = bz.multiply(a, b)
    sq = bz.multiply(a, a)
    out.append(dict(n=c["n"], q=c["q"], t=c["t"], aux=len(bz.B), a=m.to_rns(a), b=m.to_rns(b),
                    sizes=list(c["sizes"]), product=m.to_rns(prod), square=m.to_rns(sq)))
    print(c["n"], len(c["q"]), len(bz.B))
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "behz_golden.json"), "w") as f:
    json.dump({"generator": "tests/golden/make_golden_behz.py", "cases": out}, f)

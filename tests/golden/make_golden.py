"""Generates tests/golden/oracle_golden.json from the CPU oracle.

PARITY UNPINNED: the reference (SEAL absent, print-only tests) offers no ciphertext-level golden
vectors, so these fixtures pin the ORACLE's outputs (checksums of encrypted_dct [+quantize_fhe]
on seeded uniform inputs) so later rounds cannot drift silently.  The GPU parity tests compare
against the live oracle AND against these committed checksums.
Run:  python tests/golden/make_golden.py
"""
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)
import oracle_py  # noqa: E402

pkg = importlib.import_module("fully-homomorphic-image-processing_b200")

CASES = [
    dict(n=4096, nblocks=2, quant=True, seed=0x5EA1B200, first_ct=0),
    dict(n=4096, nblocks=1, quant=False, seed=0x5EA1B200, first_ct=640),
    dict(n=2048, nblocks=2, quant=True, seed=0x1234, first_ct=0),
    dict(n=8192, nblocks=1, quant=True, seed=0x5EA1B200, first_ct=0),
    dict(n=1024, nblocks=3, quant=True, seed=99, first_ct=7, q=[0x7fffffff380001, 0x7ffffffef00001, 0x7ffffffeac0001]),
]

out = []
for c in CASES:
    q = c.get("q") or oracle_py.default_coeff_modulus(c["n"])
    o = oracle_py.Oracle(c["n"], q, 1 << 14)
    d = o.fill_uniform(c["first_ct"], 64 * c["nblocks"], 2, c["seed"])
    rec = dict(n=c["n"], q=q, t=1 << 14, nblocks=c["nblocks"], quant=c["quant"], seed=c["seed"],
               first_ct=c["first_ct"], input_checksum=hex(o.checksum(d)))
    W = 64 * o.ct_words()
    for b in range(c["nblocks"]):
        blk = d[b * W:(b + 1) * W]
        o.encrypted_dct(blk)
        if c["quant"]:
            o.quantize_fhe(blk, pkg.YQT)
    rec["output_checksum"] = hex(o.checksum(d))
    rec["first_words"] = [int(x) for x in d[:4]]
    out.append(rec)
    print(rec["n"], rec["output_checksum"])

with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_golden.json"), "w") as f:
    json.dump({"generator": "tests/golden/make_golden.py", "cases": out}, f, indent=1)

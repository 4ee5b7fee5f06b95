#!/usr/bin/env python
"""Freezes known-answer vectors of the oracle's Halton machinery into tests/golden/halton_kat.json (run once; the test
tests/test_halton_cpu.py::test_halton_known_answer_vectors compares the oracle against the frozen values from then on).
The reference ships no vectors for this path and its low-discrepancy code is device code: these pin the oracle against itself."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import front  # noqa: E402

L = front.lib()
f32 = np.float32
out = {"radical_inverse": [], "primes_tail": [L.orc_prime(i) for i in (0, 1, 2, 10, 100, 500, 999)]}
for base_index in (0, 1, 2, 7, 31, 200, 999):
    for a, h in ((0, 0), (1, 1), (12345, 0x9E3779B97F4A7C15), (2 ** 40 + 17, 42), (987654321987, 0xDEADBEEF)):
        v = f32(L.orc_owen_scrambled_radical_inverse(base_index, a, h))
        out["radical_inverse"].append({"base_index": base_index, "a": str(a), "hash": str(h), "bits": int(np.array([v], f32).view(np.uint32)[0])})
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "halton_kat.json"), "w"), indent=1)
print("wrote", len(out["radical_inverse"]), "vectors")

#!/usr/bin/env python
"""Generate tests/golden/*.npz from the CPU oracle (oracle/), run in the build container.

The reference ships no golden vectors or tests.  tools/make_ref_golden.py pins the oracle to the
reference's own compiled code (tests/golden/ref_pin.npz); the fixtures written here freeze the ORACLE's outputs: tests/test_golden.py checks (CPU) that the oracle still reproduces them
bit for bit and (GPU) that the CUDA path reproduces them.  Regenerate only when the oracle's definition changes:
    python tools/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import oracle_lib as O  # noqa: E402
import golden_cases as G  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    os.makedirs(OUT, exist_ok=True)
    O.build()
    for name, fn in G.CASES.items():
        data = fn(O)
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **data)
        print(f"{path}: {os.path.getsize(path)} bytes, keys {sorted(data)}")


if __name__ == "__main__":
    main()

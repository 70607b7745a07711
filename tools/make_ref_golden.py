#!/usr/bin/env python
"""Generate tests/golden/ref_pin.npz from the REFERENCE'S OWN compiled code (oracle/_ref/libcrt_ref.so).

Runs only where /root/reference exists (the build container): `make -C oracle ref` compiles the reference's headers and
pbrv4 sources unmodified (oracle/ref_harness.cpp, oracle/refshim/), tests/ref_pin_cases.py evaluates the fixed cases with
it, and the outputs are committed so that the oracle stays pinned to the reference where the reference cannot travel.
    python tools/make_ref_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import ref_lib as R  # noqa: E402
import ref_pin_cases as P  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "ref_pin.npz")


def main():
    R.build(force=True)
    data = P.run("ref")
    np.savez_compressed(OUT, **data)
    print(f"{OUT}: {os.path.getsize(OUT)} bytes, {len(data)} arrays; source: {R.lib().ref_describe().decode()}")


if __name__ == "__main__":
    main()

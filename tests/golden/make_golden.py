"""Generates the committed golden fixtures (run from the repo root, CPU only).

The reference ships no golden vectors and its dependencies
(faiss, openai/CLIP) are not installable offline, so the fixtures pin the oracle
restatements against each other / against independent implementations:

  flatip_golden.npz : seeded 4096x512 database (+ duplicate rows for ties), 6 queries;
                      expected (D, I) for k in {1, 21, 100} from oracle/flatip_ref.py, after
                      asserting that the independently written C heap restatement
                      (oracle/flatip_ref.c) returns the same ids.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cli-p_b200"))

from oracle import flatip_ref as F  # noqa: E402
from clipb200 import synth  # noqa: E402


def flatip_inputs():
    xb = synth.unit_rows(4096, seed=11, clip_like=True)
    xb[100:108] = xb[7]          # exact duplicates -> exact score ties
    xb[4000] = xb[3]
    xq = synth.unit_rows(6, seed=12, clip_like=True)
    xq[5] = xb[7]                # a query identical to a stored row (image-similarity query)
    return xb.astype(np.float16), xq


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_build", "liboracle_flatip.so"))
    xb, xq = flatip_inputs()
    out = {}
    for k in (1, 21, 100):
        D, I = F.search(xq, xb, k)
        D2 = np.empty_like(D)
        I2 = np.empty_like(I)
        rc = lib.oracle_flatip_search(C.c_void_p(xb.ctypes.data), 1, C.c_int64(xb.shape[0]), 512,
                                      C.c_void_p(xq.ctypes.data), C.c_int64(xq.shape[0]), C.c_int64(k),
                                      C.c_void_p(D2.ctypes.data), C.c_void_p(I2.ctypes.data), 1)
        assert rc == 0
        ok, exempt, msg = F.ids_match_with_tolerance(D, I, D2, I2)
        assert ok, msg
        assert np.abs(D - D2).max() < 1e-6
        out[f"D{k}"] = D
        out[f"I{k}"] = I
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "flatip_golden.npz"), **out)
    print("wrote flatip_golden.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()

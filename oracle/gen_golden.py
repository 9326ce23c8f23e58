"""Freeze golden vectors for the hot path into tests/golden/ (run from the repo root).

    python oracle/gen_golden.py

For every case in tests/cases.py the C oracle (fennec_oracle.c) and the independent NumPy
restatement (np_restatement.py) are both run and REQUIRED to agree — scores to <= 1e-12, pixel
buffers bit for bit — before the value is frozen.  golden.json records, per case, the SHA-256 of
the inputs (so a drifting generator is caught), and either the float64 score (with its hex
representation) or the SHA-256 + shape of the output pixels; small outputs are also stored raw in
golden_pixels.npz so a test can diff them without any oracle.

PARITY UNPINNED: no Go toolchain exists here, so these vectors are restatement-derived, not
reference-derived (see oracle/fennec_oracle.h).
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import np_restatement as N  # noqa: E402
from oracle import pyoracle as O  # noqa: E402
from tests import cases  # noqa: E402

OUT_DIR = os.path.join(ROOT, "tests", "golden")
RAW_LIMIT = 64 * 1024  # outputs up to this many bytes are stored raw


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main() -> None:
    O.set_procs(8)
    os.makedirs(OUT_DIR, exist_ok=True)
    golden = {"procs": 8, "scores": {}, "pixels": {}}
    raw = {}
    for name, (op, build) in cases.SCORE_CASES.items():
        a, b = build()
        c = getattr(O, op)(a, b)
        n = getattr(N, op)(a, b, 8)
        if not abs(c - n) <= 1e-12:
            raise SystemExit(f"{name}: C oracle {c!r} != NumPy restatement {n!r}")
        golden["scores"][name] = {
            "op": op, "value": c, "hex": float(c).hex(), "np_minus_c": n - c,
            "inputs_sha256": [sha(a), sha(b)], "shape": list(a.shape[:2]),
        }
        print(f"{name:32s} {op:10s} {c!r}")
    for name, (op, build, kw) in cases.PIXEL_CASES.items():
        src = build()
        c = getattr(O, op)(src, *kw.values())
        n = getattr(N, op)(src, *kw.values())
        if c.shape != n.shape or not np.array_equal(c, n):
            raise SystemExit(f"{name}: C oracle and NumPy restatement differ "
                             f"({c.shape} vs {n.shape}, {(c != n).sum() if c.shape == n.shape else '?'} bytes)")
        golden["pixels"][name] = {
            "op": op, "kwargs": kw, "sha256": sha(c), "shape": list(c.shape),
            "input_sha256": sha(src), "raw": c.nbytes <= RAW_LIMIT,
        }
        if c.nbytes <= RAW_LIMIT:
            raw[name] = c
        print(f"{name:32s} {op:16s} {c.shape} {sha(c)[:16]}")
    # weight tables (host side of the ABI)
    k8 = O.gaussian_kernel(8, 1.5)
    assert np.array_equal(k8, N.gaussian_kernel(8, 1.5))
    golden["tables"] = {
        "ssim_kernel_8x8_sha256": sha(k8), "ssim_kernel_center": float(k8[36]), "ssim_kernel_corner": float(k8[0]),
    }
    for dst, src_n in ((1920, 7680), (100, 400), (333, 120)):
        st, ix, wt = O.lanczos_weights(dst, src_n)
        tab = N.lanczos_weights(dst, src_n)
        assert [len(t[0]) for t in tab] == list(np.diff(st))
        assert np.array_equal(np.concatenate([np.array(t[1]) for t in tab]), wt)
        golden["tables"][f"lanczos_{src_n}_to_{dst}"] = {"entries": int(len(wt)), "weights_sha256": sha(wt),
                                                         "index_sha256": sha(ix.astype(np.int32))}
    with open(os.path.join(OUT_DIR, "golden.json"), "w") as f:
        json.dump(golden, f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(OUT_DIR, "golden_pixels.npz"), **raw)
    print(f"wrote {len(golden['scores'])} scores, {len(golden['pixels'])} pixel cases "
          f"({len(raw)} raw) to {OUT_DIR}")


if __name__ == "__main__":
    main()

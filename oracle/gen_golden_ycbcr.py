"""Freeze golden vectors for SURVEY §8(f1) — convertToNRGBA on decoded YCbCr / Gray images — into
tests/golden/ycbcr_golden.json (run from the repo root):

    python oracle/gen_golden_ycbcr.py

As in gen_golden.py the C oracle and the independent NumPy restatement must agree bit for bit before a
hash is frozen.  The arithmetic is Go's standard library (image/color/ycbcr.go, Go 1.25.5 — not under
/root/reference), restated from its published source: PARITY UNPINNED against a running Go toolchain; pinned
instead by the cross-restatement agreement over ALL 2^24 (Y, Cb, Cr) triples, by agreement within 1 with the
JFIF floating-point definition, and by the round-trip triples Go's own RGBToYCbCr documentation implies
(pure red/green/blue -> (254,0,0), (0,255,1), (0,0,254)).
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from fennec_b200 import synth as S  # noqa: E402
from oracle import np_restatement as N  # noqa: E402
from oracle import pyoracle as O  # noqa: E402
from tests import cases_ycbcr as CY  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    out = {"cases": {}, "exhaustive_444_sha256": None, "known": {}}
    for name, (build, ratio) in CY.CASES.items():
        y, cb, cr = build()
        c = O.ycbcr_to_nrgba(y, cb, cr, ratio)
        n = N.ycbcr_to_nrgba(y, cb, cr, ratio)
        if not np.array_equal(c, n):
            raise SystemExit(f"{name}: C oracle and NumPy restatement differ")
        out["cases"][name] = {"ratio": ratio, "shape": list(c.shape), "sha256": sha(c), "inputs_sha256": [sha(y), sha(cb), sha(cr)]}
        print(f"{name:28s} ratio {ratio} {c.shape} {sha(c)[:16]}")
    y, cb, cr = CY.exhaustive_planes()
    c = O.ycbcr_to_nrgba(y, cb, cr, 0)
    assert np.array_equal(c, N.ycbcr_to_nrgba(y, cb, cr, 0))
    yf, bf, rf = y.astype(np.float64), cb.astype(np.float64) - 128, cr.astype(np.float64) - 128
    f = lambda v: np.clip(np.round(v), 0, 255)  # noqa: E731
    ref = np.stack([f(yf + 1.402 * rf), f(yf - 0.34414 * bf - 0.71414 * rf), f(yf + 1.772 * bf)], -1)
    assert np.abs(c[..., :3].astype(int) - ref).max() <= 1
    out["exhaustive_444_sha256"] = sha(c)
    for k, (Y, B, R) in CY.KNOWN.items():
        px = O.ycbcr_to_nrgba(np.array([[Y]], np.uint8), np.array([[B]], np.uint8), np.array([[R]], np.uint8), 0)[0, 0]
        out["known"][k] = [int(v) for v in px]
    g = S.noise_image(37, 21, 5)[..., 0].copy()
    assert np.array_equal(O.gray_to_nrgba(g), N.gray_to_nrgba(g))
    out["gray_37x21_sha256"] = sha(O.gray_to_nrgba(g))
    with open(os.path.join(ROOT, "tests", "golden", "ycbcr_golden.json"), "w") as fp:
        json.dump(out, fp, indent=1, sort_keys=True)
    print("wrote tests/golden/ycbcr_golden.json")


if __name__ == "__main__":
    main()

"""Checks a subset of the golden cases through whichever kernel variants the FB_* env vars select."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fennec_b200 import api
from tests import cases
gold = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "golden.json")))
pix = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "golden_pixels.npz"))
API = {"ssim": api.SSIM, "ssim_fast": api.SSIMFast, "msssim": api.MSSSIM, "box_downsample": api.box_downsample,
       "gaussian_blur": api.GaussianBlur, "blur3x3": api.blur3x3, "sharpen": api.Sharpen,
       "adaptive_sharpen": api.AdaptiveSharpen, "lanczos_resize": api.lanczos_resize}
worst = 0.0
for name, (op, build) in cases.SCORE_CASES.items():
    if "1920" in name or "2016" in name:
        continue
    a, b = build()
    worst = max(worst, abs(API[op](a, b) - gold["scores"][name]["value"]))
assert worst <= 1e-5, worst
import hashlib
for name, (op, build, kw) in cases.PIXEL_CASES.items():
    out = API[op](build(), *kw.values())
    assert hashlib.sha256(np.ascontiguousarray(out).tobytes()).hexdigest() == gold["pixels"][name]["sha256"], name
print(f"variant ok worst_score_err={worst:.2e}")

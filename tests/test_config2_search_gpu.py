"""BASELINE.json config 2 as a workload (VERDICT r1 "missing 2"): the SSIM-guided quality search of compress.go:21-88 on a
4032x3024 image at Balanced, SSIM step on the GPU — must walk the same bisection path and choose the same quality as the
search scored by the CPU oracle."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.mark.timeout(600)
@pytest.mark.parametrize("w,h,target", [(4032, 3024, 0.94), (1600, 1200, 0.97)])
def test_quality_search_picks_the_same_q_as_the_oracle(w, h, target, lib, oracle):
    pytest.importorskip("PIL")
    import config2_search as C2
    from fennec_b200 import api, synth
    src = synth.gradient_noise_image(w, h, 5)
    cache = {}
    with api.SSIMReference(src) as ref:
        q, s, n, tr = C2.quality_search(src, target, lambda y, cb, cr: ref.score_ycbcr(y, cb, cr, 0), cache)
        # the NRGBA upload path of the session scores the same candidates identically (same thumbnail bytes, same kernel)
        y, cb, cr = cache[q][1]
        assert abs(ref.score_nrgba(api.ycbcr_to_nrgba(y, cb, cr, 0)) - s) <= 2e-7
    q2, s2, n2, tr2 = C2.quality_search(src, target, lambda y, cb, cr: oracle.ssim_fast(src, oracle.ycbcr_to_nrgba(y, cb, cr, 0)), cache)
    assert [m for m, _ in tr] == [m for m, _ in tr2], (tr, tr2)           # same bisection path
    assert q == q2 and n == n2                                            # "same chosen Q"
    assert max(abs(a[1] - b[1]) for a, b in zip(tr, tr2)) <= 1e-5
    assert s >= min(target, 0.999) and 30 <= q <= 100

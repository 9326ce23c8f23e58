"""convertToNRGBA (convert.go:34-64) on decoded YCbCr / Gray images, and the reference-image
session of the quality search (compress.go:45-74).

CPU part (not gpu): the C oracle against the golden vectors, the independent NumPy restatement, known answers and
the JFIF definition.  GPU part: the CUDA path through the C ABI, bit-exact, including all 2^24 colour triples.
"""
import hashlib
import json
import os

import numpy as np
import pytest

from fennec_b200 import synth as S
from tests import cases_ycbcr as CY

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ycbcr_golden.json")))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# ---- CPU: the oracle is pinned ----------------------------------------------------------------------------

@pytest.mark.parametrize("name", sorted(CY.CASES))
def test_oracle_matches_golden_and_numpy(name, oracle):
    from oracle import np_restatement as N
    build, ratio = CY.CASES[name]
    y, cb, cr = build()
    g = GOLD["cases"][name]
    assert [sha(y), sha(cb), sha(cr)] == g["inputs_sha256"], "input generator drifted"
    out = oracle.ycbcr_to_nrgba(y, cb, cr, ratio)
    assert list(out.shape) == g["shape"] and sha(out) == g["sha256"]
    assert np.array_equal(out, N.ycbcr_to_nrgba(y, cb, cr, ratio))
    assert np.all(out[..., 3] == 255)


def test_oracle_known_answers(oracle):
    for k, (Y, B, R) in CY.KNOWN.items():
        px = oracle.ycbcr_to_nrgba(np.array([[Y]], np.uint8), np.array([[B]], np.uint8), np.array([[R]], np.uint8), 0)[0, 0]
        assert tuple(int(v) for v in px[:3]) == CY.KNOWN_RGB[k] and px[3] == 255
        assert [int(v) for v in px] == GOLD["known"][k]


def test_oracle_all_triples_within_one_of_jfif(oracle):
    y, cb, cr = CY.exhaustive_planes()
    out = oracle.ycbcr_to_nrgba(y, cb, cr, 0)
    assert sha(out) == GOLD["exhaustive_444_sha256"]
    yf, bf, rf = y.astype(np.float64), cb.astype(np.float64) - 128, cr.astype(np.float64) - 128
    f = lambda v: np.clip(np.round(v), 0, 255)  # noqa: E731
    ref = np.stack([f(yf + 1.402 * rf), f(yf - 0.34414 * bf - 0.71414 * rf), f(yf + 1.772 * bf)], -1)
    assert np.abs(out[..., :3].astype(int) - ref).max() <= 1


def test_oracle_chroma_addressing_is_go_coffset(oracle):
    # every pixel of a cell shares the cell's chroma sample: convert with a 1-sample-per-cell plane and with the
    # plane expanded to 4:4:4 by hand
    sub = {0: (1, 1), 1: (2, 1), 2: (2, 2), 3: (1, 2), 4: (4, 1), 5: (4, 2)}
    for ratio, (dx, dy) in sub.items():
        y, cb, cr = S.noise_planes(19, 11, ratio, 30 + ratio)
        up = lambda p: np.ascontiguousarray(np.repeat(np.repeat(p, dy, 0), dx, 1)[:11, :19])  # noqa: E731
        assert np.array_equal(oracle.ycbcr_to_nrgba(y, cb, cr, ratio), oracle.ycbcr_to_nrgba(y, up(cb), up(cr), 0))


def test_oracle_gray(oracle):
    g = S.noise_image(37, 21, 5)[..., 0].copy()
    out = oracle.gray_to_nrgba(g)
    assert sha(out) == GOLD["gray_37x21_sha256"]
    assert np.array_equal(out[..., 0], g) and np.array_equal(out[..., 1], g) and np.array_equal(out[..., 2], g)


# ---- GPU: the CUDA path through the C ABI -------------------------------------------------------------------

@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CY.CASES))
def test_gpu_ycbcr_bit_exact(name, lib, oracle):
    from fennec_b200 import api
    build, ratio = CY.CASES[name]
    y, cb, cr = build()
    out = api.ycbcr_to_nrgba(y, cb, cr, ratio)
    assert sha(out) == GOLD["cases"][name]["sha256"]
    assert np.array_equal(out, oracle.ycbcr_to_nrgba(y, cb, cr, ratio))


@pytest.mark.gpu
def test_gpu_ycbcr_all_triples(lib):
    from fennec_b200 import api
    y, cb, cr = CY.exhaustive_planes()
    assert sha(api.ycbcr_to_nrgba(y, cb, cr, 0)) == GOLD["exhaustive_444_sha256"]


@pytest.mark.gpu
def test_gpu_ycbcr_strided_planes_and_batch(lib, oracle):
    import torch
    from fennec_b200 import api, batch
    # planes that are views into wider buffers (image.YCbCr.YStride / CStride > width), odd width → scalar tail
    yb, cbb, crb = S.noise_planes(90, 41, 2, 77)
    y, cb, cr = yb[:, 3:80], cbb[:, 1:40], crb[:, 1:40]   # 77 wide; chroma 39 = ceil(77/2)
    assert np.array_equal(api.ycbcr_to_nrgba(y, cb, cr, 2), oracle.ycbcr_to_nrgba(y, cb, cr, 2))
    planes = [S.noise_planes(128, 64, 2, 80 + i) for i in range(3)]
    ty = torch.from_numpy(np.stack([p[0] for p in planes])).cuda()
    tcb = torch.from_numpy(np.stack([p[1] for p in planes])).cuda()
    tcr = torch.from_numpy(np.stack([p[2] for p in planes])).cuda()
    got = batch.ycbcr_to_nrgba_batch(ty, tcb, tcr, 2).cpu().numpy()
    for i, p in enumerate(planes):
        assert np.array_equal(got[i], oracle.ycbcr_to_nrgba(*p, 2))


@pytest.mark.gpu
def test_gpu_gray(lib, oracle):
    from fennec_b200 import api
    for w, h in ((37, 21), (64, 8), (1, 1)):
        g = S.noise_image(w, h, w)[..., 0].copy()
        assert np.array_equal(api.gray_to_nrgba(g), oracle.gray_to_nrgba(g))


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,ratio", [(1300, 700, 2), (640, 480, 2), (300, 200, 0), (4032, 3024, 2), (7, 5, 2)])
def test_gpu_reference_session_scores(w, h, ratio, lib, oracle):
    """compress.go:45-74 with the source cached on the device: SSIMFast(src, convertToNRGBA(decoded)) per iteration."""
    from fennec_b200 import api
    src = S.gradient_noise_image(w, h, w + h)
    with api.SSIMReference(src) as ref:
        for it in range(3):   # three "search iterations" with different candidates
            y, cb, cr = S.ycbcr_planes_from_nrgba(src, ratio, seed=it, amp=2 + 3 * it)
            cand = oracle.ycbcr_to_nrgba(y, cb, cr, ratio)
            got = ref.score_ycbcr(y, cb, cr, ratio)
            assert abs(got - api.SSIMFast(src, cand)) <= 2e-7          # same kernels, same bytes
            assert abs(ref.score_nrgba(cand) - got) <= 2e-7
            if w * h <= 1300 * 700:
                assert abs(got - oracle.ssim_fast(src, cand)) <= 3e-6   # contract: 1e-5
        assert ref.score_nrgba(src) >= 0.999999

"""A compact walk over every kernel family with awkward shapes (ragged widths, unaligned strides, tiny images),
meant to run under `compute-sanitizer --tool memcheck`. Also checks results against the oracle."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fennec_b200 import api, synth as S
from oracle import pyoracle as O

def padded(img, pad):
    h, w = img.shape[:2]
    buf = np.full((h, w + pad, 4), 0xCD, dtype=np.uint8); buf[:, :w] = img
    return buf[:, :w]

for (w, h) in [(9, 9), (13, 40), (131, 77), (248, 64), (250, 33), (512, 20), (641, 19)]:
    a = S.noise_image(w, h, w + h, alpha="random"); b = S.perturb(a, 7, 9)
    for aa, bb in ((a, b), (padded(a, 3), padded(b, 5))):
        assert abs(api.SSIM(aa, bb) - O.ssim(a, b)) <= 1e-5
        assert abs(api.SSIMFast(aa, bb) - O.ssim_fast(a, b)) <= 1e-5
        assert abs(api.MSSSIM(aa, bb) - O.msssim(a, b)) <= 1e-5
    for src in (a, padded(a, 1)):
        assert np.array_equal(api.GaussianBlur(src, 2.0), O.gaussian_blur(a, 2.0))
        assert np.array_equal(api.GaussianBlur(src, 0.7), O.gaussian_blur(a, 0.7))
        assert np.array_equal(api.GaussianBlur(src, 4.0), O.gaussian_blur(a, 4.0))
        assert np.array_equal(api.Sharpen(src, 0.5), O.sharpen(a, 0.5))
        assert np.array_equal(api.Sharpen(src, 0.3), O.sharpen(a, 0.3))
        assert np.array_equal(api.AdaptiveSharpen(src, 0.6), O.adaptive_sharpen(a, 0.6))
        assert np.array_equal(api.blur3x3(src), O.blur3x3(a))
        for (dw, dh) in ((max(1, w // 3), max(1, h // 2)), (w + 5, h + 3), (1, 1)):
            assert np.array_equal(api.lanczos_resize(src, dw, dh), O.lanczos_resize(a, dw, dh))
            assert np.array_equal(api.box_downsample(src, dw, dh), O.box_downsample(a, dw, dh))
big = S.noise_image(2051, 517, 3, alpha="random")
assert np.array_equal(api.box_downsample(big, 512, 129), O.box_downsample(big, 512, 129))
assert np.array_equal(api.lanczos_resize(big, 513, 130), O.lanczos_resize(big, 513, 130))
assert np.array_equal(api.GaussianBlur(big, 2.0), O.gaussian_blur(big, 2.0))
assert abs(api.SSIMFast(big, S.perturb(big, 1, 5)) - O.ssim_fast(big, S.perturb(big, 1, 5))) <= 1e-5
print("sanitize subset ok")

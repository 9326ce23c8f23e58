"""Seeded inputs for SURVEY §8(f1) (convertToNRGBA on decoded YCbCr images): shared by the golden generator,
the CPU oracle tests and the GPU parity tests."""
import numpy as np

from fennec_b200 import synth as S

# name -> (builder of (y, cb, cr), image.YCbCrSubsampleRatio constant)
CASES = {
    "noise_444_64x48": (lambda: S.noise_planes(64, 48, 0, 1), 0),
    "noise_422_65x47": (lambda: S.noise_planes(65, 47, 1, 2), 1),
    "noise_420_67x45": (lambda: S.noise_planes(67, 45, 2, 3), 2),
    "noise_440_33x31": (lambda: S.noise_planes(33, 31, 3, 4), 3),
    "noise_411_70x20": (lambda: S.noise_planes(70, 20, 4, 5), 4),
    "noise_410_71x21": (lambda: S.noise_planes(71, 21, 5, 6), 5),
    "noise_420_1x1": (lambda: S.noise_planes(1, 1, 2, 7), 2),
    "noise_420_3x2": (lambda: S.noise_planes(3, 2, 2, 8), 2),
    "photo_420_640x480": (lambda: S.ycbcr_planes_from_nrgba(S.gradient_noise_image(640, 480, 9), 2, 10, 3), 2),
    "photo_444_320x200": (lambda: S.ycbcr_planes_from_nrgba(S.make_test_image(320, 200), 0), 0),
    "photo_420_1300x700": (lambda: S.ycbcr_planes_from_nrgba(S.gradient_noise_image(1300, 700, 11), 2, 12, 2), 2),
}

# (Y, Cb, Cr) triples whose conversion is known: white, black, and Go's RGBToYCbCr images of pure red / green / blue
KNOWN = {"white": (255, 128, 128), "black": (0, 128, 128), "red": (76, 85, 255), "green": (150, 44, 21), "blue": (29, 255, 107)}
KNOWN_RGB = {"white": (255, 255, 255), "black": (0, 0, 0), "red": (254, 0, 0), "green": (0, 255, 1), "blue": (0, 0, 254)}


def exhaustive_planes():
    """All 2^24 (Y, Cb, Cr) triples as three 4096x4096 planes (4:4:4)."""
    Y, B, R = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), indexing="ij")
    return (np.ascontiguousarray(Y.reshape(4096, 4096)), np.ascontiguousarray(B.reshape(4096, 4096)),
            np.ascontiguousarray(R.reshape(4096, 4096)))

"""convertToNRGBA (convert.go:34-64) on the decoder output types besides YCbCr / Gray: *image.RGBA, RGBA64, NRGBA64,
Gray16, CMYK, Paletted.  CPU: C oracle vs the independent NumPy restatement + hand-computed values from the Go
formulas.  GPU: bit-exact through the C ABI."""
import numpy as np
import pytest

from tests import cases_pixfmt as K

FMTS = sorted(K.NAMES)


@pytest.mark.parametrize("fmt", FMTS, ids=[K.NAMES[f] for f in FMTS])
@pytest.mark.parametrize("kind", ["valid", "wild"])
def test_oracle_matches_numpy(fmt, kind, oracle):
    from oracle import np_restatement as N
    for (w, h, seed) in ((1, 1, 1), (7, 5, 2), (64, 33, 3), (257, 19, 4)):
        pix, pal = K.make(fmt, w, h, seed, kind)
        a = oracle.convert_to_nrgba(fmt, pix, pal)
        b = N.convert_to_nrgba(fmt, pix, pal)
        assert np.array_equal(a, b)


def test_oracle_rgba_exhaustive_and_known_values(oracle):
    from oracle import np_restatement as N
    img = K.exhaustive_rgba()
    out = oracle.convert_to_nrgba(K.FMT_RGBA, img)
    assert np.array_equal(out, N.convert_to_nrgba(K.FMT_RGBA, img))
    # convert.go:42-60 by hand: a == 0 -> zeros; a == 255 -> the byte itself; R=100, A=200: r = 100*257, a = 200*257,
    # (r*0xffff)/a = 32767 -> >> 8 = 127; alpha 200
    assert out[0].max() == 0
    assert np.array_equal(out[255, :, 0], np.arange(256))
    assert out[200, 100].tolist() == [127, 127, 127, 200]
    # R = A (white premultiplied) un-premultiplies to 255 for every alpha >= 1
    assert all(out[a, a, 0] == 255 for a in range(1, 256))
    # c > a is not a valid premultiplied colour: Go's uint8() truncates, e.g. R=255, A=1: (65535*65535/257)>>8 = 65280 -> 0x00
    assert out[1, 255, 0] == (((255 * 257 * 0xFFFF) // 257) >> 8) & 0xFF


def test_oracle_other_formats_known_values(oracle):
    def be(*v):
        return np.array([[sum(([x >> 8, x & 0xFF] for x in v), [])]], np.uint8)
    # NRGBA64: R=0x8000, A=0x8000: r = 0x8000*0x8000/0xffff = 0x4000; (0x4000*0xffff)/0x8000 = 0x7fff -> 0x7f; alpha 0x80
    assert oracle.convert_to_nrgba(K.FMT_NRGBA64, be(0x8000, 0xFFFF, 0, 0x8000))[0, 0].tolist() == [0x7F, 0xFF, 0, 0x80]
    # RGBA64 opaque: high bytes
    assert oracle.convert_to_nrgba(K.FMT_RGBA64, be(0x1234, 0xABCD, 0x00FF, 0xFFFF))[0, 0].tolist() == [0x12, 0xAB, 0x00, 0xFF]
    # Gray16: high byte, opaque
    assert oracle.convert_to_nrgba(K.FMT_GRAY16, be(0xBEEF))[0, 0].tolist() == [0xBE, 0xBE, 0xBE, 0xFF]
    # CMYK: C=0, M=255, Y=128, K=64: w = 0xffff - 64*257 = 49087; r = 0xffff*w/0xffff = 49087 -> 0xBF; g = 0;
    # b = (0xffff - 128*257) * w / 0xffff = 32639*49087/65535 = 24447 -> 0x5F
    cm = np.array([[[0, 255, 128, 64]]], np.uint8)
    assert oracle.convert_to_nrgba(K.FMT_CMYK, cm)[0, 0].tolist() == [0xBF, 0x00, 0x5F, 0xFF]
    # Paletted: entry lookup, then the same rule; an index past the palette is Go's panic
    pal = np.array([[0xFFFF, 0, 0, 0xFFFF], [0x4000, 0x4000, 0x4000, 0x8000]], np.uint16)
    out = oracle.convert_to_nrgba(K.FMT_PALETTED, np.array([[0, 1]], np.uint8), pal)
    assert out[0].tolist() == [[255, 0, 0, 255], [0x7F, 0x7F, 0x7F, 0x80]]
    with pytest.raises(IndexError):
        oracle.convert_to_nrgba(K.FMT_PALETTED, np.array([[2]], np.uint8), pal)


def test_oracle_strided_rows(oracle):
    from oracle import np_restatement as N
    pix, _ = K.make(K.FMT_RGBA, 40, 9, 5)
    wide = np.zeros((9, 64, 4), np.uint8)
    wide[:, :40] = pix
    assert np.array_equal(oracle.convert_to_nrgba(K.FMT_RGBA, wide[:, :40]), N.convert_to_nrgba(K.FMT_RGBA, pix))


# ---- GPU: bit-exact through the C ABI --------------------------------------------------------------------------

@pytest.mark.gpu
@pytest.mark.parametrize("fmt", FMTS, ids=[K.NAMES[f] for f in FMTS])
@pytest.mark.parametrize("kind", ["valid", "wild"])
def test_gpu_convert_bit_exact(fmt, kind, lib, oracle):
    from fennec_b200 import api
    for (w, h, seed) in ((1, 1, 1), (7, 5, 2), (64, 33, 3), (257, 19, 4), (1023, 131, 6)):
        pix, pal = K.make(fmt, w, h, seed, kind)
        got = api.convert_to_nrgba(fmt, pix, pal)
        want = oracle.convert_to_nrgba(fmt, pix, pal)
        assert np.array_equal(got, want), (K.NAMES[fmt], w, h)


@pytest.mark.gpu
def test_gpu_convert_rgba_every_colour_alpha_pair(lib, oracle):
    from fennec_b200 import api
    img = K.exhaustive_rgba()
    assert np.array_equal(api.convert_to_nrgba(K.FMT_RGBA, img), oracle.convert_to_nrgba(K.FMT_RGBA, img))


@pytest.mark.gpu
def test_gpu_convert_strided_rows_and_unaligned_base(lib, oracle):
    """Go sub-images: the row stride exceeds the row, and Pix may start at any byte."""
    from fennec_b200 import api
    for fmt in FMTS:
        bpp = K.BPP[fmt]
        pix, pal = K.make(fmt, 93, 17, 8)
        if pix.ndim == 2:
            pix = pix[..., None]
        raw = np.zeros(17 * (93 * bpp + 13) + 3, np.uint8)
        view = np.lib.stride_tricks.as_strided(raw[3:], shape=(17, 93, bpp), strides=(93 * bpp + 13, bpp, 1))
        view[...] = pix
        got = api.convert_to_nrgba(fmt, view, pal)
        assert np.array_equal(got, oracle.convert_to_nrgba(fmt, np.ascontiguousarray(pix), pal)), K.NAMES[fmt]


@pytest.mark.gpu
def test_gpu_convert_paletted_index_out_of_range_is_an_error(lib):
    from fennec_b200 import api, _lib
    pal = np.array([[0xFFFF, 0, 0, 0xFFFF], [0, 0xFFFF, 0, 0xFFFF]], np.uint16)
    ok = api.convert_to_nrgba(K.FMT_PALETTED, np.array([[0, 1, 1, 0, 1]], np.uint8), pal)
    assert ok[0, :2].tolist() == [[255, 0, 0, 255], [0, 255, 0, 255]]
    with pytest.raises(_lib.FennecError):
        api.convert_to_nrgba(K.FMT_PALETTED, np.array([[0, 2]], np.uint8), pal)


@pytest.mark.gpu
def test_gpu_convert_batch_dev(lib, oracle):
    import torch
    from fennec_b200 import batch
    for fmt in FMTS:
        items = [K.make(fmt, 640, 96, 20 + 2 * i) for i in range(3)]            # even seeds: 256-entry palettes
        pix = torch.from_numpy(np.stack([p for p, _ in items])).cuda()
        pal_t, ncol = None, 0
        if fmt == K.FMT_PALETTED:
            pal_t = torch.from_numpy(np.stack([q for _, q in items]).view(np.int16)).cuda()
            ncol = 256
        out = batch.convert_to_nrgba_batch(fmt, pix, pal_t, ncol)
        for i, (p, q) in enumerate(items):
            assert np.array_equal(out[i].cpu().numpy(), oracle.convert_to_nrgba(fmt, p, q)), K.NAMES[fmt]

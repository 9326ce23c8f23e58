"""applyPalette + palettedToNRGBA (targetsize.go:479-545): nearest palette entry by squared RGB
distance, first minimum on ties.  CPU: C oracle (the reference's loop) vs NumPy argmin.  GPU: bit-exact indices
and reconstruction through the C ABI."""
import numpy as np
import pytest

from fennec_b200 import synth as S


def palettes():
    rng = np.random.Generator(np.random.PCG64(3))
    out = {}
    for n in (1, 2, 7, 16, 255, 256):
        p = rng.integers(0, 256, (n, 4), dtype=np.uint8)
        p[:, 3] = 255
        out[f"random_{n}"] = p
    dup = rng.integers(0, 256, (64, 4), dtype=np.uint8)
    dup[:, 3] = 255
    out["duplicates_128"] = np.concatenate([dup, dup[::-1]])            # every entry twice: ties must pick the first
    g = np.arange(0, 256, 17, dtype=np.uint8)
    out["gray_ramp_16"] = np.stack([g, g, g, np.full_like(g, 255)], 1)    # equidistant pixels between neighbours
    return out


IMAGES = {
    "noise_97x61": lambda: S.noise_image(97, 61, 1, alpha="random"),
    "photo_320x200": lambda: S.gradient_noise_image(320, 200, 2),
    "gray_64x64": lambda: np.repeat(S.noise_image(64, 64, 3)[..., :1], 4, axis=2),
}


@pytest.mark.parametrize("pname", sorted(palettes()))
def test_oracle_matches_numpy(pname, oracle):
    from oracle import np_restatement as N
    pal = palettes()[pname]
    for build in IMAGES.values():
        img = np.ascontiguousarray(build())
        ia, oa = oracle.apply_palette(img, pal)
        ib, ob = N.apply_palette(img, pal)
        assert np.array_equal(ia, ib) and np.array_equal(oa, ob)
        assert np.all(oa[..., 3] == 255)


def test_oracle_ties_take_first_entry(oracle):
    pal = np.array([[10, 10, 10, 255], [30, 30, 30, 255], [10, 10, 10, 255]], np.uint8)
    img = np.zeros((1, 3, 4), np.uint8)
    img[0, 0, :3] = 20          # equidistant from entries 0 and 1 → 0
    img[0, 1, :3] = 10          # exact match of entries 0 and 2 → 0
    img[0, 2, :3] = 31
    idx, out = oracle.apply_palette(img, pal)
    assert idx.tolist() == [[0, 0, 1]] and out[0, 2, :3].tolist() == [30, 30, 30]


@pytest.mark.gpu
@pytest.mark.parametrize("pname", sorted(palettes()))
def test_gpu_apply_palette_bit_exact(pname, lib, oracle):
    from fennec_b200 import api
    pal = palettes()[pname]
    for build in IMAGES.values():
        img = np.ascontiguousarray(build())
        gi, go = api.apply_palette(img, pal)
        oi, oo = oracle.apply_palette(img, pal)
        assert np.array_equal(gi, oi) and np.array_equal(go, oo)


@pytest.mark.gpu
def test_gpu_apply_palette_batch_and_large(lib, oracle):
    import torch
    from fennec_b200 import api, batch
    pals = list(palettes().values())
    imgs = [S.gradient_noise_image(200, 120, 30 + i) for i in range(3)]
    pal_t = torch.zeros((3, 256, 4), dtype=torch.uint8)
    use = [pals[3][:16], pals[3][:16], pals[3][:16]]
    for i in range(3):
        pal_t[i, :16] = torch.from_numpy(use[i])
    idx, out = batch.apply_palette_batch(torch.from_numpy(np.stack(imgs)).cuda(), pal_t.cuda(), 16)
    for i, img in enumerate(imgs):
        oi, oo = oracle.apply_palette(img, use[i])
        assert np.array_equal(idx[i].cpu().numpy(), oi) and np.array_equal(out[i].cpu().numpy(), oo)
    big = S.gradient_noise_image(1301, 703, 9)
    gi, go = api.apply_palette(big, pals[5])
    oi, oo = oracle.apply_palette(big, pals[5])
    assert np.array_equal(gi, oi) and np.array_equal(go, oo)
    # idempotence: quantising the reconstruction changes nothing (no duplicate entries in this palette)
    gi2, go2 = api.apply_palette(go, pals[5])
    assert np.array_equal(go2, go)

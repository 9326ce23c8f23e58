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


def cell_palettes():
    """Palettes aimed at the cell-list path (palette.cu): crowded cells (overflow → full scan), a dense diagonal
    (long lists), duplicates (ties across list positions), entries on cell borders, and a single entry."""
    rng = np.random.Generator(np.random.PCG64(11))
    out = {"random_256": palettes()["random_256"], "single": palettes()["random_1"]}
    crowd = rng.integers(100, 112, (256, 4), dtype=np.uint8)             # 256 entries inside a 12^3 cube
    crowd[:, 3] = 255
    out["crowded_256"] = crowd
    g = np.arange(256, dtype=np.uint8)
    out["gray_256"] = np.stack([g, g, g, np.full_like(g, 255)], 1)
    dup = rng.integers(0, 256, (128, 4), dtype=np.uint8)
    dup[:, 3] = 255
    out["duplicates_256"] = np.concatenate([dup, dup[::-1]])
    border = rng.integers(0, 32, (200, 4), dtype=np.uint8) * 8
    border[::2, :3] += 7                                                  # first / last colour of a cell
    border[:, 3] = 255
    out["cell_borders_200"] = border
    half = np.concatenate([crowd[:40], rng.integers(0, 256, (60, 4), dtype=np.uint8)])   # one crowded cell + sparse rest
    half[:, 3] = 255
    out["mixed_100"] = half
    return out


def colour_sweep():
    """4 Mi pixels: every (r, g) with the 64 blue values that sit first or last in a cell (cell edges on every axis
    are hit through r and g anyway) — 2048 x 2048."""
    b = np.array([v for v in range(256) if v % 8 in (0, 7)], np.uint8)
    r, g, bb = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), b, indexing="ij")
    img = np.stack([r, g, bb, np.full_like(r, 255)], -1).reshape(2048, 2048, 4)
    return np.ascontiguousarray(img)


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


def _cell_lists(pal, shift=3):
    """NumPy statement of palette_cells_kernel's rule: keep p when dmin(p, cell) <= min_q dmax(q, cell)."""
    axis = 256 >> shift
    c = np.arange(axis ** 3)
    lo = np.stack([c % axis, (c // axis) % axis, c // (axis * axis)], 1).astype(np.int64) << shift      # (cells, 3)
    hi = lo + (1 << shift) - 1
    p = pal[:, :3].astype(np.int64)[None]                                                                # (1, n, 3)
    near = np.maximum(np.maximum(lo[:, None] - p, p - hi[:, None]), 0)
    far = np.maximum(p - lo[:, None], hi[:, None] - p)
    dmin, dmax = (near ** 2).sum(-1), (far ** 2).sum(-1)
    return dmin <= dmax.min(1, keepdims=True)                                                            # (cells, n)


@pytest.mark.parametrize("pname", sorted(cell_palettes()))
def test_cell_list_rule_never_drops_the_winner(pname):
    """The pruning argument of palette.cu, checked independently of CUDA: over 0.36 M colours (random + every cell
    corner) the first-minimum entry of the full scan is on its cell's list."""
    pal = cell_palettes()[pname]
    cand = _cell_lists(pal)
    rng = np.random.Generator(np.random.PCG64(5))
    corners = np.array([[x * 8 + dx, y * 8 + dy, z * 8 + dz] for x in range(0, 32, 3) for y in range(0, 32, 3) for z in range(32)
                        for dx in (0, 7) for dy in (0, 7) for dz in (0, 7)], np.int64)
    cols = np.concatenate([rng.integers(0, 256, (250_000, 3)), corners])
    p = pal[:, :3].astype(np.int64)
    for lo in range(0, len(cols), 1 << 16):
        x = cols[lo:lo + (1 << 16)]
        d = ((x[:, None, :] - p[None]) ** 2).sum(-1)
        win = d.argmin(1)                                       # first minimum, as the reference's `<` scan
        cell = (x[:, 0] >> 3) | ((x[:, 1] >> 3) << 5) | ((x[:, 2] >> 3) << 10)
        assert cand[cell, win].all()
        # and the minimum over the list alone is the same entry (ties resolved by index inside the list too)
        masked = np.where(cand[cell], d, np.iinfo(np.int64).max)
        assert np.array_equal(masked.argmin(1), win)
    if pname == "crowded_256":
        assert cand.sum(1).max() > 32                            # really exercises the overflow → full-scan branch
    if pname == "random_256":
        assert cand.sum(1).max() <= 32 and cand.sum(1).mean() < 12


@pytest.mark.gpu
@pytest.mark.parametrize("pname", sorted(cell_palettes()))
def test_gpu_apply_palette_cell_lists_colour_sweep(pname, lib, oracle):
    """Images above the cell-list threshold: 4 Mi distinct colours incl. every cell edge, bit-exact vs the oracle."""
    from fennec_b200 import api
    pal = cell_palettes()[pname]
    img = colour_sweep()
    gi, go = api.apply_palette(img, pal)
    oi, oo = oracle.apply_palette(img, pal)
    assert np.array_equal(gi, oi) and np.array_equal(go, oo)


@pytest.mark.gpu
def test_gpu_apply_palette_cell_lists_ragged_and_batch(lib, oracle):
    """Odd width / unaligned tail through the cell path, and two images with different palettes in one launch."""
    import torch
    from fennec_b200 import api, batch
    cp = cell_palettes()
    img = S.gradient_noise_image(1023, 517, 4)
    for name in ("random_256", "crowded_256", "mixed_100"):
        gi, go = api.apply_palette(img, cp[name])
        oi, oo = oracle.apply_palette(img, cp[name])
        assert np.array_equal(gi, oi) and np.array_equal(go, oo), name
    imgs = [S.gradient_noise_image(640, 400, 50), S.noise_image(640, 400, 51)]
    pal_t = torch.zeros((2, 256, 4), dtype=torch.uint8)
    pal_t[0] = torch.from_numpy(cp["random_256"])
    pal_t[1] = torch.from_numpy(cp["gray_256"])
    idx, out = batch.apply_palette_batch(torch.from_numpy(np.stack(imgs)).cuda(), pal_t.cuda(), 256)
    for i, name in enumerate(("random_256", "gray_256")):
        oi, oo = oracle.apply_palette(imgs[i], cp[name])
        assert np.array_equal(idx[i].cpu().numpy(), oi) and np.array_equal(out[i].cpu().numpy(), oo)


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

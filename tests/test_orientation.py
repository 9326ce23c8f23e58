"""ApplyOrientation (exif.go:176-203): the C oracle (the reference's index loops) against NumPy array
operations, then the CUDA path through the C ABI — pure permutations, bit-exact."""
import numpy as np
import pytest

from fennec_b200 import synth as S

SIZES = [(64, 48), (33, 65), (1, 7), (7, 1), (100, 100), (257, 31)]


@pytest.mark.parametrize("orient", range(0, 10))
def test_oracle_matches_numpy(orient, oracle):
    from oracle import np_restatement as N
    for w, h in SIZES:
        img = S.noise_image(w, h, w * 100 + h, alpha="random")
        a, b = oracle.apply_orientation(img, orient), N.apply_orientation(img, orient)
        assert a.shape == b.shape and np.array_equal(a, b)
        if orient < 2 or orient > 8:
            assert a is img          # exif.go:178-179, 201: the input itself
        if orient >= 5 and orient <= 8:
            assert a.shape[:2] == (w, h)


def test_reference_orientation_test_on_oracle(oracle):   # fennec_test.go:802-825
    img = S.make_test_image(100, 50)
    assert oracle.apply_orientation(img, 1) is img
    assert oracle.apply_orientation(img, 6).shape[:2] == (100, 50)      # 50 wide, 100 tall
    assert oracle.apply_orientation(img, 3).shape[:2] == (50, 100)


def test_oracle_group_structure(oracle):
    img = S.noise_image(37, 23, 9, alpha="random")
    r90 = lambda a: oracle.apply_orientation(a, 6)  # noqa: E731
    assert np.array_equal(r90(r90(r90(r90(img)))), img)
    assert np.array_equal(r90(r90(img)), oracle.apply_orientation(img, 3))
    assert np.array_equal(oracle.apply_orientation(oracle.apply_orientation(img, 6), 8), img)
    for o in (2, 3, 4, 5, 7):       # involutions
        assert np.array_equal(oracle.apply_orientation(oracle.apply_orientation(img, o), o), img)


@pytest.mark.gpu
@pytest.mark.parametrize("orient", range(0, 10))
def test_gpu_orientation_bit_exact(orient, lib, oracle):
    from fennec_b200 import api
    for w, h in SIZES + [(640, 480), (1301, 703)]:
        img = S.noise_image(w, h, w * 100 + h, alpha="random")
        got = api.ApplyOrientation(img, orient)
        if orient < 2 or orient > 8:
            assert got is img
        else:
            assert np.array_equal(got, oracle.apply_orientation(img, orient))


@pytest.mark.gpu
def test_gpu_orientation_batch_and_strided(lib, oracle):
    import torch
    from fennec_b200 import api, batch
    imgs = [S.noise_image(130, 70, 50 + i, alpha="random") for i in range(3)]
    d = torch.from_numpy(np.stack(imgs)).cuda()
    for o in range(2, 9):
        got = batch.apply_orientation_batch(d, o).cpu().numpy()
        for i, img in enumerate(imgs):
            assert np.array_equal(got[i], oracle.apply_orientation(img, o))
    assert batch.apply_orientation_batch(d, 1) is d
    wide = S.noise_image(200, 60, 60, alpha="random")
    view = wide[:, 5:104]
    for o in (2, 6, 7):
        assert np.array_equal(api.ApplyOrientation(view, o), oracle.apply_orientation(np.ascontiguousarray(view), o))


@pytest.mark.gpu
def test_gpu_orientation_4k_properties(lib):
    import torch
    from fennec_b200 import batch
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randint(0, 256, (2, 2160, 3840, 4), dtype=torch.uint8, device="cuda", generator=g)
    r = batch.apply_orientation_batch
    assert torch.equal(r(r(r(r(x, 6), 6), 6), 6), x)
    assert torch.equal(r(x, 6), torch.rot90(x, k=-1, dims=(1, 2)))
    assert torch.equal(r(x, 7), x.transpose(1, 2))
    assert torch.equal(r(x, 2), torch.flip(x, dims=[2]))

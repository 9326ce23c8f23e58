"""GPU suite: the alternate code paths stay correct — the generic (first-version) kernels behind the fast ones,
the 2-columns-per-lane and the warp-specialised SSIM kernels — and the C ABI is safe under concurrent callers."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from fennec_b200 import api
from fennec_b200 import synth as S

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("env", [
    {"FB_SSIM_WS": "1"}, {"FB_SSIM_CPL": "2"}, {"FB_SSIM_CPL": "3"},
    {"FB_LZ_OLD": "1"}, {"FB_LZ_NO_OPAQUE": "1"}, {"FB_LZ_KO": "8"}, {"FB_FX_NOFAST": "1"},
    {"FB_SSIM_MODE": "0", "FB_SSIM_WPB": "4"}, {"FB_SSIM_MODE": "0", "FB_SSIM_WPB": "1"},
    {"FB_SSIM_MODE": "2", "FB_SSIM_WPB": "4"}, {"FB_SSIM_MODE": "1"}, {"FB_SSIM_MODE": "3"}, {"FB_SSIM_MODE": "4"}, {"FB_SSIM_MODE": "5"}, {"FB_SSIM_MODE": "6"},
    {"FB_BLUR_WPB": "4", "FB_BOX_NOFUSE2": "1", "FB_NO_STAGING": "1"},
    {"FB_BLUR_GENERIC": "1", "FB_FX_GENERIC": "1", "FB_RESIZE_GENERIC": "1"},
])
def test_kernel_variant_matches_golden(env, lib):
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "variant_check.py")], env=e, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0 and "variant ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_concurrent_callers_get_serial_results(lib):
    # CompressBatch runs NumCPU workers through the hot path at once (batch.go:84-124): each calling thread
    # has its own stream and arenas, results must not depend on interleaving.
    imgs = [S.noise_image(300 + 16 * i, 200 + 8 * i, i, alpha="random") for i in range(12)]
    pert = [S.perturb(x, 100 + i, 8) for i, x in enumerate(imgs)]

    def work(i):
        return (api.SSIM(imgs[i], pert[i]), api.MSSSIM(imgs[i], pert[i]), api.GaussianBlur(imgs[i], 2.0),
                api.Sharpen(imgs[i], 0.5), api.lanczos_resize(imgs[i], 97, 61), api.box_downsample(imgs[i], 50, 40))

    serial = [work(i) for i in range(len(imgs))]
    with ThreadPoolExecutor(max_workers=8) as ex:
        for _ in range(3):
            conc = list(ex.map(work, range(len(imgs))))
            for s_, c_ in zip(serial, conc):
                assert s_[0] == c_[0] and s_[1] == c_[1]
                for x, y in zip(s_[2:], c_[2:]):
                    assert np.array_equal(x, y)


def test_cpp_host_mirror(lib):
    """include/fennec.hpp — the C++ mirror of the Go API over the C ABI — built and run as its own program."""
    exe = os.path.join(ROOT, "tests", "cpp", "_build", "test_host")
    src = os.path.join(ROOT, "tests", "cpp", "test_host.cpp")
    if not os.path.exists(exe) or os.path.getmtime(exe) < os.path.getmtime(src):
        subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "cpp"), "-s"], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


def test_lanczos_pipelined_sub_batches_match(lib):
    """FB_LZ_PIPE=1: the batch runs as sub-batches with the vertical pass on a side stream (csrc/api.cu resize_on_device, the
    path peer-gathered destinations take above four ranks) — same bytes as the per-image host entry point."""
    code = (
        "import numpy as np, torch\n"
        "from fennec_b200 import api, batch, synth as S\n"
        "for n, (w, h), (dw, dh) in [(5, (512, 128), (128, 32)), (2, (300, 90), (75, 40)), (9, (256, 64), (64, 16))]:\n"
        "    imgs = [S.noise_image(w, h, 7 * n + i, alpha=('opaque', 'random')[i % 2]) for i in range(n)]\n"
        "    d = torch.from_numpy(np.stack(imgs)).cuda()\n"
        "    for rep in range(3):\n"
        "        got = batch.lanczos_resize_batch(d, dw, dh).cpu().numpy()\n"
        "        assert all(np.array_equal(got[i], api.lanczos_resize(imgs[i], dw, dh)) for i in range(n)), (n, rep)\n"
        "print('pipelined ok')\n")
    e = dict(os.environ)
    e["FB_LZ_PIPE"] = "1"
    e["PYTHONPATH"] = ROOT + os.pathsep + e.get("PYTHONPATH", "")
    r = subprocess.run([sys.executable, "-c", code], env=e, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "pipelined ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]

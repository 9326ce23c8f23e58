"""Device-resident batch entry points and the batch sharder (batch.go:58-128), over the C ABI.

torch is plumbing only: device memory (uint8 tensors), the current CUDA stream handed to the
library, and torch.distributed for the result gather.  Batches are tensors of shape (n, h, w, 4),
uint8, on a CUDA device; image i is the NRGBA buffer at data_ptr() + i*stride(0).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import FB_IDENTITY, check, dp


def _batch(t: torch.Tensor):
    if not (t.is_cuda and t.dtype == torch.uint8 and t.dim() == 4 and t.shape[3] == 4):
        raise TypeError("expected a CUDA uint8 tensor of shape (n, h, w, 4)")
    if t.stride(3) != 1 or t.stride(2) != 4:
        raise ValueError("pixels must be interleaved NRGBA bytes")
    n, h, w, _ = t.shape
    return t.data_ptr(), int(t.stride(0)), int(t.stride(1)), w, h, n


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _dev(t: torch.Tensor) -> int:
    return t.device.index if t.device.index is not None else torch.cuda.current_device()


def _scores(fn, a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor]) -> torch.Tensor:
    pa, ia, ra, w, h, n = _batch(a)
    pb, ib, rb, wb, hb, nb = _batch(b)
    if (ia, ra, w, h, n) != (ib, rb, wb, hb, nb):
        raise ValueError("a and b must have identical shapes and strides")
    if out is None:
        out = torch.empty(n, dtype=torch.float64, device=a.device)
    check(fn(_dev(a), _stream(a), pa, pb, ia, ra, w, h, n, out.data_ptr()))
    return out


def ssim_batch(a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fennec.SSIM for n device-resident pairs (ssim.go:24-43) → float64 scores on the device."""
    return _scores(_lib.load().fb_ssim_batch_dev, a, b, out)


def ssim_fast_batch(a, b, out=None) -> torch.Tensor:
    """fennec.SSIMFast per pair (ssim.go:48-70)."""
    return _scores(_lib.load().fb_ssim_fast_batch_dev, a, b, out)


def msssim_batch(a, b, out=None) -> torch.Tensor:
    """fennec.MSSSIM per pair (ssim.go:313-365)."""
    return _scores(_lib.load().fb_msssim_batch_dev, a, b, out)


def box_downsample_batch(src: torch.Tensor, dst_w: int, dst_h: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    ps, i_s, rs, w, h, n = _batch(src)
    if out is None:
        out = torch.zeros((n, dst_h, dst_w, 4), dtype=torch.uint8, device=src.device)
    pd, i_d, rd, _, _, _ = _batch(out)
    check(_lib.load().fb_box_downsample_batch_dev(_dev(src), _stream(src), ps, i_s, rs, w, h, pd, i_d, rd, dst_w, dst_h, n))
    return out


def msssim_level_batch(a: torch.Tensor, b: torch.Tensor, tw: int, th: int):
    """One MS-SSIM level step (ssim.go:57-58 + 354-360): (thumbA, thumbB, halfA, halfB) from one read of each image."""
    pa, i_s, rs, w, h, n = _batch(a)
    pb = _batch(b)[0]
    thumbs = [torch.zeros((n, th, tw, 4), dtype=torch.uint8, device=a.device) for _ in range(2)]
    halves = [torch.zeros((n, h // 2, w // 2, 4), dtype=torch.uint8, device=a.device) for _ in range(2)]
    _, ti, tr, _, _, _ = _batch(thumbs[0])
    _, hi, hr, _, _, _ = _batch(halves[0])
    check(_lib.load().fb_msssim_level_batch_dev(_dev(a), _stream(a), pa, pb, i_s, rs, w, h, n, thumbs[0].data_ptr(),
                                                thumbs[1].data_ptr(), ti, tr, tw, th, halves[0].data_ptr(),
                                                halves[1].data_ptr(), hi, hr))
    return thumbs[0], thumbs[1], halves[0], halves[1]


def msssim_level2_batch(a: torch.Tensor, b: torch.Tensor, tw: int, th: int):
    """Two MS-SSIM level steps from one read: (thumb0A, thumb0B, thumb1A, thumb1B, quarterA, quarterB), or None when the
    geometry has no common box period (fb_msssim_level2_batch_dev answers 1)."""
    pa, i_s, rs, w, h, n = _batch(a)
    pb = _batch(b)[0]
    thumbs = [torch.zeros((n, th, tw, 4), dtype=torch.uint8, device=a.device) for _ in range(4)]
    quarters = [torch.zeros((n, h // 4, w // 4, 4), dtype=torch.uint8, device=a.device) for _ in range(2)]
    _, ti, tr, _, _, _ = _batch(thumbs[0])
    _, qi, qr, _, _, _ = _batch(quarters[0])
    rc = check(_lib.load().fb_msssim_level2_batch_dev(_dev(a), _stream(a), pa, pb, i_s, rs, w, h, n, thumbs[0].data_ptr(),
                                                      thumbs[1].data_ptr(), thumbs[2].data_ptr(), thumbs[3].data_ptr(), ti, tr, tw, th,
                                                      quarters[0].data_ptr(), quarters[1].data_ptr(), qi, qr))
    if rc == 1:
        return None
    return thumbs[0], thumbs[1], thumbs[2], thumbs[3], quarters[0], quarters[1]


def ycbcr_to_nrgba_batch(y: torch.Tensor, cb: torch.Tensor, cr: torch.Tensor, ratio: int,
                         out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """convertToNRGBA (convert.go:34-64) for n device-resident YCbCr images: y (n,h,w), cb/cr (n,ch,cw), uint8."""
    for t in (y, cb, cr):
        if not (t.is_cuda and t.dtype == torch.uint8 and t.dim() == 3 and t.stride(2) == 1):
            raise TypeError("expected CUDA uint8 planes of shape (n, rows, cols)")
    if cb.shape != cr.shape or cb.stride() != cr.stride():
        raise ValueError("Cb and Cr must share shape and strides")
    n, h, w = y.shape
    if out is None:
        out = torch.empty((n, h, w, 4), dtype=torch.uint8, device=y.device)
    pd, i_d, rd, _, _, _ = _batch(out)
    check(_lib.load().fb_ycbcr_to_nrgba_batch_dev(_dev(y), _stream(y), y.data_ptr(), int(y.stride(0)), int(y.stride(1)),
                                                  cb.data_ptr(), cr.data_ptr(), int(cb.stride(0)), int(cb.stride(1)), w, h,
                                                  ratio, pd, i_d, rd, n))
    return out


def convert_to_nrgba_batch(fmt: int, pix: torch.Tensor, palettes16: Optional[torch.Tensor] = None, ncolors: int = 0,
                           out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """convertToNRGBA (convert.go:34-64) for n device-resident decoded images of one type (api.FMT_*): pix is
    (n, h, w, bytes-per-pixel) uint8 — (n, h, w) for Paletted, with palettes16 (n, 256, 4) int16/uint16 bit patterns."""
    if pix.dim() == 3:
        pix = pix.unsqueeze(-1)
    if not (pix.is_cuda and pix.dtype == torch.uint8 and pix.dim() == 4):
        raise TypeError("expected a CUDA uint8 tensor of shape (n, h, w, bytes-per-pixel)")
    n, h, w, bpp = pix.shape
    if (bpp > 1 and pix.stride(3) != 1) or (w > 1 and pix.stride(2) != bpp):
        raise ValueError("pixels must be packed within a row")
    pal_ptr = 0
    if palettes16 is not None:
        if not (palettes16.is_cuda and palettes16.is_contiguous() and palettes16.element_size() == 2
                and tuple(palettes16.shape) == (n, 256, 4)):
            raise TypeError("palettes16 must be a contiguous CUDA 16-bit tensor of shape (n, 256, 4)")
        pal_ptr = palettes16.data_ptr()
    if out is None:
        out = torch.empty((n, h, w, 4), dtype=torch.uint8, device=pix.device)
    pd, i_d, rd, _, _, _ = _batch(out)
    check(_lib.load().fb_convert_to_nrgba_batch_dev(_dev(pix), _stream(pix), fmt, pix.data_ptr(), int(pix.stride(0)),
                                                    int(pix.stride(1)), w, h, n, pal_ptr, ncolors, pd, i_d, rd))
    return out


def analyze_batch(imgs: torch.Tensor) -> List[dict]:
    """fennec.Analyze (analyze.go:26-176) for n device-resident images: the scans run on the device, the raw records
    (histogram, integer sums, counts) come back in one copy and are finished with host arithmetic."""
    p, i_s, rs, w, h, n = _batch(imgs)
    L = _lib.load()
    rec = int(L.fb_analyze_raw_bytes())
    raw = torch.empty(n * rec, dtype=torch.uint8, device=imgs.device)
    check(L.fb_analyze_batch_dev(_dev(imgs), _stream(imgs), p, i_s, rs, w, h, n, raw.data_ptr()))
    host = raw.cpu().numpy()
    out = []
    for i in range(n):
        st = _lib.FbImageStats()
        check(L.fb_analyze_finish(host[i * rec:(i + 1) * rec].ctypes.data, w, h, C.byref(st)))
        out.append({k: getattr(st, k) for k, _ in _lib.FbImageStats._fields_})
    return out


def analyze_scan_batch(imgs: torch.Tensor, raw: torch.Tensor) -> None:
    """Enqueue only (what tools/bench_ops.py times): raw must hold n * fb_analyze_raw_bytes() bytes on the device."""
    p, i_s, rs, w, h, n = _batch(imgs)
    check(_lib.load().fb_analyze_batch_dev(_dev(imgs), _stream(imgs), p, i_s, rs, w, h, n, raw.data_ptr()))


def apply_orientation_batch(src: torch.Tensor, orient: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fennec.ApplyOrientation (exif.go:176-203) per image; identity orientations return `src` itself."""
    ps, i_s, rs, w, h, n = _batch(src)
    if orient < 2 or orient > 8:
        return src
    dw, dh = (h, w) if orient >= 5 else (w, h)
    if out is None:
        out = torch.empty((n, dh, dw, 4), dtype=torch.uint8, device=src.device)
    pd, i_d, rd, _, _, _ = _batch(out)
    check(_lib.load().fb_apply_orientation_batch_dev(_dev(src), _stream(src), ps, i_s, rs, w, h, orient, pd, i_d, rd, n))
    return out


def apply_palette_batch(src: torch.Tensor, palettes: torch.Tensor, ncolors: int):
    """applyPalette + palettedToNRGBA per image; palettes: (n, 256, 4) uint8 on the device → (indices (n,h,w), NRGBA)."""
    ps, i_s, rs, w, h, n = _batch(src)
    if not (palettes.is_cuda and palettes.dtype == torch.uint8 and tuple(palettes.shape) == (n, 256, 4) and palettes.is_contiguous()):
        raise TypeError("palettes must be a contiguous CUDA uint8 tensor of shape (n, 256, 4)")
    idx = torch.empty((n, h, w), dtype=torch.uint8, device=src.device)
    out = torch.empty_like(src)
    po, i_o, ro, _, _, _ = _batch(out)
    check(_lib.load().fb_apply_palette_batch_dev(_dev(src), _stream(src), ps, i_s, rs, w, h, n, palettes.data_ptr(), ncolors,
                                                 idx.data_ptr(), int(idx.stride(0)), int(idx.stride(1)), po, i_o, ro))
    return idx, out


def gaussian_blur_batch(src: torch.Tensor, sigma: float, out: Optional[torch.Tensor] = None,
                        kernel: Optional[np.ndarray] = None) -> torch.Tensor:
    """fennec.GaussianBlur per image (effects.go:146-220); sigma <= 0 returns `src` itself."""
    if sigma <= 0:
        return src
    from .api import blur_kernel
    if kernel is None:
        kernel, radius = blur_kernel(sigma)
    else:
        kernel = np.ascontiguousarray(kernel, dtype=np.float64)
        radius = (len(kernel) - 1) // 2
    ps, i_s, rs, w, h, n = _batch(src)
    out = _like(src, out)
    check(_lib.load().fb_gaussian_blur_batch_dev(_dev(src), _stream(src), ps, out.data_ptr(), i_s, rs, w, h, n,
                                                 kernel.ctypes.data_as(dp), radius))
    return out


def _like(src: torch.Tensor, out: Optional[torch.Tensor]) -> torch.Tensor:
    """Destination for the blur / sharpen entry points, which apply src's image and row strides to dst as well: a fresh
    tensor gets src's exact strides (empty_like would compact a row-padded or sliced src and the kernel would then
    write past its end); a caller-supplied one must match src in shape AND strides."""
    if out is None:
        return torch.empty_strided(tuple(src.shape), tuple(src.stride()), dtype=src.dtype, device=src.device)
    _batch(out)
    if out.device != src.device or tuple(out.shape) != tuple(src.shape) or tuple(out.stride()) != tuple(src.stride()):
        raise ValueError(f"out must match src in device, shape and strides: src {tuple(src.shape)}/{tuple(src.stride())}, "
                         f"out {tuple(out.shape)}/{tuple(out.stride())}")
    return out


def _fx(fn, src, strength, out):
    ps, i_s, rs, w, h, n = _batch(src)
    out = _like(src, out)
    if check(fn(_dev(src), _stream(src), ps, out.data_ptr(), i_s, rs, w, h, n, float(strength))) == FB_IDENTITY:
        return src
    return out


def sharpen_batch(src: torch.Tensor, strength: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fennec.Sharpen per image (effects.go:10-45)."""
    return _fx(_lib.load().fb_sharpen_batch_dev, src, strength, out)


def adaptive_sharpen_batch(src, strength, out=None) -> torch.Tensor:
    """fennec.AdaptiveSharpen per image (effects.go:49-90)."""
    return _fx(_lib.load().fb_adaptive_sharpen_batch_dev, src, strength, out)


def lanczos_resize_batch(src: torch.Tensor, dst_w: int, dst_h: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """lanczosResize per image (resize.go:37-53)."""
    ps, i_s, rs, w, h, n = _batch(src)
    if out is None:
        out = torch.zeros((n, dst_h, dst_w, 4), dtype=torch.uint8, device=src.device)
    pd, i_d, rd, _, _, _ = _batch(out)
    check(_lib.load().fb_lanczos_resize_batch_dev(_dev(src), _stream(src), ps, i_s, rs, w, h, pd, i_d, rd, dst_w, dst_h, n))
    return out


def take_launch_count() -> int:
    return int(_lib.load().fb_take_launch_count())


# ---- sharder: the part of CompressBatch (batch.go:58-128) that moves to the GPUs ----------------------

def shard_range(n_items: int, n_shards: int, shard: int):
    """Static contiguous partition: shard s owns [begin, end)."""
    b, e = C.c_int(), C.c_int()
    check(_lib.load().fb_batch_shard(n_items, n_shards, shard, C.byref(b), C.byref(e)))
    return b.value, e.value


@dataclass
class BatchResult:
    """batch.go:21-30"""
    item: object
    result: object = None
    err: Optional[BaseException] = None
    index: int = 0


def run_sharded(items: Sequence, work: Callable[[object], object], *, rank: int = 0, world: int = 1,
                cancelled: Callable[[], bool] = lambda: False,
                on_item: Optional[Callable[[int, int], None]] = None) -> List[Optional[BatchResult]]:
    """The worker loop of CompressBatch (batch.go:84-124) for ONE shard: results keep the input index,
    a cancelled context marks unstarted items with an error instead of running them (batch.go:90-98),
    one failing item does not stop the rest (batch.go:107-113), on_item(completed, total) fires after
    each item (batch.go:115-121).  Items owned by other shards are left as None for the gather."""
    total = len(items)
    out: List[Optional[BatchResult]] = [None] * total
    if total == 0:
        return out
    begin, end = shard_range(total, world, rank)
    completed = 0
    for idx in range(begin, end):
        if cancelled():
            out[idx] = BatchResult(items[idx], None, RuntimeError("context canceled"), idx)
            continue
        try:
            out[idx] = BatchResult(items[idx], work(items[idx]), None, idx)
        except Exception as e:  # per-item isolation
            out[idx] = BatchResult(items[idx], None, e, idx)
        if on_item is not None:
            completed += 1
            on_item(completed, total)
    return out


def gather_scores(local: torch.Tensor, n_items: int, world: int, rank: int) -> torch.Tensor:
    """All-gather the per-shard float64 scores into input order. The only collective of the path:
    per-image work is embarrassingly parallel. Works on nccl (CUDA) and gloo (CPU)."""
    import torch.distributed as dist
    if world == 1:
        return local
    per = (n_items + world - 1) // world
    padded = torch.zeros(per, dtype=local.dtype, device=local.device)
    padded[: local.numel()] = local
    out = torch.empty(per * world, dtype=local.dtype, device=local.device)
    if hasattr(dist, "all_gather_into_tensor") and local.is_cuda:
        dist.all_gather_into_tensor(out, padded)   # one ncclAllGather, no per-shard copies
    else:
        parts = [torch.empty_like(padded) for _ in range(world)]
        dist.all_gather(parts, padded)
        out = torch.cat(parts)
    return out[:n_items]


class PeerGather:
    """Gather of a sharded batch's PIXEL outputs, fused into the producing kernel's own stores.

    The gathered buffer (all items, input order) lives on `root`; every rank gets a view of ITS slice of that buffer,
    mapped over NVLink through torch's symmetric memory, and passes it as `out=` to a batch entry point.  The kernel's
    stores are the transfer: no collective runs, no second pass reads the outputs again, and the copy overlaps the
    arithmetic store by store.  (Measured at N=2, 8 items of Lanczos 8K->1080p per GPU: 0.876 ms against 0.867 ms of
    compute alone and 1.047 ms with ncclAllGather after the kernel; profiles/r1d_sharded_*.)  Scores keep using
    gather_scores: 8 bytes per item are not worth a mapping.
    """

    def __init__(self, total_items: int, item_shape, world: int, rank: int, root: int = 0,
                 dtype: torch.dtype = torch.uint8, device: Optional[torch.device] = None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        self.world, self.rank, self.root = world, rank, root
        self.shape = (total_items,) + tuple(item_shape)
        device = device or torch.device("cuda", torch.cuda.current_device())
        self.local = symm_mem.empty(self.shape, dtype=dtype, device=device)   # only root's copy is used
        self.handle = symm_mem.rendezvous(self.local, dist.group.WORLD)
        self.begin, self.end = shard_range(total_items, world, rank)
        self._root_view = self.handle.get_buffer(root, self.shape, dtype)

    def my_slice(self) -> torch.Tensor:
        """Destination for this rank's items: rows [begin, end) of root's buffer (peer-mapped unless rank == root)."""
        return self._root_view[self.begin:self.end]

    def finish(self) -> Optional[torch.Tensor]:
        """Wait until every rank's kernels have stored their slice; the gathered tensor on root, None elsewhere."""
        torch.cuda.current_stream().synchronize()
        self.handle.barrier()
        return self.local if self.rank == self.root else None

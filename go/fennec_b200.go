//go:build cgo && fennec_b200

// Package fennec — cgo shim binding the B200 hot path (libfennec_b200.so) behind fennec's own
// function bodies.  AUTHORED, NOT COMPILED: no Go toolchain exists in the build image (see
// INTEGRATION.md).  It is deliberately mechanical: pointer/len/stride marshalling and
// status → pure-Go fallback, nothing else.
//
// cgo rules honoured: only &img.Pix[0] (a Go pointer to pointer-free memory) crosses the boundary,
// the library retains no pointer after return, and every call is synchronous.
package fennec

/*
#cgo CFLAGS: -I${SRCDIR}/../include
#cgo LDFLAGS: -L${SRCDIR}/../fennec_b200 -lfennec_b200 -Wl,-rpath,${SRCDIR}/../fennec_b200
#include "fennec_b200.h"
*/
import "C"

import (
	"image"
	"image/color"
	"runtime"
	"unsafe"
)

func pix(img *image.NRGBA) *C.uint8_t {
	if len(img.Pix) == 0 {
		return nil
	}
	return (*C.uint8_t)(unsafe.Pointer(&img.Pix[img.PixOffset(img.Rect.Min.X, img.Rect.Min.Y)]))
}

// gpuSSIM replaces the tail of SSIM (ssim.go:35-42). ok=false → run the pure-Go body.
func gpuSSIM(a, b *image.NRGBA) (float64, bool) {
	var out C.double
	st := C.fb_ssim(pix(a), C.int(a.Stride), pix(b), C.int(b.Stride),
		C.int(a.Bounds().Dx()), C.int(a.Bounds().Dy()), &out)
	return float64(out), st == C.FB_OK
}

// gpuSSIMFast replaces the body of SSIMFast (ssim.go:48-70).
func gpuSSIMFast(a, b *image.NRGBA) (float64, bool) {
	var out C.double
	st := C.fb_ssim_fast(pix(a), C.int(a.Stride), pix(b), C.int(b.Stride),
		C.int(a.Bounds().Dx()), C.int(a.Bounds().Dy()), &out)
	return float64(out), st == C.FB_OK
}

// gpuMSSSIM replaces MSSSIM after the size check (ssim.go:324-364).
func gpuMSSSIM(a, b *image.NRGBA) (float64, bool) {
	var out C.double
	st := C.fb_msssim(pix(a), C.int(a.Stride), pix(b), C.int(b.Stride),
		C.int(a.Bounds().Dx()), C.int(a.Bounds().Dy()), &out)
	return float64(out), st == C.FB_OK
}

// gpuBoxDownsample replaces the loops of boxDownsample (ssim.go:250-283); dst is image.NewNRGBA'd by the caller.
func gpuBoxDownsample(src, dst *image.NRGBA) bool {
	st := C.fb_box_downsample(pix(src), C.int(src.Stride), C.int(src.Bounds().Dx()), C.int(src.Bounds().Dy()),
		pix(dst), C.int(dst.Stride), C.int(dst.Bounds().Dx()), C.int(dst.Bounds().Dy()))
	return st == C.FB_OK
}

// gpuGaussianBlur replaces both passes of GaussianBlur (effects.go:167-217). The kernel slice is the one
// GaussianBlur already builds with Go's math.Exp (effects.go:155-165), so pixel parity does not depend on libm.
func gpuGaussianBlur(src, dst *image.NRGBA, kernel []float64, radius int) bool {
	st := C.fb_gaussian_blur(pix(src), C.int(src.Stride), C.int(src.Bounds().Dx()), C.int(src.Bounds().Dy()),
		(*C.double)(unsafe.Pointer(&kernel[0])), C.int(radius), pix(dst), C.int(dst.Stride))
	return st == C.FB_OK
}

// gpuSharpen / gpuAdaptiveSharpen replace effects.go:24-42 / 63-87 (guards stay in Go: they return img itself).
func gpuSharpen(src, dst *image.NRGBA, strength float64, adaptive bool) bool {
	var st C.int
	if adaptive {
		st = C.fb_adaptive_sharpen(pix(src), C.int(src.Stride), C.int(src.Bounds().Dx()), C.int(src.Bounds().Dy()),
			C.double(strength), pix(dst), C.int(dst.Stride))
	} else {
		st = C.fb_sharpen(pix(src), C.int(src.Stride), C.int(src.Bounds().Dx()), C.int(src.Bounds().Dy()),
			C.double(strength), pix(dst), C.int(dst.Stride))
	}
	return st == C.FB_OK
}

// csr flattens precomputeWeights' [][]weightEntry (resize.go:164-197) for the ABI. The backing slices
// are returned so the caller keeps them alive across the call; pointers to them are pinned for the
// duration of the call because they sit inside a C struct (cgo rule: use runtime.Pinner, Go ≥ 1.21).
func csr(w [][]weightEntry) (start, index []C.int, weight []C.double) {
	start = make([]C.int, len(w)+1)
	for d, e := range w {
		start[d+1] = start[d] + C.int(len(e))
	}
	index = make([]C.int, start[len(w)])
	weight = make([]C.double, start[len(w)])
	n := 0
	for _, e := range w {
		for _, t := range e {
			index[n], weight[n] = C.int(t.index), C.double(t.weight)
			n++
		}
	}
	return
}

// gpuLanczosResize replaces resizeH+resizeV (resize.go:51-52).  The two weight tables are the ones lanczosResize's
// own precomputeWeights builds with Go's math.Sin (resize.go:164-197), flattened by csr(): bit parity of the pixels
// then does not depend on glibc's sin agreeing with Go's (SURVEY H5).  The CSR slices are Go memory referenced from a
// C struct, so they are pinned for the duration of the call (runtime.Pinner, Go >= 1.21).
func gpuLanczosResize(src, dst *image.NRGBA) bool {
	sw, sh := src.Bounds().Dx(), src.Bounds().Dy()
	dw, dh := dst.Bounds().Dx(), dst.Bounds().Dy()
	ratioX, ratioY := float64(sw)/float64(dw), float64(sh)/float64(dh)
	supX, supY := 3.0, 3.0 // resize.go:81-85 / 125-129
	if ratioX > 1 {
		supX = 3.0 * ratioX
	}
	if ratioY > 1 {
		supY = 3.0 * ratioY
	}
	sx, ix, wx := csr(precomputeWeights(dw, sw, ratioX, supX))
	sy, iy, wy := csr(precomputeWeights(dh, sh, ratioY, supY))
	var pin runtime.Pinner
	defer pin.Unpin()
	table := func(start, index []C.int, weight []C.double, n int) *C.fb_weights {
		if len(index) == 0 {
			return nil // degenerate table: let the library build it
		}
		pin.Pin(&start[0])
		pin.Pin(&index[0])
		pin.Pin(&weight[0])
		return &C.fb_weights{n: C.int(n), start: &start[0], index: &index[0], weight: &weight[0]}
	}
	tx, ty := table(sx, ix, wx, dw), table(sy, iy, wy, dh)
	st := C.fb_lanczos_resize(pix(src), C.int(src.Stride), C.int(sw), C.int(sh),
		pix(dst), C.int(dst.Stride), C.int(dw), C.int(dh), tx, ty)
	return st == C.FB_OK
}

// ---- CompressBatch (batch.go:58-128) on several GPUs -----------------------------------------------------

// gpuWorkerInit is called once at the top of each CompressBatch worker goroutine (batch.go:84-88): the worker pins
// itself to an OS thread and takes GPU `worker % nGPU`; every hot-path call CompressFile makes from this goroutine
// (SSIMFast in the quality search, smartResize, boxDownsample) then runs on that device, on the thread's own
// stream.  When the goroutine ends the OS thread dies with it (LockOSThread without Unlock) and the library hands
// the thread's stream and arenas to the next worker (a pool, not a leak).
func gpuWorkerInit(worker int) bool {
	n := int(C.fb_device_count())
	if n <= 0 {
		return false
	}
	runtime.LockOSThread()
	return C.fb_set_device(C.int(worker%n)) == C.FB_OK
}

// gpuShard is the partition CompressBatch uses when it hands whole sub-batches to per-GPU workers
// (batch.go:63-81): shard s of n owns items [begin, end).
func gpuShard(nItems, nShards, shard int) (begin, end int, ok bool) {
	var b, e C.int
	st := C.fb_batch_shard(C.int(nItems), C.int(nShards), C.int(shard), &b, &e)
	return int(b), int(e), st == C.FB_OK
}

// gpuScoreBatch scores a list of equal-sized pairs on every GPU of the box from this one call (fb_score_batch_host:
// fb_batch_shard over the devices, the library's own worker threads, results in input order).  op is C.FB_OP_SSIM,
// C.FB_OP_SSIM_FAST or C.FB_OP_MSSSIM.  status[i] < 0 marks an item that failed or was cancelled (*cancel != 0,
// the analogue of ctx.Done() in batch.go:90-98); the caller re-runs those through the pure-Go body.
func gpuScoreBatch(op C.int, as, bs []*image.NRGBA, cancel *int32) (scores []float64, status []int32, ok bool) {
	n := len(as)
	if n == 0 || n != len(bs) {
		return nil, nil, n == 0
	}
	pairs := make([]C.fb_pair, n)
	var pin runtime.Pinner
	defer pin.Unpin()
	for i := range as {
		if len(as[i].Pix) == 0 || len(bs[i].Pix) == 0 {
			return nil, nil, false
		}
		pin.Pin(&as[i].Pix[0]) // Go pointers stored inside C-visible structs must be pinned
		pin.Pin(&bs[i].Pix[0])
		pairs[i] = C.fb_pair{a: pix(as[i]), strideA: C.int(as[i].Stride), b: pix(bs[i]), strideB: C.int(bs[i].Stride),
			w: C.int(as[i].Bounds().Dx()), h: C.int(as[i].Bounds().Dy())}
	}
	scores = make([]float64, n)
	status = make([]int32, n)
	opts := C.fb_batch_opts{workers_per_device: 4}
	if cancel != nil {
		pin.Pin(cancel)
		opts.cancel = (*C.int)(unsafe.Pointer(cancel))
	}
	failed := C.fb_score_batch_host(op, &pairs[0], C.int(n), (*C.double)(unsafe.Pointer(&scores[0])),
		(*C.int)(unsafe.Pointer(&status[0])), &opts)
	return scores, status, failed >= 0
}

// ---- SURVEY §8(f1): the step before SSIMFast in the quality search (compress.go:53-62) ----------------

// gpuConvertToNRGBA replaces convertToNRGBA's pixel loop (convert.go:38-63) for the concrete types jpeg.Decode and
// png.Decode return, with Rect.Min == (0,0); anything else keeps the pure-Go loop.
func gpuConvertToNRGBA(img image.Image, dst *image.NRGBA) bool {
	switch s := img.(type) {
	case *image.YCbCr:
		// Sub-images (SubImage shares the planes, Rect.Min != 0): the planes are addressed through YOffset / COffset
		// exactly as convertToNRGBA's At() does (convert.go:34-64 walks Bounds()).  The chroma phase of the library's
		// kernel starts at (0,0), so a sub-image must start on an even luma sample in every subsampled direction.
		if len(s.Y) == 0 || s.Rect.Empty() || s.Rect.Min.X%2 != 0 || s.Rect.Min.Y%2 != 0 {
			return false
		}
		yo, co := s.YOffset(s.Rect.Min.X, s.Rect.Min.Y), s.COffset(s.Rect.Min.X, s.Rect.Min.Y)
		st := C.fb_ycbcr_to_nrgba((*C.uint8_t)(unsafe.Pointer(&s.Y[yo])), C.int(s.YStride),
			(*C.uint8_t)(unsafe.Pointer(&s.Cb[co])), (*C.uint8_t)(unsafe.Pointer(&s.Cr[co])), C.int(s.CStride),
			C.int(s.Rect.Dx()), C.int(s.Rect.Dy()), C.int(s.SubsampleRatio), pix(dst), C.int(dst.Stride))
		return st == C.FB_OK
	case *image.Gray:
		if len(s.Pix) == 0 || s.Rect.Empty() {
			return false
		}
		st := C.fb_gray_to_nrgba((*C.uint8_t)(unsafe.Pointer(&s.Pix[s.PixOffset(s.Rect.Min.X, s.Rect.Min.Y)])), C.int(s.Stride),
			C.int(s.Rect.Dx()), C.int(s.Rect.Dy()), pix(dst), C.int(dst.Stride))
		return st == C.FB_OK
	case *image.RGBA: // png.Decode of truecolour without alpha, draw targets
		return gpuConvertPix(C.FB_FMT_RGBA, s.Pix, s.PixOffset(s.Rect.Min.X, s.Rect.Min.Y), s.Stride, s.Rect, nil, dst)
	case *image.RGBA64:
		return gpuConvertPix(C.FB_FMT_RGBA64, s.Pix, s.PixOffset(s.Rect.Min.X, s.Rect.Min.Y), s.Stride, s.Rect, nil, dst)
	case *image.NRGBA64:
		return gpuConvertPix(C.FB_FMT_NRGBA64, s.Pix, s.PixOffset(s.Rect.Min.X, s.Rect.Min.Y), s.Stride, s.Rect, nil, dst)
	case *image.Gray16:
		return gpuConvertPix(C.FB_FMT_GRAY16, s.Pix, s.PixOffset(s.Rect.Min.X, s.Rect.Min.Y), s.Stride, s.Rect, nil, dst)
	case *image.CMYK: // 4-component JPEGs
		return gpuConvertPix(C.FB_FMT_CMYK, s.Pix, s.PixOffset(s.Rect.Min.X, s.Rect.Min.Y), s.Stride, s.Rect, nil, dst)
	case *image.Paletted:
		if len(s.Palette) == 0 || len(s.Palette) > 256 {
			return false
		}
		pal := make([]uint16, 4*len(s.Palette)) // the color.Color interface is evaluated here, once per entry
		for i, c := range s.Palette {
			r, g, b, a := c.RGBA()
			pal[4*i], pal[4*i+1], pal[4*i+2], pal[4*i+3] = uint16(r), uint16(g), uint16(b), uint16(a)
		}
		return gpuConvertPix(C.FB_FMT_PALETTED, s.Pix, s.PixOffset(s.Rect.Min.X, s.Rect.Min.Y), s.Stride, s.Rect, pal, dst)
	}
	return false
}

// gpuConvertPix hands a decoded image's Pix buffer to fb_convert_to_nrgba, starting at the byte offset of the
// image's Rect.Min (PixOffset): sub-images share their parent's buffer and stride, which is all the library needs.
// A palette index past the palette makes the call fail (FB_E_INVALID); the caller then runs the pure-Go loop, which
// panics exactly as before.
func gpuConvertPix(format C.int, p []uint8, off, stride int, r image.Rectangle, pal []uint16, dst *image.NRGBA) bool {
	if r.Empty() || off < 0 || off >= len(p) {
		return false
	}
	var pp *C.uint16_t
	if len(pal) > 0 {
		pp = (*C.uint16_t)(unsafe.Pointer(&pal[0]))
	}
	st := C.fb_convert_to_nrgba(format, (*C.uint8_t)(unsafe.Pointer(&p[off])), C.int(stride), C.int(r.Dx()), C.int(r.Dy()),
		pp, C.int(len(pal)/4), pix(dst), C.int(dst.Stride))
	return st == C.FB_OK
}

// ssimSession keeps `src` (its SSIMFast thumbnail) on the device for the whole binary search of
// compressJPEG (compress.go:45-74).  Create it before the loop, Close it after; inside the loop
//
//	decoded, _ := jpeg.Decode(...)
//	ssim, ok := sess.score(decoded)          // instead of toNRGBARef(decoded) + SSIMFast(src, ·)
//	if !ok { ssim = SSIMFast(src, toNRGBARef(decoded)) }
//
// A session is bound to the OS thread's device; use it from one goroutine (runtime.LockOSThread).
type ssimSession struct{ h *C.fb_ssim_ref }

func newSSIMSession(src *image.NRGBA) (*ssimSession, bool) {
	var h *C.fb_ssim_ref
	st := C.fb_ssim_ref_create(pix(src), C.int(src.Stride), C.int(src.Bounds().Dx()), C.int(src.Bounds().Dy()), &h)
	if st != C.FB_OK {
		return nil, false
	}
	return &ssimSession{h}, true
}

func (s *ssimSession) score(decoded image.Image) (float64, bool) {
	var out C.double
	switch d := decoded.(type) {
	case *image.YCbCr:
		if d.Rect.Empty() || d.Rect.Min.X%2 != 0 || d.Rect.Min.Y%2 != 0 {
			return 0, false
		}
		yo, co := d.YOffset(d.Rect.Min.X, d.Rect.Min.Y), d.COffset(d.Rect.Min.X, d.Rect.Min.Y)
		st := C.fb_ssim_ref_score_ycbcr(s.h, (*C.uint8_t)(unsafe.Pointer(&d.Y[yo])), C.int(d.YStride),
			(*C.uint8_t)(unsafe.Pointer(&d.Cb[co])), (*C.uint8_t)(unsafe.Pointer(&d.Cr[co])), C.int(d.CStride),
			C.int(d.SubsampleRatio), &out)
		return float64(out), st == C.FB_OK
	case *image.NRGBA:
		st := C.fb_ssim_ref_score_nrgba(s.h, pix(d), C.int(d.Stride), &out)
		return float64(out), st == C.FB_OK
	}
	return 0, false
}

func (s *ssimSession) Close() { C.fb_ssim_ref_destroy(s.h); s.h = nil }

// ---- SURVEY §8(f2): Analyze (analyze.go:26-113) ---------------------------------------------------------

// gpuAnalyze replaces the three scans of Analyze; ok=false → run the pure-Go body.  The recommendation rules
// (analyze.go:183-232) are evaluated inside the library with the same thresholds, and the Go-typed values are
// rebuilt here from their numeric constants.
func gpuAnalyze(src *image.NRGBA) (ImageStats, bool) {
	var st C.fb_image_stats
	rc := C.fb_analyze(pix(src), C.int(src.Stride), C.int(src.Bounds().Dx()), C.int(src.Bounds().Dy()), &st)
	if rc != C.FB_OK {
		return ImageStats{}, false
	}
	return ImageStats{
		Width: int(st.width), Height: int(st.height),
		HasAlpha: st.has_alpha != 0, IsGrayscale: st.is_grayscale != 0, UniqueColors: int(st.unique_colors),
		Entropy: float64(st.entropy), EdgeDensity: float64(st.edge_density),
		MeanBrightness: float64(st.mean_brightness), Contrast: float64(st.contrast),
		RecommendedFormat: Format(st.recommended_format), RecommendedQuality: Quality(st.recommended_quality),
		EstimatedCompression: float64(st.estimated_compression),
	}, true
}

// ---- SURVEY §8(f3, f4) ------------------------------------------------------------------------------------

// gpuApplyPalette replaces the pixel loops of applyPalette and palettedToNRGBA (targetsize.go:479-545).
// indexed.Palette must be what medianCut returned (color.NRGBA entries with A == 255).
func gpuApplyPalette(src *image.NRGBA, indexed *image.Paletted, recon *image.NRGBA) bool {
	n := len(indexed.Palette)
	if n == 0 || n > 256 {
		return false
	}
	pal := make([]byte, 4*n)
	for i, c := range indexed.Palette {
		v, ok := c.(color.NRGBA)
		if !ok || v.A != 255 {
			return false
		}
		pal[4*i], pal[4*i+1], pal[4*i+2], pal[4*i+3] = v.R, v.G, v.B, v.A
	}
	var dst *C.uint8_t
	dstStride := 0
	if recon != nil {
		dst, dstStride = pix(recon), recon.Stride
	}
	st := C.fb_apply_palette(pix(src), C.int(src.Stride), C.int(src.Bounds().Dx()), C.int(src.Bounds().Dy()),
		(*C.uint8_t)(unsafe.Pointer(&pal[0])), C.int(n), (*C.uint8_t)(unsafe.Pointer(&indexed.Pix[0])), C.int(indexed.Stride),
		dst, C.int(dstStride))
	return st == C.FB_OK
}

// gpuApplyOrientation replaces the rotate / flip loops behind ApplyOrientation (exif.go:176-203).
// The identity orientations never reach it (the switch returns img first).
func gpuApplyOrientation(img *image.NRGBA, orient Orientation) (*image.NRGBA, bool) {
	var dw, dh C.int
	if C.fb_orientation_dims(C.int(orient), C.int(img.Bounds().Dx()), C.int(img.Bounds().Dy()), &dw, &dh) != C.FB_OK {
		return nil, false
	}
	dst := image.NewNRGBA(image.Rect(0, 0, int(dw), int(dh)))
	st := C.fb_apply_orientation(pix(img), C.int(img.Stride), C.int(img.Bounds().Dx()), C.int(img.Bounds().Dy()),
		C.int(orient), pix(dst), C.int(dst.Stride))
	return dst, st == C.FB_OK
}

// test_host.cpp — the reference's own tests for the hot path (fennec_test.go), run through the C++ host mirror
// (include/fennec.hpp → libfennec_b200.so), plus exact comparisons against the CPU oracle (test-only linkage).
#include <cmath>
#include <cstdio>
#include <cstring>

#include "../../include/fennec.hpp"
#include "../../oracle/fennec_oracle.h"

using namespace fennec;

static int failures = 0;
#define EXPECT(cond)                                                        \
    do {                                                                    \
        if (!(cond)) { printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); failures++; } \
    } while (0)

static Image makeTestImage(int w, int h) {  // fennec_test.go:20-32
    Image img = NewNRGBA(w, h);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            uint8_t *p = img->at(x, y);
            p[0] = (uint8_t)(x * 255 / w); p[1] = (uint8_t)(y * 255 / h); p[2] = (uint8_t)((x + y) % 256); p[3] = 0xff;
        }
    return img;
}
static Image makeSolidImage(int w, int h, uint8_t r, uint8_t g, uint8_t b, uint8_t a) {  // fennec_test.go:45-54
    Image img = NewNRGBA(w, h);
    for (size_t i = 0; i < img->Pix.size(); i += 4) { img->Pix[i] = r; img->Pix[i + 1] = g; img->Pix[i + 2] = b; img->Pix[i + 3] = a; }
    return img;
}
static Image makeStripedImage(int w, int h, int sw) {  // fennec_test.go:58-76
    Image img = NewNRGBA(w, h);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            uint8_t *p = img->at(x, y);
            bool even = (x / sw) % 2 == 0;
            p[0] = even ? 200 : 50; p[1] = even ? 50 : 200; p[2] = 100; p[3] = 255;
        }
    return img;
}
static Image minusRed(const Image &src, int delta) {
    Image out = std::make_shared<NRGBA>(*src);
    for (size_t i = 0; i < out->Pix.size(); i += 4)
        if (out->Pix[i] > delta) out->Pix[i] -= (uint8_t)delta;
    return out;
}
static bool differs(const Image &a, const Image &b) { return a->Pix != b->Pix; }

int main() {
    fo_set_procs(8);
    // TestSSIMIdentical / Different / Similar / Fast / SmallImage (fennec_test.go:82-129)
    Image img = makeTestImage(100, 100);
    EXPECT(SSIM(img, img) >= 0.999);
    EXPECT(SSIM(makeSolidImage(100, 100, 0, 0, 0, 255), makeSolidImage(100, 100, 255, 255, 255, 255)) <= 0.1);
    Image mod = minusRed(img, 10);
    double s = SSIM(img, mod);
    EXPECT(s >= 0.85 && s <= 0.999);
    EXPECT(std::fabs(s - fo_ssim(img->data(), img->Stride, mod->data(), mod->Stride, 100, 100)) <= 1e-5);
    Image big = makeTestImage(500, 500);
    EXPECT(SSIMFast(big, big) >= 0.999);
    Image small = makeTestImage(4, 4);
    EXPECT(SSIM(small, small) >= 0.999);
    // TestMSSSIM* (fennec_test.go:131-163)
    Image m128 = makeTestImage(128, 128);
    EXPECT(MSSSIM(m128, m128) >= 0.99);
    EXPECT(MSSSIM(makeSolidImage(128, 128, 0, 0, 0, 255), makeSolidImage(128, 128, 255, 255, 255, 255)) <= 0.1);
    Image m5 = minusRed(m128, 5);
    double ms = MSSSIM(m128, m5);
    EXPECT(ms >= 0.7 && ms < 1.0);
    EXPECT(std::fabs(ms - fo_msssim(m128->data(), m128->Stride, m5->data(), m5->Stride, 128, 128)) <= 1e-5);
    // SSIM resizes a mismatched second image (ssim.go:31-33)
    EXPECT(SSIM(makeTestImage(120, 90), makeTestImage(60, 45)) > 0.5);
    // Lanczos / smartResize (fennec_test.go:510-560)
    Image src = makeTestImage(200, 100);
    Image half = lanczosResize(src, 100, 50);
    EXPECT(half->W == 100 && half->H == 50);
    EXPECT(lanczosResize(src, 0, 10)->W == 0 && lanczosResize(src, 0, 10)->H == 0);
    EXPECT(SSIM(src, lanczosResize(half, 200, 100)) >= 0.5);
    EXPECT(smartResize(src, 100, 100)->W == 100 && smartResize(src, 100, 100)->H == 50);
    EXPECT(smartResize(src, 400, 400) == src);
    {
        Image ref = NewNRGBA(100, 50);
        fo_lanczos_resize(src->data(), src->Stride, 200, 100, ref->data(), ref->Stride, 100, 50);
        EXPECT(half->Pix == ref->Pix);  // bit-exact
    }
    // Effects (fennec_test.go:612-736): changes, pointer identity, alpha passthrough
    Image st = makeStripedImage(100, 100, 10);
    EXPECT(differs(Sharpen(st, 0.8), st));
    EXPECT(Sharpen(img, 0.0) == img);
    EXPECT(Sharpen(st, 5.0)->Pix == Sharpen(st, 1.0)->Pix);
    Image tiny = makeTestImage(2, 2);
    EXPECT(Sharpen(tiny, 0.5) == tiny && AdaptiveSharpen(tiny, 0.5) == tiny);
    EXPECT(differs(AdaptiveSharpen(st, 0.5), st));
    EXPECT(AdaptiveSharpen(st, 0.0) == st);
    Image bl = GaussianBlur(img, 2.0);
    EXPECT(bl->W == 100 && bl->H == 100 && SSIM(img, bl) >= 0.3);
    EXPECT(GaussianBlur(img, 0.0) == img && GaussianBlur(img, -1.0) == img);
    EXPECT(SSIM(st, GaussianBlur(st, 20.0)) <= 0.999);
    {
        Image ref = NewNRGBA(100, 100), ref2 = NewNRGBA(100, 100), ref3 = NewNRGBA(100, 100);
        fo_gaussian_blur(img->data(), img->Stride, 100, 100, 2.0, ref->data(), ref->Stride);
        EXPECT(bl->Pix == ref->Pix);
        fo_sharpen(st->data(), st->Stride, 100, 100, 0.8, ref2->data(), ref2->Stride);
        EXPECT(Sharpen(st, 0.8)->Pix == ref2->Pix);
        fo_adaptive_sharpen(st->data(), st->Stride, 100, 100, 0.5, ref3->data(), ref3->Stride);
        EXPECT(AdaptiveSharpen(st, 0.5)->Pix == ref3->Pix);
    }
    // boxDownsample dims + exactness (fennec_test.go:1101-1115)
    Image bd = boxDownsample(src, 50, 25);
    EXPECT(bd->W == 50 && bd->H == 25);
    EXPECT(boxDownsample(src, 0, 5)->W == 0);
    {
        Image ref = NewNRGBA(50, 25);
        fo_box_downsample(src->data(), src->Stride, 200, 100, ref->data(), ref->Stride, 50, 25);
        EXPECT(bd->Pix == ref->Pix);
    }
    // ---- SURVEY §8(f1-f4) through the C++ mirror -------------------------------------------------------
    {   // TestAnalyze (fennec_test.go:564-610)
        fb_image_stats a200 = Analyze(makeTestImage(200, 200));
        EXPECT(a200.width == 200 && a200.height == 200 && !a200.has_alpha && a200.entropy >= 1);
        fb_image_stats solid = Analyze(makeSolidImage(100, 100, 128, 128, 128, 255));
        EXPECT(solid.is_grayscale && solid.entropy <= 0.01);
        fo_image_stats want;
        fo_analyze(src->data(), src->Stride, 200, 100, &want);
        fb_image_stats got = Analyze(src);
        EXPECT(got.unique_colors == want.unique_colors && got.has_alpha == want.has_alpha && got.is_grayscale == want.is_grayscale);
        EXPECT(std::fabs(got.entropy - want.entropy) <= 1e-9 && std::fabs(got.edge_density - want.edge_density) <= 1e-12);
        EXPECT(std::fabs(got.mean_brightness - want.mean_brightness) <= 1e-9 * want.mean_brightness);
        EXPECT(got.recommended_format == want.recommended_format && got.recommended_quality == want.recommended_quality);
    }
    {   // TestApplyOrientation (fennec_test.go:802-825)
        Image o = makeTestImage(100, 50);
        EXPECT(ApplyOrientation(o, 1) == o);
        Image r90 = ApplyOrientation(o, 6);
        EXPECT(r90->W == 50 && r90->H == 100);
        Image r180 = ApplyOrientation(o, 3);
        EXPECT(r180->W == 100 && r180->H == 50);
        for (int orient = 2; orient <= 8; orient++) {
            Image g = ApplyOrientation(o, orient);
            Image ref = NewNRGBA(g->W, g->H);
            EXPECT(fo_apply_orientation(o->data(), o->Stride, 100, 50, orient, ref->data(), ref->Stride) == 0);
            EXPECT(g->Pix == ref->Pix);
        }
    }
    {   // convertToNRGBA on a 4:2:0 image + the search session (compress.go:45-74)
        YCbCr yc;
        yc.W = 200; yc.H = 100; yc.YStride = 200; yc.CStride = 100; yc.SubsampleRatio = 2;
        yc.Y.resize(200 * 100); yc.Cb.resize(100 * 50); yc.Cr.resize(100 * 50);
        for (int y = 0; y < 100; y++)
            for (int x = 0; x < 200; x++) yc.Y[y * 200 + x] = (uint8_t)((0.299 * src->at(x, y)[0] + 0.587 * src->at(x, y)[1] + 0.114 * src->at(x, y)[2]) + 0.5);
        for (int y = 0; y < 50; y++)
            for (int x = 0; x < 100; x++) {
                const uint8_t *p = src->at(2 * x, 2 * y);
                yc.Cb[y * 100 + x] = (uint8_t)(128.0 - 0.168736 * p[0] - 0.331264 * p[1] + 0.5 * p[2] + 0.5);
                yc.Cr[y * 100 + x] = (uint8_t)(128.0 + 0.5 * p[0] - 0.418688 * p[1] - 0.081312 * p[2] + 0.5);
            }
        Image conv = convertToNRGBA(yc);
        Image ref = NewNRGBA(200, 100);
        EXPECT(fo_ycbcr_to_nrgba(yc.Y.data(), 200, yc.Cb.data(), yc.Cr.data(), 100, 200, 100, 2, ref->data(), ref->Stride) == 0);
        EXPECT(conv->Pix == ref->Pix);
        SSIMSession sess(src);
        double viaSession = sess.score(yc), direct = SSIMFast(src, conv);
        EXPECT(std::fabs(viaSession - direct) <= 2e-7 && viaSession > 0.5 && std::fabs(sess.score(conv) - direct) <= 2e-7);
    }
    {   // applyPalette + palettedToNRGBA (targetsize.go:479-545)
        std::vector<uint8_t> pal;
        for (int i = 0; i < 27; i++) { pal.push_back((uint8_t)((i % 3) * 120)); pal.push_back((uint8_t)(((i / 3) % 3) * 120)); pal.push_back((uint8_t)((i / 9) * 120)); pal.push_back(255); }
        Image recon;
        Paletted idx = applyPalette(src, pal, &recon);
        std::vector<uint8_t> wantIdx(200 * 100);
        Image wantRecon = NewNRGBA(200, 100);
        fo_apply_palette(src->data(), src->Stride, 200, 100, pal.data(), 27, wantIdx.data(), 200, wantRecon->data(), wantRecon->Stride);
        EXPECT(idx.Pix == wantIdx && recon->Pix == wantRecon->Pix);
    }
    // sharder keeps input order (batch.go:71,108)
    int covered = 0;
    for (int sh = 0; sh < 8; sh++) { ShardRange r = BatchShard(1024, 8, sh); EXPECT(r.begin == covered); covered = r.end; }
    EXPECT(covered == 1024);
    printf(failures ? "%d FAILURES\n" : "ALL OK (%d)\n", failures);
    return failures ? 1 : 0;
}

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
    // sharder keeps input order (batch.go:71,108)
    int covered = 0;
    for (int sh = 0; sh < 8; sh++) { ShardRange r = BatchShard(1024, 8, sh); EXPECT(r.begin == covered); covered = r.end; }
    EXPECT(covered == 1024);
    printf(failures ? "%d FAILURES\n" : "ALL OK (%d)\n", failures);
    return failures ? 1 : 0;
}

// api.cu — the C ABI of libfennec_b200.so (include/fennec_b200.h): lifecycle, per-thread contexts,
// host-buffer entry points (H2D → kernels → D2H, synchronous like the Go functions they replace),
// device-resident batch entry points, and the host-side table builders / dimension rules that sit on
// the host side of the boundary.  No CPU compute fallback lives here: without a GPU every compute
// entry point fails with FB_E_NOGPU.
#include "common.cuh"

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <utility>
#include <cmath>
#include <vector>

namespace fb {

// ---- errors ---------------------------------------------------------------------------------
static thread_local char t_err[512] = "";
thread_local long long t_launches = 0;

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof t_err, fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
    set_error("CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e), what, file, line);
    if (e == cudaErrorMemoryAllocation) return FB_E_OOM;
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return FB_E_NOGPU;
    return FB_E_CUDA;
}

// ---- global init ------------------------------------------------------------------------------
static std::mutex g_mu;
static bool g_inited = false;
static std::vector<int> g_devices;
static thread_local int t_device = 0;

static int init_locked(const int *devices, int n) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) {
        set_error("no usable CUDA device (cudaGetDeviceCount: %s); libfennec_b200 has no CPU fallback",
                  e == cudaSuccess ? "0 devices" : cudaGetErrorString(e));
        cudaGetLastError();
        return FB_E_NOGPU;
    }
    std::vector<int> devs;
    if (n <= 0 || devices == nullptr) {
        for (int i = 0; i < count; i++) devs.push_back(i);
    } else {
        for (int i = 0; i < n; i++) {
            if (devices[i] < 0 || devices[i] >= count) {
                set_error("fb_init: device %d out of range (0..%d)", devices[i], count - 1);
                return FB_E_INVALID;
            }
            devs.push_back(devices[i]);
        }
    }
    g_devices = devs;
    g_inited = true;
    return (int)g_devices.size();
}

int ensure_init() {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_inited) return (int)g_devices.size();
    return init_locked(nullptr, 0);
}

int device_count() {
    std::lock_guard<std::mutex> lk(g_mu);
    return g_inited ? (int)g_devices.size() : 0;
}

int current_device() { return t_device; }

// ---- per-thread contexts ------------------------------------------------------------------------
struct TableKey {
    int kind, a, b;
    bool operator<(const TableKey &o) const {
        if (kind != o.kind) return kind < o.kind;
        if (a != o.a) return a < o.a;
        return b < o.b;
    }
};
struct LanczosTable {
    int *start = nullptr;
    int *index = nullptr;
    double *weight = nullptr;
    float *weight32 = nullptr;
    int *first = nullptr;      // grouped layout for the horizontal fast path (nullptr: taps not contiguous)
    float *wpadT = nullptr;
    int groups = 0;
    int entries = 0, maxTaps = 0;
    double wabs = 0.0;  // max over destinations of sum |w| (error bound of the FP32 fast path)
    IntRatioInfo ir;    // ir.ratio >= 2 when the integer-ratio kernel applies
};

// Integer ratio (srcSize == R*dstSize): find the range of destinations whose taps are R*d + off with weights
// bit-identical to a reference interior destination (precomputeWeights gives that for every unclipped d).
static void detect_int_ratio(const int *start, const int *index, const double *weight, int dstSize, int srcSize,
                             IntRatioInfo *ir) {
    *ir = IntRatioInfo();
    if (dstSize <= 0 || srcSize % dstSize != 0) return;
    const int R = srcSize / dstSize;
    if (R < 2 || R > 4) return;
    const int mid = dstSize / 2;
    const int T = start[mid + 1] - start[mid];
    if (T <= 0 || T > 24) return;
    const int off = index[start[mid]] - R * mid;
    auto same = [&](int d) {
        if (start[d + 1] - start[d] != T) return false;
        for (int k = 0; k < T; k++) {
            if (index[start[d] + k] != R * d + off + k) return false;
            if ((float)weight[start[d] + k] != (float)weight[start[mid] + k]) return false;
        }
        return true;
    };
    int lo = mid, hi = mid + 1;
    while (lo > 0 && same(lo - 1)) lo--;
    while (hi < dstSize && same(hi)) hi++;
    if (hi - lo < 8) return;
    ir->ratio = R; ir->taps = T; ir->off = off; ir->dLo = lo; ir->dHi = hi;
    // binary64 rows of the interior destinations: bit-identical for an integer ratio (the tap offsets from the centre are the
    // same exact numbers for every d) — verified here, not assumed; then the exact path needs no CSR loads for them
    ir->wdExact = 1;
    for (int d = lo; d < hi && ir->wdExact; d++)
        for (int k = 0; k < T; k++)
            if (weight[start[d] + k] != weight[start[mid] + k]) { ir->wdExact = 0; break; }
    for (int k = 0; k < 24; k++) ir->wd[k] = k < T ? weight[start[mid] + k] : 0.0;
    float ws = 0.f;
    for (int k = 0; k < T; k++) { ir->w[k] = (float)weight[start[mid] + k]; ws += ir->w[k]; }
    ir->wsum = ws;
    // Fully opaque windows: the reference computes r = sum(R * (255 * w)), a = sum(255 * w), v = r * (1/a), all in
    // binary64 (resize.go:99-110) — sum(R * w) / W with W = sum(w), up to ~1e-13.  The kernel evaluates sum(R * wn) with
    // wn = fl32(w / W) as a chain of FP32 FMAs in tap order.  Bound: the weights are off by <= 2^-24 |w/W| each
    // (255 * 2^-24 * sum|wn| in the sum) and FMA k rounds a partial sum of magnitude <= 255 * P_k, P_k = sum_{s<=k} |wn_s|
    // (2^-24 each); 5 % margin plus 1e-6 for the binary64 roundings and the last-bit differences between the weight
    // rows of different destinations.  The alpha byte clampF(a) is the same for every such window unless a sits next to
    // a tie, in which case the shortcut stays off.
    double W = 0.0, a64 = 0.0;
    for (int k = 0; k < T; k++) { W += weight[start[mid] + k]; a64 += 255.0 * weight[start[mid] + k]; }
    for (int k = 0; k < 28; k++) ir->wn[k] = 0.f;
    ir->Eo = 0.f; ir->opaqueA = 0u;
    const double fracA = a64 - std::floor(a64);
    if (W > 0.5 && a64 > 1.0 && a64 < 1e6 && std::fabs(fracA - 0.5) > 1e-6) {
        double wabsn = 0.0, psum = 0.0, P = 0.0;
        for (int k = 0; k < T; k++) {
            ir->wn[k] = (float)(weight[start[mid] + k] / W);
            wabsn += std::fabs((double)ir->wn[k]);
        }
        for (int k = 0; k < T; k++) { P += std::fabs((double)ir->wn[k]); psum += P; }
        ir->Eo = (float)(255.0 * 5.9604644775390625e-08 * (wabsn + psum) * 1.05 + 1e-6);
        const double ra = std::floor(a64 + 0.5);   // clampF: half away from zero, a64 > 0
        const unsigned ab = ra >= 255.0 ? 255u : (unsigned)ra;
        if (ir->Eo < 0.05f && ab > 0u) ir->opaqueA = ab << 24;
    }
}

// Grouped layout (see resize.cu ResizeParams): returns false when some destination's taps are not contiguous.
static bool build_groups(const int *start, const int *index, const double *weight, int n, std::vector<int> &first,
                         std::vector<float> &wpadT, int *groups) {
    first.assign(n, 0);
    int G = 1;
    for (int d = 0; d < n; d++) {
        int cnt = start[d + 1] - start[d];
        if (cnt <= 0) { first[d] = 0; continue; }
        int f = index[start[d]];
        for (int k = 0; k < cnt; k++)
            if (index[start[d] + k] != f + k) return false;
        first[d] = f;
        int ng = ((f + cnt - 1) >> 2) - (f >> 2) + 1;
        if (ng > G) G = ng;
    }
    wpadT.assign((size_t)G * n * 4, 0.f);
    for (int d = 0; d < n; d++) {
        int cnt = start[d + 1] - start[d];
        int g0 = first[d] & ~3;
        for (int k = 0; k < cnt; k++) {
            int rel = first[d] + k - g0;
            wpadT[((size_t)(rel >> 2) * n + d) * 4 + (rel & 3)] = (float)weight[start[d] + k];
        }
    }
    *groups = G;
    return true;
}

// *wabs: the weight of the FP32 error bound (resize.cu launch_pass), max over destinations of 2 S + sum_k P_k with
// P_k = sum_{s<=k} |w_s| and S = P_last: FMA k rounds a partial sum of magnitude <= X * P_k (X = 255*255 or 255), and the
// FP32 weight and the product alpha*w add two relative roundings on every term.  [Round 1 used (taps + 3) * S: twice as
// wide for a Lanczos-3 row, i.e. twice as many outputs sent to the exact path.]
static void table_stats(const int *start, const double *weight, int n, int *maxTaps, double *wabs) {
    *maxTaps = 0;
    *wabs = 0.0;
    for (int d = 0; d < n; d++) {
        double sabs = 0.0, psum = 0.0;
        for (int t = start[d]; t < start[d + 1]; t++) { sabs += fabs(weight[t]); psum += sabs; }
        sabs = 2.0 * sabs + psum;
        if (sabs > *wabs) *wabs = sabs;
        if (start[d + 1] - start[d] > *maxTaps) *maxTaps = start[d + 1] - start[d];
    }
}
// Contexts (stream + arenas) live in a process-wide pool keyed by device.  A thread checks one out on first use and
// its thread_local destructor hands it back, so a worker that exits (a LockOSThread'ed goroutine ends its OS
// thread; CompressBatch starts fresh workers per call, batch.go:84-124) leaves its 100 MB-class arenas to the next
// worker instead of leaking them: the pool never holds more contexts than the peak number of concurrent threads.
struct CtxPool {
    std::mutex mu;
    std::vector<std::vector<DevCtx *>> idle;   // per device
};
static CtxPool g_pool;

static void destroy_ctx(DevCtx *c) {   // the device of *c is current
    cudaStreamSynchronize(c->stream);
    if (c->ws.base) cudaFree(c->ws.base);
    if (c->pin.base) cudaFreeHost(c->pin.base);
    if (c->stage) cudaFreeHost(c->stage);
    if (c->ev) cudaEventDestroy(c->ev);
    if (c->useEv) cudaEventDestroy(c->useEv);
    for (int i = 0; i < 2; i++) if (c->stageEv[i]) cudaEventDestroy(c->stageEv[i]);
    for (int i = 0; i < 4; i++) if (c->pipeEv[i]) cudaEventDestroy(c->pipeEv[i]);
    if (c->joinEv) cudaEventDestroy(c->joinEv);
    if (c->side) { cudaStreamSynchronize(c->side); cudaStreamDestroy(c->side); }
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

struct ThreadState {
    std::vector<DevCtx *> ctxs;
    ~ThreadState() {
        // No CUDA calls here (a thread destructor can race with runtime teardown): the contexts only change hands.
        std::lock_guard<std::mutex> lk(g_pool.mu);
        for (DevCtx *c : ctxs) {
            if (!c) continue;
            if ((int)g_pool.idle.size() <= c->dev) g_pool.idle.resize(c->dev + 1);
            g_pool.idle[c->dev].push_back(c);
        }
    }
};
static thread_local ThreadState t_state;

static int phys_device(int dev) {
    std::lock_guard<std::mutex> lk(g_mu);
    return g_devices[dev];
}

static DevCtx *ctx_phys(int phys);

DevCtx *ctx(int dev) {
    int nd = ensure_init();
    if (nd < 0) return nullptr;
    if (dev < 0 || dev >= nd) {
        set_error("device index %d out of range (fb_init selected %d device(s))", dev, nd);
        return nullptr;
    }
    // Contexts, the pool and the table cache are keyed by the PHYSICAL device, so a later fb_init with another
    // device list (which renumbers the logical indices) can never hand a thread a stream of the wrong GPU.
    return ctx_phys(phys_device(dev));
}

static DevCtx *ctx_phys(int phys) {
    if (phys < 0) return nullptr;
    if ((int)t_state.ctxs.size() <= phys) t_state.ctxs.resize(phys + 1, nullptr);
    if (cudaSetDevice(phys) != cudaSuccess) {
        cuda_fail(cudaGetLastError(), "cudaSetDevice", __FILE__, __LINE__);
        return nullptr;
    }
    DevCtx *c = t_state.ctxs[phys];
    if (c) return c;
    {
        std::lock_guard<std::mutex> lk(g_pool.mu);
        if ((int)g_pool.idle.size() > phys && !g_pool.idle[phys].empty()) {
            c = g_pool.idle[phys].back();
            g_pool.idle[phys].pop_back();
        }
    }
    if (!c) {
        c = new DevCtx();
        if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&c->ev, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&c->useEv, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&c->stageEv[0], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&c->stageEv[1], cudaEventDisableTiming) != cudaSuccess) {
            cuda_fail(cudaGetLastError(), "stream/event creation", __FILE__, __LINE__);
            destroy_ctx(c);
            return nullptr;
        }
        c->dev = phys;
    }
    t_state.ctxs[phys] = c;
    return c;
}

// Every entry point that touches a device opens one: it restores the caller's current CUDA device on return (the
// library selects its own; a *_batch_dev call for a tensor on cuda:1 must not leave the thread on cuda:1), and for
// device-resident calls it records, behind the enqueued work, the event that orders the next user of the arena.
struct ApiScope {
    int prev = -1;
    DevCtx *c = nullptr;
    cudaStream_t s = nullptr;
    ApiScope() {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); }
    }
    void bind(DevCtx *ctx_, cudaStream_t stream) { c = ctx_; s = stream; }
    ~ApiScope() {
        if (c) {
            if (cudaEventRecord(c->useEv, s) == cudaSuccess) { c->usePending = true; c->lastStream = s; }
            else cudaGetLastError();
        }
        if (prev >= 0) {
            int cur = -1;
            if (cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
        }
    }
};

int reserve(DevCtx *c, cudaStream_t s, size_t dev_bytes, size_t pinned_bytes) {
    // The pinned arena may still feed an async H2D enqueued by a previous _dev call.
    if (c->pinBusy) {
        FB_CUDA(cudaEventSynchronize(c->ev));
        c->pinBusy = false;
    }
    // The device arena may still be in use by work a previous _dev call enqueued on ANOTHER stream (host entry points
    // run on c->stream and synchronise before returning; _dev calls run on the caller's stream and do not).
    if (c->usePending && c->lastStream != s) FB_CUDA(cudaStreamWaitEvent(s, c->useEv, 0));
    if (c->usePending && c->lastStream != s) c->usePending = false;
    dev_bytes += 4096;
    pinned_bytes += 4096;
    if (c->ws.cap < dev_bytes) {
        FB_CUDA(cudaDeviceSynchronize());  // enqueued work may still use the old buffer
        if (c->ws.base) FB_CUDA(cudaFree(c->ws.base));
        c->ws.base = nullptr;
        c->ws.cap = 0;
        size_t want = dev_bytes + dev_bytes / 4;
        void *ptr = nullptr;
        cudaError_t e = cudaMalloc(&ptr, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            want = dev_bytes;
            e = cudaMalloc(&ptr, want);
        }
        if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(workspace)", __FILE__, __LINE__);
        c->ws.base = (char *)ptr;
        c->ws.cap = want;
    }
    if (c->pin.cap < pinned_bytes) {
        FB_CUDA(cudaStreamSynchronize(c->stream));
        if (c->pin.base) FB_CUDA(cudaFreeHost(c->pin.base));
        c->pin.base = nullptr;
        c->pin.cap = 0;
        void *ptr = nullptr;
        FB_CUDA(cudaMallocHost(&ptr, pinned_bytes * 2));
        c->pin.base = (char *)ptr;
        c->pin.cap = pinned_bytes * 2;
    }
    c->ws.reset();
    c->pin.reset();
    return FB_OK;
}

static void mark_pin_busy(DevCtx *c, cudaStream_t s) {
    cudaEventRecord(c->ev, s);
    c->pinBusy = true;
}

// ---- host-side rules and table builders (host side of the boundary) -----------------------------

// ssim.go:52-56
static int ssim_fast_dims(int w, int h, int *nw, int *nh) {
    const int maxDim = 512;
    if (w > maxDim || h > maxDim) {
        double scale = (double)maxDim / fmax((double)w, (double)h);
        *nw = (int)fmax(8.0, round((double)w * scale));
        *nh = (int)fmax(8.0, round((double)h * scale));
        return 1;
    }
    *nw = w;
    *nh = h;
    return 0;
}

// resize.go:55-69
static double lanczos_kernel(double x) {
    if (x == 0) return 1.0;
    if (x < 0) x = -x;
    if (x >= 3.0) return 0.0;
    double xpi = x * M_PI;
    return (3.0 * sin(xpi) * sin(xpi / 3.0)) / (xpi * xpi);
}

static void lanczos_geom(int dstSize, int srcSize, double *ratio, double *support, double *fscale) {
    *ratio = (double)srcSize / (double)dstSize;  // resize.go:81-85
    *support = 3.0;
    if (*ratio > 1) *support = 3.0 * *ratio;
    *fscale = fmax(*ratio, 1.0);  // resize.go:166
}

static void lanczos_span(int d, int srcSize, double ratio, double support, double *center, int *left, int *right) {
    *center = ((double)d + 0.5) * ratio - 0.5;  // resize.go:169-178
    *left = (int)ceil(*center - support);
    *right = (int)floor(*center + support);
    if (*left < 0) *left = 0;
    if (*right >= srcSize) *right = srcSize - 1;
}

static int lanczos_cap(int dstSize, int srcSize) {
    if (dstSize <= 0 || srcSize <= 0) return 0;
    double ratio, support, fscale, center;
    lanczos_geom(dstSize, srcSize, &ratio, &support, &fscale);
    long long cap = 0;
    for (int d = 0; d < dstSize; d++) {
        int l, r;
        lanczos_span(d, srcSize, ratio, support, &center, &l, &r);
        if (r >= l) cap += r - l + 1;
    }
    return cap > 0x7fffffff ? 0x7fffffff : (int)cap;
}

// resize.go:164-197
static int lanczos_build(int dstSize, int srcSize, int *start, int *index, double *weight) {
    double ratio, support, fscale, center;
    lanczos_geom(dstSize, srcSize, &ratio, &support, &fscale);
    int n = 0;
    for (int d = 0; d < dstSize; d++) {
        int l, r;
        lanczos_span(d, srcSize, ratio, support, &center, &l, &r);
        start[d] = n;
        double wsum = 0.0;
        int first = n;
        for (int s = l; s <= r; s++) {
            double w = lanczos_kernel(((double)s - center) / fscale);
            if (w != 0) {
                wsum += w;
                index[n] = s;
                weight[n] = w;
                n++;
            }
        }
        if (wsum != 0)
            for (int i = first; i < n; i++) weight[i] /= wsum;
    }
    start[dstSize] = n;
    return n;
}

// Library-built CSR tables live in persistent device memory, shared by every thread (they are immutable once
// uploaded) and bounded: at most kLanczosCacheMax (dstSize, srcSize) pairs per process, least recently used first out.
// Eviction synchronises the device before freeing (a kernel enqueued by another thread may still read the table);
// it only happens when a service cycles through more than kLanczosCacheMax distinct resize geometries.
// Caller-supplied tables are uploaded per call into the arena instead.
constexpr size_t kLanczosCacheMax = 64;
struct CachedTable { LanczosTable t; unsigned long long stamp; };
static std::mutex g_tab_mu;
static std::map<std::pair<int, TableKey>, CachedTable> g_lanczos;   // (device, key) → device tables
static unsigned long long g_tab_clock = 0;

static void free_table(LanczosTable &t) {
    cudaFree(t.start); cudaFree(t.index); cudaFree(t.weight); cudaFree(t.weight32); cudaFree(t.first); cudaFree(t.wpadT);
    t = LanczosTable();
}

static int lanczos_table_upload(int dstSize, int srcSize, LanczosTable *out) {
    int cap = lanczos_cap(dstSize, srcSize);
    std::vector<int> start(dstSize + 1), index(cap + 1);
    std::vector<double> weight(cap + 1);
    int n = lanczos_build(dstSize, srcSize, start.data(), index.data(), weight.data());
    LanczosTable &t = *out;
    t = LanczosTable();
    t.entries = n;
    table_stats(start.data(), weight.data(), dstSize, &t.maxTaps, &t.wabs);
    detect_int_ratio(start.data(), index.data(), weight.data(), dstSize, srcSize, &t.ir);
    std::vector<float> w32(n + 1);
    for (int i = 0; i < n; i++) w32[i] = (float)weight[i];
    FB_CUDA(cudaMalloc((void **)&t.start, sizeof(int) * (dstSize + 1)));
    FB_CUDA(cudaMalloc((void **)&t.index, sizeof(int) * (n + 1)));
    FB_CUDA(cudaMalloc((void **)&t.weight, sizeof(double) * (n + 1)));
    FB_CUDA(cudaMalloc((void **)&t.weight32, sizeof(float) * (n + 1)));
    FB_CUDA(cudaMemcpy(t.start, start.data(), sizeof(int) * (dstSize + 1), cudaMemcpyHostToDevice));
    FB_CUDA(cudaMemcpy(t.index, index.data(), sizeof(int) * n, cudaMemcpyHostToDevice));
    FB_CUDA(cudaMemcpy(t.weight, weight.data(), sizeof(double) * n, cudaMemcpyHostToDevice));
    FB_CUDA(cudaMemcpy(t.weight32, w32.data(), sizeof(float) * n, cudaMemcpyHostToDevice));
    std::vector<int> first;
    std::vector<float> wpadT;
    if (build_groups(start.data(), index.data(), weight.data(), dstSize, first, wpadT, &t.groups)) {
        FB_CUDA(cudaMalloc((void **)&t.first, sizeof(int) * dstSize));
        FB_CUDA(cudaMalloc((void **)&t.wpadT, sizeof(float) * wpadT.size()));
        FB_CUDA(cudaMemcpy(t.first, first.data(), sizeof(int) * dstSize, cudaMemcpyHostToDevice));
        FB_CUDA(cudaMemcpy(t.wpadT, wpadT.data(), sizeof(float) * wpadT.size(), cudaMemcpyHostToDevice));
    }
    return FB_OK;
}

static int lanczos_table_cached(DevCtx *c, int dstSize, int srcSize, LanczosTable *out) {
    std::pair<int, TableKey> key{c->dev, TableKey{1, dstSize, srcSize}};
    std::lock_guard<std::mutex> lk(g_tab_mu);   // the device of *c is current (ctx() selected it)
    auto it = g_lanczos.find(key);
    if (it != g_lanczos.end()) {
        it->second.stamp = ++g_tab_clock;
        *out = it->second.t;
        return FB_OK;
    }
    if (g_lanczos.size() >= kLanczosCacheMax) {
        auto victim = g_lanczos.end();
        for (auto j = g_lanczos.begin(); j != g_lanczos.end(); ++j)
            if (j->first.first == c->dev && (victim == g_lanczos.end() || j->second.stamp < victim->second.stamp)) victim = j;
        if (victim != g_lanczos.end()) {
            FB_CUDA(cudaDeviceSynchronize());
            free_table(victim->second.t);
            g_lanczos.erase(victim);
        }
    }
    LanczosTable t;
    int rc = lanczos_table_upload(dstSize, srcSize, &t);
    if (rc < 0) {   // partial allocations of a failed upload do not stay behind
        free_table(t);
        return rc;
    }
    g_lanczos[key] = CachedTable{t, ++g_tab_clock};
    *out = t;
    return FB_OK;
}

// effects.go:153-165
static int blur_kernel_host(double sigma, std::vector<double> &k) {
    int radius = (int)ceil(sigma * 3);
    int size = radius * 2 + 1;
    k.resize(size);
    double sum = 0.0;
    for (int i = 0; i < size; i++) {
        double x = (double)(i - radius);
        k[i] = exp(-(x * x) / (2 * sigma * sigma));
        sum += k[i];
    }
    for (int i = 0; i < size; i++) k[i] /= sum;
    return radius;
}

// MSSSIM level plan (ssim.go:324-362): which levels are scored, their dims and weights.
struct MsLevel { int w, h; double weight; };
static std::vector<MsLevel> msssim_plan(int w0, int h0) {
    double weights[5] = {0.0448, 0.2856, 0.3001, 0.2363, 0.1333};
    int nweights = 5, w = w0, h = h0;
    for (int i = 0; i < 4; i++) {
        int minDim = (int)fmin((double)w, (double)h);
        if (minDim < 8) {
            nweights = i + 1;
            double sum = 0.0;
            for (int j = 0; j < nweights; j++) sum += weights[j];
            for (int j = 0; j < nweights; j++) weights[j] /= sum;
            break;
        }
        w /= 2;
        h /= 2;
    }
    std::vector<MsLevel> plan;
    int cw = w0, ch = h0;
    for (int i = 0; i < nweights; i++) {
        plan.push_back({cw, ch, weights[i]});
        if (i < nweights - 1) {
            int nw = cw / 2, nh = ch / 2;
            if (nw < 8 || nh < 8) break;  // ssim.go:356-358 — later weights are simply never used
            cw = nw;
            ch = nh;
        }
    }
    return plan;
}

// ---- device pipelines (shared by host and _dev entry points) --------------------------------------

struct ImgBatch {  // n images of identical dims in device memory
    const uint8_t *p;
    long long imgStride;
    int rowStride;
};

static size_t ssim_fast_scratch(int w, int h, int n) {
    int nw, nh;
    size_t bytes = 0;
    if (ssim_fast_dims(w, h, &nw, &nh)) bytes += 2 * align_up((size_t)dev_pitch(nw) * nh * n, 256);
    bytes += ssim_scratch_bytes(nw, nh, n);
    return bytes + 1024;
}

// SSIMFast on device images: scores[i*scoreStride] (ssim.go:48-70).
static int pipeline_ssim_fast(DevCtx *c, cudaStream_t s, ImgBatch a, ImgBatch b, int w, int h, int n,
                              double *scores, long long scoreStride) {
    int nw, nh;
    if (ssim_fast_dims(w, h, &nw, &nh)) {
        int pitch = dev_pitch(nw);
        long long imgBytes = (long long)pitch * nh;
        uint8_t *da = (uint8_t *)c->ws.take((size_t)imgBytes * n);
        uint8_t *db = (uint8_t *)c->ws.take((size_t)imgBytes * n);
        if (!da || !db) { set_error("internal: workspace under-reserved (ssim_fast)"); return FB_E_INVALID; }
        FB_TRY(launch_box_pair(s, a.p, a.imgStride, a.rowStride, b.p, b.imgStride, b.rowStride, w, h, da, db, imgBytes, pitch, nw, nh, n));
        a = ImgBatch{da, imgBytes, pitch};
        b = ImgBatch{db, imgBytes, pitch};
        w = nw;
        h = nh;
    }
    void *scratch = c->ws.take(ssim_scratch_bytes(w, h, n));
    if (!scratch) { set_error("internal: workspace under-reserved (ssim)"); return FB_E_INVALID; }
    return launch_ssim(c, s, a.p, b.p, a.imgStride, b.imgStride, a.rowStride, b.rowStride, w, h, n, scores,
                       scoreStride, scratch);
}

// Level l can take the fused step (thumbnail + next level from one read) as far as dims go; *tw/*th = thumbnail dims.
static bool msssim_level_fusable(const std::vector<MsLevel> &plan, int l, int *tw, int *th) {
    const int L = (int)plan.size();
    return l + 1 < L && ssim_fast_dims(plan[l].w, plan[l].h, tw, th) && plan[l + 1].w * 2 == plan[l].w && plan[l + 1].h * 2 == plan[l].h;
}
// Consecutive fusable levels from l on that share l's thumbnail dims: what one batched K1 launch can score.
static int msssim_run_capacity(const std::vector<MsLevel> &plan, int l) {
    int tw0, th0, tw, th, k = 0;
    if (!msssim_level_fusable(plan, l, &tw0, &th0)) return 0;
    while (l + k < (int)plan.size() && msssim_level_fusable(plan, l + k, &tw, &th) && tw == tw0 && th == th0) k++;
    return k;
}

static size_t msssim_scratch(int w, int h, int n) {
    std::vector<MsLevel> plan = msssim_plan(w, h);
    const size_t L = plan.size();
    size_t bytes = 0, fast = 0;
    for (size_t l = 0; l < L; l++) {
        if (l > 0) bytes += 2 * align_up((size_t)dev_pitch(plan[l].w) * plan[l].h * n, 256);
        fast = std::max(fast, ssim_fast_scratch(plan[l].w, plan[l].h, n));
    }
    // the arena is bump-only within a call.  A run of fused levels reserves its thumbnails when it opens and scores
    // them with one batched K1 launch; if the fused kernel declines a level (unaligned caller buffers at level 0),
    // the next level opens a new run, so every level's capacity is budgeted once.
    bytes += fast * L;
    for (size_t l = 0; l < L; l++) {
        int tw, th;
        const int cap = msssim_run_capacity(plan, (int)l);
        if (cap > 0 && msssim_level_fusable(plan, (int)l, &tw, &th))
            bytes += 2 * align_up((size_t)dev_pitch(tw) * th * n * cap, 256) + ssim_scratch_bytes(tw, th, n * cap) + 1024;
    }
    bytes += align_up(sizeof(double) * (size_t)n * 5, 256) * 2 + 1024;
    return bytes;
}

// MSSSIM on device images (ssim.go:313-365); weights table uploaded through the pinned arena.
static int pipeline_msssim(DevCtx *c, cudaStream_t s, ImgBatch a, ImgBatch b, int w, int h, int n, double *out) {
    std::vector<MsLevel> plan = msssim_plan(w, h);
    const int L = (int)plan.size();
    double *levelScores = (double *)c->ws.take(sizeof(double) * (size_t)n * L);   // level-major: score(l, i) at [l*n + i]
    double *wdev = (double *)c->ws.take(sizeof(double) * 8);
    double *wpin = (double *)c->pin.take(sizeof(double) * 8);
    if (!levelScores || !wdev || !wpin) { set_error("internal: workspace under-reserved (msssim)"); return FB_E_INVALID; }
    for (int l = 0; l < L; l++) wpin[l] = plan[l].weight;
    FB_CUDA(cudaMemcpyAsync(wdev, wpin, sizeof(double) * L, cudaMemcpyHostToDevice, s));
    mark_pin_busy(c, s);
    // Thumbnails of consecutive fused levels that share dims (8K: levels 0-3 are all 512x288) are kept side by side
    // and scored by ONE K1 launch at the end of the run instead of one small launch per level.
    struct Run { int l0 = 0, count = 0, tw = 0, th = 0, tpitch = 0; long long tBytes = 0; uint8_t *ta = nullptr, *tb = nullptr; } run;
    auto flush_run = [&]() -> int {
        if (run.count == 0) return FB_OK;
        const int pairs = run.count * n;
        void *scratch = c->ws.take(ssim_scratch_bytes(run.tw, run.th, pairs));
        if (!scratch) { set_error("internal: workspace under-reserved (ssim batch)"); return FB_E_INVALID; }
        int rc = launch_ssim(c, s, run.ta, run.tb, run.tBytes, run.tBytes, run.tpitch, run.tpitch, run.tw, run.th, pairs,
                             levelScores + (size_t)run.l0 * n, 1, scratch);
        run.count = 0;
        return rc;
    };
    ImgBatch ca = a, cb = b;
    auto open_run = [&](int l, int tw, int th) -> int {   // room for the thumbnails of the levels the run can cover
        const int cap = msssim_run_capacity(plan, l);
        run.l0 = l; run.tw = tw; run.th = th; run.tpitch = dev_pitch(tw);
        run.tBytes = (long long)run.tpitch * th;
        run.ta = (uint8_t *)c->ws.take((size_t)run.tBytes * n * cap);
        run.tb = (uint8_t *)c->ws.take((size_t)run.tBytes * n * cap);
        if (!run.ta || !run.tb) { set_error("internal: workspace under-reserved (msssim thumbs)"); return FB_E_INVALID; }
        return FB_OK;
    };
    auto level_images = [&](int l, ImgBatch *ia, ImgBatch *ib) -> int {
        int pitch = dev_pitch(plan[l].w);
        long long imgBytes = (long long)pitch * plan[l].h;
        uint8_t *pa = (uint8_t *)c->ws.take((size_t)imgBytes * n);
        uint8_t *pb = (uint8_t *)c->ws.take((size_t)imgBytes * n);
        if (!pa || !pb) { set_error("internal: workspace under-reserved (msssim level)"); return FB_E_INVALID; }
        *ia = ImgBatch{pa, imgBytes, pitch};
        *ib = ImgBatch{pb, imgBytes, pitch};
        return FB_OK;
    };
    for (int l = 0; l < L; l++) {
        const int lw = plan[l].w, lh = plan[l].h;
        int tw, th, tw1, th1;
        // Two levels from one read (box.cu: box_fused2_kernel): levels l and l+1 both need a thumbnail and level l+2
        // exists; level l+1 then never exists in HBM.
        // (Levels below ~4 MP stay on the single-level step: per CTA the two thumbnails' outputs then weigh more than the
        // read they save — 8K pairs: 0.178 ms for the fused 1920x1080 + 960x540 step against ~0.10 ms for the two single ones.)
        static const long long f2MinPixels = [] { const char *e = getenv("FB_F2_MINPX"); return e ? atoll(e) : 4000000LL; }();
        if ((long long)lw * lh >= f2MinPixels && msssim_level_fusable(plan, l, &tw, &th) && msssim_level_fusable(plan, l + 1, &tw1, &th1) &&
            tw1 == tw && th1 == th) {
            if (run.count > 0 && (run.tw != tw || run.th != th)) FB_TRY(flush_run());
            const size_t mark = c->ws.off;
            if (run.count == 0) FB_TRY(open_run(l, tw, th));
            ImgBatch l2a, l2b;
            FB_TRY(level_images(l + 2, &l2a, &l2b));
            uint8_t *t0a = run.ta + (size_t)run.tBytes * n * run.count, *t0b = run.tb + (size_t)run.tBytes * n * run.count;
            uint8_t *t1a = t0a + (size_t)run.tBytes * n, *t1b = t0b + (size_t)run.tBytes * n;
            int rc = launch_box_fused2(s, ca.p, ca.imgStride, ca.rowStride, cb.p, cb.imgStride, cb.rowStride, lw, lh, t0a, t0b,
                                       run.tBytes, run.tpitch, tw, th, t1a, t1b, run.tBytes, run.tpitch, tw1, th1,
                                       (uint8_t *)l2a.p, (uint8_t *)l2b.p, l2a.imgStride, l2a.rowStride, n);
            if (rc < 0) return rc;
            if (rc == FB_OK) {
                run.count += 2;
                ca = l2a;
                cb = l2b;
                l += 1;     // level l+1 is done too
                continue;
            }
            c->ws.off = mark;   // nothing launched: give back what this attempt reserved (a run opened here has count 0 and is re-opened below)
        }
        ImgBatch na{nullptr, 0, 0}, nb{nullptr, 0, 0};
        if (l + 1 < L) FB_TRY(level_images(l + 1, &na, &nb));   // 2x box cascade (ssim.go:354-360)
        // Fused level step: when this level needs both a thumbnail (SSIMFast, > 512 px) and the next level's
        // image, one kernel reads it once and writes both (box.cu: box_fused_kernel).
        bool fused = false;
        if (msssim_level_fusable(plan, l, &tw, &th)) {
            if (run.count > 0 && (run.tw != tw || run.th != th)) FB_TRY(flush_run());
            if (run.count == 0) FB_TRY(open_run(l, tw, th));
            uint8_t *ta = run.ta + (size_t)run.tBytes * n * run.count, *tb = run.tb + (size_t)run.tBytes * n * run.count;
            int rc = launch_box_fused(s, ca.p, ca.imgStride, ca.rowStride, cb.p, cb.imgStride, cb.rowStride, lw, lh, ta, tb,
                                      run.tBytes, run.tpitch, tw, th, (uint8_t *)na.p, (uint8_t *)nb.p, na.imgStride, na.rowStride, n);
            if (rc < 0) return rc;
            if (rc == FB_OK) {
                run.count++;
                fused = true;
            }
        }
        if (!fused) {
            FB_TRY(flush_run());
            FB_TRY(pipeline_ssim_fast(c, s, ca, cb, lw, lh, n, levelScores + (size_t)l * n, 1));
            if (l + 1 < L) {
                FB_TRY(launch_box(s, ca.p, ca.imgStride, ca.rowStride, lw, lh, (uint8_t *)na.p, na.imgStride, na.rowStride, plan[l + 1].w, plan[l + 1].h, n, nullptr));
                FB_TRY(launch_box(s, cb.p, cb.imgStride, cb.rowStride, lw, lh, (uint8_t *)nb.p, nb.imgStride, nb.rowStride, plan[l + 1].w, plan[l + 1].h, n, nullptr));
            }
        }
        ca = na;
        cb = nb;
    }
    FB_TRY(flush_run());
    return launch_msssim_combine(s, levelScores, L, n, wdev, out);
}

// ---- host-buffer plumbing ---------------------------------------------------------------------------

static int check_img(const char *fn, const void *p, int stride, int w, int h) {
    if (w < 0 || h < 0) { set_error("%s: negative dimensions %dx%d", fn, w, h); return FB_E_INVALID; }
    if (w > 0 && h > 0) {
        if (!p) { set_error("%s: null pixel pointer", fn); return FB_E_INVALID; }
        if (stride < w * 4) { set_error("%s: stride %d < 4*w (%d)", fn, stride, w * 4); return FB_E_INVALID; }
    }
    return FB_OK;
}

// Pageable caller memory (a Go Pix slice, a numpy array) cannot be DMA'ed directly: cudaMemcpyAsync stages it
// inside the driver, one copy at a time per process (measured: 12 GB/s, and concurrent callers serialise).  The library
// stages it itself instead — rows are packed into one of two pinned chunks by the calling thread while the previous
// chunk's DMA is in flight — so every caller thread overlaps its own memcpy with its own DMA and N threads stage in
// parallel.  Pinned or registered caller memory (fb_alloc_pinned) takes the direct path.
constexpr size_t kStageChunk = 4u << 20;

static bool host_is_pageable(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

static int stage_acquire(DevCtx *c, int *slot, char **buf) {
    if (!c->stage) {
        void *p = nullptr;
        FB_CUDA(cudaMallocHost(&p, 2 * kStageChunk));
        c->stage = (char *)p;
    }
    const int k = (int)(c->stageNext++ & 1);
    if (c->stageBusy[k]) {
        FB_CUDA(cudaEventSynchronize(c->stageEv[k]));
        c->stageBusy[k] = false;
    }
    *slot = k;
    *buf = c->stage + (size_t)k * kStageChunk;
    return FB_OK;
}

// rows x rowBytes from host (stride hostStride) to device (pitch devPitch) on c->stream.
static int upload_rows(DevCtx *c, const uint8_t *host, size_t hostStride, uint8_t *dev, size_t devPitch, size_t rowBytes,
                       int rows) {
    static const bool noStage = getenv("FB_NO_STAGING") != nullptr;
    if (rows <= 0 || rowBytes == 0) return FB_OK;
    if (noStage || rowBytes > kStageChunk || !host_is_pageable(host)) {
        FB_CUDA(cudaMemcpy2DAsync(dev, devPitch, host, hostStride, rowBytes, rows, cudaMemcpyHostToDevice, c->stream));
        return FB_OK;
    }
    const int per = (int)(kStageChunk / rowBytes);
    for (int r0 = 0; r0 < rows; r0 += per) {
        const int nr = rows - r0 < per ? rows - r0 : per;
        int slot;
        char *buf;
        FB_TRY(stage_acquire(c, &slot, &buf));
        const uint8_t *src = host + (size_t)r0 * hostStride;
        if (hostStride == rowBytes) memcpy(buf, src, rowBytes * nr);
        else for (int r = 0; r < nr; r++) memcpy(buf + (size_t)r * rowBytes, src + (size_t)r * hostStride, rowBytes);
        FB_CUDA(cudaMemcpy2DAsync(dev + (size_t)r0 * devPitch, devPitch, buf, rowBytes, rowBytes, nr, cudaMemcpyHostToDevice, c->stream));
        FB_CUDA(cudaEventRecord(c->stageEv[slot], c->stream));
        c->stageBusy[slot] = true;
    }
    return FB_OK;
}

// device → host; returns after the caller's buffer holds the data when it was staged (pageable), otherwise the copy is
// only enqueued and the entry point's cudaStreamSynchronize completes it.
static int download_rows(DevCtx *c, const uint8_t *dev, size_t devPitch, uint8_t *host, size_t hostStride, size_t rowBytes,
                         int rows) {
    static const bool noStage = getenv("FB_NO_STAGING") != nullptr;
    if (rows <= 0 || rowBytes == 0) return FB_OK;
    if (noStage || rowBytes > kStageChunk || !host_is_pageable(host)) {
        FB_CUDA(cudaMemcpy2DAsync(host, hostStride, dev, devPitch, rowBytes, rows, cudaMemcpyDeviceToHost, c->stream));
        return FB_OK;
    }
    const int per = (int)(kStageChunk / rowBytes);
    int pendSlot = -1, pendR0 = 0, pendN = 0;
    char *pendBuf = nullptr;
    auto drain = [&]() -> int {
        if (pendSlot < 0) return FB_OK;
        FB_CUDA(cudaEventSynchronize(c->stageEv[pendSlot]));
        c->stageBusy[pendSlot] = false;
        uint8_t *dst = host + (size_t)pendR0 * hostStride;
        if (hostStride == rowBytes) memcpy(dst, pendBuf, rowBytes * pendN);
        else for (int r = 0; r < pendN; r++) memcpy(dst + (size_t)r * hostStride, pendBuf + (size_t)r * rowBytes, rowBytes);
        pendSlot = -1;
        return FB_OK;
    };
    for (int r0 = 0; r0 < rows; r0 += per) {
        const int nr = rows - r0 < per ? rows - r0 : per;
        int slot;
        char *buf;
        FB_TRY(stage_acquire(c, &slot, &buf));
        FB_CUDA(cudaMemcpy2DAsync(buf, rowBytes, dev + (size_t)r0 * devPitch, devPitch, rowBytes, nr, cudaMemcpyDeviceToHost, c->stream));
        FB_CUDA(cudaEventRecord(c->stageEv[slot], c->stream));
        c->stageBusy[slot] = true;
        FB_TRY(drain());   // the previous chunk is copied out while this one is in flight
        pendSlot = slot; pendR0 = r0; pendN = nr; pendBuf = buf;
    }
    return drain();
}

static int upload(DevCtx *c, const uint8_t *host, int stride, int w, int h, uint8_t **dev, int *pitch) {
    *pitch = dev_pitch(w);
    *dev = (uint8_t *)c->ws.take((size_t)*pitch * h + 16);
    if (!*dev) { set_error("internal: workspace under-reserved (upload)"); return FB_E_INVALID; }
    return upload_rows(c, host, (size_t)stride, *dev, (size_t)*pitch, (size_t)w * 4, h);
}

static int download(DevCtx *c, const uint8_t *dev, int pitch, uint8_t *host, int stride, int w, int h) {
    return download_rows(c, dev, (size_t)pitch, host, (size_t)stride, (size_t)w * 4, h);
}

static int finish_score(DevCtx *c, const double *dscore, double *out) {
    double *pin = (double *)c->pin.take(sizeof(double));
    if (!pin) { set_error("internal: pinned arena under-reserved"); return FB_E_INVALID; }
    FB_CUDA(cudaMemcpyAsync(pin, dscore, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    FB_CUDA(cudaStreamSynchronize(c->stream));
    *out = *pin;
    return FB_OK;
}

enum ScoreOp { OP_SSIM, OP_SSIM_FAST, OP_MSSSIM, OP_PIXEL };

static int host_score(const char *fn, ScoreOp op, const uint8_t *a, int strideA, const uint8_t *b, int strideB,
                      int w, int h, double *out) {
    if (!out) { set_error("%s: null output pointer", fn); return FB_E_INVALID; }
    FB_TRY(check_img(fn, a, strideA, w, h));
    FB_TRY(check_img(fn, b, strideB, w, h));
    if (w == 0 || h == 0) {  // pixelSSIM with n == 0 (ssim.go:172-175); MSSSIM: exp(1.0 * ln(1)) = 1
        *out = 1.0;
        return FB_OK;
    }
    ApiScope scope_;
    DevCtx *c = ctx(current_device());
    if (!c) return ensure_init() < 0 ? FB_E_NOGPU : FB_E_CUDA;
    size_t img = (size_t)dev_pitch(w) * h + 512;
    size_t need = 2 * img + 1024;
    if (op == OP_SSIM || op == OP_PIXEL) need += ssim_scratch_bytes(w, h, 1);
    if (op == OP_SSIM_FAST) need += ssim_fast_scratch(w, h, 1);
    if (op == OP_MSSSIM) need += msssim_scratch(w, h, 1);
    FB_TRY(reserve(c, c->stream, need, 256));
    uint8_t *da, *db;
    int pa, pb;
    FB_TRY(upload(c, a, strideA, w, h, &da, &pa));
    FB_TRY(upload(c, b, strideB, w, h, &db, &pb));
    double *dscore = (double *)c->ws.take(sizeof(double) * 2);
    ImgBatch A{da, 0, pa}, B{db, 0, pb};
    switch (op) {
        case OP_SSIM: {
            void *scratch = c->ws.take(ssim_scratch_bytes(w, h, 1));
            FB_TRY(launch_ssim(c, c->stream, da, db, 0, 0, pa, pb, w, h, 1, dscore, 1, scratch));
            break;
        }
        case OP_PIXEL: {
            // force the global-statistics path regardless of size
            void *scratch = c->ws.take(256);
            (void)scratch;
            FB_TRY(launch_pixel_ssim(c->stream, da, db, 0, 0, pa, pb, w, h, 1, dscore, 1));
            break;
        }
        case OP_SSIM_FAST:
            FB_TRY(pipeline_ssim_fast(c, c->stream, A, B, w, h, 1, dscore, 1));
            break;
        case OP_MSSSIM:
            FB_TRY(pipeline_msssim(c, c->stream, A, B, w, h, 1, dscore));
            break;
    }
    return finish_score(c, dscore, out);
}

static int dev_ctx_for(const char *fn, int device, DevCtx **out) {
    DevCtx *c = ctx(device);
    if (!c) return ensure_init() < 0 ? FB_E_NOGPU : FB_E_INVALID;
    (void)fn;
    *out = c;
    return FB_OK;
}

}  // namespace fb

using namespace fb;

// =====================================================================================================
// C ABI
// =====================================================================================================
extern "C" {

int fb_init(const int *devices, int n) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_inited && n <= 0) return (int)g_devices.size();
    return init_locked(devices, n);
}

void fb_shutdown(void) {
    // Releases the calling thread's contexts, every pooled (idle) context and the shared tables.  Contexts still
    // checked out by other live threads go with the process (call fb_shutdown after the workers have finished).
    std::vector<DevCtx *> victims;
    for (auto &c : t_state.ctxs) if (c) victims.push_back(c);
    t_state.ctxs.clear();
    {
        std::lock_guard<std::mutex> lk(g_pool.mu);
        for (auto &v : g_pool.idle) { victims.insert(victims.end(), v.begin(), v.end()); v.clear(); }
    }
    int prev = -1;
    if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); }
    for (DevCtx *c : victims) {
        cudaSetDevice(c->dev);   // physical id
        destroy_ctx(c);
    }
    {
        std::lock_guard<std::mutex> lk(g_tab_mu);
        for (auto &kv : g_lanczos) {
            cudaSetDevice(kv.first.first);   // physical id
            cudaDeviceSynchronize();
            free_table(kv.second.t);
        }
        g_lanczos.clear();
    }
    if (prev >= 0) cudaSetDevice(prev);
    std::lock_guard<std::mutex> lk(g_mu);
    g_inited = false;
    g_devices.clear();
}

// Pinned host memory for callers that want their pixel buffers DMA-able (a Go caller can wrap it with unsafe.Slice and
// decode into it): uploads from such buffers skip the staging copy that pageable memory needs.
void *fb_alloc_pinned(size_t bytes) {
    if (ensure_init() < 0) return nullptr;
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        cuda_fail(cudaGetLastError(), "cudaMallocHost", __FILE__, __LINE__);
        return nullptr;
    }
    return p;
}
void fb_free_pinned(void *p) {
    if (p) cudaFreeHost(p);
}

// Idle contexts held by the pool / live Lanczos tables: what tests/test_batch_host_gpu.py watches.
int fb_debug_pool_size(void) {
    std::lock_guard<std::mutex> lk(g_pool.mu);
    int n = 0;
    for (auto &v : g_pool.idle) n += (int)v.size();
    return n;
}
int fb_debug_table_count(void) {
    std::lock_guard<std::mutex> lk(g_tab_mu);
    return (int)g_lanczos.size();
}

int fb_device_count(void) {
    int n = ensure_init();
    return n < 0 ? 0 : n;
}

int fb_set_device(int device) {
    int n = ensure_init();
    if (n < 0) return n;
    if (device < 0 || device >= n) { set_error("fb_set_device: %d out of range (0..%d)", device, n - 1); return FB_E_INVALID; }
    t_device = device;
    return FB_OK;
}

const char *fb_last_error(void) { return t_err; }
const char *fb_version(void) { return "fennec-b200 0.1.0 (sm_100a; parity target shamspias/fennec 1.0.2 @ 98234f2c)"; }
long long fb_take_launch_count(void) { long long v = t_launches; t_launches = 0; return v; }

// ---- SSIM family --------------------------------------------------------------------------------
int fb_ssim(const uint8_t *a, int strideA, const uint8_t *b, int strideB, int w, int h, double *out) {
    return host_score("fb_ssim", OP_SSIM, a, strideA, b, strideB, w, h, out);
}
int fb_ssim_fast(const uint8_t *a, int strideA, const uint8_t *b, int strideB, int w, int h, double *out) {
    return host_score("fb_ssim_fast", OP_SSIM_FAST, a, strideA, b, strideB, w, h, out);
}
int fb_msssim(const uint8_t *a, int strideA, const uint8_t *b, int strideB, int w, int h, double *out) {
    return host_score("fb_msssim", OP_MSSSIM, a, strideA, b, strideB, w, h, out);
}
int fb_pixel_ssim(const uint8_t *a, int strideA, const uint8_t *b, int strideB, int w, int h, double *out) {
    return host_score("fb_pixel_ssim", OP_PIXEL, a, strideA, b, strideB, w, h, out);
}
int fb_ssim_fast_dims(int w, int h, int *newW, int *newH) {
    int nw, nh;
    int did = ssim_fast_dims(w, h, &nw, &nh);
    if (newW) *newW = nw;
    if (newH) *newH = nh;
    return did;
}

int fb_box_downsample(const uint8_t *src, int srcStride, int srcW, int srcH, uint8_t *dst, int dstStride,
                      int dstW, int dstH) {
    if (srcW <= 0 || srcH <= 0 || dstW <= 0 || dstH <= 0) return FB_IDENTITY;  // ssim.go:246-248
    FB_TRY(check_img("fb_box_downsample", src, srcStride, srcW, srcH));
    FB_TRY(check_img("fb_box_downsample", dst, dstStride, dstW, dstH));
    ApiScope scope_;
    DevCtx *c = ctx(current_device());
    if (!c) return ensure_init() < 0 ? FB_E_NOGPU : FB_E_CUDA;
    FB_TRY(reserve(c, c->stream, (size_t)dev_pitch(srcW) * srcH + (size_t)dev_pitch(dstW) * dstH + 4096, 256));
    uint8_t *ds;
    int ps;
    FB_TRY(upload(c, src, srcStride, srcW, srcH, &ds, &ps));
    int pd = dev_pitch(dstW);
    uint8_t *dd = (uint8_t *)c->ws.take((size_t)pd * dstH);
    FB_TRY(launch_box(c->stream, ds, 0, ps, srcW, srcH, dd, 0, pd, dstW, dstH, 1, nullptr));
    FB_TRY(download(c, dd, pd, dst, dstStride, dstW, dstH));
    FB_CUDA(cudaStreamSynchronize(c->stream));
    return FB_OK;
}

// ---- effects ---------------------------------------------------------------------------------------
int fb_blur_kernel(double sigma, double *kernel, int cap) {
    if (!(sigma > 0)) { set_error("fb_blur_kernel: sigma must be > 0"); return FB_E_INVALID; }
    std::vector<double> k;
    int radius = blur_kernel_host(sigma, k);
    if (!kernel || cap < (int)k.size()) return -(int)k.size();
    memcpy(kernel, k.data(), sizeof(double) * k.size());
    return radius;
}

static int blur_on_device(DevCtx *c, cudaStream_t s, const uint8_t *dsrc, uint8_t *ddst, long long imgStride,
                          int rowStride, int w, int h, int n, const double *kernel_host, int radius) {
    int taps = 2 * radius + 1;
    double *kpin = (double *)c->pin.take(sizeof(double) * taps);
    float *fpin = (float *)c->pin.take(sizeof(float) * taps);
    double *kdev = (double *)c->ws.take(sizeof(double) * taps);
    float *fdev = (float *)c->ws.take(sizeof(float) * taps);
    int tpitch = dev_pitch(w);
    long long timg = (long long)tpitch * h;
    uint8_t *tmp = (uint8_t *)c->ws.take((size_t)timg * n);
    if (!kpin || !fpin || !kdev || !fdev || !tmp) { set_error("internal: workspace under-reserved (blur)"); return FB_E_INVALID; }
    // The FP32 fast path rounds values known to lie in [0, 255] (a convex combination of bytes).  A caller-supplied
    // table with negative taps or a gain above 1 leaves that range: it takes the exact FP64 kernels only (wabs huge).
    double wabs = 0.0;
    bool convex = true;
    for (int i = 0; i < taps; i++) {
        kpin[i] = kernel_host[i];
        fpin[i] = (float)kernel_host[i];
        wabs += fabs(kernel_host[i]);
        if (!(kernel_host[i] >= 0.0)) convex = false;
    }
    if (!convex || !(wabs <= 1.000001)) wabs = 1e9;
    FB_CUDA(cudaMemcpyAsync(kdev, kpin, sizeof(double) * taps, cudaMemcpyHostToDevice, s));
    FB_CUDA(cudaMemcpyAsync(fdev, fpin, sizeof(float) * taps, cudaMemcpyHostToDevice, s));
    mark_pin_busy(c, s);
    return launch_gaussian_blur(s, dsrc, ddst, imgStride, rowStride, w, h, n, kdev, fdev, fpin, radius, wabs, tmp, timg, tpitch);
}

int fb_gaussian_blur(const uint8_t *src, int srcStride, int w, int h, const double *kernel, int radius,
                     uint8_t *dst, int dstStride) {
    FB_TRY(check_img("fb_gaussian_blur", src, srcStride, w, h));
    FB_TRY(check_img("fb_gaussian_blur", dst, dstStride, w, h));
    if (!kernel || radius < 0) { set_error("fb_gaussian_blur: kernel table missing or radius < 0"); return FB_E_INVALID; }
    if (w == 0 || h == 0) return FB_OK;
    ApiScope scope_;
    DevCtx *c = ctx(current_device());
    if (!c) return ensure_init() < 0 ? FB_E_NOGPU : FB_E_CUDA;
    int taps = 2 * radius + 1;
    size_t img = (size_t)dev_pitch(w) * h + 512;
    FB_TRY(reserve(c, c->stream, 3 * img + 16 * (size_t)taps + 4096, 16 * (size_t)taps + 256));
    uint8_t *ds;
    int ps;
    FB_TRY(upload(c, src, srcStride, w, h, &ds, &ps));
    uint8_t *dd = (uint8_t *)c->ws.take((size_t)ps * h);
    FB_TRY(blur_on_device(c, c->stream, ds, dd, 0, ps, w, h, 1, kernel, radius));
    FB_TRY(download(c, dd, ps, dst, dstStride, w, h));
    FB_CUDA(cudaStreamSynchronize(c->stream));
    return FB_OK;
}

int fb_gaussian_blur_sigma(const uint8_t *src, int srcStride, int w, int h, double sigma, uint8_t *dst,
                           int dstStride) {
    if (sigma <= 0) return FB_IDENTITY;  // effects.go:147-149
    std::vector<double> k;
    int radius = blur_kernel_host(sigma, k);
    return fb_gaussian_blur(src, srcStride, w, h, k.data(), radius, dst, dstStride);
}

static int host_fx(const char *fn, int mode, const uint8_t *src, int srcStride, int w, int h, double strength,
                   uint8_t *dst, int dstStride) {
    if (mode != 0) {
        if (strength <= 0) return FB_IDENTITY;  // effects.go:11-13 / 50-52
        if (strength > 1) strength = 1;
        if (w < 3 || h < 3) return FB_IDENTITY;  // effects.go:20-22 / 59-61
    }
    FB_TRY(check_img(fn, src, srcStride, w, h));
    FB_TRY(check_img(fn, dst, dstStride, w, h));
    if (w == 0 || h == 0) return FB_OK;
    ApiScope scope_;
    DevCtx *c = ctx(current_device());
    if (!c) return ensure_init() < 0 ? FB_E_NOGPU : FB_E_CUDA;
    size_t img = (size_t)dev_pitch(w) * h + 512;
    FB_TRY(reserve(c, c->stream, 2 * img + 4096, 256));
    uint8_t *ds;
    int ps;
    FB_TRY(upload(c, src, srcStride, w, h, &ds, &ps));
    uint8_t *dd = (uint8_t *)c->ws.take((size_t)ps * h);
    if (mode == 0) FB_TRY(launch_blur3x3(c->stream, ds, dd, 0, ps, w, h, 1, 0, ps));
    else {
        double amount = mode == 1 ? 1.0 + strength * 1.5 : 1.0 + strength * 2.0;  // effects.go:26 / 65
        FB_TRY(launch_sharpen(c->stream, ds, dd, 0, ps, w, h, 1, 0, ps, amount, mode == 2));
    }
    FB_TRY(download(c, dd, ps, dst, dstStride, w, h));
    FB_CUDA(cudaStreamSynchronize(c->stream));
    return FB_OK;
}

int fb_blur3x3(const uint8_t *src, int srcStride, int w, int h, uint8_t *dst, int dstStride) {
    return host_fx("fb_blur3x3", 0, src, srcStride, w, h, 0.0, dst, dstStride);
}
int fb_sharpen(const uint8_t *src, int srcStride, int w, int h, double strength, uint8_t *dst, int dstStride) {
    return host_fx("fb_sharpen", 1, src, srcStride, w, h, strength, dst, dstStride);
}
int fb_adaptive_sharpen(const uint8_t *src, int srcStride, int w, int h, double strength, uint8_t *dst,
                        int dstStride) {
    return host_fx("fb_adaptive_sharpen", 2, src, srcStride, w, h, strength, dst, dstStride);
}

// ---- Lanczos ----------------------------------------------------------------------------------------
int fb_lanczos_weights_cap(int dstSize, int srcSize) { return lanczos_cap(dstSize, srcSize); }

int fb_build_lanczos_weights(int dstSize, int srcSize, int *start, int *index, double *weight) {
    if (dstSize <= 0 || srcSize <= 0 || !start || !index || !weight) {
        set_error("fb_build_lanczos_weights: bad arguments");
        return FB_E_INVALID;
    }
    return lanczos_build(dstSize, srcSize, start, index, weight);
}

static int upload_weights(DevCtx *c, cudaStream_t s, const fb_weights *w, LanczosTable *t) {
    int n = w->n, entries = w->start[n];
    int *pstart = (int *)c->pin.take(sizeof(int) * (n + 1));
    int *pindex = (int *)c->pin.take(sizeof(int) * (entries + 1));
    double *pweight = (double *)c->pin.take(sizeof(double) * (entries + 1));
    float *pw32 = (float *)c->pin.take(sizeof(float) * (entries + 1));
    t->start = (int *)c->ws.take(sizeof(int) * (n + 1));
    t->index = (int *)c->ws.take(sizeof(int) * (entries + 1));
    t->weight = (double *)c->ws.take(sizeof(double) * (entries + 1));
    t->weight32 = (float *)c->ws.take(sizeof(float) * (entries + 1));
    if (!pstart || !pindex || !pweight || !pw32 || !t->start || !t->index || !t->weight || !t->weight32) {
        set_error("internal: workspace under-reserved (weights)");
        return FB_E_INVALID;
    }
    memcpy(pstart, w->start, sizeof(int) * (n + 1));
    memcpy(pindex, w->index, sizeof(int) * entries);
    memcpy(pweight, w->weight, sizeof(double) * entries);
    for (int i = 0; i < entries; i++) pw32[i] = (float)w->weight[i];
    FB_CUDA(cudaMemcpyAsync(t->weight32, pw32, sizeof(float) * entries, cudaMemcpyHostToDevice, s));
    FB_CUDA(cudaMemcpyAsync(t->start, pstart, sizeof(int) * (n + 1), cudaMemcpyHostToDevice, s));
    FB_CUDA(cudaMemcpyAsync(t->index, pindex, sizeof(int) * entries, cudaMemcpyHostToDevice, s));
    FB_CUDA(cudaMemcpyAsync(t->weight, pweight, sizeof(double) * entries, cudaMemcpyHostToDevice, s));
    mark_pin_busy(c, s);
    t->entries = entries;
    table_stats(w->start, w->weight, n, &t->maxTaps, &t->wabs);
    return FB_OK;
}

static int check_weights(const char *fn, const fb_weights *w, int dstSize, int srcSize) {
    if (!w) return FB_OK;
    if (w->n != dstSize || !w->start || !w->index || !w->weight) {
        set_error("%s: weight table has n=%d, expected %d", fn, w->n, dstSize);
        return FB_E_INVALID;
    }
    for (int d = 0; d < dstSize; d++)
        if (w->start[d + 1] < w->start[d]) { set_error("%s: weight table start[] not monotone", fn); return FB_E_INVALID; }
    int entries = w->start[dstSize];
    for (int i = 0; i < entries; i++)
        if (w->index[i] < 0 || w->index[i] >= srcSize) { set_error("%s: tap index %d out of range", fn, w->index[i]); return FB_E_INVALID; }
    return FB_OK;
}

// Is `p` memory of another device than `dev` (a peer-mapped destination: batch.PeerGather, a symmetric-memory view)?
static bool on_other_device(const void *p, int dev) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeDevice && at.device != dev;
}

static int resize_on_device(DevCtx *c, cudaStream_t s, const uint8_t *dsrc, long long srcImgStride, int srcRowStride,
                            int srcW, int srcH, uint8_t *ddst, long long dstImgStride, int dstRowStride, int dstW,
                            int dstH, int n, const LanczosTable &tx, const LanczosTable &ty) {
    int tpitch = dev_pitch(dstW);
    long long timg = (long long)tpitch * srcH;
    uint8_t *tmp = (uint8_t *)c->ws.take((size_t)timg * n);
    if (!tmp) { set_error("internal: workspace under-reserved (resize tmp)"); return FB_E_INVALID; }
    // Peer destination (the gather fused into the vertical pass's stores, batch.PeerGather): with many ranks storing into ONE GPU
    // the V pass is bound by the NVLink ingest of that GPU, not by this one — 8 ranks: 465 MB at ~900 GB/s = 0.52 ms behind
    // 0.57 ms of compute.  The batch then runs as up to four sub-batches, V of sub-batch k on a high-priority side stream while
    // H of sub-batch k+1 (compute-bound, local) runs on the caller's stream; the side stream is joined before returning.
    // Measured (8 images per rank, opaque 8K -> 1080p): 8 GPUs 1.146 -> 0.927 ms, 4 GPUs 0.758 -> 0.773 (the ingest is not
    // the bound yet), local destinations 0.569 -> 0.68-0.72 (four small launches, tails) — so it is opt-in:
    // FB_LZ_PIPE=peer pipelines batches whose destination lives on another device (bench.py sets it above four ranks),
    // FB_LZ_PIPE=1 every batch (tests), unset / 0: never.
    static const int pipeEnv = [] { const char *e = getenv("FB_LZ_PIPE"); return !e ? 0 : (e[0] == '1' ? 1 : (e[0] == 'p' ? 2 : 0)); }();
    const bool pipelined = n >= 2 && (pipeEnv == 1 || (pipeEnv == 2 && on_other_device(ddst, c->dev)));
    if (pipelined) {
        if (!c->side) {
            // highest priority: V blocks (few, stalled on remote stores) must be dispatched ahead of the next sub-batch's H blocks,
            // which would otherwise fill every SM first and serialise the two
            int prLo = 0, prHi = 0;
            cudaDeviceGetStreamPriorityRange(&prLo, &prHi);
            bool ok = cudaStreamCreateWithPriority(&c->side, cudaStreamNonBlocking, prHi) == cudaSuccess &&
                      cudaEventCreateWithFlags(&c->joinEv, cudaEventDisableTiming) == cudaSuccess;
            for (int i = 0; ok && i < 4; i++) ok = cudaEventCreateWithFlags(&c->pipeEv[i], cudaEventDisableTiming) == cudaSuccess;
            if (!ok) { cudaGetLastError(); set_error("resize: cannot create the side stream"); return FB_E_CUDA; }
        }
        const int chunks = n < 4 ? n : 4, per = (n + chunks - 1) / chunks;
        FB_CUDA(cudaEventRecord(c->joinEv, s));               // the side stream starts behind whatever precedes this call on s
        FB_CUDA(cudaStreamWaitEvent(c->side, c->joinEv, 0));
        for (int k = 0, i0 = 0; i0 < n; k++, i0 += per) {
            const int m = n - i0 < per ? n - i0 : per;
            FB_TRY(launch_resize_h(s, dsrc + (size_t)i0 * srcImgStride, srcImgStride, srcRowStride, srcW, srcH, tmp + (size_t)i0 * timg, timg,
                                   tpitch, dstW, m, tx.start, tx.index, tx.weight, tx.weight32, tx.maxTaps, tx.wabs, tx.first, tx.wpadT,
                                   tx.groups, &tx.ir));
            FB_CUDA(cudaEventRecord(c->pipeEv[k & 3], s));
            FB_CUDA(cudaStreamWaitEvent(c->side, c->pipeEv[k & 3], 0));
            FB_TRY(launch_resize_v(c->side, tmp + (size_t)i0 * timg, timg, tpitch, dstW, srcH, ddst + (size_t)i0 * dstImgStride, dstImgStride,
                                   dstRowStride, dstH, m, ty.start, ty.index, ty.weight, ty.weight32, ty.maxTaps, ty.wabs, &ty.ir));
        }
        FB_CUDA(cudaEventRecord(c->joinEv, c->side));
        FB_CUDA(cudaStreamWaitEvent(s, c->joinEv, 0));
        return FB_OK;
    }
    FB_TRY(launch_resize_h(s, dsrc, srcImgStride, srcRowStride, srcW, srcH, tmp, timg, tpitch, dstW, n, tx.start,
                           tx.index, tx.weight, tx.weight32, tx.maxTaps, tx.wabs, tx.first, tx.wpadT, tx.groups, &tx.ir));
    return launch_resize_v(s, tmp, timg, tpitch, dstW, srcH, ddst, dstImgStride, dstRowStride, dstH, n, ty.start,
                           ty.index, ty.weight, ty.weight32, ty.maxTaps, ty.wabs, &ty.ir);
}

int fb_lanczos_resize(const uint8_t *src, int srcStride, int srcW, int srcH, uint8_t *dst, int dstStride, int dstW,
                      int dstH, const fb_weights *wx, const fb_weights *wy) {
    if (srcW <= 0 || srcH <= 0 || dstW <= 0 || dstH <= 0) return FB_IDENTITY;  // resize.go:41-43
    FB_TRY(check_img("fb_lanczos_resize", src, srcStride, srcW, srcH));
    FB_TRY(check_img("fb_lanczos_resize", dst, dstStride, dstW, dstH));
    if (srcW == dstW && srcH == dstH) {  // resize.go:45-49 — plain copy, no device work needed
        for (int y = 0; y < srcH; y++) memcpy(dst + (size_t)y * dstStride, src + (size_t)y * srcStride, (size_t)srcW * 4);
        return FB_OK;
    }
    FB_TRY(check_weights("fb_lanczos_resize", wx, dstW, srcW));
    FB_TRY(check_weights("fb_lanczos_resize", wy, dstH, srcH));
    ApiScope scope_;
    DevCtx *c = ctx(current_device());
    if (!c) return ensure_init() < 0 ? FB_E_NOGPU : FB_E_CUDA;
    size_t tabBytes = 0;
    if (wx) tabBytes += 24 * (size_t)(wx->start[dstW] + dstW + 8);
    if (wy) tabBytes += 24 * (size_t)(wy->start[dstH] + dstH + 8);
    size_t need = (size_t)dev_pitch(srcW) * srcH + (size_t)dev_pitch(dstW) * srcH + (size_t)dev_pitch(dstW) * dstH +
                  tabBytes + 8192;
    FB_TRY(reserve(c, c->stream, need, tabBytes + 4096));
    LanczosTable tx, ty;
    if (wx) FB_TRY(upload_weights(c, c->stream, wx, &tx)); else FB_TRY(lanczos_table_cached(c, dstW, srcW, &tx));
    if (wy) FB_TRY(upload_weights(c, c->stream, wy, &ty)); else FB_TRY(lanczos_table_cached(c, dstH, srcH, &ty));
    uint8_t *ds;
    int ps;
    FB_TRY(upload(c, src, srcStride, srcW, srcH, &ds, &ps));
    int pd = dev_pitch(dstW);
    uint8_t *dd = (uint8_t *)c->ws.take((size_t)pd * dstH);
    FB_TRY(resize_on_device(c, c->stream, ds, 0, ps, srcW, srcH, dd, 0, pd, dstW, dstH, 1, tx, ty));
    FB_TRY(download(c, dd, pd, dst, dstStride, dstW, dstH));
    FB_CUDA(cudaStreamSynchronize(c->stream));
    return FB_OK;
}

int fb_smart_resize_dims(int srcW, int srcH, int maxW, int maxH, int *dstW, int *dstH) {
    if (maxW <= 0) maxW = srcW;  // resize.go:16-21
    if (maxH <= 0) maxH = srcH;
    int dw = srcW, dh = srcH, noop = 1;
    if (!(srcW <= maxW && srcH <= maxH)) {
        double ratio = fmin((double)maxW / (double)srcW, (double)maxH / (double)srcH);
        dw = (int)fmax(1.0, round((double)srcW * ratio));
        dh = (int)fmax(1.0, round((double)srcH * ratio));
        noop = 0;
    }
    if (dstW) *dstW = dw;
    if (dstH) *dstH = dh;
    return noop;
}

// ---- device-resident batch entry points ---------------------------------------------------------------
static int check_batch(const char *fn, const void *p, long long imgStride, int rowStride, int w, int h, int n) {
    if (n < 0 || w <= 0 || h <= 0) { set_error("%s: bad batch dims n=%d %dx%d", fn, n, w, h); return FB_E_INVALID; }
    if (!p && n > 0) { set_error("%s: null device pointer", fn); return FB_E_INVALID; }
    if (rowStride < w * 4 || (rowStride & 3) || ((uintptr_t)p & 3) || (imgStride & 3)) {
        set_error("%s: rowStride %d / imgStride %lld / base must be >= 4*w and 4-byte aligned", fn, rowStride, imgStride);
        return FB_E_INVALID;
    }
    if (n > 1 && imgStride < (long long)rowStride * (h - 1) + (long long)w * 4) {
        set_error("%s: imgStride %lld smaller than one image", fn, imgStride);
        return FB_E_INVALID;
    }
    return FB_OK;
}

int fb_ssim_batch_dev(int device, void *stream, const uint8_t *a, const uint8_t *b, int64_t imgStride,
                      int rowStride, int w, int h, int n, double *scores) {
    FB_TRY(check_batch("fb_ssim_batch_dev", a, imgStride, rowStride, w, h, n));
    FB_TRY(check_batch("fb_ssim_batch_dev", b, imgStride, rowStride, w, h, n));
    if (n == 0) return FB_OK;
    DevCtx *c;
    ApiScope scope_;
    FB_TRY(dev_ctx_for("fb_ssim_batch_dev", device, &c));
    scope_.bind(c, (cudaStream_t)stream);
    FB_TRY(reserve(c, (cudaStream_t)stream, ssim_scratch_bytes(w, h, n) + 1024, 256));
    void *scratch = c->ws.take(ssim_scratch_bytes(w, h, n));
    return launch_ssim(c, (cudaStream_t)stream, a, b, imgStride, imgStride, rowStride, rowStride, w, h, n, scores, 1,
                       scratch);
}

int fb_ssim_fast_batch_dev(int device, void *stream, const uint8_t *a, const uint8_t *b, int64_t imgStride,
                           int rowStride, int w, int h, int n, double *scores) {
    FB_TRY(check_batch("fb_ssim_fast_batch_dev", a, imgStride, rowStride, w, h, n));
    FB_TRY(check_batch("fb_ssim_fast_batch_dev", b, imgStride, rowStride, w, h, n));
    if (n == 0) return FB_OK;
    DevCtx *c;
    ApiScope scope_;
    FB_TRY(dev_ctx_for("fb_ssim_fast_batch_dev", device, &c));
    scope_.bind(c, (cudaStream_t)stream);
    FB_TRY(reserve(c, (cudaStream_t)stream, ssim_fast_scratch(w, h, n), 256));
    return pipeline_ssim_fast(c, (cudaStream_t)stream, ImgBatch{a, imgStride, rowStride}, ImgBatch{b, imgStride, rowStride},
                              w, h, n, scores, 1);
}

int fb_msssim_batch_dev(int device, void *stream, const uint8_t *a, const uint8_t *b, int64_t imgStride,
                        int rowStride, int w, int h, int n, double *scores) {
    FB_TRY(check_batch("fb_msssim_batch_dev", a, imgStride, rowStride, w, h, n));
    FB_TRY(check_batch("fb_msssim_batch_dev", b, imgStride, rowStride, w, h, n));
    if (n == 0) return FB_OK;
    DevCtx *c;
    ApiScope scope_;
    FB_TRY(dev_ctx_for("fb_msssim_batch_dev", device, &c));
    scope_.bind(c, (cudaStream_t)stream);
    FB_TRY(reserve(c, (cudaStream_t)stream, msssim_scratch(w, h, n), 1024));
    return pipeline_msssim(c, (cudaStream_t)stream, ImgBatch{a, imgStride, rowStride}, ImgBatch{b, imgStride, rowStride},
                           w, h, n, scores);
}

int fb_box_downsample_batch_dev(int device, void *stream, const uint8_t *src, int64_t srcImgStride, int srcRowStride,
                                int srcW, int srcH, uint8_t *dst, int64_t dstImgStride, int dstRowStride, int dstW,
                                int dstH, int n) {
    if (srcW <= 0 || srcH <= 0 || dstW <= 0 || dstH <= 0) return FB_IDENTITY;
    FB_TRY(check_batch("fb_box_downsample_batch_dev", src, srcImgStride, srcRowStride, srcW, srcH, n));
    FB_TRY(check_batch("fb_box_downsample_batch_dev", dst, dstImgStride, dstRowStride, dstW, dstH, n));
    DevCtx *c;
    ApiScope scope_;
    FB_TRY(dev_ctx_for("fb_box_downsample_batch_dev", device, &c));
    scope_.bind(c, (cudaStream_t)stream);
    return launch_box((cudaStream_t)stream, src, srcImgStride, srcRowStride, srcW, srcH, dst, dstImgStride,
                      dstRowStride, dstW, dstH, n, nullptr);
}

int fb_msssim_level_batch_dev(int device, void *stream, const uint8_t *a, const uint8_t *b, int64_t imgStride,
                              int rowStride, int w, int h, int n, uint8_t *thumbA, uint8_t *thumbB,
                              int64_t thumbImgStride, int thumbRowStride, int tw, int th, uint8_t *halfA,
                              uint8_t *halfB, int64_t halfImgStride, int halfRowStride) {
    const char *fn = "fb_msssim_level_batch_dev";
    if (w < 2 || h < 2 || tw <= 0 || th <= 0) { set_error("fb_msssim_level_batch_dev: bad dims"); return FB_E_INVALID; }
    FB_TRY(check_batch(fn, a, imgStride, rowStride, w, h, n));
    FB_TRY(check_batch(fn, b, imgStride, rowStride, w, h, n));
    FB_TRY(check_batch(fn, thumbA, thumbImgStride, thumbRowStride, tw, th, n));
    FB_TRY(check_batch(fn, thumbB, thumbImgStride, thumbRowStride, tw, th, n));
    FB_TRY(check_batch(fn, halfA, halfImgStride, halfRowStride, w / 2, h / 2, n));
    FB_TRY(check_batch(fn, halfB, halfImgStride, halfRowStride, w / 2, h / 2, n));
    if (n == 0) return FB_OK;
    DevCtx *c;
    ApiScope scope_;
    FB_TRY(dev_ctx_for(fn, device, &c));
    scope_.bind(c, (cudaStream_t)stream);
    cudaStream_t s = (cudaStream_t)stream;
    int rc = launch_box_fused(s, a, imgStride, rowStride, b, imgStride, rowStride, w, h, thumbA, thumbB, thumbImgStride,
                              thumbRowStride, tw, th, halfA, halfB, halfImgStride, halfRowStride, n);
    if (rc <= 0) return rc;
    // preconditions of the fused kernel not met: the four separate downsamples (same bytes)
    FB_TRY(launch_box(s, a, imgStride, rowStride, w, h, thumbA, thumbImgStride, thumbRowStride, tw, th, n, nullptr));
    FB_TRY(launch_box(s, b, imgStride, rowStride, w, h, thumbB, thumbImgStride, thumbRowStride, tw, th, n, nullptr));
    FB_TRY(launch_box(s, a, imgStride, rowStride, w, h, halfA, halfImgStride, halfRowStride, w / 2, h / 2, n, nullptr));
    return launch_box(s, b, imgStride, rowStride, w, h, halfB, halfImgStride, halfRowStride, w / 2, h / 2, n, nullptr);
}

int fb_msssim_level2_batch_dev(int device, void *stream, const uint8_t *a, const uint8_t *b, int64_t imgStride, int rowStride,
                               int w, int h, int n, uint8_t *thumb0A, uint8_t *thumb0B, uint8_t *thumb1A, uint8_t *thumb1B,
                               int64_t thumbImgStride, int thumbRowStride, int tw, int th, uint8_t *quarterA, uint8_t *quarterB,
                               int64_t quarterImgStride, int quarterRowStride) {
    const char *fn = "fb_msssim_level2_batch_dev";
    if (w < 4 || h < 4 || tw <= 0 || th <= 0) { set_error("fb_msssim_level2_batch_dev: bad dims"); return FB_E_INVALID; }
    FB_TRY(check_batch(fn, a, imgStride, rowStride, w, h, n));
    FB_TRY(check_batch(fn, b, imgStride, rowStride, w, h, n));
    FB_TRY(check_batch(fn, thumb0A, thumbImgStride, thumbRowStride, tw, th, n));
    FB_TRY(check_batch(fn, thumb0B, thumbImgStride, thumbRowStride, tw, th, n));
    FB_TRY(check_batch(fn, thumb1A, thumbImgStride, thumbRowStride, tw, th, n));
    FB_TRY(check_batch(fn, thumb1B, thumbImgStride, thumbRowStride, tw, th, n));
    FB_TRY(check_batch(fn, quarterA, quarterImgStride, quarterRowStride, w / 4, h / 4, n));
    FB_TRY(check_batch(fn, quarterB, quarterImgStride, quarterRowStride, w / 4, h / 4, n));
    if (n == 0) return FB_OK;
    DevCtx *c;
    ApiScope scope_;
    FB_TRY(dev_ctx_for(fn, device, &c));
    scope_.bind(c, (cudaStream_t)stream);
    // FB_IDENTITY-style answer 1: the geometry has no common period (or buffers are unaligned) — nothing was written;
    // the caller runs fb_msssim_level_batch_dev twice instead.
    return launch_box_fused2((cudaStream_t)stream, a, imgStride, rowStride, b, imgStride, rowStride, w, h, thumb0A, thumb0B,
                             thumbImgStride, thumbRowStride, tw, th, thumb1A, thumb1B, thumbImgStride, thumbRowStride, tw, th,
                             quarterA, quarterB, quarterImgStride, quarterRowStride, n);
}

int fb_gaussian_blur_batch_dev(int device, void *stream, const uint8_t *src, uint8_t *dst, int64_t imgStride,
                               int rowStride, int w, int h, int n, const double *kernel_host, int radius) {
    FB_TRY(check_batch("fb_gaussian_blur_batch_dev", src, imgStride, rowStride, w, h, n));
    FB_TRY(check_batch("fb_gaussian_blur_batch_dev", dst, imgStride, rowStride, w, h, n));
    if (!kernel_host || radius < 0) { set_error("fb_gaussian_blur_batch_dev: kernel table missing"); return FB_E_INVALID; }
    if (n == 0) return FB_OK;
    DevCtx *c;
    ApiScope scope_;
    FB_TRY(dev_ctx_for("fb_gaussian_blur_batch_dev", device, &c));
    scope_.bind(c, (cudaStream_t)stream);
    int taps = 2 * radius + 1;
    FB_TRY(reserve(c, (cudaStream_t)stream, (size_t)dev_pitch(w) * h * n + 16 * (size_t)taps + 4096, 16 * (size_t)taps + 256));
    return blur_on_device(c, (cudaStream_t)stream, src, dst, imgStride, rowStride, w, h, n, kernel_host, radius);
}

static int fx_batch(const char *fn, int adaptive, int device, void *stream, const uint8_t *src, uint8_t *dst,
                    int64_t imgStride, int rowStride, int w, int h, int n, double strength) {
    if (strength <= 0 || w < 3 || h < 3) return FB_IDENTITY;
    if (strength > 1) strength = 1;
    FB_TRY(check_batch(fn, src, imgStride, rowStride, w, h, n));
    FB_TRY(check_batch(fn, dst, imgStride, rowStride, w, h, n));
    if (n == 0) return FB_OK;
    DevCtx *c;
    ApiScope scope_;
    FB_TRY(dev_ctx_for(fn, device, &c));
    scope_.bind(c, (cudaStream_t)stream);
    double amount = adaptive ? 1.0 + strength * 2.0 : 1.0 + strength * 1.5;
    return launch_sharpen((cudaStream_t)stream, src, dst, imgStride, rowStride, w, h, n, imgStride, rowStride, amount,
                          adaptive);
}

int fb_sharpen_batch_dev(int device, void *stream, const uint8_t *src, uint8_t *dst, int64_t imgStride, int rowStride,
                         int w, int h, int n, double strength) {
    return fx_batch("fb_sharpen_batch_dev", 0, device, stream, src, dst, imgStride, rowStride, w, h, n, strength);
}
int fb_adaptive_sharpen_batch_dev(int device, void *stream, const uint8_t *src, uint8_t *dst, int64_t imgStride,
                                  int rowStride, int w, int h, int n, double strength) {
    return fx_batch("fb_adaptive_sharpen_batch_dev", 1, device, stream, src, dst, imgStride, rowStride, w, h, n, strength);
}

int fb_lanczos_resize_batch_dev(int device, void *stream, const uint8_t *src, int64_t srcImgStride, int srcRowStride,
                                int srcW, int srcH, uint8_t *dst, int64_t dstImgStride, int dstRowStride, int dstW,
                                int dstH, int n) {
    if (srcW <= 0 || srcH <= 0 || dstW <= 0 || dstH <= 0) return FB_IDENTITY;
    FB_TRY(check_batch("fb_lanczos_resize_batch_dev", src, srcImgStride, srcRowStride, srcW, srcH, n));
    FB_TRY(check_batch("fb_lanczos_resize_batch_dev", dst, dstImgStride, dstRowStride, dstW, dstH, n));
    if (n == 0) return FB_OK;
    DevCtx *c;
    ApiScope scope_;
    FB_TRY(dev_ctx_for("fb_lanczos_resize_batch_dev", device, &c));
    scope_.bind(c, (cudaStream_t)stream);
    cudaStream_t s = (cudaStream_t)stream;
    if (srcW == dstW && srcH == dstH) {
        for (int i = 0; i < n; i++)
            FB_CUDA(cudaMemcpy2DAsync(dst + (size_t)i * dstImgStride, dstRowStride, src + (size_t)i * srcImgStride,
                                      srcRowStride, (size_t)srcW * 4, srcH, cudaMemcpyDeviceToDevice, s));
        return FB_OK;
    }
    FB_TRY(reserve(c, (cudaStream_t)stream, (size_t)dev_pitch(dstW) * srcH * n + 8192, 256));
    LanczosTable tx, ty;
    FB_TRY(lanczos_table_cached(c, dstW, srcW, &tx));
    FB_TRY(lanczos_table_cached(c, dstH, srcH, &ty));
    return resize_on_device(c, s, src, srcImgStride, srcRowStride, srcW, srcH, dst, dstImgStride, dstRowStride, dstW,
                            dstH, n, tx, ty);
}

size_t fb_workspace_bytes(const char *op, int w, int h, int dstW, int dstH, int n) {
    if (!op) return 0;
    if (!strcmp(op, "ssim")) return ssim_scratch_bytes(w, h, n);
    if (!strcmp(op, "ssim_fast")) return ssim_fast_scratch(w, h, n);
    if (!strcmp(op, "msssim")) return msssim_scratch(w, h, n);
    if (!strcmp(op, "gaussian_blur")) return (size_t)dev_pitch(w) * h * n;
    if (!strcmp(op, "lanczos_resize")) return (size_t)dev_pitch(dstW) * h * n;
    if (!strcmp(op, "apply_palette")) return palette_scratch_bytes(w, h, n);
    (void)dstH;
    return 0;
}

// ---- batch sharder (batch.go:58-128) ----------------------------------------------------------------------
int fb_batch_shard(int n_items, int n_shards, int shard, int *begin, int *end) {
    if (n_items < 0 || n_shards <= 0 || shard < 0 || shard >= n_shards || !begin || !end) {
        set_error("fb_batch_shard: bad arguments (n_items=%d n_shards=%d shard=%d)", n_items, n_shards, shard);
        return FB_E_INVALID;
    }
    int per = (n_items + n_shards - 1) / n_shards;  // owner(i) = i / ceil(n/G)
    long long b = (long long)shard * per, e = b + per;
    if (b > n_items) b = n_items;
    if (e > n_items) e = n_items;
    *begin = (int)b;
    *end = (int)e;
    return FB_OK;
}

}  // extern "C"

namespace fb {

// ---- convertToNRGBA on decoded images, and the search loop's reference-image session ----

static int check_planes(const char *fn, const uint8_t *y, int yStride, const uint8_t *cb, const uint8_t *cr, int cStride,
                        int w, int h, int ratio, int *cw, int *ch) {
    int xs, ys;
    if (!ycbcr_ratio_shifts(ratio, &xs, &ys)) { set_error("%s: unknown subsample ratio %d", fn, ratio); return FB_E_INVALID; }
    if (w <= 0 || h <= 0) { set_error("%s: bad dims %dx%d", fn, w, h); return FB_E_INVALID; }
    *cw = (w + (1 << xs) - 1) >> xs;
    *ch = (h + (1 << ys) - 1) >> ys;
    if (!y || !cb || !cr) { set_error("%s: null plane pointer", fn); return FB_E_INVALID; }
    if (yStride < w || cStride < *cw) { set_error("%s: plane stride too small (y %d < %d or c %d < %d)", fn, yStride, w, cStride, *cw); return FB_E_INVALID; }
    return FB_OK;
}

// Upload the three planes into the workspace and convert; *out is a workspace NRGBA image of pitch *pitch.
static int ycbcr_upload_convert(DevCtx *c, const uint8_t *y, int yStride, const uint8_t *cb, const uint8_t *cr, int cStride,
                                int w, int h, int ratio, int cw, int ch, uint8_t **out, int *pitch) {
    const int yp = (int)align_up((size_t)w, 16), cp = (int)align_up((size_t)cw, 16);
    uint8_t *dy = (uint8_t *)c->ws.take((size_t)yp * h);
    uint8_t *dcb = (uint8_t *)c->ws.take((size_t)cp * ch);
    uint8_t *dcr = (uint8_t *)c->ws.take((size_t)cp * ch);
    *pitch = dev_pitch(w);
    *out = (uint8_t *)c->ws.take((size_t)*pitch * h + 16);
    if (!dy || !dcb || !dcr || !*out) { set_error("internal: workspace under-reserved (ycbcr)"); return FB_E_INVALID; }
    FB_TRY(upload_rows(c, y, (size_t)yStride, dy, (size_t)yp, (size_t)w, h));
    FB_TRY(upload_rows(c, cb, (size_t)cStride, dcb, (size_t)cp, (size_t)cw, ch));
    FB_TRY(upload_rows(c, cr, (size_t)cStride, dcr, (size_t)cp, (size_t)cw, ch));
    return launch_ycbcr_to_nrgba(c->stream, dy, 0, yp, dcb, dcr, 0, cp, w, h, ratio, *out, 0, *pitch, 1);
}

static size_t ycbcr_scratch(int w, int h, int cw, int ch) {
    return align_up((size_t)w, 16) * h + 2 * align_up((size_t)cw, 16) * ch + (size_t)dev_pitch(w) * h + 4096;
}

}  // namespace fb

struct fb_ssim_ref {
    int dev, w, h;       // PHYSICAL device and dims of the reference image
    int tw, th, pitch;   // what is kept: the SSIMFast thumbnail (or the image itself when <= 512 px)
    uint8_t *img;        // cudaMalloc'ed, owned
};

namespace fb {

// SSIMFast(ref, img) with ref's downsample cached: box `img` if needed, then K1 (ssim.go:48-70).
static int score_against_ref(DevCtx *c, const fb_ssim_ref *r, const uint8_t *dimg, int pitch, double *out) {
    const uint8_t *b = dimg;
    int bp = pitch;
    if (r->tw != r->w || r->th != r->h) {
        int tp = dev_pitch(r->tw);
        uint8_t *t = (uint8_t *)c->ws.take((size_t)tp * r->th + 16);
        if (!t) { set_error("internal: workspace under-reserved (ref thumb)"); return FB_E_INVALID; }
        FB_TRY(launch_box(c->stream, dimg, 0, pitch, r->w, r->h, t, 0, tp, r->tw, r->th, 1, nullptr));
        b = t;
        bp = tp;
    }
    double *dscore = (double *)c->ws.take(sizeof(double) * 2);
    void *scratch = c->ws.take(ssim_scratch_bytes(r->tw, r->th, 1));
    if (!dscore || !scratch) { set_error("internal: workspace under-reserved (ref score)"); return FB_E_INVALID; }
    FB_TRY(launch_ssim(c, c->stream, r->img, b, 0, 0, r->pitch, bp, r->tw, r->th, 1, dscore, 1, scratch));
    return finish_score(c, dscore, out);
}

static size_t ref_score_scratch(const fb_ssim_ref *r) {
    return (size_t)dev_pitch(r->tw) * r->th + ssim_scratch_bytes(r->tw, r->th, 1) + 4096;
}

}  // namespace fb

using namespace fb;

extern "C" {

int fb_ycbcr_to_nrgba(const uint8_t *y, int yStride, const uint8_t *cb, const uint8_t *cr, int cStride, int w, int h,
                      int ratio, uint8_t *dst, int dstStride) {
    int cw, ch;
    FB_TRY(check_planes("fb_ycbcr_to_nrgba", y, yStride, cb, cr, cStride, w, h, ratio, &cw, &ch));
    FB_TRY(check_img("fb_ycbcr_to_nrgba", dst, dstStride, w, h));
    ApiScope scope_;
    DevCtx *c = ctx(current_device());
    if (!c) return ensure_init() < 0 ? FB_E_NOGPU : FB_E_CUDA;
    FB_TRY(reserve(c, c->stream, ycbcr_scratch(w, h, cw, ch), 256));
    uint8_t *d;
    int pitch;
    FB_TRY(ycbcr_upload_convert(c, y, yStride, cb, cr, cStride, w, h, ratio, cw, ch, &d, &pitch));
    FB_TRY(download(c, d, pitch, dst, dstStride, w, h));
    FB_CUDA(cudaStreamSynchronize(c->stream));
    return FB_OK;
}

int fb_gray_to_nrgba(const uint8_t *g, int gStride, int w, int h, uint8_t *dst, int dstStride) {
    if (w <= 0 || h <= 0 || !g || gStride < w) { set_error("fb_gray_to_nrgba: bad plane"); return FB_E_INVALID; }
    FB_TRY(check_img("fb_gray_to_nrgba", dst, dstStride, w, h));
    ApiScope scope_;
    DevCtx *c = ctx(current_device());
    if (!c) return ensure_init() < 0 ? FB_E_NOGPU : FB_E_CUDA;
    const int gp = (int)align_up((size_t)w, 16), pitch = dev_pitch(w);
    FB_TRY(reserve(c, c->stream, (size_t)gp * h + (size_t)pitch * h + 4096, 256));
    uint8_t *dg = (uint8_t *)c->ws.take((size_t)gp * h);
    uint8_t *d = (uint8_t *)c->ws.take((size_t)pitch * h + 16);
    FB_TRY(upload_rows(c, g, (size_t)gStride, dg, (size_t)gp, (size_t)w, h));
    FB_TRY(launch_gray_to_nrgba(c->stream, dg, 0, gp, w, h, d, 0, pitch, 1));
    FB_TRY(download(c, d, pitch, dst, dstStride, w, h));
    FB_CUDA(cudaStreamSynchronize(c->stream));
    return FB_OK;
}

// convertToNRGBA (convert.go:34-64) for *image.RGBA / RGBA64 / NRGBA64 / Gray16 / CMYK / Paletted — pixfmt.cu
static int check_pixfmt(const char *fn, int format, const void *pix, int stride, int w, int h, const void *palette16, int ncolors) {
    const int bpp = pixfmt_bytes_per_pixel(format);
    if (!bpp) { set_error("%s: unknown format %d (1 RGBA, 2 RGBA64, 3 NRGBA64, 4 Gray16, 5 CMYK, 6 Paletted)", fn, format); return FB_E_INVALID; }
    if (w < 0 || h < 0) { set_error("%s: negative dimensions %dx%d", fn, w, h); return FB_E_INVALID; }
    if (w > 0 && h > 0) {
        if (!pix) { set_error("%s: null pixel pointer", fn); return FB_E_INVALID; }
        if ((long long)stride < (long long)w * bpp) { set_error("%s: stride %d < %d bytes per row", fn, stride, w * bpp); return FB_E_INVALID; }
    }
    if (format == FB_FMT_PALETTED && (!palette16 || ncolors < 1 || ncolors > 256)) {
        set_error("%s: a Paletted image needs 1..256 palette entries", fn);
        return FB_E_INVALID;
    }
    return FB_OK;
}

int fb_convert_to_nrgba(int format, const uint8_t *pix, int stride, int w, int h, const uint16_t *palette16, int ncolors,
                        uint8_t *dst, int dstStride) {
    FB_TRY(check_pixfmt("fb_convert_to_nrgba", format, pix, stride, w, h, palette16, ncolors));
    FB_TRY(check_img("fb_convert_to_nrgba", dst, dstStride, w, h));
    if (w == 0 || h == 0) return FB_OK;
    ApiScope scope_;
    DevCtx *c = ctx(current_device());
    if (!c) return ensure_init() < 0 ? FB_E_NOGPU : FB_E_CUDA;
    const int bpp = pixfmt_bytes_per_pixel(format);
    const size_t sp = align_up((size_t)w * bpp, 256);
    const int pitch = dev_pitch(w);
    FB_TRY(reserve(c, c->stream, sp * h + (size_t)pitch * h + 2048 + 256 + 4096, 2048 + 256));
    uint8_t *ds = (uint8_t *)c->ws.take(sp * h);
    uint8_t *d = (uint8_t *)c->ws.take((size_t)pitch * h + 16);
    uint16_t *dpal = (uint16_t *)c->ws.take(2048);
    unsigned int *dbad = (unsigned int *)c->ws.take(16);
    uint16_t *ppal = (uint16_t *)c->pin.take(2048);
    unsigned int *pbad = (unsigned int *)c->pin.take(16);
    if (!ds || !d || !dpal || !dbad || !ppal || !pbad) { set_error("internal: workspace under-reserved (convert)"); return FB_E_INVALID; }
    FB_TRY(upload_rows(c, pix, (size_t)stride, ds, (size_t)sp, (size_t)w * bpp, h));
    if (format == FB_FMT_PALETTED) {
        memset(ppal, 0, 2048);
        memcpy(ppal, palette16, (size_t)ncolors * 8);
        FB_CUDA(cudaMemcpyAsync(dpal, ppal, 2048, cudaMemcpyHostToDevice, c->stream));
        FB_CUDA(cudaMemsetAsync(dbad, 0, 4, c->stream));
    }
    FB_TRY(launch_pixfmt_to_nrgba(c->stream, format, ds, 0, (int)sp, w, h, dpal, ncolors, d, 0, pitch, 1, dbad));
    FB_TRY(download(c, d, pitch, dst, dstStride, w, h));
    if (format == FB_FMT_PALETTED) FB_CUDA(cudaMemcpyAsync(pbad, dbad, 4, cudaMemcpyDeviceToHost, c->stream));
    FB_CUDA(cudaStreamSynchronize(c->stream));
    if (format == FB_FMT_PALETTED && *pbad) {
        set_error("fb_convert_to_nrgba: a pixel index is >= the palette length %d (Go panics here); those pixels were written as 0", ncolors);
        return FB_E_INVALID;
    }
    return FB_OK;
}

int fb_convert_to_nrgba_batch_dev(int device, void *stream, int format, const uint8_t *pix, int64_t imgStride, int rowStride,
                                  int w, int h, int n, const uint16_t *palettes16, int ncolors, uint8_t *dst,
                                  int64_t dstImgStride, int dstRowStride) {
    FB_TRY(check_pixfmt("fb_convert_to_nrgba_batch_dev", format, pix, rowStride, w, h, palettes16, ncolors));
    if (n < 0) { set_error("fb_convert_to_nrgba_batch_dev: negative batch size"); return FB_E_INVALID; }
    if (n > 1 && imgStride < (int64_t)rowStride * h) { set_error("fb_convert_to_nrgba_batch_dev: image stride smaller than an image"); return FB_E_INVALID; }
    FB_TRY(check_batch("fb_convert_to_nrgba_batch_dev", dst, dstImgStride, dstRowStride, w, h, n));
    if (n == 0 || w == 0 || h == 0) return FB_OK;
    DevCtx *c;
    ApiScope scope_;
    FB_TRY(dev_ctx_for("fb_convert_to_nrgba_batch_dev", device, &c));
    scope_.bind(c, (cudaStream_t)stream);
    return launch_pixfmt_to_nrgba((cudaStream_t)stream, format, pix, imgStride, rowStride, w, h, palettes16, ncolors, dst,
                                  dstImgStride, dstRowStride, n, nullptr);
}

int fb_ycbcr_to_nrgba_batch_dev(int device, void *stream, const uint8_t *y, int64_t yImgStride, int yStride,
                                const uint8_t *cb, const uint8_t *cr, int64_t cImgStride, int cStride, int w, int h,
                                int ratio, uint8_t *dst, int64_t dstImgStride, int dstRowStride, int n) {
    int cw, ch;
    FB_TRY(check_planes("fb_ycbcr_to_nrgba_batch_dev", y, yStride, cb, cr, cStride, w, h, ratio, &cw, &ch));
    FB_TRY(check_batch("fb_ycbcr_to_nrgba_batch_dev", dst, dstImgStride, dstRowStride, w, h, n));
    if (n == 0) return FB_OK;
    DevCtx *c;
    ApiScope scope_;
    FB_TRY(dev_ctx_for("fb_ycbcr_to_nrgba_batch_dev", device, &c));
    scope_.bind(c, (cudaStream_t)stream);
    return launch_ycbcr_to_nrgba((cudaStream_t)stream, y, yImgStride, yStride, cb, cr, cImgStride, cStride, w, h, ratio, dst,
                                 dstImgStride, dstRowStride, n);
}

int fb_ssim_ref_create(const uint8_t *src, int stride, int w, int h, fb_ssim_ref **out) {
    if (!out) { set_error("fb_ssim_ref_create: null output"); return FB_E_INVALID; }
    *out = nullptr;
    FB_TRY(check_img("fb_ssim_ref_create", src, stride, w, h));
    if (w <= 0 || h <= 0) { set_error("fb_ssim_ref_create: empty image"); return FB_E_INVALID; }
    ApiScope scope_;
    DevCtx *c = ctx(current_device());
    if (!c) return ensure_init() < 0 ? FB_E_NOGPU : FB_E_CUDA;
    fb_ssim_ref r;
    r.dev = c->dev; r.w = w; r.h = h; r.tw = w; r.th = h;
    const bool down = ssim_fast_dims(w, h, &r.tw, &r.th) != 0;
    r.pitch = dev_pitch(r.tw);
    FB_TRY(reserve(c, c->stream, (size_t)dev_pitch(w) * h + 4096, 256));
    uint8_t *d;
    int pitch;
    FB_TRY(upload(c, src, stride, w, h, &d, &pitch));
    void *own = nullptr;
    cudaError_t e = cudaMalloc(&own, (size_t)r.pitch * r.th + 16);
    if (e != cudaSuccess) { cudaGetLastError(); set_error("fb_ssim_ref_create: cudaMalloc failed"); return FB_E_OOM; }
    r.img = (uint8_t *)own;
    int rc = down ? launch_box(c->stream, d, 0, pitch, w, h, r.img, 0, r.pitch, r.tw, r.th, 1, nullptr)
                  : (cudaMemcpy2DAsync(r.img, r.pitch, d, pitch, (size_t)w * 4, h, cudaMemcpyDeviceToDevice, c->stream) == cudaSuccess ? FB_OK : FB_E_CUDA);
    if (rc == FB_OK && cudaStreamSynchronize(c->stream) != cudaSuccess) rc = FB_E_CUDA;
    if (rc != FB_OK) { cudaFree(own); if (rc == FB_E_CUDA) set_error("fb_ssim_ref_create: CUDA failure"); return rc; }
    *out = new fb_ssim_ref(r);
    return FB_OK;
}

void fb_ssim_ref_destroy(fb_ssim_ref *ref) {
    if (!ref) return;
    ApiScope scope_;
    if (DevCtx *c = ctx_phys(ref->dev)) { (void)c; cudaFree(ref->img); }
    delete ref;
}

int fb_ssim_ref_score_nrgba(const fb_ssim_ref *ref, const uint8_t *img, int stride, double *score) {
    if (!ref || !score) { set_error("fb_ssim_ref_score_nrgba: null argument"); return FB_E_INVALID; }
    FB_TRY(check_img("fb_ssim_ref_score_nrgba", img, stride, ref->w, ref->h));
    ApiScope scope_;
    DevCtx *c = ctx_phys(ref->dev);
    if (!c) return ensure_init() < 0 ? FB_E_NOGPU : FB_E_CUDA;
    FB_TRY(reserve(c, c->stream, (size_t)dev_pitch(ref->w) * ref->h + ref_score_scratch(ref) + 4096, 256));
    uint8_t *d;
    int pitch;
    FB_TRY(upload(c, img, stride, ref->w, ref->h, &d, &pitch));
    return score_against_ref(c, ref, d, pitch, score);
}

int fb_ssim_ref_score_ycbcr(const fb_ssim_ref *ref, const uint8_t *y, int yStride, const uint8_t *cb, const uint8_t *cr,
                            int cStride, int ratio, double *score) {
    if (!ref || !score) { set_error("fb_ssim_ref_score_ycbcr: null argument"); return FB_E_INVALID; }
    int cw, ch;
    FB_TRY(check_planes("fb_ssim_ref_score_ycbcr", y, yStride, cb, cr, cStride, ref->w, ref->h, ratio, &cw, &ch));
    ApiScope scope_;
    DevCtx *c = ctx_phys(ref->dev);
    if (!c) return ensure_init() < 0 ? FB_E_NOGPU : FB_E_CUDA;
    FB_TRY(reserve(c, c->stream, ycbcr_scratch(ref->w, ref->h, cw, ch) + ref_score_scratch(ref), 256));
    uint8_t *d;
    int pitch;
    FB_TRY(ycbcr_upload_convert(c, y, yStride, cb, cr, cStride, ref->w, ref->h, ratio, cw, ch, &d, &pitch));
    return score_against_ref(c, ref, d, pitch, score);
}

}  // extern "C"

// ---- Analyze (analyze.go:26-176) ----------------------------------------------------------------

namespace fb {

// math.Log2 as Go defines it (Frexp; 0.5 -> exp-1; else Log(frac)*(1/Ln2) + exp), with libm's log.
static double go_log2(double x) {
    int e;
    double frac = frexp(x, &e);
    if (frac == 0.5) return (double)(e - 1);
    return log(frac) * (1.0 / 0.693147180559945309417232121458176568) + (double)e;
}

static void analyze_finish_host(const AnalyzeRaw &r, int w, int h, fb_image_stats *st) {
    memset(st, 0, sizeof *st);
    st->width = w;
    st->height = h;
    if (w <= 0 || h <= 0) return;   // analyze.go:36-38
    AnalyzeSteps sp;
    analyze_steps(w, h, &sp);
    const double n = (double)((long long)w * h);
    st->has_alpha = r.hasAlpha != 0;
    st->is_grayscale = r.hasColour == 0;
    st->unique_colors = (int)(r.uniqueSampled < 1024u ? r.uniqueSampled : 1024u);   // the map stops growing at 1024
    st->mean_brightness = (double)r.sumL / 1000.0 / n;
    const long long samples = (long long)sp.contrastNx * sp.contrastNy;
    if (samples > 0) st->contrast = sqrt(r.varSum / (double)samples);
    double entropy = 0.0;           // computeEntropy analyze.go:116-128, bins in ascending order
    for (int i = 0; i < 256; i++)
        if (r.hist[i] > 0) {
            double p = (double)r.hist[i] / n;
            entropy -= p * go_log2(p);
        }
    st->entropy = entropy;
    const long long total = (long long)sp.edgeNx * sp.edgeNy;
    st->edge_density = total > 0 ? (double)r.edges / (double)total : 0.0;
    // recommendFormat / recommendQuality / estimateCompression, analyze.go:183-232
    const int JPEG = 1, PNG = 2, Balanced = 0, High = 3, Aggressive = 4;
    if (st->has_alpha) st->recommended_format = PNG;
    else if (st->unique_colors <= 256) st->recommended_format = PNG;
    else if (st->edge_density > 0.3 && st->unique_colors < 1000) st->recommended_format = PNG;
    else st->recommended_format = JPEG;
    if (st->entropy > 6 && st->edge_density < 0.15) st->recommended_quality = Balanced;
    else if (st->entropy < 4) st->recommended_quality = Aggressive;
    else if (st->edge_density > 0.25) st->recommended_quality = High;
    else st->recommended_quality = Balanced;
    if (st->recommended_format == PNG) {
        if (st->unique_colors <= 256) st->estimated_compression = 5.0 + (256 - (double)st->unique_colors) / 50;
        else if (st->is_grayscale) st->estimated_compression = 3.0;
        else st->estimated_compression = 2.0;
    } else {
        double base = 10.0;
        if (st->entropy > 7) base = 5.0;
        else if (st->entropy > 5) base = 8.0;
        if (st->edge_density > 0.2) base *= 0.7;
        st->estimated_compression = base;
    }
}

}  // namespace fb

extern "C" {

size_t fb_analyze_raw_bytes(void) { return sizeof(AnalyzeRaw); }

int fb_analyze_finish(const void *raw_host, int w, int h, fb_image_stats *out) {
    if (!raw_host || !out) { set_error("fb_analyze_finish: null argument"); return FB_E_INVALID; }
    AnalyzeRaw r;
    memcpy(&r, raw_host, sizeof r);
    analyze_finish_host(r, w, h, out);
    return FB_OK;
}

int fb_analyze_batch_dev(int device, void *stream, const uint8_t *imgs, int64_t imgStride, int rowStride, int w, int h,
                         int n, void *raw) {
    FB_TRY(check_batch("fb_analyze_batch_dev", imgs, imgStride, rowStride, w, h, n));
    if (!raw) { set_error("fb_analyze_batch_dev: null output"); return FB_E_INVALID; }
    if (n == 0) return FB_OK;
    DevCtx *c;
    ApiScope scope_;
    FB_TRY(dev_ctx_for("fb_analyze_batch_dev", device, &c));
    scope_.bind(c, (cudaStream_t)stream);
    FB_TRY(reserve(c, (cudaStream_t)stream, analyze_scratch_bytes(w, h, n), 256));
    void *scratch = c->ws.take(analyze_scratch_bytes(w, h, n) - 512);
    if (!scratch) { set_error("internal: workspace under-reserved (analyze)"); return FB_E_INVALID; }
    return launch_analyze((cudaStream_t)stream, imgs, imgStride, rowStride, w, h, n, (AnalyzeRaw *)raw, scratch);
}

int fb_analyze(const uint8_t *pix, int stride, int w, int h, fb_image_stats *out) {
    if (!out) { set_error("fb_analyze: null output"); return FB_E_INVALID; }
    FB_TRY(check_img("fb_analyze", pix, stride, w, h));
    if (w == 0 || h == 0) {          // analyze.go:36-38: only the dimensions are filled in
        memset(out, 0, sizeof *out);
        out->width = w;
        out->height = h;
        return FB_OK;
    }
    ApiScope scope_;
    DevCtx *c = ctx(current_device());
    if (!c) return ensure_init() < 0 ? FB_E_NOGPU : FB_E_CUDA;
    FB_TRY(reserve(c, c->stream, (size_t)dev_pitch(w) * h + analyze_scratch_bytes(w, h, 1) + sizeof(AnalyzeRaw) + 4096, sizeof(AnalyzeRaw) + 256));
    uint8_t *d;
    int pitch;
    FB_TRY(upload(c, pix, stride, w, h, &d, &pitch));
    AnalyzeRaw *draw = (AnalyzeRaw *)c->ws.take(sizeof(AnalyzeRaw));
    void *scratch = c->ws.take(analyze_scratch_bytes(w, h, 1) - 512);
    AnalyzeRaw *pin = (AnalyzeRaw *)c->pin.take(sizeof(AnalyzeRaw));
    if (!draw || !scratch || !pin) { set_error("internal: workspace under-reserved (analyze)"); return FB_E_INVALID; }
    FB_TRY(launch_analyze(c->stream, d, 0, pitch, w, h, 1, draw, scratch));
    FB_CUDA(cudaMemcpyAsync(pin, draw, sizeof(AnalyzeRaw), cudaMemcpyDeviceToHost, c->stream));
    FB_CUDA(cudaStreamSynchronize(c->stream));
    analyze_finish_host(*pin, w, h, out);
    return FB_OK;
}

}  // extern "C"

// ---- ApplyOrientation (exif.go:176-203) --------------------------------------------------------

extern "C" {

int fb_orientation_dims(int orient, int w, int h, int *dstW, int *dstH) {
    if (!dstW || !dstH) { set_error("fb_orientation_dims: null output"); return FB_E_INVALID; }
    *dstW = w;
    *dstH = h;
    return orient_dims(orient, w, h, dstW, dstH) ? FB_OK : FB_IDENTITY;
}

int fb_apply_orientation(const uint8_t *src, int srcStride, int w, int h, int orient, uint8_t *dst, int dstStride) {
    int dw, dh;
    if (!orient_dims(orient, w, h, &dw, &dh)) return FB_IDENTITY;
    FB_TRY(check_img("fb_apply_orientation", src, srcStride, w, h));
    FB_TRY(check_img("fb_apply_orientation", dst, dstStride, dw, dh));
    if (w == 0 || h == 0) return FB_OK;
    ApiScope scope_;
    DevCtx *c = ctx(current_device());
    if (!c) return ensure_init() < 0 ? FB_E_NOGPU : FB_E_CUDA;
    FB_TRY(reserve(c, c->stream, (size_t)dev_pitch(w) * h + (size_t)dev_pitch(dw) * dh + 4096, 256));
    uint8_t *d;
    int pitch;
    FB_TRY(upload(c, src, srcStride, w, h, &d, &pitch));
    const int opitch = dev_pitch(dw);
    uint8_t *o = (uint8_t *)c->ws.take((size_t)opitch * dh + 16);
    if (!o) { set_error("internal: workspace under-reserved (orientation)"); return FB_E_INVALID; }
    FB_TRY(launch_orient(c->stream, d, 0, pitch, w, h, orient, o, 0, opitch, 1));
    FB_TRY(download(c, o, opitch, dst, dstStride, dw, dh));
    FB_CUDA(cudaStreamSynchronize(c->stream));
    return FB_OK;
}

int fb_apply_orientation_batch_dev(int device, void *stream, const uint8_t *src, int64_t srcImgStride, int srcRowStride,
                                   int w, int h, int orient, uint8_t *dst, int64_t dstImgStride, int dstRowStride, int n) {
    int dw, dh;
    if (!orient_dims(orient, w, h, &dw, &dh)) return FB_IDENTITY;
    FB_TRY(check_batch("fb_apply_orientation_batch_dev", src, srcImgStride, srcRowStride, w, h, n));
    FB_TRY(check_batch("fb_apply_orientation_batch_dev", dst, dstImgStride, dstRowStride, dw, dh, n));
    if (n == 0) return FB_OK;
    DevCtx *c;
    ApiScope scope_;
    FB_TRY(dev_ctx_for("fb_apply_orientation_batch_dev", device, &c));
    scope_.bind(c, (cudaStream_t)stream);
    return launch_orient((cudaStream_t)stream, src, srcImgStride, srcRowStride, w, h, orient, dst, dstImgStride, dstRowStride, n);
}

}  // extern "C"

// ---- applyPalette + palettedToNRGBA (targetsize.go:479-545) ------------------------------------

extern "C" {

static int check_palette(const char *fn, const uint8_t *palette, int ncolors, bool host) {
    if (!palette || ncolors < 1 || ncolors > 256) { set_error("%s: palette must hold 1..256 entries", fn); return FB_E_INVALID; }
    if (host)
        for (int i = 0; i < ncolors; i++)
            if (palette[4 * i + 3] != 255) { set_error("%s: palette entry %d has alpha %d (medianCut produces 255)", fn, i, palette[4 * i + 3]); return FB_E_INVALID; }
    return FB_OK;
}

int fb_apply_palette(const uint8_t *src, int srcStride, int w, int h, const uint8_t *palette, int ncolors,
                     uint8_t *indices, int idxStride, uint8_t *dst, int dstStride) {
    FB_TRY(check_palette("fb_apply_palette", palette, ncolors, true));
    FB_TRY(check_img("fb_apply_palette", src, srcStride, w, h));
    if (dst) FB_TRY(check_img("fb_apply_palette", dst, dstStride, w, h));
    if (indices && idxStride < w) { set_error("fb_apply_palette: index stride %d < w", idxStride); return FB_E_INVALID; }
    if (w == 0 || h == 0) return FB_OK;
    ApiScope scope_;
    DevCtx *c = ctx(current_device());
    if (!c) return ensure_init() < 0 ? FB_E_NOGPU : FB_E_CUDA;
    const int ipitch = (int)align_up((size_t)w, 16);
    FB_TRY(reserve(c, c->stream, 2 * ((size_t)dev_pitch(w) * h + 512) + (size_t)ipitch * h + palette_scratch_bytes(w, h, 1) + 4096, 2048));
    uint8_t *d;
    int pitch;
    FB_TRY(upload(c, src, srcStride, w, h, &d, &pitch));
    uint8_t *dpal = (uint8_t *)c->ws.take(1024);
    uint8_t *ppin = (uint8_t *)c->pin.take(1024);
    uint8_t *didx = indices ? (uint8_t *)c->ws.take((size_t)ipitch * h) : nullptr;
    uint8_t *dout = dst ? (uint8_t *)c->ws.take((size_t)pitch * h + 16) : nullptr;
    const size_t cellBytes = palette_scratch_bytes(w, h, 1);
    void *cells = cellBytes ? c->ws.take(cellBytes - 256) : nullptr;
    if (!dpal || !ppin || (indices && !didx) || (dst && !dout) || (cellBytes && !cells)) { set_error("internal: workspace under-reserved (palette)"); return FB_E_INVALID; }
    memset(ppin, 0, 1024);
    memcpy(ppin, palette, (size_t)ncolors * 4);
    FB_CUDA(cudaMemcpyAsync(dpal, ppin, 1024, cudaMemcpyHostToDevice, c->stream));
    FB_TRY(launch_apply_palette(c->stream, d, 0, pitch, w, h, dpal, ncolors, didx, 0, ipitch, dout, 0, pitch, 1, cells));
    if (indices) FB_TRY(download_rows(c, didx, (size_t)ipitch, indices, (size_t)idxStride, (size_t)w, h));
    if (dst) FB_TRY(download(c, dout, pitch, dst, dstStride, w, h));
    FB_CUDA(cudaStreamSynchronize(c->stream));
    return FB_OK;
}

int fb_apply_palette_batch_dev(int device, void *stream, const uint8_t *src, int64_t imgStride, int rowStride, int w, int h,
                               int n, const uint8_t *palettes, int ncolors, uint8_t *indices, int64_t idxImgStride,
                               int idxRowStride, uint8_t *dst, int64_t dstImgStride, int dstRowStride) {
    FB_TRY(check_palette("fb_apply_palette_batch_dev", palettes, ncolors, false));
    FB_TRY(check_batch("fb_apply_palette_batch_dev", src, imgStride, rowStride, w, h, n));
    if (dst) FB_TRY(check_batch("fb_apply_palette_batch_dev", dst, dstImgStride, dstRowStride, w, h, n));
    if (indices && idxRowStride < w) { set_error("fb_apply_palette_batch_dev: index stride %d < w", idxRowStride); return FB_E_INVALID; }
    if (n == 0) return FB_OK;
    DevCtx *c;
    ApiScope scope_;
    FB_TRY(dev_ctx_for("fb_apply_palette_batch_dev", device, &c));
    scope_.bind(c, (cudaStream_t)stream);
    const size_t cellBytes = palette_scratch_bytes(w, h, n);
    void *cells = nullptr;
    if (cellBytes) {
        FB_TRY(reserve(c, (cudaStream_t)stream, cellBytes + 1024, 256));
        cells = c->ws.take(cellBytes - 256);
        if (!cells) { set_error("internal: workspace under-reserved (palette cells)"); return FB_E_INVALID; }
    }
    return launch_apply_palette((cudaStream_t)stream, src, imgStride, rowStride, w, h, palettes, ncolors, indices, idxImgStride,
                                idxRowStride, dst, dstImgStride, dstRowStride, n, cells);
}

}  // extern "C"


// =====================================================================================================
// Host-buffer batches: the worker pool of CompressBatch (batch.go:58-128) behind ONE call.
// The item list is split over the initialised devices with fb_batch_shard (static contiguous blocks, input order kept);
// inside a device's block `workers_per_device` library threads pull item indices from an atomic counter — the
// reference's buffered channel of indices (batch.go:72-81) — and run the ordinary host entry point for each item on
// their own stream, so one worker's staging / H2D overlaps another's kernels.  Results land at their input index
// (batch.go:108-113); one failing item does not stop the rest; a raised cancel flag marks every item not yet started
// with FB_E_CANCELLED (batch.go:90-98); on_item(completed, total) fires after each item, possibly concurrently
// (batch.go:115-121).  No collective is involved: per-item work is independent and the outputs go straight to the
// caller's buffers.
// =====================================================================================================
namespace fb {

template <typename F>
static int run_batch_host(const char *fn, int n, const fb_batch_opts *opts, int *status, F item_fn) {
    if (n < 0) { set_error("%s: n = %d", fn, n); return FB_E_INVALID; }
    if (n == 0) return 0;   // batch.go:59-61: nothing to do
    const int nd = ensure_init();
    if (nd < 0) return nd;
    const int wpd = (opts && opts->workers_per_device > 0) ? opts->workers_per_device : 4;
    const volatile int *cancel = opts ? opts->cancel : nullptr;
    std::unique_ptr<std::atomic<int>[]> next(new std::atomic<int>[nd]);
    std::atomic<int> completed{0}, failed{0};
    std::atomic<long long> launches{0};
    std::mutex errMu;
    std::string firstErr;
    std::vector<std::thread> workers;
    for (int d = 0; d < nd; d++) {
        int b = 0, e = 0;
        fb_batch_shard(n, nd, d, &b, &e);
        next[d].store(b);
        const int nw = e - b < wpd ? e - b : wpd;
        for (int k = 0; k < nw; k++) {
            workers.emplace_back([&, d, e]() {
                fb_set_device(d);
                for (;;) {
                    const int i = next[d].fetch_add(1);
                    if (i >= e) break;
                    int rc;
                    if (cancel && *cancel) {
                        rc = FB_E_CANCELLED;
                    } else {
                        rc = item_fn(i);
                        if (rc < 0) {
                            std::lock_guard<std::mutex> lk(errMu);
                            if (firstErr.empty()) firstErr = std::string("item ") + std::to_string(i) + ": " + fb_last_error();
                        }
                    }
                    if (status) status[i] = rc;
                    if (rc < 0) failed.fetch_add(1);
                    const int c = completed.fetch_add(1) + 1;
                    if (opts && opts->on_item) opts->on_item(c, n, opts->user);
                }
                launches.fetch_add(t_launches);   // the workers' kernel launches count for the calling thread
            });
        }
    }
    for (auto &t : workers) t.join();
    t_launches += launches.load();
    if (!firstErr.empty()) set_error("%s: %s", fn, firstErr.c_str());
    return failed.load();
}

}  // namespace fb

extern "C" {

int fb_score_batch_host(int op, const fb_pair *pairs, int n, double *scores, int *status, const fb_batch_opts *opts) {
    if (op != FB_OP_SSIM && op != FB_OP_SSIM_FAST && op != FB_OP_MSSSIM) { set_error("fb_score_batch_host: unknown op %d", op); return FB_E_INVALID; }
    if (n > 0 && (!pairs || !scores)) { set_error("fb_score_batch_host: null pairs / scores"); return FB_E_INVALID; }
    return run_batch_host("fb_score_batch_host", n, opts, status, [&](int i) {
        const fb_pair &p = pairs[i];
        switch (op) {
            case FB_OP_SSIM: return fb_ssim(p.a, p.strideA, p.b, p.strideB, p.w, p.h, &scores[i]);
            case FB_OP_SSIM_FAST: return fb_ssim_fast(p.a, p.strideA, p.b, p.strideB, p.w, p.h, &scores[i]);
            default: return fb_msssim(p.a, p.strideA, p.b, p.strideB, p.w, p.h, &scores[i]);
        }
    });
}

int fb_lanczos_resize_batch_host(const fb_resize_item *items, int n, int *status, const fb_batch_opts *opts) {
    if (n > 0 && !items) { set_error("fb_lanczos_resize_batch_host: null items"); return FB_E_INVALID; }
    return run_batch_host("fb_lanczos_resize_batch_host", n, opts, status, [&](int i) {
        const fb_resize_item &t = items[i];
        return fb_lanczos_resize(t.src, t.srcStride, t.srcW, t.srcH, t.dst, t.dstStride, t.dstW, t.dstH, nullptr, nullptr);
    });
}

int fb_effect_batch_host(int effect, double param, const fb_effect_item *items, int n, int *status, const fb_batch_opts *opts) {
    if (effect != FB_FX_GAUSSIAN_BLUR && effect != FB_FX_SHARPEN && effect != FB_FX_ADAPTIVE_SHARPEN) {
        set_error("fb_effect_batch_host: unknown effect %d", effect);
        return FB_E_INVALID;
    }
    if (n > 0 && !items) { set_error("fb_effect_batch_host: null items"); return FB_E_INVALID; }
    return run_batch_host("fb_effect_batch_host", n, opts, status, [&](int i) {
        const fb_effect_item &t = items[i];
        switch (effect) {
            case FB_FX_GAUSSIAN_BLUR: return fb_gaussian_blur_sigma(t.src, t.srcStride, t.w, t.h, param, t.dst, t.dstStride);
            case FB_FX_SHARPEN: return fb_sharpen(t.src, t.srcStride, t.w, t.h, param, t.dst, t.dstStride);
            default: return fb_adaptive_sharpen(t.src, t.srcStride, t.w, t.h, param, t.dst, t.dstStride);
        }
    });
}

}  // extern "C"

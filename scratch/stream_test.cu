// Development harness (not shipped): streaming cascade (stream.cuh, cp.async input form) vs the tile cascade (k_cascade):
// bitwise comparison on a set of image sizes, then timing at the 4K / 1080p octave-0 sizes.  Call pyramid_init() first
// when the TMA form is wanted (it resolves cuTensorMapEncodeTiled).
#ifndef PYR_SRC
#define PYR_SRC "../sift_project_b200/csrc/pyramid.cu"
#endif
#include PYR_SRC
#include <cstdio>
#include <cmath>
#include <cstring>
#include <vector>
namespace sb {
static BlurTaps mk(double sigma) {
    BlurTaps t{}; int n = (int)ceil(3 * sigma) + 1; double k[64], tot = 0;
    for (int i = 0; i < n; ++i) { k[i] = exp(-i * i / (2 * sigma * sigma)); tot += i ? 2 * k[i] : k[i]; }
    t.radius = n - 1; for (int i = 0; i < n; ++i) t.w[i] = (float)(k[i] / tot); return t;
}
static const double sig[6] = {1.6, 1.2262734984654078, 1.5450077936447955, 1.9465878414647133, 2.4525469969308156, 3.090015587289591};

struct Planes {
    float *in, *g[3], *d[3], *dec;
    size_t n;
    void alloc(size_t n_) { n = n_; cudaMalloc(&in, n * 4); for (int i = 0; i < 3; ++i) { cudaMalloc(&g[i], n * 4); cudaMalloc(&d[i], n * 4); } cudaMalloc(&dec, n * 4); }
    void clear() { for (int i = 0; i < 3; ++i) { cudaMemset(g[i], 0xff, n * 4); cudaMemset(d[i], 0xff, n * 4); } cudaMemset(dec, 0xff, n * 4); }
    void release() { cudaFree(in); for (int i = 0; i < 3; ++i) { cudaFree(g[i]); cudaFree(d[i]); } cudaFree(dec); }
};

static CascadeArgs args_a(const Planes& p, int w, int h, int pitch) {
    CascadeArgs a{};
    a.in = p.in; a.w = w; a.h = h; a.pitch = pitch;
    for (int i = 0; i < 3; ++i) { a.g[i] = p.g[i]; a.d[i] = p.d[i]; a.taps[i] = mk(sig[i + 1]); }
    a.dec = p.dec; a.dec_w = w / 2; a.dec_h = h / 2; a.dec_pitch = ((w / 2) + 31) & ~31;
    if (a.dec_w == 0 || a.dec_h == 0) a.dec = nullptr;
    return a;
}
static CascadeArgs args_b(const Planes& p, int w, int h, int pitch, bool keep) {
    CascadeArgs a{};
    a.in = p.in; a.w = w; a.h = h; a.pitch = pitch;
    a.g[0] = keep ? p.g[0] : nullptr; a.g[1] = keep ? p.g[1] : nullptr; a.g[2] = nullptr;
    a.d[0] = p.d[0]; a.d[1] = p.d[1]; a.d[2] = nullptr;
    a.taps[0] = mk(sig[4]); a.taps[1] = mk(sig[5]); a.taps[2] = a.taps[1];
    return a;
}

static long long compare(const float* x, const float* y, int w, int h, int pitch, const char* name) {
    std::vector<float> a((size_t)pitch * h), b((size_t)pitch * h);
    cudaMemcpy(a.data(), x, a.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(b.data(), y, b.size() * 4, cudaMemcpyDeviceToHost);
    long long bad = 0; int fx = -1, fy = -1;
    for (int r = 0; r < h; ++r)
        for (int c = 0; c < w; ++c)
            if (memcmp(&a[(size_t)r * pitch + c], &b[(size_t)r * pitch + c], 4)) { if (!bad) { fx = c; fy = r; } ++bad; }
    if (bad) printf("    MISMATCH %s: %lld px, first at (%d,%d): %g vs %g\n", name, bad, fx, fy,
                    a[(size_t)fy * pitch + fx], b[(size_t)fy * pitch + fx]);
    return bad;
}

static int check(int w, int h, int segs, bool onewarp) {
    const int pitch = (w + 31) & ~31;
    const size_t n = (size_t)pitch * h;
    Planes ref, neu; ref.alloc(n); neu.alloc(n);
    std::vector<float> hbuf(n);
    unsigned st = 12345u + w * 31 + h;
    for (size_t i = 0; i < n; ++i) { st = st * 1664525u + 1013904223u; hbuf[i] = (float)(st >> 24) + (float)((st >> 8) & 0xff) / 256.f; }
    cudaMemcpy(ref.in, hbuf.data(), n * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(neu.in, hbuf.data(), n * 4, cudaMemcpyHostToDevice);
    long long bad = 0;
    for (int variant = 0; variant < 2; ++variant) {
        ref.clear(); neu.clear();
        CascadeArgs ra = variant == 0 ? args_a(ref, w, h, pitch) : args_b(ref, w, h, pitch, true);
        CascadeArgs na = variant == 0 ? args_a(neu, w, h, pitch) : args_b(neu, w, h, pitch, true);
        cudaError_t e = variant == 0 ? launch_cascade_t<4, 5, 6>(ra, 148, 0) : launch_cascade_t<8, 10, 0>(ra, 148, 0);
        using CA = StreamGeom<3, 4, 5, 6, 2, 96, 12, 4, true, 1>;
        using CB = StreamGeom<2, 8, 10, 0, 2, 104, 6, 5, true, 2>;
        if (e == cudaSuccess) {
            if (onewarp) e = variant == 0 ? launch_stream_t<CA>(na, 148, 0, segs) : launch_stream_t<CB>(na, 148, 0, segs);
            else e = variant == 0 ? launch_stream_t<StreamA>(na, 148, 0, segs) : launch_stream_t<StreamB>(na, 148, 0, segs);
        }
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("  %dx%d variant %d: CUDA error %s\n", w, h, variant, cudaGetErrorString(e)); return 1; }
        const int nl = variant == 0 ? 3 : 2;
        for (int i = 0; i < nl; ++i) {
            char nm[32];
            snprintf(nm, sizeof nm, "v%d g[%d]", variant, i); bad += compare(ref.g[i], neu.g[i], w, h, pitch, nm);
            snprintf(nm, sizeof nm, "v%d d[%d]", variant, i); bad += compare(ref.d[i], neu.d[i], w, h, pitch, nm);
        }
        if (variant == 0 && ra.dec) bad += compare(ref.dec, neu.dec, ra.dec_w, ra.dec_h, ra.dec_pitch, "dec");
    }
    printf("  %5d x %5d ctas %d %s: %s\n", w, h, segs, onewarp ? "2 columns/thread" : "default", bad ? "FAIL" : "bit-identical");
    ref.release(); neu.release();
    return bad != 0;
}

template <class GA, class GB>
static void time_variant(const char* name, Planes& p, int w, int h, int pitch, int segs) {
    cudaFuncSetAttribute(k_stream<GA, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)StreamLayout<GA, false>::kSmem);
    cudaFuncSetAttribute(k_stream<GB, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)StreamLayout<GB, false>::kSmem);
    cudaFuncSetAttribute(k_stream<GA, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)StreamLayout<GA, false>::kSmem);
    cudaFuncSetAttribute(k_stream<GB, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)StreamLayout<GB, false>::kSmem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best[2] = {1e9f, 1e9f};
    for (int variant = 0; variant < 2; ++variant) {
        CascadeArgs a = variant == 0 ? args_a(p, w, h, pitch) : args_b(p, w, h, pitch, false);
        for (int rep = 0; rep < 5; ++rep) {
            cudaEventRecord(e0);
            cudaError_t e = variant == 0 ? launch_stream_t<GA>(a, 148, 0, segs) : launch_stream_t<GB>(a, 148, 0, segs);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            if (e != cudaSuccess || cudaGetLastError() != cudaSuccess) { printf("%s: launch failed\n", name); return; }
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep && ms < best[variant]) best[variant] = ms;
        }
    }
    int occA = 0, occB = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occA, k_stream<GA, true, false>, GA::THREADS, StreamLayout<GA, false>::kSmem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occB, k_stream<GB, false, false>, GB::THREADS, StreamLayout<GB, false>::kSmem);
    printf("  %-34s segs %2d: A %7.1f us  B %7.1f us  sum %7.1f us  (CTAs/SM %d / %d, smem %zu / %zu)\n", name, segs,
           best[0] * 1e3, best[1] * 1e3, (best[0] + best[1]) * 1e3, occA, occB, StreamLayout<GA, false>::kSmem, StreamLayout<GB, false>::kSmem);
}

static void time_size(int w, int h) {
    const int pitch = (w + 31) & ~31;
    const size_t n = (size_t)pitch * h;
    Planes p; p.alloc(n);
    std::vector<float> hbuf(n); for (size_t i = 0; i < n; ++i) hbuf[i] = (float)((i * 2654435761u) >> 24);
    cudaMemcpy(p.in, hbuf.data(), n * 4, cudaMemcpyHostToDevice);
    printf("timing %d x %d\n", w, h);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int variant = 0; variant < 2; ++variant) {
        CascadeArgs a = variant == 0 ? args_a(p, w, h, pitch) : args_b(p, w, h, pitch, false);
        float best = 1e9f;
        for (int rep = 0; rep < 5; ++rep) {
            cudaEventRecord(e0);
            if (variant == 0) launch_cascade_t<4, 5, 6>(a, 148, 0); else launch_cascade_t<8, 10, 0>(a, 148, 0);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep && ms < best) best = ms;
        }
        printf("  tile cascade %c: %7.1f us\n", variant ? 'B' : 'A', best * 1e3);
    }
    time_variant<StreamA, StreamB>("stream default (auto split)", p, w, h, pitch, 0);
    time_variant<StreamA, StreamB>("stream default, flattened split", p, w, h, pitch, 1);
    time_variant<StreamA, StreamB>("stream default, 8 bands", p, w, h, pitch, -8);
    time_variant<StreamA, StreamB>("stream default, 12 bands", p, w, h, pitch, -12);
    using A_k2 = StreamGeom<3, 4, 5, 6, 4, 96, 6, 4, true, 2>;
    using B_k1 = StreamGeom<2, 8, 10, 0, 4, 232, 12, 3, true, 1>;
    time_variant<A_k2, B_k1>("A K=2 4/SM PF6, B(232) K=1 PF12", p, w, h, pitch, 0);
    using A_c2b = StreamGeom<3, 4, 5, 6, 2, 96, 12, 4, true, 1>;
    using B_c2d = StreamGeom<2, 8, 10, 0, 2, 104, 6, 5, true, 2>;
    time_variant<A_c2b, B_c2d>("C=2: A 4/SM, B(104) 5/SM K=2", p, w, h, pitch, 0);
#ifdef QUICK
    p.release();
    return;
#endif
    using A_np = StreamGeom<3, 4, 5, 6, 4, 96, 12, 6, false>;
    using B_np = StreamGeom<2, 8, 10, 0, 4, 104, 12, 6, false>;
    time_variant<A_np, B_np>("stream no FFMA2", p, w, h, pitch, 0);
    using A_pf6 = StreamGeom<3, 4, 5, 6, 4, 96, 6, 7, true>;
    using B_pf6 = StreamGeom<2, 8, 10, 0, 4, 104, 6, 8, true>;
    time_variant<A_pf6, B_pf6>("stream PF 6, MINB 7 / 8", p, w, h, pitch, 0);
    using A_8 = StreamGeom<3, 4, 5, 6, 8, 224, 12, 4, true>;
    using B_232 = StreamGeom<2, 8, 10, 0, 4, 232, 12, 3, true>;
    time_variant<A_8, B_232>("stream A(8col,224) B(4col,232)", p, w, h, pitch, 0);
    using A_5 = StreamGeom<3, 4, 5, 6, 4, 96, 12, 5, true>;
    using B_5 = StreamGeom<2, 8, 10, 0, 4, 104, 12, 5, true>;
    time_variant<A_5, B_5>("stream MINB 5 / 5", p, w, h, pitch, 0);
    using A_4 = StreamGeom<3, 4, 5, 6, 4, 96, 12, 4, true>;
    using B_7 = StreamGeom<2, 8, 10, 0, 4, 104, 8, 7, true>;
    time_variant<A_4, B_7>("stream MINB 4 / 7 (PF 8)", p, w, h, pitch, 0);
    p.release();
}

int run(int argc, char** argv) {
    pyramid_init();
    {
        using CA = StreamGeom<3, 4, 5, 6, 2, 96, 12, 4, true, 1>;
        using CB = StreamGeom<2, 8, 10, 0, 2, 104, 6, 5, true, 2>;
        cudaFuncSetAttribute(k_stream<CA, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)StreamLayout<CA, false>::kSmem);
        cudaFuncSetAttribute(k_stream<CB, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)StreamLayout<CB, false>::kSmem);
    }
    if (argc > 1 && !strcmp(argv[1], "prof")) {
        const int w = 7680, h = 4320, pitch = 7680;
        Planes p; p.alloc((size_t)pitch * h);
        cudaMemset(p.in, 0, (size_t)pitch * h * 4);
        for (int rep = 0; rep < 2; ++rep) {
            CascadeArgs a = args_a(p, w, h, pitch), b = args_b(p, w, h, pitch, false);
            launch_stream_t<StreamA>(a, 148, 0, 0);
            launch_stream_t<StreamB>(b, 148, 0, 0);
            cudaDeviceSynchronize();
        }
        return 0;
    }
    int fails = 0;
    const int sizes[][3] = {{40, 30, 0}, {230, 50, 0}, {300, 200, 0}, {300, 200, 3}, {1000, 700, 0}, {1000, 700, 5},
                            {225, 131, 2}, {7, 9, 0}, {1, 1, 0}, {2, 300, 4}, {960, 540, 0}, {1920, 1080, 0}, {449, 64, 1},
                            {300, 200, -3}, {1000, 700, -2}, {225, 131, -1}, {2, 300, -4}, {7, 9, -2}, {100, 500, -7}};
#ifndef QUICK
    for (auto& s : sizes) { fails += check(s[0], s[1], s[2], false); fails += check(s[0], s[1], s[2], true); }
#endif
    printf("%s\n", fails ? "SOME CHECKS FAILED" : "all checks bit-identical");
    if (argc > 1 && !strcmp(argv[1], "check")) return fails;

    time_size(7680, 4320);
    time_size(3840, 2160);
    return fails;
}
}  // namespace sb
int main(int argc, char** argv) { return sb::run(argc, argv); }

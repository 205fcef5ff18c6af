// Per-phase cycle profile of the fused cascade kernels (development harness, not shipped).
#define SB_PHASE_TIMING
#ifndef PYR_SRC
#define PYR_SRC "../sift_project_b200/csrc/pyramid.cu"
#endif
#include PYR_SRC
#include <cstdio>
#include <cmath>
#include <vector>
#ifdef V2
#define LAUNCH(a1,a2,a3) launch_cascade_t<a1,a2,a3>(a, 148, 0)
#else
#define LAUNCH(a1,a2,a3) launch_cascade_t<a1,a2,a3>(a, 148, 0)
#endif
namespace sb {
static BlurTaps mk(double sigma) {
    BlurTaps t{}; int n = (int)ceil(3 * sigma) + 1; double k[64], tot = 0;
    for (int i = 0; i < n; ++i) { k[i] = exp(-i * i / (2 * sigma * sigma)); tot += i ? 2 * k[i] : k[i]; }
    t.radius = n - 1; for (int i = 0; i < n; ++i) t.w[i] = (float)(k[i] / tot); return t;
}
int run() {
    const int w = 7680, h = 4320, pitch = 7680;
    const size_t n = (size_t)pitch * h;
    float *in, *g[3], *d[3], *dec;
    cudaMalloc(&in, n * 4); for (int i = 0; i < 3; ++i) { cudaMalloc(&g[i], n * 4); cudaMalloc(&d[i], n * 4); }
    cudaMalloc(&dec, n);
    std::vector<float> hbuf(n); for (size_t i = 0; i < n; ++i) hbuf[i] = (float)((i * 2654435761u) >> 24);
    cudaMemcpy(in, hbuf.data(), n * 4, cudaMemcpyHostToDevice);
    pyramid_init();
    const double sig[6] = {1.6, 1.2262734984654078, 1.5450077936447955, 1.9465878414647133, 2.4525469969308156, 3.090015587289591};
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int variant = 0; variant < 2; ++variant) {
        CascadeArgs a{};
        a.in = in; a.w = w; a.h = h; a.pitch = pitch;
        if (variant == 0) {
            for (int i = 0; i < 3; ++i) { a.g[i] = g[i]; a.d[i] = d[i]; a.taps[i] = mk(sig[i + 1]); }
            a.dec = dec; a.dec_w = w / 2; a.dec_h = h / 2; a.dec_pitch = pitch / 2;
        } else {
            a.g[0] = a.g[1] = a.g[2] = nullptr; a.d[0] = d[0]; a.d[1] = d[1]; a.d[2] = nullptr;
            a.taps[0] = mk(sig[4]); a.taps[1] = mk(sig[5]); a.taps[2] = a.taps[1];
        }
        unsigned long long zero[16] = {0};
        float ms = 0;
        for (int rep = 0; rep < 4; ++rep) {
            cudaMemcpyToSymbol(g_phase, zero, sizeof zero);
            cudaEventRecord(e0);
            cudaError_t e = variant == 0 ? LAUNCH(4, 5, 6) : LAUNCH(8, 10, 0);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            if (e != cudaSuccess || cudaGetLastError() != cudaSuccess) { printf("launch failed\n"); return 1; }
            cudaEventElapsedTime(&ms, e0, e1);
        }
        unsigned long long ph[16]; cudaMemcpyFromSymbol(ph, g_phase, sizeof ph);
        const double tiles = ((w + CTW - 1) / CTW) * ((h + TH - 1) / TH);
        double tot = 0; for (int i = 0; i < 16; ++i) tot += ph[i];
        printf("variant %d: %.1f us, %.0f tiles, avg cycles/tile %.0f\n", variant, ms * 1e3, tiles, tot / tiles);
        for (int i = 0; i < 16; ++i) if (ph[i]) printf("   phase %2d: %8.0f cycles/tile (%4.1f%%)\n", i, ph[i] / tiles, 100.0 * ph[i] / tot);
    }
    return 0;
}
}  // namespace sb
int main() { return sb::run(); }

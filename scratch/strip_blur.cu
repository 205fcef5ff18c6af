// Experiment: single-level separable blur in "register scatter" form -- no shared memory, no
// barriers.  A warp streams down a 128-column strip; per input row each lane forms the horizontal
// pass of its 4 columns from an L1-cached window and scatters it into 2R+1 rotating vertical
// accumulators.  Compared against k_blur<R> (tile / shared-memory / gather form) for time and bits.
#include "../sift_project_b200/csrc/pyramid.cu"
#include <cstdio>
#include <cmath>
#include <vector>
namespace sb {
constexpr int SEG = 128;  // output rows per warp

template <int R>
__global__ void __launch_bounds__(128) k_strip(const float* __restrict__ in, float* __restrict__ out,
                                               float* __restrict__ dog, int w, int h, int pitch, const BlurTaps taps) {
    constexpr int HXR = (R + 3) & ~3;
    constexpr int NV = (4 + 2 * HXR) / 4;
    constexpr int P = 2 * R + 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int strip = blockIdx.x * 4 + warp;
    const int x0 = strip * 128 + 4 * lane;          // this lane's 4 columns
    const int y0 = blockIdx.y * SEG;
    if (x0 >= w || y0 >= h) return;
    const int y1 = min(y0 + SEG, h);
    const bool inner = x0 - HXR >= 0 && x0 + 4 + HXR <= w;
    float4 acc[P];                                   // acc[j]: pending output row with (row % P) == j
#pragma unroll
    for (int j = 0; j < P; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    // rows r = y0 - R .. y1 - 1 + R (clamped), processed in groups of P so that indices are static
    const int r_begin = y0 - R;
    const int n_rows = (y1 - y0) + 2 * R;
    for (int base = 0; base < n_rows; base += P) {
#pragma unroll
        for (int j = 0; j < P; ++j) {
            const int i = base + j;                  // i-th fed row; rr = r_begin + i
            if (i < n_rows) {
                const int rr = r_begin + i;
                const int rc = min(max(rr, 0), h - 1);
                const float* row = in + (size_t)rc * pitch;
                float v[4 * NV];
                if (inner) {
                    const float4* src = reinterpret_cast<const float4*>(row + x0 - HXR);
#pragma unroll
                    for (int k = 0; k < NV; ++k) {
                        const float4 t = __ldg(src + k);
                        v[4 * k] = t.x; v[4 * k + 1] = t.y; v[4 * k + 2] = t.z; v[4 * k + 3] = t.w;
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 4 * NV; ++k) v[k] = __ldg(row + min(max(x0 - HXR + k, 0), w - 1));
                }
                float4 hq;
                {
                    float o[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        float a = 0.f;
#pragma unroll
                        for (int u = R; u >= 1; --u) a = fmaf(taps.w[u], v[HXR + k - u] + v[HXR + k + u], a);
                        o[k] = fmaf(taps.w[0], v[HXR + k], a);
                    }
                    hq = make_float4(o[0], o[1], o[2], o[3]);
                }
                // fed row i contributes to outputs yo = y0 + (i - R) + d, d in [-R, R]; slot = (i - R + d) mod P
                // NOTE: the gather form adds top row first; to stay bit-identical the scatter must add
                // contributions in increasing input-row order, which it does (rows arrive in order).
#pragma unroll
                for (int d = -R; d <= R; ++d) {
                    const int slot = ((j - R + d) % P + 2 * P) % P;   // static after unrolling (base % P == 0)
                    const float wt = taps.w[d < 0 ? -d : d];
                    acc[slot].x = fmaf(wt, hq.x, acc[slot].x);
                    acc[slot].y = fmaf(wt, hq.y, acc[slot].y);
                    acc[slot].z = fmaf(wt, hq.z, acc[slot].z);
                    acc[slot].w = fmaf(wt, hq.w, acc[slot].w);
                }
                // output row yo = rr - R is complete after this row; it sits in slot (j - 2R) mod P = (j + 1) mod P
                const int yo = rr - R;
                constexpr int dummy = 0; (void)dummy;
                const int oslot = (j + 1) % P;
                if (yo >= y0 && yo < y1) {
                    const float4 o = acc[oslot];
                    *reinterpret_cast<float4*>(out + (size_t)yo * pitch + x0) = o;
                    const float4 c = __ldg(reinterpret_cast<const float4*>(in + (size_t)yo * pitch + x0));
                    *reinterpret_cast<float4*>(dog + (size_t)yo * pitch + x0) = make_float4(o.x - c.x, o.y - c.y, o.z - c.z, o.w - c.w);
                }
                acc[oslot] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
}

static BlurTaps mk(double sigma) {
    BlurTaps t{}; int n = (int)ceil(3 * sigma) + 1; double k[64], tot = 0;
    for (int i = 0; i < n; ++i) { k[i] = exp(-i * i / (2 * sigma * sigma)); tot += i ? 2 * k[i] : k[i]; }
    t.radius = n - 1; for (int i = 0; i < n; ++i) t.w[i] = (float)(k[i] / tot); return t;
}

template <int R>
int run_one(double sigma) {
    const int w = 7680, h = 4320, pitch = 7680;
    const size_t n = (size_t)pitch * h;
    float *in, *o1, *d1, *o2, *d2;
    cudaMalloc(&in, n * 4); cudaMalloc(&o1, n * 4); cudaMalloc(&d1, n * 4); cudaMalloc(&o2, n * 4); cudaMalloc(&d2, n * 4);
    std::vector<float> hb(n); for (size_t i = 0; i < n; ++i) hb[i] = (float)((i * 2654435761u) >> 24);
    cudaMemcpy(in, hb.data(), n * 4, cudaMemcpyHostToDevice);
    pyramid_init();
    BlurTaps t = mk(sigma);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms_a = 0, ms_b = 0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        launch_blur(in, o1, d1, nullptr, w, h, pitch, 0, 0, 0, t, 0);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms_a, e0, e1);
        dim3 grid((w + 511) / 512, (h + SEG - 1) / SEG);
        cudaEventRecord(e0);
        k_strip<R><<<grid, 128>>>(in, o2, d2, w, h, pitch, t);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms_b, e0, e1);
    }
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> a(n), b(n);
    cudaMemcpy(a.data(), o1, n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(b.data(), o2, n * 4, cudaMemcpyDeviceToHost);
    size_t bad = 0; for (size_t i = 0; i < n; ++i) if (a[i] != b[i]) ++bad;
    cudaMemcpy(a.data(), d1, n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(b.data(), d2, n * 4, cudaMemcpyDeviceToHost);
    size_t badd = 0; for (size_t i = 0; i < n; ++i) if (a[i] != b[i]) ++badd;
    printf("R=%d: k_blur %.1f us, k_strip %.1f us (%.2f TB/s), mismatches G %zu D %zu (%s)\n", R, ms_a * 1e3, ms_b * 1e3,
           12.0 * n / (ms_b * 1e-3) / 1e12, bad, badd, cudaGetErrorString(e));
    cudaFree(in); cudaFree(o1); cudaFree(d1); cudaFree(o2); cudaFree(d2);
    return 0;
}
}  // namespace sb
int main() {
    sb::run_one<4>(1.2262734984654078);
    sb::run_one<6>(1.9465878414647133);
    sb::run_one<10>(3.090015587289591);
    return 0;
}

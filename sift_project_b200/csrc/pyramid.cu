// Gaussian scale-space kernels (sm_100a).
//
// Replaces, per level, the reference's apply_gaussian_blur_fast -> apply_double_convolution_1d
// (image.cpp:156-238: horizontal then vertical clamp-to-edge symmetric convolution, renormalised),
// with subtract (image.cpp:30-36, DoG, sift.cpp:214-219) and resize_inter_nearest
// (image.cpp:41-55, next octave base from G[3], sift.cpp:195-196) fused into the epilogue, and
// convert_to_grayscale + resize_inter_bilinear (image.cpp:8-24, 62-88) as the input stage.
//
// HBM-bound by design: one read of G[i-1], one write of G[i], one write of D[i-1] per level.
#include "common.cuh"
#include "kernels.h"

namespace sb {

namespace {

constexpr int TW = 128;  // output tile width  (one warp row = 32 lanes x float4)
constexpr int TH = 64;   // output tile height
constexpr int NT = 256;  // threads per CTA

template <int R>
struct BlurGeom {
    static constexpr int HX = (R + 3) & ~3;    // x halo padded to a float4 boundary
    static constexpr int IW = TW + 2 * HX;     // staged tile width
    static constexpr int IH = TH + 2 * R;      // staged tile height
    static constexpr size_t kSmem = (size_t)(IH * IW + IH * TW) * sizeof(float);
};

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n" ::);
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
}

// One CTA produces a TW x TH tile of G[i] from G[i-1]:
//   stage (TW+2HX) x (TH+2R) input tile in shared memory (replicate padding == the reference's
//   index clamping, both passes), horizontal pass -> shared, vertical pass -> registers,
//   epilogue writes G[i], D[i-1] = G[i] - G[i-1] and (for G[3]) the decimated next-octave base.
template <int R>
__global__ void __launch_bounds__(NT, 2)
k_blur(const float* __restrict__ in, float* __restrict__ out, float* __restrict__ dog,
       float* __restrict__ dec, int w, int h, int pitch, int dec_w, int dec_h, int dec_pitch,
       const BlurTaps taps) {
    using G = BlurGeom<R>;
    extern __shared__ __align__(16) float smem[];
    float* s_in = smem;
    float* s_tmp = smem + G::IH * G::IW;

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int tx0 = blockIdx.x * TW, ty0 = blockIdx.y * TH;
    const int gx0 = tx0 - G::HX, gy0 = ty0 - R;

    const bool interior = gx0 >= 0 && gy0 >= 0 && (tx0 + TW + G::HX) <= w && (ty0 + TH + R) <= h;
    if (interior) {
        constexpr int V = G::IW / 4;
        for (int idx = tid; idx < G::IH * V; idx += NT) {
            int r = idx / V, c4 = idx - r * V;
            cp_async16(s_in + r * G::IW + 4 * c4, in + (size_t)(gy0 + r) * pitch + gx0 + 4 * c4);
        }
        cp_async_wait_all();
    } else {
        for (int idx = tid; idx < G::IH * G::IW; idx += NT) {
            int r = idx / G::IW, c = idx - r * G::IW;
            int gy = min(max(gy0 + r, 0), h - 1), gx = min(max(gx0 + c, 0), w - 1);
            s_in[idx] = __ldg(in + (size_t)gy * pitch + gx);
        }
    }
    __syncthreads();

    // ---- horizontal pass: each lane produces 4 adjacent outputs of one staged row ----
    for (int r = warp; r < G::IH; r += NT / 32) {
        const float4* row = reinterpret_cast<const float4*>(s_in + r * G::IW) + lane;
        float v[4 + 2 * G::HX];
#pragma unroll
        for (int k = 0; k < (4 + 2 * G::HX) / 4; ++k) {
            float4 q = row[k];
            v[4 * k + 0] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
        }
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float acc = 0.f;
#pragma unroll
            for (int u = R; u >= 1; --u)  // outermost (smallest) taps first
                acc = fmaf(taps.w[u], v[G::HX + k - u] + v[G::HX + k + u], acc);
            o[k] = fmaf(taps.w[0], v[G::HX + k], acc);
        }
        *reinterpret_cast<float4*>(s_tmp + r * TW + 4 * lane) = make_float4(o[0], o[1], o[2], o[3]);
    }
    __syncthreads();

    // ---- vertical pass: each lane owns 4 columns x 4 rows, streaming down the staged rows ----
    for (int g = warp; g < TH / 4; g += NT / 32) {
        float4 acc[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < 4 + 2 * R; ++r) {
            const float4 t = *reinterpret_cast<const float4*>(s_tmp + (4 * g + r) * TW + 4 * lane);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int d = (r - k - R) < 0 ? (k + R - r) : (r - k - R);
                if (d <= R) {
                    const float wt = taps.w[d];
                    acc[k].x = fmaf(wt, t.x, acc[k].x);
                    acc[k].y = fmaf(wt, t.y, acc[k].y);
                    acc[k].z = fmaf(wt, t.z, acc[k].z);
                    acc[k].w = fmaf(wt, t.w, acc[k].w);
                }
            }
        }
        const int gx = tx0 + 4 * lane;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int y = 4 * g + k, gy = ty0 + y;
            if (gy < h && gx < w) {
                const float4 o = acc[k];
                *reinterpret_cast<float4*>(out + (size_t)gy * pitch + gx) = o;
                if (dog != nullptr) {
                    const float4 c = *reinterpret_cast<const float4*>(s_in + (y + R) * G::IW + G::HX + 4 * lane);
                    *reinterpret_cast<float4*>(dog + (size_t)gy * pitch + gx) =
                        make_float4(o.x - c.x, o.y - c.y, o.z - c.z, o.w - c.w);
                }
                if (dec != nullptr && !(gy & 1)) {
                    const int dy = gy >> 1, dx = gx >> 1;
                    if (dy < dec_h) {
                        if (dx < dec_w) dec[(size_t)dy * dec_pitch + dx] = o.x;
                        if (dx + 1 < dec_w) dec[(size_t)dy * dec_pitch + dx + 1] = o.z;
                    }
                }
            }
        }
    }
}

// Input stage: (RGB ->) gray, optional 2x bilinear with the reference's right/bottom clamp.
// u8 gray input is exact in FP32 (weights 0, 1/2, 1/4 of integers); everything else is formed in
// FP64 with the reference's association order and rounded once.
template <typename T>
__global__ void k_prepare(const T* __restrict__ src, int sw, int sh, int ch, float* __restrict__ dst,
                          int dw, int dh, int dpitch, int doubled) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dw || y >= dh) return;
    auto gray = [&](int sx, int sy) -> double {
        const T* p = src + ((size_t)sy * sw + sx) * ch;
        if (ch == 1) return (double)p[0];
        return 0.2126 * (double)p[0] + 0.7152 * (double)p[1] + 0.0722 * (double)p[2];
    };
    double v;
    if (!doubled) {
        v = gray(x, y);
    } else {
        const int x0 = x >> 1, y0 = y >> 1;
        const int x1 = min(x0 + 1, sw - 1), y1 = min(y0 + 1, sh - 1);
        const double fx = (x & 1) ? 0.5 : 0.0, fy = (y & 1) ? 0.5 : 0.0;
        const double top = gray(x0, y0) * (1 - fx) + gray(x1, y0) * fx;
        const double bot = gray(x0, y1) * (1 - fx) + gray(x1, y1) * fx;
        v = top * (1 - fy) + bot * fy;
    }
    dst[(size_t)y * dpitch + x] = (float)v;
}

__global__ void k_prepare_gray_u8(const uint8_t* __restrict__ src, int sw, int sh,
                                  float* __restrict__ dst, int dw, int dh, int dpitch, int doubled) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dw || y >= dh) return;
    float v;
    if (!doubled) {
        v = (float)src[(size_t)y * sw + x];
    } else {
        const int x0 = x >> 1, y0 = y >> 1;
        const int x1 = min(x0 + 1, sw - 1), y1 = min(y0 + 1, sh - 1);
        const float fx = (x & 1) ? 0.5f : 0.0f, fy = (y & 1) ? 0.5f : 0.0f;
        const float a = src[(size_t)y0 * sw + x0], b = src[(size_t)y0 * sw + x1];
        const float c = src[(size_t)y1 * sw + x0], d = src[(size_t)y1 * sw + x1];
        const float top = a * (1.f - fx) + b * fx, bot = c * (1.f - fx) + d * fx;
        v = top * (1.f - fy) + bot * fy;  // exact: quarter-integers below 2^10
    }
    dst[(size_t)y * dpitch + x] = v;
}

template <int R>
cudaError_t launch_blur_r(const float* in, float* out, float* dog, float* dec, int w, int h,
                          int pitch, int dec_w, int dec_h, int dec_pitch, const BlurTaps& taps,
                          cudaStream_t s) {
    dim3 grid((w + TW - 1) / TW, (h + TH - 1) / TH);
    k_blur<R><<<grid, NT, BlurGeom<R>::kSmem, s>>>(in, out, dog, dec, w, h, pitch, dec_w, dec_h,
                                                  dec_pitch, taps);
    return cudaGetLastError();
}

template <int R>
cudaError_t init_blur_r() {
    return cudaFuncSetAttribute(k_blur<R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)BlurGeom<R>::kSmem);
}

}  // namespace

// Per-device one-time setup (opt-in to > 48 KB dynamic shared memory); called by context creation.
cudaError_t pyramid_init() {
    cudaError_t e;
#define SB_INIT(R) if ((e = init_blur_r<R>()) != cudaSuccess) return e;
    SB_INIT(1) SB_INIT(2) SB_INIT(3) SB_INIT(4) SB_INIT(5) SB_INIT(6) SB_INIT(7) SB_INIT(8)
    SB_INIT(9) SB_INIT(10) SB_INIT(11) SB_INIT(12)
#undef SB_INIT
    return cudaSuccess;
}

cudaError_t launch_blur(const float* in, float* out, float* dog, float* dec, int w, int h, int pitch,
                        int dec_w, int dec_h, int dec_pitch, const BlurTaps& taps, cudaStream_t s) {
#define SB_CASE(R) \
    case R: return launch_blur_r<R>(in, out, dog, dec, w, h, pitch, dec_w, dec_h, dec_pitch, taps, s);
    switch (taps.radius) {
        SB_CASE(1) SB_CASE(2) SB_CASE(3) SB_CASE(4) SB_CASE(5) SB_CASE(6) SB_CASE(7) SB_CASE(8)
        SB_CASE(9) SB_CASE(10) SB_CASE(11) SB_CASE(12)
        default: return cudaErrorInvalidValue;
    }
#undef SB_CASE
}

cudaError_t launch_prepare_u8(const uint8_t* src, int sw, int sh, int ch, float* dst, int dw, int dh,
                              int dpitch, int doubled, cudaStream_t s) {
    dim3 block(32, 8), grid((dw + 31) / 32, (dh + 7) / 8);
    if (ch == 1)
        k_prepare_gray_u8<<<grid, block, 0, s>>>(src, sw, sh, dst, dw, dh, dpitch, doubled);
    else
        k_prepare<uint8_t><<<grid, block, 0, s>>>(src, sw, sh, ch, dst, dw, dh, dpitch, doubled);
    return cudaGetLastError();
}

cudaError_t launch_prepare_f32(const float* src, int sw, int sh, int ch, float* dst, int dw, int dh,
                               int dpitch, int doubled, cudaStream_t s) {
    dim3 block(32, 8), grid((dw + 31) / 32, (dh + 7) / 8);
    k_prepare<float><<<grid, block, 0, s>>>(src, sw, sh, ch, dst, dw, dh, dpitch, doubled);
    return cudaGetLastError();
}

}  // namespace sb

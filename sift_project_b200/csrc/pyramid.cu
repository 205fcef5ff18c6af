// Gaussian scale-space kernels (sm_100a).
//
// Replaces, per level, the reference's apply_gaussian_blur_fast -> apply_double_convolution_1d
// (image.cpp:156-238: horizontal then vertical clamp-to-edge symmetric convolution, renormalised),
// with subtract (image.cpp:30-36, DoG, sift.cpp:214-219) and resize_inter_nearest
// (image.cpp:41-55, next octave base from G[3], sift.cpp:195-196) fused into the epilogue, and
// convert_to_grayscale + resize_inter_bilinear (image.cpp:8-24, 62-88) as the input stage.
//
// HBM-bound by design: one read of G[i-1], one write of G[i], one write of D[i-1] per level.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <type_traits>
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

// ------------------------------------------------------------------------------------------
// Fused cascade: several consecutive levels of one octave per tile, intermediate levels staged in
// shared memory only.  Two instantiations cover the reference's default scale space
// (sift.cpp:143-155 -> radii 4,5,6,8,10):
//   k_cascade<4,5,6>  G0 -> G1,G2,G3 (stored), D0,D1,D2, decimated next-octave base
//   k_cascade<8,10,0> G3 -> (G4, G5 stay on chip) -> D3, D4
// HBM traffic per octave pixel: 4 + 24 and 4 + 8 bytes (the algorithmic minimum is 36 + 1).
// Tile = 128 x 64 outputs; the staged input carries the summed halo (rounded so that every
// shared-memory row stays float4-aligned); each level is a horizontal pass smem -> smem and a
// vertical pass smem -> smem (or -> registers -> HBM for the last level).
// Clamp-to-edge semantics of every level (image.cpp:170-184): out-of-image positions of an
// intermediate level must hold that level's edge value, not a blur of replicated input, so
// border tiles re-replicate the edge after each intermediate level.
// ------------------------------------------------------------------------------------------
constexpr int CT = 512;  // threads per CTA of the fused kernels
constexpr int CCT = 384; // threads per cascade CTA: the 64x64 tile's pass sizes (2304/528/1760/380/1216/256 items)
                         // split into whole rounds of 384 far better than of 512
constexpr int CTW = 64;  // cascade tile width: 64 x 64 tiles need ~110 KB, so TWO CTAs share an SM and one
                         // CTA's load / store phases overlap the other's arithmetic
__host__ __device__ constexpr int ru4(int v) { return (v + 3) & ~3; }
__host__ __device__ constexpr int ru2(int v) { return (v + 1) & ~1; }

template <int R1, int R2, int R3, int TWP, int THP>
struct CascadeGeom {
    static constexpr int NL = R3 > 0 ? 3 : 2;
    static constexpr int HX2 = NL == 3 ? ru4(R3) : 0, HY2 = NL == 3 ? ru2(R3) : 0;  // halo kept around level 2
    static constexpr int HX1 = ru4(HX2 + R2), HY1 = ru2(HY2 + R2);                  // ... around level 1
    static constexpr int HX0 = ru4(HX1 + R1), HY0 = ru2(HY1 + R1);                  // ... around the input
    static constexpr int W0 = TWP + 2 * HX0, H0 = THP + 2 * HY0;
    static constexpr int W1 = TWP + 2 * HX1, H1 = THP + 2 * HY1;
    static constexpr int W2 = TWP + 2 * HX2, H2 = THP + 2 * HY2;
    static constexpr int A_FLOATS = W0 * H0;   // input, later level 2
    static constexpr int T_FLOATS = W1 * H0;   // horizontal-pass scratch (largest: level 1)
    static constexpr int B_FLOATS = W1 * H1;   // level 1
    static constexpr size_t kSmem = (size_t)(A_FLOATS + T_FLOATS + B_FLOATS) * sizeof(float);
};

struct CascadePlanes {
    const float* in;      // level 0 of this launch (G0 or G3)
    float* g[3];          // level outputs (nullable: not stored)
    float* d[3];          // d[l] = level(l+1) - level(l)
    float* dec;           // decimated copy of the LAST level (nullable)
    int w, h, pitch;
    int dec_w, dec_h, dec_pitch;
};
struct CascadeArgs : CascadePlanes {
    BlurTaps taps[3];
};

// horizontal pass: out[r][4q..4q+3] (width OUT_W) from in (width IN_W); OFF = x offset of the output
// region inside the input region (a multiple of 4)
template <int NTH, int R, int IN_W, int OUT_W, int OFF>
__device__ __forceinline__ void cascade_hpass(const float* __restrict__ in, float* __restrict__ out, int rows,
                                              const BlurTaps& taps) {
    constexpr int HXR = ru4(R);
    constexpr int Q = OUT_W / 4;
    for (int idx = threadIdx.x; idx < rows * Q; idx += NTH) {
        const int r = idx / Q, q = idx - r * Q;
        const float4* src = reinterpret_cast<const float4*>(in + r * IN_W + OFF + 4 * q - HXR);
        float v[4 + 2 * HXR];
#pragma unroll
        for (int k = 0; k < (4 + 2 * HXR) / 4; ++k) {
            const float4 t = src[k];
            v[4 * k] = t.x; v[4 * k + 1] = t.y; v[4 * k + 2] = t.z; v[4 * k + 3] = t.w;
        }
        // (packed FADD2 / FFMA2 measured slower here: the shifted windows are not register-pair aligned)
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float acc = 0.f;
#pragma unroll
            for (int u = R; u >= 1; --u) acc = fmaf(taps.w[u], v[HXR + k - u] + v[HXR + k + u], acc);
            o[k] = fmaf(taps.w[0], v[HXR + k], acc);
        }
        *reinterpret_cast<float4*>(out + r * OUT_W + 4 * q) = make_float4(o[0], o[1], o[2], o[3]);
    }
}

// vertical pass: 4 columns x 4 rows per item, streaming down 4 + 2R scratch rows; emit(r, q, acc[4])
template <int NTH, int R, int W, typename Emit>
__device__ __forceinline__ void cascade_vpass(const float* __restrict__ tmp, int out_rows, const BlurTaps& taps,
                                              Emit emit) {
    constexpr int Q = W / 4;
    // packed FP32 FMAs (FFMA2, sm_100): two IEEE fmas per instruction -- same results, half the issue slots
    float2 w2[R + 1];
#pragma unroll
    for (int d = 0; d <= R; ++d) w2[d] = make_float2(taps.w[d], taps.w[d]);
    for (int idx = threadIdx.x; idx < (out_rows / 4) * Q; idx += NTH) {
        const int g = idx / Q, q = idx - g * Q;
        float2 lo[4], hi[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) lo[k] = hi[k] = make_float2(0.f, 0.f);
#pragma unroll
        for (int r = 0; r < 4 + 2 * R; ++r) {
            const float4 t = *reinterpret_cast<const float4*>(tmp + (4 * g + r) * W + 4 * q);
            const float2 tl = make_float2(t.x, t.y), th = make_float2(t.z, t.w);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int d = (r - k - R) < 0 ? (k + R - r) : (r - k - R);
                if (d <= R) {
                    lo[k] = __ffma2_rn(w2[d], tl, lo[k]);
                    hi[k] = __ffma2_rn(w2[d], th, hi[k]);
                }
            }
        }
        float4 acc[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k] = make_float4(lo[k].x, lo[k].y, hi[k].x, hi[k].y);
        emit(4 * g, q, acc);
    }
}

// re-replicate the image edge into the out-of-image part of a staged level (border tiles only)
template <int NTH, int W, int H>
__device__ __forceinline__ void cascade_fix_edges(float* __restrict__ buf, int gx0, int gy0, int w, int h) {
    for (int idx = threadIdx.x; idx < W * H; idx += NTH) {
        const int r = idx / W, c = idx - r * W;
        const int gx = gx0 + c, gy = gy0 + r;
        const int cx = min(max(gx, 0), w - 1), cy = min(max(gy, 0), h - 1);
        if (cx != gx || cy != gy) buf[idx] = buf[(cy - gy0) * W + (cx - gx0)];
    }
}

#ifdef SB_PHASE_TIMING   // scratch/casc_prof.cu: per-phase cycle counters (thread 0 of every CTA)
__device__ unsigned long long g_phase[16];
#define SB_PHASE(p)                                                            \
    do {                                                                       \
        if (threadIdx.x == 0) {                                                \
            const long long now__ = clock64();                                 \
            atomicAdd(&g_phase[p], (unsigned long long)(now__ - phase_t0__));  \
            phase_t0__ = now__;                                                \
        }                                                                      \
    } while (0)
#define SB_PHASE_INIT long long phase_t0__ = clock64();
#else
#define SB_PHASE(p)
#define SB_PHASE_INIT
#endif

// TWP x THP = tile, NTP = threads, MINB = CTAs per SM.  64 x 64 / 384 / 2 for large octaves; 32 x 32 / 128 / 4
// for octaves that would not fill the GPU with 64 x 64 tiles (a lone CTA takes 10-18 us per tile).
// One tile.  `a`: the planes of this launch; `taps`: its three tap sets (kernel-parameter space: the FMAs take them
// as constant-bank operands); (bx, by): the tile.  COHERENT: the input may have been written by another CTA of the
// SAME launch (k_tail), so it must not come through the non-coherent (LDG.NC) path.
template <int R1, int R2, int R3, int TWP, int THP, int NTP, bool COHERENT>
__device__ __forceinline__ void cascade_tile(const CascadePlanes& a, const BlurTaps* __restrict__ taps, int bx, int by,
                                             float* __restrict__ smem) {
    using G = CascadeGeom<R1, R2, R3, TWP, THP>;
    SB_PHASE_INIT
    float* sA = smem;
    float* sT = sA + G::A_FLOATS;
    float* sB = sT + G::T_FLOATS;
    const int tid = threadIdx.x;
    const int w = a.w, h = a.h, pitch = a.pitch;
    const int tx0 = bx * TWP, ty0 = by * THP;
    const int gx0 = tx0 - G::HX0, gy0 = ty0 - G::HY0;
    const bool interior = gx0 >= 0 && gy0 >= 0 && (tx0 + TWP + G::HX0) <= w && (ty0 + THP + G::HY0) <= h;

    // ---- stage the input tile (replicate padding == the reference's index clamping) ----
    if (interior) {
        constexpr int V = G::W0 / 4;
        for (int idx = tid; idx < G::H0 * V; idx += NTP) {
            const int r = idx / V, c4 = idx - r * V;
            cp_async16(sA + r * G::W0 + 4 * c4, a.in + (size_t)(gy0 + r) * pitch + gx0 + 4 * c4);
        }
        cp_async_wait_all();
    } else {
        for (int idx = tid; idx < G::H0 * G::W0; idx += NTP) {
            const int r = idx / G::W0, c = idx - r * G::W0;
            const int gy = min(max(gy0 + r, 0), h - 1), gx = min(max(gx0 + c, 0), w - 1);
            const float* src = a.in + (size_t)gy * pitch + gx;
            sA[idx] = COHERENT ? __ldcg(src) : __ldg(src);
        }
    }
    __syncthreads();
    SB_PHASE(0);

    // store helper: 4 rows x 4 columns of a level and of its DoG against `prev` (shared memory)
    auto store_level = [&](float* gout, float* dout, const float* prev, int prev_w, int prev_ox, int prev_oy,
                           int y, int q, const float4 (&acc)[4], bool decimate) {
        const int gx = tx0 + 4 * q;
        if (gx >= w) return;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int gy = ty0 + y + k;
            if (gy >= h) break;
            const float4 o = acc[k];
            if (gout != nullptr) *reinterpret_cast<float4*>(gout + (size_t)gy * pitch + gx) = o;
            const float4 c = *reinterpret_cast<const float4*>(prev + (prev_oy + y + k) * prev_w + prev_ox + 4 * q);
            *reinterpret_cast<float4*>(dout + (size_t)gy * pitch + gx) =
                make_float4(o.x - c.x, o.y - c.y, o.z - c.z, o.w - c.w);
            if (decimate && a.dec != nullptr && !(gy & 1)) {
                const int dy = gy >> 1, dx = gx >> 1;
                if (dy < a.dec_h) {
                    if (dx + 1 < a.dec_w)
                        *reinterpret_cast<float2*>(a.dec + (size_t)dy * a.dec_pitch + dx) = make_float2(o.x, o.z);
                    else if (dx < a.dec_w)
                        a.dec[(size_t)dy * a.dec_pitch + dx] = o.x;
                }
            }
        }
    };

    // ---- level 1: sA -> sT -> sB ----
    cascade_hpass<NTP, R1, G::W0, G::W1, G::HX0 - G::HX1>(sA, sT, G::H0, taps[0]);
    __syncthreads();
    SB_PHASE(1);
    cascade_vpass<NTP, R1, G::W1>(sT + (G::HY0 - G::HY1 - R1) * G::W1, G::H1, taps[0],
                             [&](int y, int q, const float4 (&acc)[4]) {
#pragma unroll
                                 for (int k = 0; k < 4; ++k)
                                     *reinterpret_cast<float4*>(sB + (y + k) * G::W1 + 4 * q) = acc[k];
                             });
    __syncthreads();
    SB_PHASE(2);
    // emit the centre of level 1 (+ DoG against the input centre)
    for (int idx = tid; idx < (THP / 4) * (TWP / 4); idx += NTP) {
        const int g = idx / (TWP / 4), q = idx - g * (TWP / 4);
        float4 acc[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            acc[k] = *reinterpret_cast<const float4*>(sB + (G::HY1 + 4 * g + k) * G::W1 + G::HX1 + 4 * q);
        store_level(a.g[0], a.d[0], sA, G::W0, G::HX0, G::HY0, 4 * g, q, acc, false);
    }
    // interior tiles go straight on: the next pass only reads what the emission reads
    if (!interior) {
        __syncthreads();
        cascade_fix_edges<NTP, G::W1, G::H1>(sB, tx0 - G::HX1, ty0 - G::HY1, w, h);
        __syncthreads();
    }
    SB_PHASE(4);

    if (G::NL == 3) {
        // ---- level 2: sB -> sT -> sA (the input is dead by now) ----
        cascade_hpass<NTP, R2, G::W1, G::W2, G::HX1 - G::HX2>(sB, sT, G::H1, taps[1]);
        __syncthreads();
    SB_PHASE(5);
        cascade_vpass<NTP, R2, G::W2>(sT + (G::HY1 - G::HY2 - R2) * G::W2, G::H2, taps[1],
                                 [&](int y, int q, const float4 (&acc)[4]) {
#pragma unroll
                                     for (int k = 0; k < 4; ++k)
                                         *reinterpret_cast<float4*>(sA + (y + k) * G::W2 + 4 * q) = acc[k];
                                 });
        __syncthreads();
    SB_PHASE(6);
        for (int idx = tid; idx < (THP / 4) * (TWP / 4); idx += NTP) {
            const int g = idx / (TWP / 4), q = idx - g * (TWP / 4);
            float4 acc[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                acc[k] = *reinterpret_cast<const float4*>(sA + (G::HY2 + 4 * g + k) * G::W2 + G::HX2 + 4 * q);
            store_level(a.g[1], a.d[1], sB, G::W1, G::HX1, G::HY1, 4 * g, q, acc, false);
        }
        if (!interior) {
            __syncthreads();
            cascade_fix_edges<NTP, G::W2, G::H2>(sA, tx0 - G::HX2, ty0 - G::HY2, w, h);
            __syncthreads();
        }
    SB_PHASE(8);
        // ---- level 3: sA -> sT -> registers -> HBM ----
        cascade_hpass<NTP, (R3 > 0 ? R3 : 1), G::W2, TWP, G::HX2>(sA, sT, G::H2, taps[2]);
        __syncthreads();
    SB_PHASE(9);
        cascade_vpass<NTP, (R3 > 0 ? R3 : 1), TWP>(sT + (G::HY2 - R3) * TWP, THP, taps[2],
                                           [&](int y, int q, const float4 (&acc)[4]) {
                                               store_level(a.g[2], a.d[2], sA, G::W2, G::HX2, G::HY2, y, q, acc, true);
                                           });
    } else {
        // ---- two-level variant: level 2 is the last: sB -> sT -> registers -> HBM ----
        cascade_hpass<NTP, R2, G::W1, TWP, G::HX1>(sB, sT, G::H1, taps[1]);
        __syncthreads();
    SB_PHASE(10);
        cascade_vpass<NTP, R2, TWP>(sT + (G::HY1 - R2) * TWP, THP, taps[1],
                              [&](int y, int q, const float4 (&acc)[4]) {
                                  store_level(a.g[1], a.d[1], sB, G::W1, G::HX1, G::HY1, y, q, acc, true);
                              });
    }
    SB_PHASE(15);
}

template <int R1, int R2, int R3, int TWP, int THP, int NTP, int MINB>
__global__ void __launch_bounds__(NTP, MINB) k_cascade(const CascadeArgs a) {
    extern __shared__ __align__(16) float smem[];
    cascade_tile<R1, R2, R3, TWP, THP, NTP, false>(a, a.taps, blockIdx.x, blockIdx.y, smem);
}

// ------------------------------------------------------------------------------------------
// The tail of the pyramid in ONE launch.  Octaves of at most one 32 x 32 tile per SM are latency-bound: a launch
// lasts one tile's serial phases (~5 us) plus the launch gap, two launches per octave, seven such octaves below a
// 4K image, and each octave needs the previous one's decimated G3 (sift.cpp:187-199).  k_tail runs the same tile
// code (cascade_tile<.., 32, 32, 512>, hence the same bits) for all of them from a ticket counter: CTAs draw work
// items in an order in which every item depends only on items with LOWER tickets -- group g = the first-kernel
// tiles of tail octave g, then the second-kernel tiles of octave g - 1 -- so a waiting CTA only ever waits for
// CTAs that are already running: no co-residency requirement (not a cooperative launch), no deadlock whatever
// else shares the GPU.  Hand-over between octaves: per-octave "tiles done" counters, release = CTA barrier +
// __threadfence + atomicAdd by one thread, acquire = ld.acquire polled by one thread + CTA barrier; consumers read
// the fresh planes through L2 (cp.async.cg / ld.cg).
// ------------------------------------------------------------------------------------------
constexpr int kMaxTail = 12;
constexpr int TAIL_NT = 512;
struct TailOct {
    CascadePlanes a, b;   // G0 -> G1..G3, D0..D2, next base | G3 -> (G4, G5) D3, D4
    int tiles_x, tiles;
};
struct TailArgs {
    int n, total;
    int begin[kMaxTail + 2];   // first ticket of group g
    TailOct oct[kMaxTail];
    BlurTaps taps_a[3], taps_b[3];
    int* sync;                 // [0] ticket counter, [1 + o] first-kernel tiles of tail octave o done; zeroed per call
};

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(TAIL_NT, 2) k_tail(const TailArgs t) {
    extern __shared__ __align__(16) float smem[];
    __shared__ int s_ticket;
    for (;;) {
        __syncthreads();   // the previous item's shared-memory reads and its s_ticket reads are over
        if (threadIdx.x == 0) s_ticket = atomicAdd(t.sync, 1);
        __syncthreads();
        const int k = s_ticket;
        if (k >= t.total) return;
        int g = 0;
        while (k >= t.begin[g + 1]) ++g;
        int i = k - t.begin[g];
        const int na = g < t.n ? t.oct[g].tiles : 0;
        const bool first = i < na;
        const int o = first ? g : g - 1;
        if (!first) i -= na;
        const int dep = first ? o - 1 : o;   // the octave whose first-kernel tiles this item reads
        if (dep >= 0) {
            if (threadIdx.x == 0) {
                const int need = t.oct[dep].tiles;
                int spins = 0;
                while (ld_acquire_gpu(t.sync + 1 + dep) < need) {
                    __nanosleep(32);
                    if (++spins > (1 << 21)) __trap();   // seconds: a lost hand-over must fail, not hang
                }
            }
            __syncthreads();
        }
        const int tiles_x = t.oct[o].tiles_x;
        const int by = i / tiles_x, bx = i - by * tiles_x;
        if (first) {
            const CascadePlanes p = t.oct[o].a;
            cascade_tile<4, 5, 6, 32, 32, TAIL_NT, true>(p, t.taps_a, bx, by, smem);
            __syncthreads();
            if (threadIdx.x == 0) {
                __threadfence();
                atomicAdd(t.sync + 1 + o, 1);
            }
        } else {
            const CascadePlanes p = t.oct[o].b;
            cascade_tile<8, 10, 0, 32, 32, TAIL_NT, true>(p, t.taps_b, bx, by, smem);
        }
    }
}

constexpr size_t kTailSmem = CascadeGeom<4, 5, 6, 32, 32>::kSmem > CascadeGeom<8, 10, 0, 32, 32>::kSmem
                                 ? CascadeGeom<4, 5, 6, 32, 32>::kSmem
                                 : CascadeGeom<8, 10, 0, 32, 32>::kSmem;

template <int R1, int R2, int R3>
cudaError_t launch_cascade_t(const CascadeArgs& a, int sm_count, cudaStream_t s) {
    const int big_tiles = ((a.w + CTW - 1) / CTW) * ((a.h + TH - 1) / TH);
    const int small_tiles = ((a.w + 31) / 32) * ((a.h + 31) / 32);
    if (big_tiles >= 2 * sm_count) {
        dim3 grid((a.w + CTW - 1) / CTW, (a.h + TH - 1) / TH);
        k_cascade<R1, R2, R3, CTW, TH, CCT, 2><<<grid, CCT, CascadeGeom<R1, R2, R3, CTW, TH>::kSmem, s>>>(a);
    } else if (small_tiles > sm_count) {
        dim3 grid((a.w + 31) / 32, (a.h + 31) / 32);
        k_cascade<R1, R2, R3, 32, 32, 128, 4><<<grid, 128, CascadeGeom<R1, R2, R3, 32, 32>::kSmem, s>>>(a);
    } else {
        // at most one tile per SM: the kernel time is one tile's serial phases, so spread each over 512 threads
        dim3 grid((a.w + 31) / 32, (a.h + 31) / 32);
        k_cascade<R1, R2, R3, 32, 32, 512, 1><<<grid, 512, CascadeGeom<R1, R2, R3, 32, 32>::kSmem, s>>>(a);
    }
    return cudaGetLastError();
}

#include "stream.cuh"

PFN_cuTensorMapEncodeTiled g_stream_encode = nullptr;

cudaError_t stream_make_map(CUtensorMap* map, const float* base, int w, int h, int pitch, int boxw) {
    const cuuint64_t dims[2] = {(cuuint64_t)w, (cuuint64_t)h};
    const cuuint64_t strides[1] = {(cuuint64_t)pitch * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)boxw, 1u};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = g_stream_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides,
                                       box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

// SIFT_B200_STREAM_TMA=0 selects the cp.async input path of the streaming kernels (comparison runs)
bool stream_use_tma() {
    static const bool on = getenv("SIFT_B200_STREAM_TMA") && atoi(getenv("SIFT_B200_STREAM_TMA")) == 1;   // off until verified
    return on;
}

// ------------------------------------------------------------------------------------------
// Fused input stage for 8-bit gray images: resize_inter_bilinear(2,2) (image.cpp:62-88, optional)
// + the initial blur (sift.cpp:124, radius 4) in one pass: the up-sampled tile is formed in shared
// memory straight from the u8 pixels (exact in FP32: quarter-integers), blurred, and only G[0][0]
// is written.  Reads 1 byte, writes 16 bytes per input pixel instead of 4 + 32 + 16.
// ------------------------------------------------------------------------------------------
constexpr int IN_R = 4;
template <int TWI, int THI>
struct InputGeom {
    static constexpr int W0 = TWI + 2 * IN_R, H0 = THI + 2 * IN_R;
    static constexpr size_t kSmem = (size_t)(W0 * H0 + TWI * H0) * sizeof(float);
};

// convert_to_grayscale (image.cpp:14-22): 0.2126 R + 0.7152 G + 0.0722 B in FP64, products and sums rounded one by
// one like the reference's (no FMA contraction), so the gray value is the reference's own double.
__device__ __forceinline__ double gray_bt709(double r, double g, double b) {
    return __dadd_rn(__dadd_rn(__dmul_rn(0.2126, r), __dmul_rn(0.7152, g)), __dmul_rn(0.0722, b));
}

// TWI x THI = output tile, NTI = threads, MINBI = CTAs per SM, CH = channels of the u8 input (1 gray, 3 RGB).
// Gray pixels are integers: the 2x bilinear values are quarter-integers, exact in FP32.  RGB pixels go through the
// FP64 gray value and the FP64 bilinear form of k_prepare (rounded to FP32 once, after the centre is subtracted), so
// the fused and the unfused input paths give the same bytes.
template <bool DOUBLED, int TWI, int THI, int NTI, int MINBI, int CH>
__global__ void __launch_bounds__(NTI, MINBI)
k_input_u8(const uint8_t* __restrict__ src, int sw, int sh, float* __restrict__ dst, int w, int h, int pitch,
           const BlurTaps taps, const float centre) {
    constexpr int IN_W0 = InputGeom<TWI, THI>::W0, IN_H0 = InputGeom<TWI, THI>::H0;
    constexpr int TW = TWI, TH = THI, CT = NTI;
    using V = typename std::conditional<CH == 1, float, double>::type;
    extern __shared__ __align__(16) float smem[];
    float* sA = smem;
    float* sT = smem + IN_W0 * IN_H0;
    const int tx0 = blockIdx.x * TW, ty0 = blockIdx.y * TH;
    const bool inside = tx0 - IN_R >= 0 && ty0 - IN_R >= 0 && tx0 + TW + IN_R <= w && ty0 + TH + IN_R <= h;
    auto px = [&](int x, int y) -> V {   // gray level of source pixel (x, y)
        // (32-bit indices: the context's image has < 2^31 bytes even as RGB; checked at creation)
        if (CH == 1) return (V)__ldg(src + (y * sw + x));
        const uint8_t* p = src + (y * sw + x) * 3;
        return (V)gray_bt709((double)__ldg(p), (double)__ldg(p + 1), (double)__ldg(p + 2));
    };
    const V c = (V)centre, half = (V)0.5;
    if (DOUBLED && inside) {
        // interior tiles: one source pixel neighbourhood -> a 2x2 block of the up-sampled tile
        // (p, (p+q)/2; (p+t)/2, ((p+q)/2+(t+u)/2)/2 -- resize_inter_bilinear's values, image.cpp:64-86)
        const int sx0 = (tx0 - IN_R) >> 1, sy0 = (ty0 - IN_R) >> 1;   // tile origin is even
        for (int idx = threadIdx.x; idx < (IN_W0 / 2) * (IN_H0 / 2); idx += CT) {
            const int br = idx / (IN_W0 / 2), bc = idx - br * (IN_W0 / 2);
            const int x0 = sx0 + bc, y0 = sy0 + br;
            const int x1 = min(x0 + 1, sw - 1), y1 = min(y0 + 1, sh - 1);
            const V p = px(x0, y0), q = px(x1, y0), t = px(x0, y1), u = px(x1, y1);
            const V top = p * half + q * half, bot = t * half + u * half;
            *reinterpret_cast<float2*>(sA + (2 * br) * IN_W0 + 2 * bc) = make_float2((float)(p - c), (float)(top - c));
            *reinterpret_cast<float2*>(sA + (2 * br + 1) * IN_W0 + 2 * bc) =
                make_float2((float)((p * half + t * half) - c), (float)((top * half + bot * half) - c));
        }
    } else
    for (int idx = threadIdx.x; idx < IN_W0 * IN_H0; idx += CT) {
        const int r = idx / IN_W0, cc = idx - r * IN_W0;
        const int X = min(max(tx0 - IN_R + cc, 0), w - 1), Y = min(max(ty0 - IN_R + r, 0), h - 1);
        V v;
        if (DOUBLED) {
            const int x0 = X >> 1, y0 = Y >> 1;
            const int x1 = min(x0 + 1, sw - 1), y1 = min(y0 + 1, sh - 1);
            const V fx = (X & 1) ? half : (V)0, fy = (Y & 1) ? half : (V)0;
            const V p = px(x0, y0), q = px(x1, y0), t = px(x0, y1), u = px(x1, y1);
            const V top = p * ((V)1 - fx) + q * fx, bot = t * ((V)1 - fx) + u * fx;
            v = top * ((V)1 - fy) + bot * fy;
        } else {
            v = px(X, Y);
        }
        sA[idx] = (float)(v - c);
    }
    __syncthreads();
    cascade_hpass<CT, IN_R, IN_W0, TW, IN_R>(sA, sT, IN_H0, taps);
    __syncthreads();
    cascade_vpass<CT, IN_R, TW>(sT, TH, taps, [&](int y, int q, const float4 (&acc)[4]) {
        const int gx = tx0 + 4 * q;
        if (gx >= w) return;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int gy = ty0 + y + k;
            if (gy < h) *reinterpret_cast<float4*>(dst + (size_t)gy * pitch + gx) = acc[k];
        }
    });
}

// Input stage: (RGB ->) gray, optional 2x bilinear with the reference's right/bottom clamp.
// u8 gray input is exact in FP32 (weights 0, 1/2, 1/4 of integers); everything else is formed in
// FP64 with the reference's association order and rounded once.
template <typename T>
__global__ void k_prepare(const T* __restrict__ src, int sw, int sh, int ch, float* __restrict__ dst,
                          int dw, int dh, int dpitch, int doubled, float centre_const,
                          const float* __restrict__ centre_range) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dw || y >= dh) return;
    const double centre = centre_range ? 0.5 * ((double)centre_range[0] + (double)centre_range[1]) : (double)centre_const;
    auto gray = [&](int sx, int sy) -> double {
        const T* p = src + ((size_t)sy * sw + sx) * ch;
        if (ch == 1) return (double)p[0];
        return gray_bt709((double)p[0], (double)p[1], (double)p[2]);
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
    dst[(size_t)y * dpitch + x] = (float)(v - centre);
}

__global__ void k_prepare_gray_u8(const uint8_t* __restrict__ src, int sw, int sh,
                                  float* __restrict__ dst, int dw, int dh, int dpitch, int doubled, float centre) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dw || y >= dh) return;
    float v;
    if (!doubled) {
        v = (float)src[(size_t)y * sw + x] - centre;
    } else {
        const int x0 = x >> 1, y0 = y >> 1;
        const int x1 = min(x0 + 1, sw - 1), y1 = min(y0 + 1, sh - 1);
        const float fx = (x & 1) ? 0.5f : 0.0f, fy = (y & 1) ? 0.5f : 0.0f;
        const float a = src[(size_t)y0 * sw + x0] - centre, b = src[(size_t)y0 * sw + x1] - centre;
        const float c = src[(size_t)y1 * sw + x0] - centre, d = src[(size_t)y1 * sw + x1] - centre;
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
#define SB_IN_ATTR(D, CH)                                                                                       \
    if ((e = cudaFuncSetAttribute(k_input_u8<D, 128, 32, 512, 3, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                  (int)InputGeom<128, 32>::kSmem)) != cudaSuccess) return e;
    SB_IN_ATTR(true, 1) SB_IN_ATTR(false, 1) SB_IN_ATTR(true, 3) SB_IN_ATTR(false, 3)
#undef SB_IN_ATTR
#define SB_CASC_ATTR(R1, R2, R3, TWP, THP, NTP, MINB)                                                       \
    if ((e = cudaFuncSetAttribute(k_cascade<R1, R2, R3, TWP, THP, NTP, MINB>,                             \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize,                            \
                                  (int)CascadeGeom<R1, R2, R3, TWP, THP>::kSmem)) != cudaSuccess) return e;
    SB_CASC_ATTR(4, 5, 6, CTW, TH, CCT, 2) SB_CASC_ATTR(8, 10, 0, CTW, TH, CCT, 2)
    SB_CASC_ATTR(4, 5, 6, 32, 32, 128, 4) SB_CASC_ATTR(8, 10, 0, 32, 32, 128, 4)
    SB_CASC_ATTR(4, 5, 6, 32, 32, 512, 1) SB_CASC_ATTR(8, 10, 0, 32, 32, 512, 1)
#undef SB_CASC_ATTR
    if ((e = cudaFuncSetAttribute(k_tail, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTailSmem)) != cudaSuccess)
        return e;
    if (g_stream_encode == nullptr) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if ((e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q)) != cudaSuccess) return e;
        if (q != cudaDriverEntryPointSuccess || fn == nullptr) return cudaErrorNotSupported;
        g_stream_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fn);
    }
#define SB_STREAM_ATTR(G, SG, TM)                                                                          \
    if ((e = cudaFuncSetAttribute(k_stream<G, SG, TM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)StreamLayout<G, TM>::kSmem)) != \
        cudaSuccess)                                                                                      \
        return e;
    SB_STREAM_ATTR(StreamA1, true, false) SB_STREAM_ATTR(StreamA1, false, false) SB_STREAM_ATTR(StreamA1, true, true) SB_STREAM_ATTR(StreamA1, false, true)
    SB_STREAM_ATTR(StreamA, true, true) SB_STREAM_ATTR(StreamA, false, true) SB_STREAM_ATTR(StreamB, true, true)
    SB_STREAM_ATTR(StreamB, false, true) SB_STREAM_ATTR(StreamA, true, false) SB_STREAM_ATTR(StreamA, false, false)
    SB_STREAM_ATTR(StreamB, true, false) SB_STREAM_ATTR(StreamB, false, false)
#undef SB_STREAM_ATTR
#define SB_INIT(R) if ((e = init_blur_r<R>()) != cudaSuccess) return e;
    SB_INIT(1) SB_INIT(2) SB_INIT(3) SB_INIT(4) SB_INIT(5) SB_INIT(6) SB_INIT(7) SB_INIT(8)
    SB_INIT(9) SB_INIT(10) SB_INIT(11) SB_INIT(12) SB_INIT(13) SB_INIT(14) SB_INIT(15) SB_INIT(16)
#undef SB_INIT
    return cudaSuccess;
}

cudaError_t launch_blur(const float* in, float* out, float* dog, float* dec, int w, int h, int pitch,
                        int dec_w, int dec_h, int dec_pitch, const BlurTaps& taps, cudaStream_t s) {
#define SB_CASE(R) \
    case R: return launch_blur_r<R>(in, out, dog, dec, w, h, pitch, dec_w, dec_h, dec_pitch, taps, s);
    switch (taps.radius) {
        SB_CASE(1) SB_CASE(2) SB_CASE(3) SB_CASE(4) SB_CASE(5) SB_CASE(6) SB_CASE(7) SB_CASE(8)
        SB_CASE(9) SB_CASE(10) SB_CASE(11) SB_CASE(12) SB_CASE(13) SB_CASE(14) SB_CASE(15) SB_CASE(16)
        default: return cudaErrorInvalidValue;
    }
#undef SB_CASE
}

// u8 gray or RGB input -> G[0][0] (gray + up-sample + initial blur); taps.radius must be 4
bool input_fused_supported(int channels, const BlurTaps& taps) {
    return (channels == 1 || channels == 3) && taps.radius == IN_R;
}

cudaError_t launch_input_u8(const uint8_t* src, int sw, int sh, int channels, float* dst, int w, int h, int pitch,
                            int doubled, const BlurTaps& taps, float centre, cudaStream_t s) {
    // tile shapes measured at 4K (ms): 128x64 / 512 threads / 2 per SM 0.088, 128x32 / 256 / 4 0.079,
    // 128x32 / 512 / 3 0.077, 64x64 / 256 / 4 0.077, 128x16 / 256 / 6 0.078
    dim3 grid((w + 127) / 128, (h + 31) / 32);
#define SB_IN(D, CH) k_input_u8<D, 128, 32, 512, 3, CH><<<grid, 512, InputGeom<128, 32>::kSmem, s>>>(src, sw, sh, dst, w, h, pitch, taps, centre)
    if (channels == 3) { if (doubled) SB_IN(true, 3); else SB_IN(false, 3); }
    else { if (doubled) SB_IN(true, 1); else SB_IN(false, 1); }
#undef SB_IN
    return cudaGetLastError();
}

// The fused per-octave path exists for the reference's default radii (4,5,6 | 8,10).
bool cascade_supported(const BlurTaps* taps) {
    return taps[1].radius == 4 && taps[2].radius == 5 && taps[3].radius == 6 && taps[4].radius == 8 &&
           taps[5].radius == 10;
}

// One octave: G[0] -> G[1..3], D[0..4], next base.  keep_all also stores G[4], G[5] (debug planes).
// mode: 0 = streaming kernels (stream.cuh) on octaves large enough to profit, tile kernels below; 2 = tile kernels
// only; 3 = streaming kernels on every octave (tests).  All three give bit-identical planes.
// part: 1 = the G0 -> G1..G3 kernel, 2 = the G3 -> D3, D4 kernel (the caller brackets each with profiling events).
cudaError_t launch_octave_fused(const OctaveDesc& od, const BlurTaps* taps, float* dec, int dec_w, int dec_h,
                                int dec_pitch, bool keep_all, int sm_count, int mode, int part, cudaStream_t s) {
    // Measured on B200.  Alone on the GPU (scratch/stream_test.cu): 7680 x 4320 tile 303 + 249 us, streaming
    // 232 + 166 us; 3840 x 2160 tile 94 + 76 us, streaming 74 + 72 us; below ~2 Mpx the tile kernels are faster
    // alone (the pipeline fill of ~35 rows per CTA is pure latency).  With four images in flight the streaming
    // kernels (small CTAs, 24-62 KB of shared memory) also share the SMs better with the other images' kernels:
    // 4K batch images/s by threshold 8 Mpx 580, 2 Mpx 594, 0.5 Mpx 590 against 511 with tile kernels only,
    // single-image pyramid time 0.762 / 0.789 / 0.834 ms against 0.943.
    // SIFT_B200_STREAM_MIN_PX overrides the threshold (experiments).
    static const long long min_px =
        getenv("SIFT_B200_STREAM_MIN_PX") ? atoll(getenv("SIFT_B200_STREAM_MIN_PX")) : 2000000ll;
    const bool stream = mode == 3 || (mode == 0 && (long long)od.w * od.h >= min_px);
    CascadeArgs a;
    a.in = od.G[0];
    a.g[0] = od.G[1]; a.g[1] = od.G[2]; a.g[2] = od.G[3];
    a.d[0] = od.D[0]; a.d[1] = od.D[1]; a.d[2] = od.D[2];
    a.dec = dec; a.dec_w = dec_w; a.dec_h = dec_h; a.dec_pitch = dec_pitch;
    a.w = od.w; a.h = od.h; a.pitch = od.pitch;
    a.taps[0] = taps[1]; a.taps[1] = taps[2]; a.taps[2] = taps[3];
    if (part == 1) {
        // strip geometry of the first cascade kernel: 224-column strips with two warps per level on octaves of
        // >= 16 Mpx (less x halo, fuller warps: 7680 x 4320 in 212 us against 225 us), 96-column strips with one
        // warp per level below (more CTAs for the smaller grids).  SIFT_B200_STREAM_A = 0 / 1 forces one of them.
        static const int forced = getenv("SIFT_B200_STREAM_A") ? atoi(getenv("SIFT_B200_STREAM_A")) : -1;
        if (!stream) return launch_cascade_t<4, 5, 6>(a, sm_count, s);
        const int geom = forced >= 0 ? forced : ((long long)od.w * od.h >= (16ll << 20) ? 1 : 0);
        if (geom == 1) return launch_stream_t<StreamA1>(a, sm_count, s);
        return launch_stream_t<StreamA>(a, sm_count, s);
    }
    CascadeArgs b;
    b.in = od.G[3];
    b.g[0] = keep_all ? od.G[4] : nullptr; b.g[1] = keep_all ? od.G[5] : nullptr; b.g[2] = nullptr;
    b.d[0] = od.D[3]; b.d[1] = od.D[4]; b.d[2] = nullptr;
    b.dec = nullptr; b.dec_w = b.dec_h = b.dec_pitch = 0;
    b.w = od.w; b.h = od.h; b.pitch = od.pitch;
    b.taps[0] = taps[4]; b.taps[1] = taps[5]; b.taps[2] = taps[5];
    if (!stream) return launch_cascade_t<8, 10, 0>(b, sm_count, s);
    return launch_stream_t<StreamB>(b, sm_count, s);
}

// Octaves that launch_cascade_t would run as at most one 32 x 32 tile per SM (its third form): the tail.
bool tail_eligible(const OctaveDesc& od, int sm_count) {
    static const int mult = getenv("SIFT_B200_TAIL_TILES") ? atoi(getenv("SIFT_B200_TAIL_TILES")) : 1;   // experiments
    return ((od.w + 31) / 32) * ((od.h + 31) / 32) <= mult * sm_count;
}

// Octaves first .. octaves - 1 (all tail_eligible) in one launch; `sync` = 1 + n zeroed ints.  Same planes, bit for
// bit, as launch_octave_fused(part 1), (part 2) per octave.
cudaError_t launch_tail(const OctaveDesc* octs, int first, int octaves, const BlurTaps* taps, bool keep_all, int* sync,
                        int sm_count, cudaStream_t s) {
    const int n = octaves - first;
    if (n < 1 || n > kMaxTail) return cudaErrorInvalidValue;
    TailArgs t;
    memset(&t, 0, sizeof t);
    t.n = n;
    t.sync = sync;
    for (int i = 0; i < 3; ++i) t.taps_a[i] = taps[1 + i];
    t.taps_b[0] = taps[4]; t.taps_b[1] = taps[5]; t.taps_b[2] = taps[5];
    for (int i = 0; i < n; ++i) {
        const OctaveDesc& od = octs[first + i];
        TailOct& to = t.oct[i];
        to.tiles_x = (od.w + 31) / 32;
        to.tiles = to.tiles_x * ((od.h + 31) / 32);
        CascadePlanes& a = to.a;
        a.in = od.G[0];
        a.g[0] = od.G[1]; a.g[1] = od.G[2]; a.g[2] = od.G[3];
        a.d[0] = od.D[0]; a.d[1] = od.D[1]; a.d[2] = od.D[2];
        a.w = od.w; a.h = od.h; a.pitch = od.pitch;
        if (first + i + 1 < octaves) {   // next base = G3 decimated, sift.cpp:195-196
            const OctaveDesc& nx = octs[first + i + 1];
            a.dec = nx.G[0]; a.dec_w = nx.w; a.dec_h = nx.h; a.dec_pitch = nx.pitch;
        }
        CascadePlanes& b = to.b;
        b.in = od.G[3];
        b.g[0] = keep_all ? od.G[4] : nullptr; b.g[1] = keep_all ? od.G[5] : nullptr;
        b.d[0] = od.D[3]; b.d[1] = od.D[4];
        b.w = od.w; b.h = od.h; b.pitch = od.pitch;
    }
    int total = 0;
    for (int g = 0; g <= n; ++g) {   // group g: first kernel of octave g, second kernel of octave g - 1
        t.begin[g] = total;
        if (g < n) total += t.oct[g].tiles;
        if (g >= 1) total += t.oct[g - 1].tiles;
    }
    t.begin[n + 1] = total;
    t.total = total;
    static const int ctas = getenv("SIFT_B200_TAIL_CTAS") ? atoi(getenv("SIFT_B200_TAIL_CTAS")) : 1;   // experiments
    k_tail<<<std::min(ctas * sm_count, total), TAIL_NT, kTailSmem, s>>>(t);
    return cudaGetLastError();
}

cudaError_t launch_prepare_u8(const uint8_t* src, int sw, int sh, int ch, float* dst, int dw, int dh,
                              int dpitch, int doubled, float centre, cudaStream_t s) {
    dim3 block(32, 8), grid((dw + 31) / 32, (dh + 7) / 8);
    if (ch == 1)
        k_prepare_gray_u8<<<grid, block, 0, s>>>(src, sw, sh, dst, dw, dh, dpitch, doubled, centre);
    else
        k_prepare<uint8_t><<<grid, block, 0, s>>>(src, sw, sh, ch, dst, dw, dh, dpitch, doubled, centre, nullptr);
    return cudaGetLastError();
}

cudaError_t launch_prepare_f32(const float* src, int sw, int sh, int ch, float* dst, int dw, int dh,
                               int dpitch, int doubled, const float* centre_range, cudaStream_t s) {
    dim3 block(32, 8), grid((dw + 31) / 32, (dh + 7) / 8);
    k_prepare<float><<<grid, block, 0, s>>>(src, sw, sh, ch, dst, dw, dh, dpitch, doubled, 0.f, centre_range);
    return cudaGetLastError();
}

}  // namespace sb

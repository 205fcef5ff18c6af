// Keypoint kernels (sm_100a): 3x3x3 DoG extrema + compaction, quadratic refinement, orientation
// histograms, canonical sort + de-duplication, 4x4x8 descriptors.
//
// Per-pixel / per-sample arithmetic is FP32 on the FP32 pyramid; every per-keypoint scalar
// (refinement, scale, radius, peak interpolation, final normalisation) is FP64 and follows the
// reference's formulas.  Histogram accumulation uses 2^-32 fixed-point integer atomics so that
// results are independent of the order in which lanes arrive: the whole detect call is
// bit-reproducible although list compaction uses atomics (the final order is re-established by an
// exact sort on the reference's own comparison key).
#include <cstdlib>
#include <type_traits>
#include "common.cuh"
#include "kernels.h"

namespace sb {

namespace {

__device__ __forceinline__ float ldg(const float* p) { return __ldg(p); }
// atan2 for the gradient angle: odd minimax polynomial of degree 17 on [0, 1] (Abramowitz & Stegun
// 4.4.49, |error| <= 2e-8) + quadrant fix-up, with an approximate reciprocal for the ratio; total
// error ~1e-7 rad, the same order as atan2f's, at about half the instructions and no slow path.
// MUFU without the denormal pre- / post-scaling that the non-ftz forms wrap around it (4-5 instructions each in the
// sample loops): identical bits for normal arguments; the callers treat sub-normal arguments as zero.
constexpr float kFltMin = 1.17549435e-38f;
__device__ __forceinline__ float rcp_ftz(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rsqrt_ftz(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float exp_ftz(float x) {   // __expf for results that are normal numbers
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 1.4426950216293334961f));
    return r;
}
__device__ __forceinline__ float fast_atan2(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const float a = mx >= kFltMin ? mn * rcp_ftz(mx) : 0.f;
    const float s = a * a;
    float p = 0.0028662257f;
    p = fmaf(p, s, -0.0161657367f);
    p = fmaf(p, s, 0.0429096138f);
    p = fmaf(p, s, -0.0752896400f);
    p = fmaf(p, s, 0.1065626393f);
    p = fmaf(p, s, -0.1420889944f);
    p = fmaf(p, s, 0.1999355085f);
    p = fmaf(p, s, -0.3333314528f);
    float r = fmaf(p * s, a, a);
    if (ay > ax) r = 1.57079632679489662f - r;
    if (x < 0.f) r = 3.14159265358979323846f - r;
    return copysignf(r, y);
}
// three-input FP32 min / max (FMNMX3, sm_100)
__device__ __forceinline__ float fmin3(float a, float b, float c) {
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

// (min, max) of a float image, for the fixed-point histogram scale of float inputs.  Order-free
// (min / max commute), so the result is reproducible.  range[] must hold (+inf, -inf) on entry.
__device__ __forceinline__ void atomic_min_f(float* a, float v) {
    if (v >= 0.f) atomicMin(reinterpret_cast<int*>(a), __float_as_int(v));
    else atomicMax(reinterpret_cast<unsigned*>(a), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_f(float* a, float v) {
    if (v >= 0.f) atomicMax(reinterpret_cast<int*>(a), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned*>(a), __float_as_uint(v));
}
__global__ void __launch_bounds__(256) k_range(const float* __restrict__ px, size_t n, float* __restrict__ range) {
    float lo = INFINITY, hi = -INFINITY;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v = __ldg(px + i);
        lo = fminf(lo, v);
        hi = fmaxf(hi, v);
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, d));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, d));
    }
    if ((threadIdx.x & 31) == 0 && lo <= hi) {
        atomic_min_f(range, lo);
        atomic_max_f(range + 1, hi);
    }
}

// ------------------------------------------------------------------------------------------
// Extrema scan -- sift.cpp:264-291 (detect_octave_extrema) + :227-256 (is_extremum).
// A pixel is kept iff |D| > threshold and it is >= all 26 neighbours or <= all of them (ties do
// not disqualify).  "v >= every neighbour" is evaluated as "v == max over the 27-cell cube".
// One warp walks a 30-column strip downwards: each lane loads only its own column, the 3-wide
// row min/max come from warp shuffles, the 3-tall window lives in registers.
// ------------------------------------------------------------------------------------------
// warp-aggregated, order-free append of the hits of one ballot (the list is sorted later)
__device__ __noinline__ void extrema_emit(unsigned m, bool hit, int lane, int x, int y, int z, int octave,
                                          Cand* __restrict__ cands, int cap, Counters* __restrict__ counters) {
    int slot0 = 0;
    if (lane == 0) slot0 = atomicAdd(&counters->n_extrema, __popc(m));
    slot0 = __shfl_sync(0xffffffffu, slot0, 0);
    if (hit) {
        const int slot = slot0 + __popc(m & ((1u << lane) - 1));
        if (slot < cap) cands[slot] = Cand{x, y, z, octave};
    }
}

// EX_ROWS = output rows per warp: 32 on large octaves (2 halo rows per 32), 8 on small ones, where the
// grid would otherwise not fill the GPU and a warp's serial walk down 34 rows is pure latency
// CTAs per SM: 3 (85 registers, no spills) 0.412 ms at 4K, 4 (64 registers, spills in the row loop) 0.431 ms,
// 5 0.561 ms
#ifndef SB_EXT_CTAS
#define SB_EXT_CTAS 3
#endif
template <int EX_ROWS, int ND>   // ND = DoG planes per octave (intervals + 2); planes 1 .. ND-2 are tested
__global__ void __launch_bounds__(256, SB_EXT_CTAS)
k_extrema(const OctaveDesc oct, int octave, float thr, Cand* __restrict__ cands, int cap,
          Counters* __restrict__ counters) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int w = oct.w, h = oct.h, pitch = oct.pitch;
    const int x = blockIdx.x * 30 + lane;                       // lanes 1..30 produce output
    const int ys = 1 + (blockIdx.y * 8 + warp) * EX_ROWS;       // first output row of this warp
    if (ys > h - 2) return;
    const int xc = min(x, w - 1);
    const bool lane_out = lane >= 1 && lane <= 30 && x <= w - 2;

    // 3 rotating window rows per plane: min / max over x-1..x+1, and the centre of planes 1..3
    float hmin[ND][3], hmax[ND][3], ctr[ND - 2][3];
    float raw[3][ND];  // rows loaded two iterations ahead of their use (bytes in flight hide HBM latency)
    auto fetch = [&](int y, int rs) {
        const unsigned off = (unsigned)min(y, h - 1) * (unsigned)pitch + (unsigned)xc;   // a plane has < 2^32 pixels
#pragma unroll
        for (int z = 0; z < ND; ++z) raw[rs][z] = ldg(oct.D[z] + off);
    };
    auto absorb = [&](int slot, int rs) {
#pragma unroll
        for (int z = 0; z < ND; ++z) {
            const float v = raw[rs][z];
            const float l = __shfl_up_sync(FULL, v, 1), r = __shfl_down_sync(FULL, v, 1);
            hmin[z][slot] = fmin3(v, l, r);
            hmax[z][slot] = fmax3(v, l, r);
            if (z >= 1 && z <= ND - 2) ctr[z - 1][slot] = v;
        }
    };
    fetch(ys - 1, 0); fetch(ys, 1); fetch(ys + 1, 2);
    absorb(0, 0);
    fetch(ys + 2, 0);
    absorb(1, 1);
#pragma unroll 1
    for (int k0 = 0; k0 < EX_ROWS + 2; k0 += 3) {
#pragma unroll
        for (int kj = 0; kj < 3; ++kj) {  // unrolled by the window height: slot rotation without moves
            const int y = ys + k0 + kj;
            if (k0 + kj >= EX_ROWS || y > h - 2) break;
            const int mid = (kj + 1) % 3;
            absorb((kj + 2) % 3, (kj + 2) % 3);   // row y + 1 (fetched two iterations ago)
            fetch(y + 3, (kj + 1) % 3);           // in flight while this and the next row are tested
            float vmin[ND], vmax[ND];
#pragma unroll
            for (int z = 0; z < ND; ++z) {
                vmin[z] = fmin3(hmin[z][0], hmin[z][1], hmin[z][2]);
                vmax[z] = fmax3(hmax[z][0], hmax[z][1], hmax[z][2]);
            }
            bool hit[ND - 2];
            bool any = false;
#pragma unroll
            for (int z = 1; z <= ND - 2; ++z) {
                const float cv = ctr[z - 1][mid];
                const float mx = fmax3(vmax[z - 1], vmax[z], vmax[z + 1]);
                const float mn = fmin3(vmin[z - 1], vmin[z], vmin[z + 1]);
                hit[z - 1] = lane_out && fabsf(cv) > thr && (cv == mx || cv == mn);
                any |= hit[z - 1];
            }
            if (__ballot_sync(FULL, any)) {   // one vote per row; ~5 % of the rows hold an extremum
#pragma unroll
                for (int z = 1; z <= ND - 2; ++z) {
                    const unsigned m = __ballot_sync(FULL, hit[z - 1]);
                    if (m) extrema_emit(m, hit[z - 1], lane, x, y, z, octave, cands, cap, counters);
                }
            }
        }
    }
}

// Second form of the same scan, FOUR columns per lane (the default).  The 27-cell max / min is separable, so it is
// reduced vertically first (the three window rows are the raw float4 rows themselves, kept in a 4-slot rotating
// register window: three rows in use, the fourth in flight), then across the three planes, and only the
// NZ = ND - 2 reduced rows go through the warp shuffles of the horizontal step (2 shuffles per quantity per FOUR
// pixels instead of per pixel).  A warp covers 128 loaded / 124 tested columns per row with one LDG.128 per plane;
// ~37 instructions per pixel against ~105 of the one-column form.  Same candidate set (the list is unordered).
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// CandCube slot of cube cell [dz][dx][dy] (get_pixel_cube's indexing, sift.cpp:32-44); corners are not stored (-1)
__host__ __device__ constexpr int cube_slot(int dz, int dx, int dy) {
    const int off = (dz != 1) + (dx != 1) + (dy != 1);
    if (off == 0) return 0;
    if (off == 1) return dz != 1 ? (dz == 0 ? 1 : 2) : dx != 1 ? (dx == 0 ? 3 : 4) : (dy == 0 ? 5 : 6);
    if (off == 2) {
        if (dy == 1) return 7 + (dz == 2 ? 0 : 2) + (dx == 2 ? 0 : 1);    // 7 c221, 8 c201, 9 c021, 10 c001
        if (dx == 1) return 11 + (dz == 2 ? 0 : 2) + (dy == 2 ? 0 : 1);   // 11 c212, 12 c210, 13 c012, 14 c010
        return 15 + (dy == 0 ? 0 : 2) + (dx == 0 ? 0 : 1);                // 15 c100, 16 c120, 17 c102, 18 c122
    }
    return -1;
}
// The scan hands the fit its first cube: the 19 cells are in this warp's L1 lines (it loaded rows y-1..y+1 of all
// planes within the last three iterations), where the refinement kernel would fetch them from DRAM one sector at a time.
__device__ __forceinline__ void emit_cube(const OctaveDesc& oc, int x, int y, int z, CandCube* __restrict__ out) {
    float v[20];
    v[19] = 0.f;
#pragma unroll
    for (int dz = 0; dz < 3; ++dz) {
        const float* plane = oc.D[z + dz - 1] + (unsigned)y * (unsigned)oc.pitch + (unsigned)x;
#pragma unroll
        for (int dx = 0; dx < 3; ++dx)
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
                if (cube_slot(dz, dx, dy) >= 0) v[cube_slot(dz, dx, dy)] = __ldg(plane + (dy - 1) * oc.pitch + (dx - 1));
    }
    float4* o4 = reinterpret_cast<float4*>(out->v);
#pragma unroll
    for (int q = 0; q < 5; ++q) o4[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}

constexpr int EX4_STRIP = 124;   // tested columns per warp: 128 loaded minus 2 on each side (kept a multiple of 4)
// SLOTS = rows of the rotating register window (3 in use + SLOTS - 3 in flight), CTAS = CTAs per SM
// XW = 1: the 4 warps of a CTA take consecutive row bands of one strip; XW = 4: they take 4 adjacent strips of one
// row band and walk down side by side, so that a CTA reads ~2 KB contiguous per plane row (DRAM pages)
template <int ND, int SLOTS, int XW>
__device__ __forceinline__ void extrema4_body(const OctaveDesc& oct, int octave, float thr, int rows,
                                              Cand* __restrict__ cands, int cap, Counters* __restrict__ counters,
                                              CandCube* __restrict__ cubes, int cube_cap, int bx, int by) {
    constexpr int NZ = ND - 2;
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int w = oct.w, h = oct.h, pitch = oct.pitch;
    const int strip = XW == 4 ? bx * 4 + warp : bx;
    if (strip * EX4_STRIP - 2 > w - 2) return;          // (XW = 4) a strip beyond the last tested column
    const int x0 = strip * EX4_STRIP - 4 + 4 * lane;    // this lane's first column; the strip tests lane positions 2..125
    const int ys = 1 + (XW == 4 ? by : by * 4 + warp) * rows;      // first tested row of this warp
    if (ys > h - 2) return;
    const int ye = min(ys + rows - 1, h - 2);
    // lanes hanging over the left / right end of the row load a clamped (wrong, unused) position: their columns are
    // neither tested nor neighbours of a tested column
    const unsigned xl = (unsigned)min(max(x0, 0), pitch - 4);
    unsigned ok = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const int pos = 4 * lane + c, x = x0 + c;
        if (pos >= 2 && pos <= EX4_STRIP + 1 && x >= 1 && x <= w - 2) ok |= 1u << c;
    }
    float4 win[SLOTS][ND];
    auto fetch = [&](int y, int slot) {
        const unsigned off = (unsigned)min(y, h - 1) * (unsigned)pitch + xl;   // a plane has < 2^31 pixels
#pragma unroll
        for (int z = 0; z < ND; ++z) win[slot][z] = ldg4(oct.D[z] + off);
    };
    auto test_row = [&](const float4 (&ra)[ND], const float4 (&rb)[ND], const float4 (&rc)[ND], int y) {
        float vmx[3][4], vmn[3][4];   // vertical max / min of the last three planes
        unsigned hits = 0;            // bit 4 (z - 1) + c
#pragma unroll
        for (int p = 0; p < ND; ++p) {
            float* mx = vmx[p % 3];
            float* mn = vmn[p % 3];
            mx[0] = fmax3(ra[p].x, rb[p].x, rc[p].x); mn[0] = fmin3(ra[p].x, rb[p].x, rc[p].x);
            mx[1] = fmax3(ra[p].y, rb[p].y, rc[p].y); mn[1] = fmin3(ra[p].y, rb[p].y, rc[p].y);
            mx[2] = fmax3(ra[p].z, rb[p].z, rc[p].z); mn[2] = fmin3(ra[p].z, rb[p].z, rc[p].z);
            mx[3] = fmax3(ra[p].w, rb[p].w, rc[p].w); mn[3] = fmin3(ra[p].w, rb[p].w, rc[p].w);
            if (p < 2) continue;
            const int z = p - 1;
            float zx[4], zn[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                zx[c] = fmax3(vmx[0][c], vmx[1][c], vmx[2][c]);
                zn[c] = fmin3(vmn[0][c], vmn[1][c], vmn[2][c]);
            }
            const float lx = __shfl_up_sync(FULL, zx[3], 1), rx = __shfl_down_sync(FULL, zx[0], 1);
            const float ln = __shfl_up_sync(FULL, zn[3], 1), rn = __shfl_down_sync(FULL, zn[0], 1);
            const float hx[4] = {fmax3(lx, zx[0], zx[1]), fmax3(zx[0], zx[1], zx[2]), fmax3(zx[1], zx[2], zx[3]),
                                 fmax3(zx[2], zx[3], rx)};
            const float hn[4] = {fmin3(ln, zn[0], zn[1]), fmin3(zn[0], zn[1], zn[2]), fmin3(zn[1], zn[2], zn[3]),
                                 fmin3(zn[2], zn[3], rn)};
            const float cv[4] = {rb[z].x, rb[z].y, rb[z].z, rb[z].w};
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (fabsf(cv[c]) > thr && (cv[c] == hx[c] || cv[c] == hn[c])) hits |= 1u << (4 * (z - 1) + c);
        }
        unsigned mask = 0;
#pragma unroll
        for (int z = 0; z < NZ; ++z) mask |= ok << (4 * z);
        hits &= mask;
        if (__ballot_sync(FULL, hits != 0)) {   // ~a third of the 124-pixel rows hold an extremum
            const int cnt = __popc(hits);
            int incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(FULL, incl, d);
                if (lane >= d) incl += t;
            }
            int slot = 0;
#ifdef SB_EXP_NOATOM   // timing experiment only (wrong results): what the slot round trip costs the scan
            if (lane == 31) slot = (int)((((unsigned)y * 2654435761u) ^ ((unsigned)x0 * 40503u)) % (unsigned)(cap - 256));
#else
            if (lane == 31) slot = atomicAdd(&counters->n_extrema, incl);
#endif
            slot = __shfl_sync(FULL, slot, 31) + incl - cnt;
            while (hits) {
                const int b = __ffs(hits) - 1;
                hits &= hits - 1;
                if (slot < cap) cands[slot] = Cand{x0 + (b & 3), y, 1 + (b >> 2), octave};
                if (slot < cube_cap) emit_cube(oct, x0 + (b & 3), y, 1 + (b >> 2), cubes + slot);
                ++slot;
            }
        }
    };
#pragma unroll
    for (int j = 0; j < SLOTS - 1; ++j) fetch(ys - 1 + j, j);
#pragma unroll 1
    for (int y0 = ys; y0 <= ye; y0 += SLOTS) {
#pragma unroll
        for (int j = 0; j < SLOTS; ++j) {   // unrolled by the slot count: the window rotates without register moves
            const int y = y0 + j;
            if (y > ye) break;
            fetch(y + SLOTS - 2, (j + SLOTS - 1) % SLOTS);   // in flight while this row and the next ones are tested
            test_row(win[j % SLOTS], win[(j + 1) % SLOTS], win[(j + 2) % SLOTS], y);
        }
    }
}

template <int ND, int SLOTS, int CTAS, int XW>
__global__ void __launch_bounds__(128, CTAS)
k_extrema4(const OctaveDesc oct, int octave, float thr, int rows, Cand* __restrict__ cands, int cap,
           Counters* __restrict__ counters, CandCube* __restrict__ cubes, int cube_cap) {
    extrema4_body<ND, SLOTS, XW>(oct, octave, thr, rows, cands, cap, counters, cubes, cube_cap, blockIdx.x, blockIdx.y);
}

// The scans of the small octaves (the ones k_tail produced) in one launch: a CTA finds its octave in a prefix table
// and runs the same body.  Six launches of 5-9 us, each a single partial wave, become one.
constexpr int kMaxExtremaMulti = 12;
struct ExtremaMulti {
    int n;
    int cta_begin[kMaxExtremaMulti + 1];
    int gx[kMaxExtremaMulti], rows[kMaxExtremaMulti], octave[kMaxExtremaMulti];
    OctaveDesc oct[kMaxExtremaMulti];
};
template <int ND>
__global__ void __launch_bounds__(128, 3)
k_extrema4_multi(const ExtremaMulti m, float thr, Cand* __restrict__ cands, int cap, Counters* __restrict__ counters,
                 CandCube* __restrict__ cubes, int cube_cap) {
    int i = 0;
    while ((int)blockIdx.x >= m.cta_begin[i + 1]) ++i;
    const int k = blockIdx.x - m.cta_begin[i];
    const int gx = m.gx[i];
    // (the octave is read where it lies in parameter space: a local copy would have to live in local memory, because
    // the cube hand-over indexes its plane pointers with the run-time layer)
    extrema4_body<ND, 4, 4>(m.oct[i], m.octave[i], thr, m.rows[i], cands, cap, counters, cubes, cube_cap, k % gx, k / gx);
}

// Generic window (window_size = 5, 7: border = 2, 3): the same tie-tolerant test over the
// (2 border + 1)^3 cube, thread per pixel, neighbours straight from L1/L2.  Rarely used knob; the
// 3x3x3 default takes the register-window kernel above.
__global__ void __launch_bounds__(256)
k_extrema_window(const OctaveDesc oct, int octave, int dogs, int border, float thr, Cand* __restrict__ cands,
                 int cap, Counters* __restrict__ counters) {
    const int x = border + blockIdx.x * blockDim.x + threadIdx.x;
    const int y = border + blockIdx.y;
    if (x >= oct.w - border || y >= oct.h - border) return;
    for (int z = border; z < dogs - border; ++z) {
        const float c = ldg(oct.D[z] + (size_t)y * oct.pitch + x);
        if (!(fabsf(c) > thr)) continue;
        bool mx = true, mn = true;
        for (int dz = -border; dz <= border && (mx || mn); ++dz)
            for (int dy = -border; dy <= border; ++dy) {
                const float* row = oct.D[z + dz] + (size_t)(y + dy) * oct.pitch + x;
                for (int dx = -border; dx <= border; ++dx) {
                    const float v = ldg(row + dx);
                    mx = mx && !(c < v);
                    mn = mn && !(c > v);
                }
            }
        if (mx || mn) {
            const int slot = atomicAdd(&counters->n_extrema, 1);
            if (slot < cap) cands[slot] = Cand{x, y, z, octave};
        }
    }
}

// ------------------------------------------------------------------------------------------
// Refinement -- sift.cpp:330-436 (compute_keypoints) with get_pixel_cube :32-44,
// compute_gradient :49-55, compute_hessian :60-80, fit_quadratic :86-106.  FP64, one thread per
// candidate; cube indexed [z][x][y] of D/255 like the reference.
// ------------------------------------------------------------------------------------------
struct Fit {
    double off[3], g[3], hxx, hyy, hxy, centre;
};

// load(dz, dx, dy) = DoG value of cube cell [dz][dx][dy]: from the planes, or from the cube the scan handed over
template <typename Load>
__device__ __forceinline__ Fit fit_cell(Load load) {
    double c[3][3][3];
#pragma unroll
    for (int dz = 0; dz < 3; ++dz)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx)
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
                c[dz][dx][dy] = cube_slot(dz, dx, dy) >= 0 ? (double)load(dz, dx, dy) / 255.0 : 0.0;
    Fit f;
    f.centre = c[1][1][1];
    f.g[0] = 0.5 * (c[2][1][1] - c[0][1][1]);
    f.g[1] = 0.5 * (c[1][2][1] - c[1][0][1]);
    f.g[2] = 0.5 * (c[1][1][2] - c[1][1][0]);
    const double h00 = c[0][1][1] - 2 * c[1][1][1] + c[2][1][1];
    const double h11 = c[1][0][1] - 2 * c[1][1][1] + c[1][2][1];
    const double h22 = c[1][1][0] - 2 * c[1][1][1] + c[1][1][2];
    const double h01 = 0.25 * (c[2][2][1] - c[2][0][1] - c[0][2][1] + c[0][0][1]);
    const double h02 = 0.25 * (c[2][1][2] - c[2][1][0] - c[0][1][2] + c[0][1][0]);
    const double h12 = 0.25 * (c[1][0][0] - c[1][2][0] - c[1][0][2] + c[1][2][2]);
    const double det = h00 * h11 * h22 + 2 * (h01 * h12 * h02) - h02 * h11 * h02 - h00 * h12 * h12 -
                       h01 * h01 * h22;
    const double i00 = (h11 * h22 - h12 * h12) / det;
    const double i01 = (h02 * h12 - h01 * h22) / det;
    const double i02 = (h01 * h12 - h02 * h11) / det;
    const double i11 = (h00 * h22 - h02 * h02) / det;
    const double i12 = (h02 * h01 - h00 * h12) / det;
    const double i22 = (h00 * h11 - h01 * h01) / det;
    f.off[0] = -i00 * f.g[0] - i01 * f.g[1] - i02 * f.g[2];
    f.off[1] = -i01 * f.g[0] - i11 * f.g[1] - i12 * f.g[2];
    f.off[2] = -i02 * f.g[0] - i12 * f.g[1] - i22 * f.g[2];
    f.hxx = h11; f.hyy = h22; f.hxy = h12;
    return f;
}

#ifndef SB_REFINE_CTAS
#define SB_REFINE_CTAS 10   // 48 registers: every candidate of a 4K image resident at once; measured (ms at 4K, with the cube hand-over) 6: 0.066, 8: 0.064, 10: 0.056
#endif
__global__ void __launch_bounds__(128, SB_REFINE_CTAS)
k_refine(const PyramidDesc* __restrict__ pyr, const Cand* __restrict__ cands, KpCore* __restrict__ raw,
         Counters* __restrict__ counters, const CandCube* __restrict__ cubes, int cube_cap, const StageParams sp) {
    const int n = min(counters->n_extrema, sp.cap_extrema);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const Cand e = cands[i];
        const OctaveDesc& oc = pyr->oct[e.o];
        int x = e.x, y = e.y, layer = e.z;
        Fit f;
        bool keep = false;
        for (int step = 0; step < 5; ++step) {  // MAX_CONVERGENCE_STEPS, sift.hh:7
            if (step == 0 && i < cube_cap) {   // the first cube came with the candidate
                float v[20];
                const float4* c4 = reinterpret_cast<const float4*>(cubes[i].v);
#pragma unroll
                for (int q = 0; q < 5; ++q) {
                    const float4 t = __ldg(c4 + q);
                    v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
                }
                f = fit_cell([&](int dz, int dx, int dy) { return v[cube_slot(dz, dx, dy) >= 0 ? cube_slot(dz, dx, dy) : 0]; });
            } else {
                f = fit_cell([&](int dz, int dx, int dy) {
                    return ldg(oc.D[layer + dz - 1] + (size_t)(y + dy - 1) * oc.pitch + (x + dx - 1));
                });
            }
            const double m = fmax(fabs(f.off[0]), fmax(fabs(f.off[1]), fabs(f.off[2])));
            if (m < 0.5) {  // CONVERGENCE_THR, sift.hh:8
                const double dot = f.g[0] * f.off[0] + f.g[1] * f.off[1] + f.g[2] * f.off[2];
                const double val = f.centre + 0.5 * dot;
                if (!((fabs(val) * sp.intervals) >= sp.contrast_threshold)) break;
                const double tr = f.hxx + f.hyy;
                const double det = f.hxx * f.hyy - f.hxy * f.hxy;
                if (tr <= 0) break;  // sift.cpp:385 -- rejects every DoG maximum, kept as is
                const double r = sp.eigen_ratio;
                keep = !((tr * tr * r) >= ((r + 1) * (r + 1) * det));
                break;
            }
            // a singular Hessian gives inf/NaN offsets; the reference then walks out of range
            if (!(isfinite(f.off[0]) && isfinite(f.off[1]) && isfinite(f.off[2]))) break;
            layer += (int)round(f.off[0]);
            x += (int)round(f.off[1]);
            y += (int)round(f.off[2]);
            if (x < sp.border || x >= oc.w - sp.border || y < sp.border || y >= oc.h - sp.border ||
                layer < sp.border || layer >= sp.dogs - sp.border) break;  // sift.cpp:405-410
        }
        if (!keep) continue;
        const double s = (double)(1 << e.o);  // pow(2, octave)
        KpCore kp;
        kp.octave = e.o;
        kp.layer = layer;
        kp.x = s * ((double)x + f.off[1]);
        kp.y = s * ((double)y + f.off[2]);
        kp.size = sp.init_sigma * s * pow(2.0, ((double)layer + f.off[0]) / sp.intervals);
        kp.pori = 0.0;
        const int slot = atomicAdd(&counters->n_raw, 1);
        if (slot < sp.cap_raw) raw[slot] = kp;
    }
}

// ------------------------------------------------------------------------------------------
// Orientation -- sift.cpp:447-533.  One warp per raw keypoint; lanes stride over the
// (2r+1)^2 window; 36-bin histogram in shared memory, privatised 8x to thin out bank conflicts,
// accumulated as 32-bit fixed point (integer adds commute => bit-reproducible); lane 0 smooths in
// place (sequentially, exactly like the reference) and emits one keypoint per qualifying peak.
// Fixed-point scale: a bin can receive at most sum(w) * max|grad| <= (1 + sqrt(2 pi) s)^2 * 361
// for pixel values in [0, 255], which is mapped to 2^32.
// ------------------------------------------------------------------------------------------
#ifndef SB_ORI_COPIES
#define SB_ORI_COPIES 8
#endif
#ifndef SB_ORI_STRIDE
#define SB_ORI_STRIDE 36     // words between the private copies of the 36-bin histogram (copy c sits 4 c banks on)
#endif
constexpr int ORI_COPIES = SB_ORI_COPIES;
constexpr int kMaxOriBins = 128;   // largest num_bins of the generic instantiation

// NB > 0: compile-time bin count (36, the reference default); NB == 0: sp.num_bins at run time
// (<= kMaxOriBins).  The sequential smoothing runs through shared memory in both.
// CTAs per SM (the grid is exactly one wave): 4 0.099 ms at 4K, 5 0.091 ms, 6 0.090 ms
#ifndef SB_ORI_UNROLL
#define SB_ORI_UNROLL 1
#endif
constexpr int ORI_UNROLL = SB_ORI_UNROLL;   // sample-loop unrolling of k_orient: 2 measured 0.0865 ms against 0.0849 at 1
#ifndef SB_ORI_CTAS
#define SB_ORI_CTAS 5
#endif
template <int NB>
__global__ void __launch_bounds__(256, SB_ORI_CTAS)
k_orient(const PyramidDesc* __restrict__ pyr, const KpCore* __restrict__ raw, KpCore* __restrict__ oriented,
         Counters* __restrict__ counters, const StageParams sp) {
    constexpr int CAPB = NB > 0 ? NB : kMaxOriBins;
    constexpr int COPIES = NB > 0 ? ORI_COPIES : 2;
    constexpr int HSTRIDE = NB > 0 ? (SB_ORI_STRIDE >= NB ? SB_ORI_STRIDE : NB) : CAPB + 4;
    __shared__ unsigned s_hist[8][COPIES][HSTRIDE];
    __shared__ double s_smooth[8][CAPB];
    const int nb = NB > 0 ? NB : sp.num_bins;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* smooth = s_smooth[warp];
    const int n = min(counters->n_raw, sp.cap_raw);
    const double mag_bound = sp.range ? fmax(1.4142135623730951 * ((double)sp.range[1] - (double)sp.range[0]), 1e-30) * 1.001
                                      : sp.mag_bound;
    unsigned* hist = &s_hist[warp][0][0];
    unsigned* my_hist = s_hist[warp][lane & (COPIES - 1)];
    // keypoints cost 4x more or less than one another: warps pull the next one from a shared cursor
    for (;;) {
        int i = 0;
        if (lane == 0) i = atomicAdd(&counters->next_orient, 1);
        i = __shfl_sync(0xffffffffu, i, 0);
        if (i >= n) break;
        const KpCore kp = raw[i];
        const OctaveDesc& oc = pyr->oct[kp.octave];
        const float* __restrict__ img = oc.G[kp.layer];
        const int W = oc.w, H = oc.h, pitch = oc.pitch;
        const double inv = 1.0 / (double)(1 << kp.octave);
        const int x = (int)round(kp.x * inv), y = (int)round(kp.y * inv);
        const double scale = sp.ori_sigma_factor * (kp.size * inv);
        const int radius = (int)round(3.0 * scale);
        const float neg_inv_denom = (float)(-1.0 / (2.0 * scale * scale));
        const double g1 = 1.0 + 2.5066282746310002 * scale;
        const double bound = g1 * g1 * mag_bound;
        const float fix = (float)(4294967296.0 / bound);
        const double unfix = bound / 4294967296.0;
        const float bins_per_rad = (float)nb * (1.0f / 6.283185307179586f);
        for (int b = lane; b < COPIES * HSTRIDE; b += 32) hist[b] = 0u;
        __syncwarp();
        // window clipped to the pixels whose 4-neighbourhood is inside the image (sift.cpp:473,478)
        const int i_lo = max(-radius, 1 - x), i_hi = min(radius, W - 2 - x);
        const int j_lo = max(-radius, 1 - y), j_hi = min(radius, H - 2 - y);
        const int side = i_hi - i_lo + 1;
        const int total = (side > 0 && j_hi >= j_lo) ? side * (j_hi - j_lo + 1) : 0;
        // s / side by multiplication: floor(s / side) == umulhi(s, ceil(2^32 / side)) while s * side < 2^32 (windows of
        // up to 1024 columns; anything wider -- only reachable with an absurd ori_sigma_factor -- divides)
        const unsigned magic = side > 0 ? (unsigned)((0x100000000ull + (unsigned)side - 1u) / (unsigned)side) : 0u;
        auto sample_loop = [&](auto small_tag) {
        constexpr bool SMALL = decltype(small_tag)::value;
#pragma unroll ORI_UNROLL
        for (int s = lane; s < total; s += 32) {
            const int jr = SMALL ? (int)__umulhi((unsigned)s, magic) : s / side;
            const int i_off = i_lo + (s - jr * side);
            const int j_off = j_lo + jr;
            const int ic = (y + j_off) * pitch + (x + i_off);   // 32-bit indices: a plane has < 2^31 elements
            const int iu = ic - pitch, id = ic + pitch;
            const float dx = ldg(img + ic + 1) - ldg(img + ic - 1);
            const float dy = ldg(img + iu) - ldg(img + id);  // up minus down, sift.cpp:483
            const float g2 = fmaf(dx, dx, __fmul_rn(dy, dy));
            const float mag = g2 >= kFltMin ? g2 * rsqrt_ftz(g2) : 0.f;
            const float ang = fast_atan2(dy, dx);
            const float wgt = exp_ftz((float)(i_off * i_off + j_off * j_off) * neg_inv_denom);
            int b = (int)roundf((ang + 3.14159265358979323846f) * bins_per_rad);
            b = (b < nb) ? b : 0;  // sift.cpp:489-490: bin 0 <-> angle -pi
            b = max(b, 0);
            atomicAdd(&my_hist[b], __float2uint_rn(__fmul_rn(__fmul_rn(wgt, mag), fix)));
        }
        };
        if (side <= 1024) sample_loop(std::true_type{}); else sample_loop(std::false_type{});
        __syncwarp();
        // lane 0 smooths (the reference's in-place, sequential 1/4-1/2-1/4 filter, sift.cpp:496-504:
        // bin i sees the already-updated bin i-1, and the last bin the already-updated bin 0); with
        // a compile-time bin count the loop is fully unrolled and the values stay in registers.
        // All lanes then test their bins for peaks.
        for (int b = lane; b < nb; b += 32) {
            unsigned long long t = 0;
#pragma unroll
            for (int cpy = 0; cpy < COPIES; ++cpy) t += hist[cpy * HSTRIDE + b];
            smooth[b] = (double)t * unfix;
        }
        __syncwarp();
        if (lane == 0) {
            // a sliding window of three values in registers, the rest in shared memory: the 36 doubles of the
            // histogram would otherwise cost 72 registers per thread and halve the occupancy of the sample loop
#pragma unroll 1
            for (int it = 0; it < 2; ++it) {  // ORI_SMOOTH_ITERATIONS
                double prev = smooth[nb - 1], cur = smooth[0];
#pragma unroll 6
                for (int b = 0; b + 1 < nb; ++b) {
                    const double nxt = smooth[b + 1];
                    const double v = 0.25 * prev + 0.5 * cur + 0.25 * nxt;
                    smooth[b] = v;
                    prev = v;
                    cur = nxt;
                }
                smooth[nb - 1] = 0.25 * prev + 0.5 * cur + 0.25 * smooth[0];
            }
        }
        __syncwarp();
        double top = 0.0;
        for (int b = lane; b < nb; b += 32) top = fmax(top, smooth[b]);
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) top = fmax(top, __shfl_xor_sync(0xffffffffu, top, d));
        for (int b0 = 0; b0 < nb; b0 += 32) {
            const int b = b0 + lane;
            bool peak = false;
            double ori = 0.0;
            if (b < nb) {
                const double h0 = smooth[(b + nb - 1) % nb], h1 = smooth[b], h2 = smooth[(b + 1) % nb];
                if (h1 > h0 && h1 > h2 && h1 > (sp.peak_ratio * top)) {
                    peak = true;
                    // fmod(t, m) for t in [0, 2m) is t or t - m, both exact
                    double pos = (double)b + 0.5 * (h0 - h2) / (h0 - 2 * h1 + h2);
                    pos = pos + (double)nb;
                    if (pos >= (double)nb) pos -= (double)nb;
                    ori = kTwoPi * pos / nb;
                    ori = ori + kTwoPi;
                    if (ori >= kTwoPi) ori -= kTwoPi;
                }
            }
            const unsigned m = __ballot_sync(0xffffffffu, peak);
            if (m) {
                int slot0 = 0;
                if (lane == 0) slot0 = atomicAdd(&counters->n_oriented, __popc(m));
                slot0 = __shfl_sync(0xffffffffu, slot0, 0);
                if (peak) {
                    const int slot = slot0 + __popc(m & ((1u << lane) - 1));
                    KpCore out = kp;
                    out.pori = ori;
                    if (sp.doubled) { out.x /= 2; out.y /= 2; out.size /= 2; }
                    if (slot < sp.cap_oriented) oriented[slot] = out;
                }
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// Canonical order + de-duplication -- clean_keypoints, sift.cpp:20-24 with Keypoint::operator<
// and operator== (sift.hh:25-41).  Counting sort on floor(x) buckets, exact ranking inside each
// bucket with the reference's comparator, equal records collapse to the first.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ bool kp_less(const KpCore& a, const KpCore& b) {
    if (a.x != b.x) return a.x < b.x;
    if (a.y != b.y) return a.y < b.y;
    if (a.size != b.size) return a.size > b.size;
    if (a.pori != b.pori) return a.pori < b.pori;
    return a.octave > b.octave;
}
__device__ __forceinline__ bool kp_same(const KpCore& a, const KpCore& b) {
    return a.x == b.x && a.y == b.y && a.size == b.size && a.pori == b.pori;
}
__device__ __forceinline__ int bucket_of(double x, int nb) {
    int b = (int)floor(x);
    return min(max(b, 0), nb - 1);
}

__global__ void k_sort_clear(SortScratch ss) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= ss.nb; i += gridDim.x * blockDim.x) {
        ss.bucket_cnt[i] = 0;
        ss.uniq_cnt[i] = 0;
        if (i < ss.nb) ss.bucket_fill[i] = 0;
    }
}

__global__ void k_bucket_count(const KpCore* __restrict__ kps, const Counters* __restrict__ counters,
                               SortScratch ss, int cap) {
    const int n = min(counters->n_oriented, cap);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        atomicAdd(&ss.bucket_cnt[bucket_of(kps[i].x, ss.nb)], 1);
}

// single-CTA exclusive scan of cnt[0..n) into off[0..n], off[n] = total.  Four consecutive values per thread: the
// 3842 buckets of a 4K image take one pass (three barriers) instead of four passes of five.
__global__ void __launch_bounds__(1024) k_scan(const int* __restrict__ cnt, int* __restrict__ off, int n,
                                               int* __restrict__ total_out) {
    constexpr int VPT = 4;
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < n; base += 1024 * VPT) {
        const int i0 = base + threadIdx.x * VPT;
        int v[VPT];
        int sum = 0;
#pragma unroll
        for (int k = 0; k < VPT; ++k) {
            v[k] = (i0 + k < n) ? cnt[i0 + k] : 0;
            sum += v[k];
        }
        int incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int wv = s_warp[lane];
            int winc = wv;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, winc, d);
                if (lane >= d) winc += t;
            }
            s_warp[lane] = winc - wv;  // exclusive prefix of warp totals
        }
        __syncthreads();
        int run = s_carry + s_warp[warp] + incl - sum;   // exclusive prefix of this thread's first value
#pragma unroll
        for (int k = 0; k < VPT; ++k) {
            if (i0 + k < n) off[i0 + k] = run;
            run += v[k];
        }
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = run;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        off[n] = s_carry;
        if (total_out) *total_out = s_carry;
    }
}

__global__ void k_bucket_scatter(const KpCore* __restrict__ kps, const Counters* __restrict__ counters,
                                 SortScratch ss, int cap) {
    const int n = min(counters->n_oriented, cap);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int b = bucket_of(kps[i].x, ss.nb);
        ss.perm[ss.bucket_off[b] + atomicAdd(&ss.bucket_fill[b], 1)] = i;
    }
}

// one warp per bucket: exact rank inside the bucket, duplicates dropped, survivors compacted to
// the front of the bucket's range in `sorted`.
__global__ void __launch_bounds__(256) k_bucket_rank(const KpCore* __restrict__ kps, SortScratch ss) {
    const int lane = threadIdx.x & 31;
    for (int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < ss.nb; b += (gridDim.x * blockDim.x) >> 5) {
        const int s = ss.bucket_off[b], e = ss.bucket_off[b + 1];
        if (e == s) continue;
        for (int p = s + lane; p < e; p += 32) {
            const int ip = ss.perm[p];
            const KpCore a = kps[ip];
            int rank = 0, dup = 0;
            for (int q = s; q < e; ++q) {
                if (q == p) continue;
                const int iq = ss.perm[q];
                const KpCore c = kps[iq];
                const bool lt = kp_less(c, a), gt = kp_less(a, c);
                const bool before = lt || (!gt && iq < ip);
                rank += before ? 1 : 0;
                if (before && kp_same(c, a)) dup = 1;
            }
            ss.tmp_sorted[s + rank] = ip | (dup << 30);
        }
        __syncwarp();
        int kept = 0;
        for (int t = s; t < e; t += 32) {
            const int p = t + lane;
            int v = 0;
            bool keep = false;
            if (p < e) { v = ss.tmp_sorted[p]; keep = !((v >> 30) & 1); }
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (keep) ss.sorted[s + kept + __popc(m & ((1u << lane) - 1))] = v & 0x3fffffff;
            kept += __popc(m);
        }
        if (lane == 0) ss.uniq_cnt[b] = kept;
        __syncwarp();
    }
}

__global__ void __launch_bounds__(256) k_bucket_gather(SortScratch ss) {
    const int lane = threadIdx.x & 31;
    for (int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < ss.nb; b += (gridDim.x * blockDim.x) >> 5) {
        const int s = ss.bucket_off[b], u = ss.uniq_cnt[b], o = ss.uniq_off[b];
        for (int k = lane; k < u; k += 32) ss.final_order[o + k] = ss.sorted[s + k];
    }
}

// ------------------------------------------------------------------------------------------
// Descriptors -- sift.cpp:610-682 (compute_descriptors), :541-571 (update_histogram, trilinear),
// :576-603 (convert_hist_to_desc).  One warp per final keypoint.  The reference walks the whole
// (2r+1)^2 window and rejects about half of it; here each lane first bounds, per window row, the
// column interval that can pass the rotated-bin test (a conservative superset, the exact test is
// still evaluated per sample), the intervals are flattened with a warp scan and the 32 lanes
// stride over the flattened list, so nearly every lane-iteration is a contributing sample.
// 4x4x8 histogram in shared memory as 32-bit fixed point (integer adds commute =>
// bit-reproducible), two copies per warp (odd / even lanes) to thin out conflicts.  Scale: one bin
// receives at most sum(tent_r * tent_c) * max|grad| <= (hw + 2)^2 * 361 for pixel values in
// [0, 255]; the descriptor is normalised afterwards, so the scale cancels.
// FP64 normalise / clamp 0.2 / renormalise / floor(512 x) / min 255.  Writes the 168-byte record
// and the dense 128-byte row.
// ------------------------------------------------------------------------------------------
#ifndef SB_DESC_COPIES
#define SB_DESC_COPIES 4
#endif
constexpr int DESC_COPIES = SB_DESC_COPIES;   // one copy per lane & 3 (8 copies x 4 warps measured slower: occupancy)
constexpr int DESC_WARPS = 8;    // warps (= keypoints in flight) per CTA
#ifndef SB_DESC_CTAS
#define SB_DESC_CTAS 5
#endif
constexpr int DESC_CTAS = SB_DESC_CTAS;   // CTAs per SM (the grid is exactly one wave).  4 (64 registers) 0.492 ms,
                                          // 5 (48 registers, 8 B spilled) 0.461 ms, 6 (40 registers) 0.459 ms
constexpr int DESC_GRID = 6;                            // 4x4 cells + a one-cell border that absorbs dropped bins
constexpr int DESC_WORDS = DESC_GRID * DESC_GRID * 8;   // per histogram copy
// Copy stride in words.  The 32 lanes of a sample iteration are adjacent pixels: they fall into one or two cells and
// a few orientation bins, i.e. onto 2-3 addresses.  Private copies remove the read-modify-write serialisation only
// if they ALSO sit in different banks: 288 words is a multiple of 32, so the copies of one bin shared a bank (ncu:
// 75 % of this kernel's shared-memory wavefronts were bank-conflict replays, and more copies did not help).  +8
// words per copy moves copy c by 8 c banks.  Measured at 4K (ms): 2 copies, same banks 0.462; 2 copies + 4 / 8 / 16
// words 0.376 / 0.382 / 0.385; 4 copies + 1 / 4 / 8 / 12 / 20 words 0.371 / 0.369 / 0.370 / 0.370 / 0.371.
#ifndef SB_DESC_PAD
#define SB_DESC_PAD 4
#endif
constexpr int DESC_STRIDE = DESC_WORDS + SB_DESC_PAD;

// position of the (n + 1)-th set bit of m, 32 if there are fewer
__device__ __forceinline__ int nth_set_bit(unsigned m, int n) {
    int pos = 0;
#pragma unroll
    for (int w = 16; w >= 1; w >>= 1) {
        const int c = __popc(m & ((1u << w) - 1u));
        if (n >= c) { n -= c; pos += w; m >>= w; }
    }
    return (m & 1u) && n == 0 ? pos : 32;
}

__global__ void __launch_bounds__(DESC_WARPS * 32, DESC_CTAS)
k_describe(const PyramidDesc* __restrict__ pyr, const KpCore* __restrict__ oriented,
           const int* __restrict__ final_order, Counters* __restrict__ counters,
           uint8_t* __restrict__ records, uint8_t* __restrict__ desc, int cap_final, const StageParams sp) {
    __shared__ unsigned s_hist[DESC_WARPS][DESC_COPIES][DESC_STRIDE];
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned* hist = &s_hist[warp][0][0];
    unsigned* my_hist = s_hist[warp][lane & (DESC_COPIES - 1)];
    const unsigned lt_mask = (1u << lane) - 1u;
    const int n = min(counters->n_final, cap_final);
    const double mag_bound = sp.range ? fmax(1.4142135623730951 * ((double)sp.range[1] - (double)sp.range[0]), 1e-30) * 1.001
                                      : sp.mag_bound;
    for (;;) {   // work stealing: the window area varies 4x between keypoints
        int i = 0;
        if (lane == 0) i = atomicAdd(&counters->next_describe, 1);
        i = __shfl_sync(0xffffffffu, i, 0);
        if (i >= n) break;
        const KpCore kp = oriented[final_order[i]];
        const OctaveDesc& oc = pyr->oct[kp.octave];
        const float* __restrict__ img = oc.G[kp.layer];
        const int W = oc.w, H = oc.h, pitch = oc.pitch;
        const double inv = sp.doubled ? (kp.octave >= 1 ? 1.0 / (double)(1 << (kp.octave - 1)) : 2.0)
                                      : 1.0 / (double)(1 << kp.octave);
        const int x = (int)(kp.x * inv), y = (int)(kp.y * inv);  // truncation, sift.cpp:623-624
        const double size = kp.size * inv;
        const double hw = sp.desc_scale_factor * size;
        const double tmp_r = round(hw * 0.5 * sqrt(2.0) * (4 + 1.0) + 0.5);
        const int radius = (int)fmin(tmp_r, sqrt((double)(W * W + H * H)));
        const float ca = (float)cos(kp.pori), sa = (float)sin(kp.pori);
        const float inv_hw = (float)(1.0 / hw);
        const float pori = (float)kp.pori;
        const float fix = (float)(4294967296.0 / ((hw + 2.0) * (hw + 2.0) * mag_bound));
        for (int b = lane; b < DESC_COPIES * DESC_STRIDE; b += 32) hist[b] = 0u;
        __syncwarp();
        // |col*sa + row*ca| < 2.5 hw  and  |col*ca - row*sa| < 2.5 hw  (bins in (-1, 4)), widened
        const float lim = 2.5f * (float)hw + 0.5f;
        for (int row0 = -radius; row0 <= radius; row0 += 32) {
            const int row = row0 + lane;
            int lo = -radius, hi = radius;
            const int ny = row + y;
            if (row > radius || !(ny > 0 && ny < H - 1)) hi = lo - 1;
            lo = max(lo, 1 - x);
            hi = min(hi, W - 2 - x);
            {
                const float b1 = (float)row * ca, b2 = -(float)row * sa;
                // strip a*col + b in (-lim, lim)
                auto clip = [&](float a, float b) {
                    if (fabsf(a) < 1e-6f) {
                        if (!(fabsf(b) < lim + 1.0f)) hi = lo - 1;
                        return;
                    }
                    const float t0 = (-lim - b) / a, t1 = (lim - b) / a;
                    const float mn = fminf(t0, t1), mx = fmaxf(t0, t1);
                    if (mn > (float)lo) lo = max(lo, (int)floorf(mn));
                    if (mx < (float)hi) hi = min(hi, (int)ceilf(mx));
                };
                clip(sa, b1);
                clip(ca, b2);
            }
            const int cnt = max(0, hi - lo + 1);
            int off = cnt;  // inclusive scan
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(FULL, off, d);
                if (lane >= d) off += t;
            }
            const int total = __shfl_sync(FULL, off, 31);
            off -= cnt;  // exclusive
            // The non-empty rows, compacted: lane c then holds the c-th of them (first sample `c_off`, strictly
            // increasing; first column and row slot packed in `c_row`).  Inside the sample loop a lane finds its row
            // without a search: the rows that start in (s0, s0 + 32] set one bit each (REDUX.OR), a lane's row is the
            // row of sample s0 plus the bits below its own position.
            // (compaction by shuffles: shared memory for it would push five CTAs over the 196 KB carve-out step)
            const unsigned nonempty = __ballot_sync(FULL, cnt > 0);
            const int src = nth_set_bit(nonempty, lane);
            const int src_off = __shfl_sync(FULL, off, src & 31);
            const int c_row = __shfl_sync(FULL, lo * 32 + lane, src & 31);
            const int c_off = src < 32 ? src_off : 0x7fffffff;
            int kb = 0;   // compact row of sample s0
            for (int s0 = 0; s0 < total; s0 += 32) {
                const int s = s0 + lane;
                const unsigned p = (unsigned)(c_off - s0 - 1);
                const unsigned starts = __reduce_or_sync(FULL, p < 32u ? 1u << p : 0u);
                const int k = kb + __popc(starts & lt_mask);
                kb += __popc(starts);
                const int off_k = __shfl_sync(FULL, c_off, k);
                const int row_k = __shfl_sync(FULL, c_row, k);
                if (s >= total) continue;
                const int col = (row_k >> 5) + (s - off_k);
                const int rw = row0 + (row_k & 31);
                // (contractions written out: which product is rounded first must not depend on the compiler's mood --
                // it flipped once and moved one descriptor byte in 3327 keypoints)
                const float fcol = (float)col, frow = (float)rw;
                const float rr = fmaf(ca, frow, __fmul_rn(sa, fcol)) * inv_hw;
                const float cr = fmaf(ca, fcol, -__fmul_rn(sa, frow)) * inv_hw;
                const float rb = rr + 1.5f, cb = cr + 1.5f;  // + DESC_HIST_WIDTH/2 - 0.5 (integer 4/2)
                if (!(rb > -1.0f && rb < 4.0f && cb > -1.0f && cb < 4.0f)) continue;
                const int ic = (rw + y) * pitch + (col + x);   // 32-bit indices: a plane has < 2^31 elements
                const int iu = ic - pitch, id = ic + pitch;
                const float dx = ldg(img + ic + 1) - ldg(img + ic - 1);
                const float dy = ldg(img + iu) - ldg(img + id);
                const float g2 = fmaf(dx, dx, __fmul_rn(dy, dy));
                const float mag = g2 >= kFltMin ? g2 * rsqrt_ftz(g2) : 0.f;
#ifndef SB_DESC_DIET
#define SB_DESC_DIET 0
#endif
#if SB_DESC_DIET
                // Conversion-pipe diet -- MEASURED SLOWER, kept off (0.382 ms against 0.369 ms at 4K: the kernel is
                // issue-bound at 77 %, and the rounded 64-bit split costs more slots than the conversions it saves).
                // Once the histogram copies sat in different banks, the conversion / special-function pipe (16 lanes
                // per SM) became this kernel's busiest unit (ncu: xu 59 %): 4 FRND + 11 F2I + 3 MUFU per sample.  Here
                // the angle is wrapped with compares, the floors come
                // from the round-to-nearest of (v - 0.5 + 1.5 * 2^23) -- floor(v) except on exact integers, where the
                // trilinear weights make either choice the same histogram -- and the split between the two
                // orientation bins is integer and ROUNDED (a1 = (u * fo32 + 2^31) >> 32, a0 = u - a1: exact mass, no
                // bias; a truncating split moved 5e-6 of every sample from bin o+1 to bin o and cost 0.8 % of the
                // bit-exact descriptors), leaving 5 F2I + 3 MUFU.
                float ang = fast_atan2(dy, dx) - pori;  // in (-3pi, pi]
                if (ang < 0.f) ang += 6.283185307179586f;
                if (ang < 0.f) ang += 6.283185307179586f;
                if (ang >= 6.283185307179586f) ang -= 6.283185307179586f;
                const float ob = ang * (8.0f / 6.283185307179586f);
                const float wgt = exp_ftz(-(rr * rr + cr * cr) * 0.125f);
                const float m = mag * wgt * fix;
                constexpr float kMagic = 12582912.0f;          // 1.5 * 2^23, bit pattern 0x4B400000
                const float tr = (rb - 0.5f) + kMagic, tc = (cb - 0.5f) + kMagic, to = (ob - 0.5f) + kMagic;
                const int br = min(max(__float_as_int(tr) - 0x4B400000, -1), 4);
                const int bc = min(max(__float_as_int(tc) - 0x4B400000, -1), 4);
                const int bo = __float_as_int(to) - 0x4B400000;
                const float fr = rb - (float)br, fc = cb - (float)bc;
                const float fo = fminf(fmaxf(ob - (to - kMagic), 0.f), 1.f);
                // trilinear spread (sift.cpp:541-571) into the 6x6 padded grid: rows / columns -1 and 4
                // (dropped by the reference) land in the border, so no range tests are needed
                unsigned* cell = my_hist + ((br + 1) * DESC_GRID + (bc + 1)) * 8;
                const int o0 = bo & 7, o1 = (bo + 1) & 7;
                const float mr0 = m * (1.0f - fr), mr1 = m * fr;
                const unsigned u00 = __float2uint_rn(mr0 * (1.0f - fc)), u01 = __float2uint_rn(mr0 * fc);
                const unsigned u10 = __float2uint_rn(mr1 * (1.0f - fc)), u11 = __float2uint_rn(mr1 * fc);
                const unsigned long long fo32 = (unsigned long long)__float2uint_rn(fo * 4294967296.0f);   // saturates at 2^32 - 1
                const unsigned a00 = (unsigned)((u00 * fo32 + 0x80000000ull) >> 32);
                const unsigned a01 = (unsigned)((u01 * fo32 + 0x80000000ull) >> 32);
                const unsigned a10 = (unsigned)((u10 * fo32 + 0x80000000ull) >> 32);
                const unsigned a11 = (unsigned)((u11 * fo32 + 0x80000000ull) >> 32);
                atomicAdd(cell + o0, u00 - a00);
                atomicAdd(cell + o1, a00);
                atomicAdd(cell + 8 + o0, u01 - a01);
                atomicAdd(cell + 8 + o1, a01);
                atomicAdd(cell + DESC_GRID * 8 + o0, u10 - a10);
                atomicAdd(cell + DESC_GRID * 8 + o1, a10);
                atomicAdd(cell + DESC_GRID * 8 + 8 + o0, u11 - a11);
                atomicAdd(cell + DESC_GRID * 8 + 8 + o1, a11);
#else
                // floors as float -> int (round down) -> float: one conversion-pipe (XU) operation each instead of two
                // (FRND + F2I); the int -> float direction runs on the FP32 pipe.  The XU pipe is this kernel's busiest
                // unit (ncu: 69 %), the values are small integers, the results are the same bits.
                float ang = fast_atan2(dy, dx) - pori;  // in (-3pi, pi]
                ang = fmaf(-6.283185307179586f, (float)__float2int_rd(ang * (1.0f / 6.283185307179586f)), ang);
                if (ang < 0.f) ang = 0.f;
                if (ang >= 6.283185307179586f) ang -= 6.283185307179586f;
                const float ob = ang * (8.0f / 6.283185307179586f);
                const float wgt = exp_ftz(-fmaf(rr, rr, __fmul_rn(cr, cr)) * 0.125f);
                const float m = __fmul_rn(__fmul_rn(mag, wgt), fix);
                const int br = __float2int_rd(rb), bc = __float2int_rd(cb), bo = __float2int_rd(ob);
                const float fbr = (float)br, fbc = (float)bc, fbo = (float)bo;
                const float fr = rb - fbr, fc = cb - fbc, fo = ob - fbo;
                // trilinear spread (sift.cpp:541-571) into the 6x6 padded grid: rows / columns -1 and 4
                // (dropped by the reference) land in the border, so no range tests are needed
                unsigned* cell = my_hist + ((br + 1) * DESC_GRID + (bc + 1)) * 8;
                const int o0 = bo & 7, o1 = (bo + 1) & 7;
                const float w0 = 1.0f - fo;
                const float v00 = m * (1.0f - fr) * (1.0f - fc), v01 = m * (1.0f - fr) * fc;
                const float v10 = m * fr * (1.0f - fc), v11 = m * fr * fc;
                atomicAdd(cell + o0, __float2uint_rn(v00 * w0));
                atomicAdd(cell + o1, __float2uint_rn(v00 * fo));
                atomicAdd(cell + 8 + o0, __float2uint_rn(v01 * w0));
                atomicAdd(cell + 8 + o1, __float2uint_rn(v01 * fo));
                atomicAdd(cell + DESC_GRID * 8 + o0, __float2uint_rn(v10 * w0));
                atomicAdd(cell + DESC_GRID * 8 + o1, __float2uint_rn(v10 * fo));
                atomicAdd(cell + DESC_GRID * 8 + 8 + o0, __float2uint_rn(v11 * w0));
                atomicAdd(cell + DESC_GRID * 8 + 8 + o1, __float2uint_rn(v11 * fo));
#endif
            }
        }
        __syncwarp();
        // lane L owns bins 4L..4L+3 (consecutive bytes of the descriptor)
        double hv[4];
        double ss = 0.0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            unsigned long long t = 0;
            // bin 4L+k = (row L/8, column (L/2)%4, orientation 4(L%2)+k) of the inner 4x4 cells
            const int cellw = (((lane >> 3) + 1) * DESC_GRID + ((lane >> 1) & 3) + 1) * 8 + 4 * (lane & 1) + k;
#pragma unroll
            for (int cpy = 0; cpy < DESC_COPIES; ++cpy) t += hist[cpy * DESC_STRIDE + cellw];
            hv[k] = (double)t;
            ss += hv[k] * hv[k];
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) ss += __shfl_xor_sync(FULL, ss, d);
        double inv_n = 1.0 / sqrt(ss);
        ss = 0.0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            hv[k] *= inv_n;
            if (hv[k] > 0.2) hv[k] = 0.2;  // DESC_MAGNITUDE_THR
            ss += hv[k] * hv[k];
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) ss += __shfl_xor_sync(FULL, ss, d);
        inv_n = 1.0 / sqrt(ss);
        uint32_t packed = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int q = (int)floor(512.0 * hv[k] * inv_n);  // INT_DESCR_FCTR
            q = min(q, 255);
            q = max(q, 0);  // NaN (empty histogram) -> 0; the reference's cast of NaN is undefined
            packed |= (uint32_t)q << (8 * k);
        }
        uint8_t* rec = records + (size_t)i * 168;
        *reinterpret_cast<uint32_t*>(rec + 40 + 4 * lane) = packed;
        *reinterpret_cast<uint32_t*>(desc + (size_t)i * 128 + 4 * lane) = packed;
        if (lane < 5) {
            // the 40-byte head of the record is copied from the keypoint list again (L2 hit) instead of being kept in
            // 10 registers through the sample loop; the pointer is laundered so that the loads are not merged with
            // the ones at the top
            const unsigned long long* src = reinterpret_cast<const unsigned long long*>(oriented + final_order[i]);
            asm volatile("" : "+l"(src));
            *reinterpret_cast<unsigned long long*>(rec + 8 * lane) = src[lane];
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// Canary audit (debug): the arena is filled with a NaN pattern before a detect call; afterwards every element of a
// plane the pipeline writes must have been overwritten inside the image (x < w) and must still hold the pattern in
// the row padding (w <= x < pitch).  Catches out-of-bounds and missed stores of the scale-space kernels without a
// sanitizer.  counts[0] = padding elements overwritten, counts[1] = image elements never written.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_canary_fill(unsigned* __restrict__ p, size_t n, unsigned pattern) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = pattern;
}
__global__ void __launch_bounds__(256) k_canary_plane(const unsigned* __restrict__ plane, int w, int h, int pitch,
                                                       unsigned pattern, int expect_written,
                                                       unsigned long long* __restrict__ counts) {
    unsigned long long bad_pad = 0, unwritten = 0;
    const size_t n = (size_t)pitch * h;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % (size_t)pitch);
        const bool is_pattern = plane[i] == pattern;
        // (a row's last float4 store may spill into the first 1-3 padding floats when w is not a multiple of 4:
        // by design, the pitch is a multiple of 32 floats; everything from round_up(w, 4) on must be untouched)
        if (x >= ((w + 3) & ~3)) bad_pad += is_pattern ? 0 : 1;
        else if (x >= w) continue;
        else if (expect_written) unwritten += is_pattern ? 1 : 0;
        else bad_pad += is_pattern ? 0 : 1;          // a plane nobody may write at all
    }
    if (bad_pad) atomicAdd(&counts[0], bad_pad);
    if (unwritten) atomicAdd(&counts[1], unwritten);
}
__global__ void __launch_bounds__(256) k_canary_tail(const unsigned* __restrict__ p, size_t n, unsigned pattern,
                                                      unsigned long long* __restrict__ counts) {
    unsigned long long bad = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        bad += p[i] == pattern ? 0 : 1;
    if (bad) atomicAdd(&counts[0], bad);
}

}  // namespace

cudaError_t launch_canary_fill(void* p, size_t words, unsigned pattern, int sm_count, cudaStream_t s) {
    k_canary_fill<<<sm_count * 8, 256, 0, s>>>(static_cast<unsigned*>(p), words, pattern);
    return cudaGetLastError();
}
cudaError_t launch_canary_plane(const float* plane, int w, int h, int pitch, unsigned pattern, int expect_written,
                                unsigned long long* counts, int sm_count, cudaStream_t s) {
    k_canary_plane<<<sm_count * 4, 256, 0, s>>>(reinterpret_cast<const unsigned*>(plane), w, h, pitch, pattern,
                                                 expect_written, counts);
    return cudaGetLastError();
}
cudaError_t launch_canary_tail(const void* p, size_t words, unsigned pattern, unsigned long long* counts, int sm_count,
                               cudaStream_t s) {
    k_canary_tail<<<sm_count * 4, 256, 0, s>>>(static_cast<const unsigned*>(p), words, pattern, counts);
    return cudaGetLastError();
}

cudaError_t launch_range(const float* px, size_t n, float* range, int sm_count, cudaStream_t s) {
    const float init[2] = {INFINITY, -INFINITY};
    cudaError_t e = cudaMemcpyAsync(range, init, sizeof init, cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return e;
    k_range<<<sm_count * 4, 256, 0, s>>>(px, n, range);
    return cudaGetLastError();
}

static int extrema4_rows(long long px);

cudaError_t launch_extrema(const OctaveDesc& oct, int octave, int dogs, int border, int threshold, Cand* cands,
                           int cap, Counters* counters, CandCube* cubes, int cube_cap, int form, cudaStream_t s) {
    if (oct.w < 2 * border + 1 || oct.h < 2 * border + 1) return cudaSuccess;
    if (border != 1) {
        dim3 grid((oct.w - 2 * border + 255) / 256, oct.h - 2 * border);
        k_extrema_window<<<grid, 256, 0, s>>>(oct, octave, dogs, border, (float)threshold, cands, cap, counters);
        return cudaGetLastError();
    }
    const long long px = (long long)oct.w * oct.h;
    if (form != 1) {   // form 1: the one-column-per-lane kernel (kept for comparison; same candidate set)
        // rows per warp: long walks on large octaves (2 halo rows each), short ones where the grid would not fill the GPU
        const int rows = extrema4_rows(px);
        dim3 grid(oct.w / EX4_STRIP + 1, (oct.h - 2 + 4 * rows - 1) / (4 * rows));
        const int strips = oct.w / EX4_STRIP + 1;
        const dim3 grid_x((strips + 3) / 4, (oct.h - 2 + rows - 1) / rows);   // XW = 4: warps side by side
#define SB_EX4(ND, SL, CT) k_extrema4<ND, SL, CT, 1><<<grid, 128, 0, s>>>(oct, octave, (float)threshold, rows, cands, cap, counters, cubes, cube_cap)
#define SB_EX4X(ND, SL, CT) k_extrema4<ND, SL, CT, 4><<<grid_x, 128, 0, s>>>(oct, octave, (float)threshold, rows, cands, cap, counters, cubes, cube_cap)
        // measured at 4K (all octaves, ms): warps side by side 0.279, warps stacked 0.290; a deeper window (5 / 6
        // slots) 0.307 / 0.349, 4 CTAs per SM 0.290 / 0.279 (no gain), no prefetch slot 0.316
        switch (dogs) {
            case 4: SB_EX4X(4, 4, 3); break;
            case 5:
                if (form == 2) SB_EX4(5, 4, 3);        // experiments: warps stacked in y instead of side by side
                else SB_EX4X(5, 4, 3);
                break;
            case 6: SB_EX4X(6, 4, 3); break;
            case 7: SB_EX4X(7, 4, 3); break;
            default: return cudaErrorInvalidValue;
        }
#undef SB_EX4
#undef SB_EX4X
        return cudaGetLastError();
    }
    // rows per warp: 32 on large octaves, 8 on mid-size ones, 2 on tiny ones (more warps, shorter serial walks)
    const int rows = px >= (16ll << 20) ? 32 : px >= (1ll << 18) ? 8 : 2;
    dim3 grid((oct.w - 2 + 29) / 30, (oct.h - 2 + 8 * rows - 1) / (8 * rows));
#define SB_EX(ND)                                                                                                  \
    case ND:                                                                                                       \
        if (rows == 32) k_extrema<32, ND><<<grid, 256, 0, s>>>(oct, octave, (float)threshold, cands, cap, counters); \
        else if (rows == 8) k_extrema<8, ND><<<grid, 256, 0, s>>>(oct, octave, (float)threshold, cands, cap, counters); \
        else k_extrema<2, ND><<<grid, 256, 0, s>>>(oct, octave, (float)threshold, cands, cap, counters);            \
        break;
    switch (dogs) {
        SB_EX(4) SB_EX(5) SB_EX(6) SB_EX(7)
        default: return cudaErrorInvalidValue;
    }
#undef SB_EX
    return cudaGetLastError();
}

// rows per warp of the four-column scan: long walks on large octaves (2 halo rows each), short ones where the grid
// would not fill the GPU
static int extrema4_rows(long long px) {
    static const int rows_mid = getenv("SIFT_B200_EX_ROWS_MID") ? atoi(getenv("SIFT_B200_EX_ROWS_MID")) : 16;   // experiments
    return px >= (16ll << 20) ? 32 : px >= (1ll << 20) ? rows_mid : px >= (1ll << 16) ? 4 : 2;
}

bool extrema_multi_supported(int border, int form) { return border == 1 && form != 1 && form != 2; }

// octaves first .. first + n - 1 in one launch (window 3 only); same candidate set as n launch_extrema calls
cudaError_t launch_extrema_multi(const OctaveDesc* octs, int first, int n, int dogs, int threshold, Cand* cands, int cap,
                                 Counters* counters, CandCube* cubes, int cube_cap, cudaStream_t s) {
    if (n > kMaxExtremaMulti) return cudaErrorInvalidValue;
    ExtremaMulti m;
    memset(&m, 0, sizeof m);
    int total = 0;
    for (int i = 0; i < n; ++i) {
        const OctaveDesc& od = octs[first + i];
        if (od.w < 3 || od.h < 3) continue;
        const int j = m.n++;
        const int rows = extrema4_rows((long long)od.w * od.h);
        const int strips = od.w / EX4_STRIP + 1;
        m.gx[j] = (strips + 3) / 4;
        m.rows[j] = rows;
        m.octave[j] = first + i;
        m.oct[j] = od;
        m.cta_begin[j] = total;
        total += m.gx[j] * ((od.h - 2 + rows - 1) / rows);
    }
    m.cta_begin[m.n] = total;
    if (total == 0) return cudaSuccess;
    switch (dogs) {
        case 4: k_extrema4_multi<4><<<total, 128, 0, s>>>(m, (float)threshold, cands, cap, counters, cubes, cube_cap); break;
        case 5: k_extrema4_multi<5><<<total, 128, 0, s>>>(m, (float)threshold, cands, cap, counters, cubes, cube_cap); break;
        case 6: k_extrema4_multi<6><<<total, 128, 0, s>>>(m, (float)threshold, cands, cap, counters, cubes, cube_cap); break;
        case 7: k_extrema4_multi<7><<<total, 128, 0, s>>>(m, (float)threshold, cands, cap, counters, cubes, cube_cap); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

bool extrema_hands_cubes(int border, int form) { return border == 1 && form != 1; }

cudaError_t launch_refine(const PyramidDesc* d_pyr, const Cand* cands, KpCore* raw, Counters* counters,
                          const CandCube* cubes, int cube_cap, const StageParams& sp, int sm_count, cudaStream_t s) {
    k_refine<<<sm_count * 16, 128, 0, s>>>(d_pyr, cands, raw, counters, cubes, cube_cap, sp);
    return cudaGetLastError();
}

cudaError_t launch_orient(const PyramidDesc* d_pyr, const KpCore* raw, KpCore* oriented, Counters* counters,
                          const StageParams& sp, int sm_count, cudaStream_t s) {
    if (sp.num_bins == kOriBins)
        k_orient<kOriBins><<<sm_count * SB_ORI_CTAS, 256, 0, s>>>(d_pyr, raw, oriented, counters, sp);
    else
        k_orient<0><<<sm_count * SB_ORI_CTAS, 256, 0, s>>>(d_pyr, raw, oriented, counters, sp);
    return cudaGetLastError();
}

cudaError_t launch_sort_dedup(const KpCore* oriented, Counters* counters, const SortScratch& ss,
                              const StageParams& sp, int sm_count, cudaStream_t s, int* launches) {
    cudaError_t e;
    k_sort_clear<<<(ss.nb + 256) / 256, 256, 0, s>>>(ss);
    k_bucket_count<<<sm_count * 2, 256, 0, s>>>(oriented, counters, ss, sp.cap_oriented);
    k_scan<<<1, 1024, 0, s>>>(ss.bucket_cnt, ss.bucket_off, ss.nb, nullptr);
    k_bucket_scatter<<<sm_count * 2, 256, 0, s>>>(oriented, counters, ss, sp.cap_oriented);
    k_bucket_rank<<<sm_count * 4, 256, 0, s>>>(oriented, ss);
    k_scan<<<1, 1024, 0, s>>>(ss.uniq_cnt, ss.uniq_off, ss.nb, &counters->n_final);
    k_bucket_gather<<<sm_count * 2, 256, 0, s>>>(ss);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (launches) *launches += 7;
    return cudaSuccess;
}

cudaError_t launch_describe(const PyramidDesc* d_pyr, const KpCore* oriented, const int* final_order,
                            Counters* counters, uint8_t* records, uint8_t* desc, int cap_final,
                            const StageParams& sp, int sm_count, cudaStream_t s) {
    k_describe<<<sm_count * DESC_CTAS, DESC_WARPS * 32, 0, s>>>(d_pyr, oriented, final_order, counters, records, desc, cap_final, sp);
    return cudaGetLastError();
}

}  // namespace sb

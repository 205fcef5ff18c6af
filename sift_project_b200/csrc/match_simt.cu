// Brute-force descriptor matcher, SIMT path (sm_100a): exact integer squared distances with
// __dp4a, strict-'<' top-2 (lowest j wins ties), Lowe ratio test and ordered emission.
// Replaces euclid_dist (sift.cpp:688-695) + match_keypoints (sift.cpp:783-815) for problems too
// small to fill tcgen05 tiles; the tensor-core path (match_tc.cu) shares the emit kernel.
#include <limits.h>

#include "common.cuh"
#include "kernels.h"

namespace sb {

namespace {

constexpr int MT = 128;  // A rows per CTA (one per thread)
constexpr int BT = 32;   // B rows staged per tile

__global__ void __launch_bounds__(256) k_norms(const uint8_t* __restrict__ d, int n, int* __restrict__ norms) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4* p = reinterpret_cast<const uint4*>(d + (size_t)i * 128);
    unsigned s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint4 q = __ldg(p + k);
        s = __dp4a(q.x, q.x, s); s = __dp4a(q.y, q.y, s); s = __dp4a(q.z, q.z, s); s = __dp4a(q.w, q.w, s);
    }
    norms[i] = (int)s;
}

__global__ void __launch_bounds__(MT)
k_match_simt(const uint8_t* __restrict__ a, int na, const uint8_t* __restrict__ b, int nb,
             const int* __restrict__ norms_a, const int* __restrict__ norms_b, int rows_per_split,
             int* __restrict__ part_idx, int* __restrict__ part_d1, int* __restrict__ part_d2) {
    __shared__ __align__(16) unsigned s_b[BT][32];
    __shared__ int s_nb[BT];
    const int i = blockIdx.x * MT + threadIdx.x;
    const int irow = min(i, na - 1);
    unsigned ar[32];
    {
        const uint4* p = reinterpret_cast<const uint4*>(a + (size_t)irow * 128);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint4 q = __ldg(p + k);
            ar[4 * k] = q.x; ar[4 * k + 1] = q.y; ar[4 * k + 2] = q.z; ar[4 * k + 3] = q.w;
        }
    }
    const int j_begin = blockIdx.y * rows_per_split;
    const int j_end = min(nb, j_begin + rows_per_split);
    int best = INT_MAX, second = INT_MAX, best_j = -1;
    for (int j0 = j_begin; j0 < j_end; j0 += BT) {
        const int rows = min(BT, j_end - j0);
        __syncthreads();
        for (int t = threadIdx.x; t < rows * 32; t += MT)
            s_b[t >> 5][t & 31] = __ldg(reinterpret_cast<const unsigned*>(b + (size_t)j0 * 128) + t);
        if (threadIdx.x < rows) s_nb[threadIdx.x] = norms_b[j0 + threadIdx.x];
        __syncthreads();
        for (int r = 0; r < rows; ++r) {
            unsigned dot = 0;
            const uint4* row = reinterpret_cast<const uint4*>(s_b[r]);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint4 q = row[k];
                dot = __dp4a(ar[4 * k], q.x, dot);
                dot = __dp4a(ar[4 * k + 1], q.y, dot);
                dot = __dp4a(ar[4 * k + 2], q.z, dot);
                dot = __dp4a(ar[4 * k + 3], q.w, dot);
            }
            const int t = s_nb[r] - 2 * (int)dot;  // ||b||^2 - 2ab; ||a||^2 is added at the end
            if (t < best) { second = best; best = t; best_j = j0 + r; }
            else if (t < second) second = t;
        }
    }
    if (i < na) {
        const int n2 = norms_a[i];
        const size_t o = (size_t)blockIdx.y * na + i;
        part_idx[o] = best_j;
        part_d1[o] = best == INT_MAX ? INT_MAX : best + n2;
        part_d2[o] = second == INT_MAX ? INT_MAX : second + n2;
    }
}

// Merge the per-split partial results in ascending-j order (split s covers lower j than s+1).
__global__ void __launch_bounds__(256)
k_match_merge(const int* __restrict__ part_idx, const int* __restrict__ part_d1, const int* __restrict__ part_d2,
              int splits, int na, int* __restrict__ best_idx, int* __restrict__ best_d2, int* __restrict__ second_d2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= na) return;
    int best = INT_MAX, second = INT_MAX, idx = -1;
    for (int s = 0; s < splits; ++s) {
        const size_t o = (size_t)s * na + i;
        const int b1 = part_d1[o], b2 = part_d2[o];
        if (b1 < best) {
            second = min(best, b2);
            best = b1;
            idx = part_idx[o];
        } else {
            second = min(second, b1);
        }
    }
    best_idx[i] = idx;
    best_d2[i] = best;
    second_d2[i] = second;
}

// Lowe ratio test exactly as the reference evaluates it (FP64 sqrt of the integer squared
// distances, sift.cpp:694, :808) and emission in ascending i.  Single CTA, ordered compaction.
__global__ void __launch_bounds__(1024)
k_match_emit(const int* __restrict__ best_idx, const int* __restrict__ best_d2, const int* __restrict__ second_d2,
             int na, int nb, double ratio, int* __restrict__ out_ia, int* __restrict__ out_ib,
             double* __restrict__ out_dist, int cap, int* __restrict__ out_count) {
    __shared__ int s_warp[32];
    __shared__ int s_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int i0 = 0; i0 < na; i0 += 1024) {
        const int i = i0 + threadIdx.x;
        bool hit = false;
        double d1 = 0.0;
        int j = -1;
        if (i < na && nb > 0) {
            j = best_idx[i];
            d1 = sqrt((double)best_d2[i]);
            const int s2 = second_d2[i];
            const double d2 = (s2 == INT_MAX) ? 1.7976931348623157e308 : sqrt((double)s2);
            hit = j >= 0 && d1 < (ratio * d2);
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) s_warp[warp] = __popc(m);
        __syncthreads();
        if (warp == 0) {
            const int v = s_warp[lane];
            int inc = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc += t;
            }
            s_warp[lane] = inc - v;
        }
        __syncthreads();
        const int base = s_base;
        if (hit) {
            const int slot = base + s_warp[warp] + __popc(m & ((1u << lane) - 1));
            if (slot < cap) { out_ia[slot] = i; out_ib[slot] = j; out_dist[slot] = d1; }
        }
        __syncthreads();
        if (threadIdx.x == 1023) s_base = base + s_warp[31] + __popc(m);
        __syncthreads();
    }
    if (threadIdx.x == 0) *out_count = s_base;
}

}  // namespace

cudaError_t launch_norms(const uint8_t* d, int n, int* norms, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    k_norms<<<(n + 255) / 256, 256, 0, s>>>(d, n, norms);
    return cudaGetLastError();
}

cudaError_t launch_match_simt(const uint8_t* a, int na, const uint8_t* b, int nb, int* best_idx,
                              int* best_d2, int* second_d2, const MatchScratch& ms, int sm_count,
                              cudaStream_t s, int* launches) {
    if (na <= 0) return cudaSuccess;
    cudaError_t e;
    if ((e = launch_norms(a, na, ms.norms_a, s)) != cudaSuccess) return e;
    if ((e = launch_norms(b, nb, ms.norms_b, s)) != cudaSuccess) return e;
    const int row_blocks = (na + MT - 1) / MT;
    int splits = (2 * sm_count + row_blocks - 1) / row_blocks;
    splits = max(1, min(splits, ms.max_splits));
    int rows_per_split = ((nb + splits - 1) / splits + BT - 1) / BT * BT;
    if (rows_per_split <= 0) rows_per_split = BT;
    splits = max(1, (nb + rows_per_split - 1) / rows_per_split);
    dim3 grid(row_blocks, splits);
    k_match_simt<<<grid, MT, 0, s>>>(a, na, b, nb, ms.norms_a, ms.norms_b, rows_per_split, ms.part_idx,
                                     ms.part_d1, ms.part_d2);
    k_match_merge<<<(na + 255) / 256, 256, 0, s>>>(ms.part_idx, ms.part_d1, ms.part_d2, splits, na, best_idx,
                                                  best_d2, second_d2);
    if (launches) *launches += 2 + (na > 0) + (nb > 0);
    return cudaGetLastError();
}

cudaError_t launch_match_merge(const MatchScratch& ms, int splits, int na, int* best_idx, int* best_d2,
                               int* second_d2, cudaStream_t s) {
    k_match_merge<<<(na + 255) / 256, 256, 0, s>>>(ms.part_idx, ms.part_d1, ms.part_d2, splits, na, best_idx,
                                                  best_d2, second_d2);
    return cudaGetLastError();
}

cudaError_t launch_match_emit(const int* best_idx, const int* best_d2, const int* second_d2, int na, int nb,
                              double ratio, int* out_ia, int* out_ib, double* out_dist, int cap,
                              int* out_count, cudaStream_t s) {
    k_match_emit<<<1, 1024, 0, s>>>(best_idx, best_d2, second_d2, na, nb, ratio, out_ia, out_ib, out_dist, cap,
                                    out_count);
    return cudaGetLastError();
}

}  // namespace sb

// Brute-force descriptor matcher, tensor-core path (sm_100a): tcgen05.mma kind::i8 + TMA + TMEM.
//
// Replaces the N1 x N2 x 128 loop of match_keypoints / euclid_dist (sift.cpp:783-815, :688-695).
// ||a - b||^2 = ||a||^2 + ||b||^2 - 2 a.b : the a.b term is a dense u8 x u8 -> s32 contraction
// (exact integer arithmetic on the tensor cores), the norms are folded in by the epilogue, and the
// top-2 reduction over j runs straight out of tensor memory -- the N1 x N2 distance matrix is
// never materialised.
//
// One persistent CTA per SM, 18 warps:
//   warp 0       TMA producer: A tile (128 descriptors, resident per segment) and a 4-stage ring of
//                B tiles (256 descriptors = 32 KB each); 128-byte rows land 128B-swizzled, which is
//                exactly the K-major SWIZZLE_128B operand layout of tcgen05.mma.
//   warp 1       MMA issuer: per B tile four tcgen05.mma.kind::i8 (M128 x N256 x K32) into one of two
//                256-column TMEM accumulators; tcgen05.commit releases the smem stage and publishes
//                the accumulator.
//   warps 2..17  epilogue: each warp owns 32 TMEM lanes (rows) x 64 columns, tcgen05.ld 16 columns at a
//                time (double-buffered).  K is only 128, so the kernel is epilogue-paced and every
//                accumulator must cost as few issue slots as possible:
//                  * pre-filter on the raw dot product: key_j < bound  =>  dot_j > theta (a 3-input
//                    max tree over 16 values and one compare, no shared-memory traffic);
//                  * only groups that pass form exact keys ((||b_j||^2 - 2 a.b_j) << 8) | (j & 255)
//                    (one IMAD each) and merge their two smallest into the row's running
//                    (best, second) with a short tournament.  Keys order by (distance, column), so
//                    the reference's "lowest j wins ties" falls out of integer min();
//                  * the four column slices of a row exchange (best, second) once per tile so that
//                    each filters against the row's second best over ALL columns seen so far.
// The flattened (row block x B tile) grid is cut into one contiguous range per CTA; a range is
// walked as segments (runs inside one row block) whose partial (index, d1, d2) are merged in column
// order by k_match_tc_merge, then the shared emit kernel applies Lowe's ratio test.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <limits.h>

#include "common.cuh"
#include "kernels.h"

namespace sb {

namespace {

constexpr int BM = 128;            // rows of A per tile (TMEM lanes)
constexpr int BN = 256;            // rows of B per tile (TMEM columns of one accumulator)
constexpr int KB = 128;            // descriptor bytes = K
constexpr int STAGES = 4;          // B ring depth
constexpr int A_BYTES = BM * KB;   // 16 KB
constexpr int B_BYTES = BN * KB;   // 32 KB
constexpr int EPI_WARPS = 16;          // 4 per scheduler: TMEM-lane quarter x 64-column slice
constexpr int SLICE = BN / (EPI_WARPS / 4);  // columns per epilogue warp
constexpr int NTHREADS = 32 * (2 + EPI_WARPS);
constexpr int PAD_NORM = (1 << 23) - 1;  // > 128 * 255^2: padded columns never beat a real one

struct __align__(1024) SmemLayout {
    uint8_t a[2][A_BYTES];
    uint8_t b[STAGES][B_BYTES];
    int cprime[EPI_WARPS][SLICE];    // per-warp staging of the per-column constants
    int merge[EPI_WARPS / 4][BM][4]; // column slices 1.. -> slice 0 hand-over
    int2 exch[2][EPI_WARPS / 4][BM]; // per-tile exchange of (best, second) between the slices of a row
    unsigned long long full_b[STAGES], empty_b[STAGES];
    unsigned long long a_full[2], a_empty[2];
    unsigned long long tmem_full[2], tmem_empty[2];
    uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(void* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(void* bar, int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(void* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded spin: a protocol bug traps (CUDA error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(void* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    for (uint32_t spin = 0;; ++spin) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(a), "r"(parity) : "memory");
        if (done) return;
        if (spin > (1u << 26)) __trap();
    }
}

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, void* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}

// K-major, 128-byte-swizzled shared-memory operand descriptor: rows of 128 bytes, 8-row groups
// 1024 bytes apart (SBO), descriptor version 1 (sm_100), layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;            // LBO (ignored for swizzled K-major), canonical value
    d |= (uint64_t)(1024 >> 4) << 32;  // SBO
    d |= (uint64_t)1 << 46;            // version
    d |= (uint64_t)2 << 61;            // SWIZZLE_128B
    return d;
}

// Instruction descriptor, kind::i8: D = s32, A = B = unsigned 8 bit, both K-major, N = 256, M = 128.
constexpr uint32_t kIdesc = (2u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(kIdesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(void* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ int max3(int a, int b, int c) { return max(max(a, b), c); }
__device__ __forceinline__ int min3(int a, int b, int c) { return min(min(a, b), c); }
// (lo, hi) <- the two smallest of two sorted pairs
__device__ __forceinline__ void merge2(int a1, int a2, int b1, int b2, int& lo, int& hi) {
    lo = min(a1, b1);
    hi = min3(max(a1, b1), a2, b2);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Per-column constants: c'_j = (||b_j||^2 << 8) | (j & 255); entries j >= n up to the tile
// boundary get PAD_NORM so that the zero rows TMA fills in can never win.  One block = one tile of
// 256 columns; it also records the smallest norm of each 128-column half (the epilogue's
// conservative pre-filter needs a lower bound of ||b_j||^2 over the columns it is about to skip).
__global__ void __launch_bounds__(BN)
k_cprime(const uint8_t* __restrict__ d, int n, int n_padded, int* __restrict__ cprime, int* __restrict__ half_min) {
    __shared__ int s_min[BN / 32];
    const int j = blockIdx.x * BN + threadIdx.x;
    int nrm = PAD_NORM;
    if (j < n) {
        const uint4* p = reinterpret_cast<const uint4*>(d + (size_t)j * 128);
        unsigned s = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint4 q = __ldg(p + k);
            s = __dp4a(q.x, q.x, s); s = __dp4a(q.y, q.y, s); s = __dp4a(q.z, q.z, s); s = __dp4a(q.w, q.w, s);
        }
        nrm = (int)s;
    }
    if (j < n_padded) cprime[j] = (nrm << 8) | (j & 255);
    int m = nrm;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) s_min[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 2) {
        const int* q = s_min + 4 * threadIdx.x;
        half_min[2 * blockIdx.x + threadIdx.x] = min(min(q[0], q[1]), min(q[2], q[3]));
    }
}

// Schedule: the m_tiles x n_tiles tile grid is flattened row-block-major and cut into gridDim.x
// contiguous ranges, one per CTA (perfect balance to within one tile, and each CTA sees long runs
// of columns for the same rows, which is what makes the "does it beat the running second best"
// test almost always false).  A range is walked as segments = maximal runs inside one row block;
// segment results go to part[c - first_cta(row block)][row] and are merged in that order.
struct Sched {
    long long T;  // total tiles
    int G;        // CTAs
    int n_tiles;
    __host__ __device__ long long range_begin(int c) const { return (long long)c * T / G; }
    __host__ __device__ int cta_of(long long tile) const { return (int)(((tile + 1) * G - 1) / T); }
};

// Merge the per-CTA partial results of each row in ascending-column order.
__global__ void __launch_bounds__(256)
k_match_tc_merge(const int* __restrict__ part_idx, const int* __restrict__ part_d1, const int* __restrict__ part_d2,
                 Sched sc, int na, int* __restrict__ best_idx, int* __restrict__ best_d2, int* __restrict__ second_d2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= na) return;
    const long long m = i / BM;
    const int parts = sc.cta_of((m + 1) * sc.n_tiles - 1) - sc.cta_of(m * sc.n_tiles) + 1;
    int best = INT_MAX, second = INT_MAX, idx = -1;
    for (int s = 0; s < parts; ++s) {
        const size_t o = (size_t)s * na + i;
        const int b1 = part_d1[o], b2 = part_d2[o];
        if (b1 < best) { second = min(best, b2); best = b1; idx = part_idx[o]; }
        else second = min(second, b1);
    }
    best_idx[i] = idx; best_d2[i] = best; second_d2[i] = second;
}

__global__ void __launch_bounds__(NTHREADS, 1)
k_match_tc(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, int na, int nb,
           const int* __restrict__ norms_a, const int* __restrict__ cprime_b, const int* __restrict__ half_min,
           Sched sc,
           int* __restrict__ part_idx, int* __restrict__ part_d1, int* __restrict__ part_d2) {
    extern __shared__ uint8_t smem_raw[];
    // align to 1024 B (swizzle atom) with plain pointer arithmetic so the compiler keeps the shared space
    SmemLayout& S = *reinterpret_cast<SmemLayout*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = sc.n_tiles;
    const long long r_begin = sc.range_begin(blockIdx.x), r_end = sc.range_begin(blockIdx.x + 1);

    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(&S.full_b[i], 1); mbar_init(&S.empty_b[i], 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&S.a_full[i], 1); mbar_init(&S.a_empty[i], 1);
            mbar_init(&S.tmem_full[i], 1); mbar_init(&S.tmem_empty[i], EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // TMEM: all 512 columns (two 128 x 256 s32 accumulators)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&S.tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = S.tmem_base;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0; int a_cnt = 0;
            for (long long seg = r_begin; seg < r_end;) {
                const int m_tile = (int)(seg / n_tiles), t0 = (int)(seg - (long long)m_tile * n_tiles);
                const int t1 = (int)min((long long)n_tiles, t0 + (r_end - seg));
                const int ab = a_cnt & 1;
                mbar_wait(&S.a_empty[ab], ((a_cnt >> 1) & 1) ^ 1);
                mbar_expect_tx(&S.a_full[ab], A_BYTES);
                tma_load_2d(S.a[ab], &map_a, 0, m_tile * BM, &S.a_full[ab]);
                ++a_cnt;
                for (int t = t0; t < t1; ++t) {
                    mbar_wait(&S.empty_b[stage], phase ^ 1);
                    mbar_expect_tx(&S.full_b[stage], B_BYTES);
                    tma_load_2d(S.b[stage], &map_b, 0, t * BN, &S.full_b[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                seg += t1 - t0;
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0; int a_cnt = 0; int acc = 0; uint32_t acc_phase = 0;
            for (long long seg = r_begin; seg < r_end;) {
                const int m_tile = (int)(seg / n_tiles), t0 = (int)(seg - (long long)m_tile * n_tiles);
                const int t1 = (int)min((long long)n_tiles, t0 + (r_end - seg));
                const int ab = a_cnt & 1;
                mbar_wait(&S.a_full[ab], (a_cnt >> 1) & 1);
                const uint64_t da = umma_desc(smem_u32(S.a[ab]));
                for (int t = t0; t < t1; ++t) {
                    mbar_wait(&S.tmem_empty[acc], acc_phase ^ 1);
                    mbar_wait(&S.full_b[stage], phase);
                    tc_fence_after();
                    const uint64_t db = umma_desc(smem_u32(S.b[stage]));
                    const uint32_t d_addr = tmem_base + (uint32_t)(acc * BN);
#pragma unroll
                    for (int k = 0; k < KB / 32; ++k)  // K = 32 bytes per instruction: +32 B = +2 in the address field
                        umma_i8(d_addr, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), k > 0 ? 1u : 0u);
                    umma_commit(&S.empty_b[stage]);   // smem stage reusable once these MMAs retire
                    umma_commit(&S.tmem_full[acc]);   // accumulator ready for the epilogue
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    acc ^= 1;
                    if (acc == 0) acc_phase ^= 1;
                }
                umma_commit(&S.a_empty[ab]);
                ++a_cnt;
                seg += t1 - t0;
            }
        }
    } else {
        // ===================== epilogue: top-2 straight out of TMEM =====================
        const int ew = warp - 2;
        const int quarter = warp & 3;        // TMEM lanes a warp may touch: 32 * (warp_id % 4)
        const int slice = ew >> 2;           // which SLICE columns of the 256
        const int row = quarter * 32 + lane; // row of the A tile == TMEM lane
        int* cp = S.cprime[ew];
        int acc = 0; uint32_t acc_phase = 0;
        for (long long seg = r_begin; seg < r_end;) {
            const int m_tile = (int)(seg / n_tiles), t0 = (int)(seg - (long long)m_tile * n_tiles);
            const int t1 = (int)min((long long)n_tiles, t0 + (r_end - seg));
            const int part = (int)blockIdx.x - sc.cta_of((long long)m_tile * n_tiles);
            int best = INT_MAX, second = INT_MAX, bj = 0;
            int bound = INT_MAX;   // upper bound of the row's second-best key over all column slices
            int2 cnext = __ldg(reinterpret_cast<const int2*>(cprime_b + (size_t)t0 * BN + slice * SLICE) + lane);
            for (int t = t0; t < t1; ++t) {
                __syncwarp();
                reinterpret_cast<int2*>(cp)[lane] = cnext;
                if (t + 1 < t1)
                    cnext = __ldg(reinterpret_cast<const int2*>(cprime_b + (size_t)(t + 1) * BN + slice * SLICE) + lane);
                // Pre-filter on the raw dot product: key_j = c'_j - 512 dot_j >= (nmin << 8) - 512 dot_j, so
                // key_j < second  =>  dot_j > ((nmin << 8) - second) / 512 =: theta (floor keeps it safe).
                const long long nmin8 = (long long)__ldg(half_min + 2 * t + (slice >> 1)) << 8;
                int theta = (int)max((nmin8 - (long long)bound) >> 9, -1ll);
                const int best_in = best;
                __syncwarp();
                mbar_wait(&S.tmem_full[acc], acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN + slice * SLICE);
                int v[2][16];
                tmem_ld16(taddr, v[0]);
#pragma unroll
                for (int ch = 0; ch < SLICE / 16; ++ch) {
                    tmem_ld_wait();
                    if (ch < SLICE / 16 - 1) {
                        tmem_ld16(taddr + (ch + 1) * 16, v[(ch + 1) & 1]);
                    } else {
                        // every accumulator value of this warp is in registers: hand the TMEM buffer back
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&S.tmem_empty[acc]);
                    }
                    const int* w = v[ch & 1];
                    const int g0 = max3(max3(w[0], w[1], w[2]), max3(w[3], w[4], w[5]), max(w[6], w[7]));
                    const int g1 = max3(max3(w[8], w[9], w[10]), max3(w[11], w[12], w[13]), max(w[14], w[15]));
                    if (max(g0, g1) > theta) {
                        // rare after warm-up: some row of this warp may have a new top-2 column among these 16
#pragma unroll
                        for (int g = 0; g < 2; ++g) {
                            if ((g == 0 ? g0 : g1) > theta) {
                                const int* u = w + 8 * g;
                                const int4 c0 = *reinterpret_cast<const int4*>(cp + ch * 16 + 8 * g);
                                const int4 c1 = *reinterpret_cast<const int4*>(cp + ch * 16 + 8 * g + 4);
                                // key = c' - 512 dot: the product alone reaches 4.3e9, so it is formed in unsigned
                                // arithmetic (defined wrap-around); the key itself is below 2^31
                                auto key = [](int dot, int c) { return (int)((unsigned)c - ((unsigned)dot << 9)); };
                                const int k0 = key(u[0], c0.x), k1 = key(u[1], c0.y);
                                const int k2 = key(u[2], c0.z), k3 = key(u[3], c0.w);
                                const int k4 = key(u[4], c1.x), k5 = key(u[5], c1.y);
                                const int k6 = key(u[6], c1.z), k7 = key(u[7], c1.w);
                                // tournament: the two smallest of the eight keys (short dependency chains) ...
                                int a1, a2, b1, b2, lo, hi;
                                merge2(min(k0, k1), max(k0, k1), min(k2, k3), max(k2, k3), a1, a2);
                                merge2(min(k4, k5), max(k4, k5), min(k6, k7), max(k6, k7), b1, b2);
                                merge2(a1, a2, b1, b2, lo, hi);
                                // ... merged into the row's running (best, second)
                                const int nb_ = min(best, lo);
                                second = min3(max(best, lo), second, hi);
                                best = nb_;
                                bound = min(bound, second);
                                theta = (int)max((nmin8 - (long long)bound) >> 9, -1ll);
                            }
                        }
                    }
                }
                // The key carries its in-tile column: recover the global column of a new best, then
                // drop the column bits -- across tiles only the distance orders candidates (an
                // earlier column wins a tie), so a later equal distance must not look "smaller".
                if (best != best_in) bj = t * BN + (best & 255);
                best &= ~255; second &= ~255;
                // The four column slices of a row each see a quarter of the columns; what matters for
                // skipping is the row's second best over ALL of them.  Exchange (best, second) once
                // per tile (double-buffered, one named barrier per TMEM-lane quarter) and take the
                // second smallest of the eight values as the bound for the next tile.
                {
                    S.exch[t & 1][slice][row] = make_int2(best, second);
                    asm volatile("bar.sync %0, 128;" ::"r"(2 + quarter) : "memory");
                    int g1 = INT_MAX, g2 = INT_MAX;
#pragma unroll
                    for (int o = 0; o < EPI_WARPS / 4; ++o) {
                        const int2 e = S.exch[t & 1][o][row];
                        // (g1, g2) <- two smallest of {g1, g2, e.x, e.y}, e.x <= e.y
                        const int n1 = min(g1, e.x);
                        g2 = min3(max(g1, e.x), g2, e.y);
                        g1 = n1;
                    }
                    bound = g2 & ~255;
                }
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
            // ---- merge the column slices of each row, add ||a||^2, store the segment's partial ----
            if (slice != 0) {
                int* mg = S.merge[slice][row];
                mg[0] = best; mg[1] = second; mg[2] = bj;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
            if (slice == 0) {
                int b = best, s2 = second, j = bj;
#pragma unroll
                for (int o = 1; o < EPI_WARPS / 4; ++o) {  // ascending column order within a tile
                    const int* mg = S.merge[o][row];
                    const int ob = mg[0], os = mg[1], oj = mg[2];
                    const int d0 = b >> 8, d1 = ob >> 8;
                    if (d1 < d0 || (d1 == d0 && oj < j)) {   // (distance, column) order
                        s2 = min(os, b); b = ob; j = oj;
                    } else {
                        s2 = min(s2, ob);
                    }
                }
                const int gi = m_tile * BM + row;
                if (gi < na) {
                    const int n2 = norms_a[gi];
                    const int db = b >> 8, ds = s2 >> 8;
                    const size_t o = (size_t)part * na + gi;
                    const bool has1 = db < PAD_NORM, has2 = ds < PAD_NORM;  // INT_MAX >> 8 == PAD_NORM
                    part_idx[o] = has1 ? j : -1;
                    part_d1[o] = has1 ? db + n2 : INT_MAX;
                    part_d2[o] = has2 ? ds + n2 : INT_MAX;
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
            seg += t1 - t0;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

PFN_cuTensorMapEncodeTiled g_encode = nullptr;

cudaError_t make_map(CUtensorMap* map, const uint8_t* base, int rows, int box_rows) {
    const cuuint64_t dims[2] = {(cuuint64_t)KB, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)KB};
    const cuuint32_t box[2] = {(cuuint32_t)KB, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t*>(base), dims, strides, box,
                                estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

constexpr size_t kSmemBytes = sizeof(SmemLayout) + 1024;

}  // namespace

cudaError_t match_tc_init() {
    if (g_encode == nullptr) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
        if (e != cudaSuccess) return e;
        if (q != cudaDriverEntryPointSuccess || fn == nullptr) return cudaErrorNotSupported;
        g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fn);
    }
    return cudaFuncSetAttribute(k_match_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
}

int match_tc_padded_rows(int nb) { return (nb + BN - 1) / BN * BN; }

// a, b: device, n x 128 u8, 16-byte aligned.  ms.norms_b must hold match_tc_padded_rows(nb) ints.
cudaError_t launch_match_tc(const uint8_t* a, int na, const uint8_t* b, int nb, int* best_idx, int* best_d2,
                            int* second_d2, const MatchScratch& ms, int sm_count, cudaStream_t s, int* launches) {
    cudaError_t e;
    CUtensorMap map_a, map_b;
    if ((e = make_map(&map_a, a, na, BM)) != cudaSuccess) return e;
    if ((e = make_map(&map_b, b, nb, BN)) != cudaSuccess) return e;
    const int n_pad = match_tc_padded_rows(nb);
    if ((e = launch_norms(a, na, ms.norms_a, s)) != cudaSuccess) return e;
    int* half_min = ms.norms_b + n_pad;  // 2 ints per tile, right behind the per-column constants
    k_cprime<<<n_pad / BN, BN, 0, s>>>(b, nb, n_pad, ms.norms_b, half_min);
    const int m_tiles = (na + BM - 1) / BM, n_tiles = n_pad / BN;
    Sched sc;
    sc.T = (long long)m_tiles * n_tiles;
    sc.n_tiles = n_tiles;
    // a row block may be cut into at most max_splits partials: G <= (max_splits - 1) * m_tiles
    sc.G = (int)min((long long)sm_count, min(sc.T, (long long)(ms.max_splits - 1) * m_tiles));
    k_match_tc<<<sc.G, NTHREADS, kSmemBytes, s>>>(map_a, map_b, na, nb, ms.norms_a, ms.norms_b, half_min, sc, ms.part_idx,
                                                   ms.part_d1, ms.part_d2);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    k_match_tc_merge<<<(na + 255) / 256, 256, 0, s>>>(ms.part_idx, ms.part_d1, ms.part_d2, sc, na, best_idx, best_d2,
                                                      second_d2);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (launches) *launches += 4;
    return cudaSuccess;
}

}  // namespace sb

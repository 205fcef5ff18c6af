// Streaming (rolling column-strip) form of the fused Gaussian cascade (sm_100a).  Included by pyramid.cu
// inside its anonymous namespace (uses CascadeArgs, BlurTaps, cp_async16, ru4).
//
// Same arithmetic as k_cascade / k_blur (image.cpp:156-238 per level, subtract image.cpp:30-36, nearest
// decimation image.cpp:41-55), bit-identical results, different data movement:
//
//   * a CTA owns a column strip of WS outputs and a band of rows and walks DOWN the rows, K rows per step;
//     there is no y halo to recompute (only a pipeline fill of 2 sum(R) rows per CTA) and no 2-D tile passing
//     through shared memory in barrier-separated passes;
//   * the levels of the cascade form a pipeline over shared-memory row rings: the warps of level l read
//     the rows that level l-1 completed in the previous step, so one CTA barrier per step is enough;
//   * per row a thread of level l (C adjacent columns) runs the horizontal pass from a register window
//     (C + 2R floats, vector LDS) and then the vertical pass in SCATTER form: the new horizontal value is
//     accumulated into the 2R+1 output rows it contributes to, which live in registers (2R x C accumulators
//     that move down one slot per row through the FMA itself: destination slot p, addend slot p+1).  Each
//     accumulator receives its terms in ascending input-row order starting from fma(w[R], v, 0) -- exactly
//     the order of cascade_vpass -- so the sums round identically;
//   * the row that completes is written to the level's ring (next level's input) and, inside the
//     strip, straight from registers to HBM together with its DoG against the previous level's row.
//
// Shared-memory traffic per pixel and level drops from ~(2 + R) floats to ~(2 + R/4); HBM traffic is the
// same 4 + 24 (+1) and 4 + 8 bytes per octave pixel as the tile kernels.
//
// Clamp-to-edge semantics of every level: rows outside the image are read as the edge row of the ring
// (virtual row index clamped); columns outside the image are read with clamped column indices in the
// strips that touch the left / right image edge (BORDER variant of the window load).
// Design history, measurements and the variants that lost: DESIGN.md, "Streaming cascade".

struct StreamSched {
    int y0, y1;            // output rows [y0, y1) of the last level
    int r0, rlast;         // input rows the strip segment needs
    int T[4], Tend[4];     // level l is active for steps T[l] .. Tend[l]
    int i0[4];             // virtual input row of level l at step T[l] (may be negative: replicated row 0)
    int f[4], e[4];        // first / last row of level l that is needed (f[0], e[0] = r0, rlast)
    int steps;
};

template <int NL_, int R1_, int R2_, int R3_, int C_, int WS_, int PF_, int MINB_, bool PACK_, int K_ = 1>
struct StreamGeom {
    static constexpr int NL = NL_, C = C_, WS = WS_, PF = PF_, MINB = MINB_;
    static constexpr int K = K_;   // rows per step (and per CTA barrier)
    static constexpr bool PACK = PACK_;
    __host__ __device__ static constexpr int R(int l) { return l == 1 ? R1_ : l == 2 ? R2_ : l == 3 ? R3_ : 0; }
    __host__ __device__ static constexpr int RA(int l) { return ru4(R(l)); }
    // x halo kept around level l (columns computed beyond the strip so that the later levels have their windows)
    __host__ __device__ static constexpr int HO(int l) { return l >= NL ? 0 : HO(l + 1) + RA(l + 1); }
    __host__ __device__ static constexpr int W(int l) { return WS + 2 * HO(l); }
    __host__ __device__ static constexpr int NTH(int l) { return W(l) / C; }
    __host__ __device__ static constexpr int WARPS(int l) { return l < 1 || l > NL ? 0 : (NTH(l) + 31) / 32; }
    __host__ __device__ static constexpr int FIRSTWARP(int l) { return l <= 1 ? 0 : FIRSTWARP(l - 1) + WARPS(l - 1); }
    static constexpr int THREADS = 32 * (FIRSTWARP(NL) + WARPS(NL));
    // ring depths: see the schedule in k_stream
    __host__ __device__ static constexpr int DEPTH(int l) {
        return l == 0 ? K * PF + 2 * R(1) + 3 * K - 1 : 2 * R(l + 1) + 3 * K - 1;
    }
    // The input ring is filled by TMA (cp.async.bulk.tensor.2d): one box per row, or two when the row is wider than
    // the 256-element box limit; every box lands 128-byte aligned, so the ring's row stride W0S is padded.
    static constexpr int NBOX = (W(0) + 255) / 256;
    static constexpr int BOXW = (((W(0) + NBOX - 1) / NBOX) + 31) & ~31;
    static_assert(WS % C == 0 && C % 2 == 0, "strip width / columns per thread");
};

// Shared-memory layout of the rings; only the TMA form pads the input ring (and carries the mbarriers): the padding
// costs 2-6 KB per CTA, which matters for how many kernels of OTHER images fit beside this one on an SM.
template <class G, bool TMA>
struct StreamLayout {
    static constexpr int W0S = TMA ? G::NBOX * G::BOXW : G::W(0);
    __host__ __device__ static constexpr int RS(int l) { return l == 0 ? W0S : G::W(l); }   // ring row stride of level l
    __host__ __device__ static constexpr int OFF(int l) { return l <= 0 ? 0 : OFF(l - 1) + G::DEPTH(l - 1) * RS(l - 1); }
    static constexpr size_t kRingBytes = (size_t)OFF(G::NL) * sizeof(float);
    static constexpr size_t kSmem = kRingBytes + (TMA ? (size_t)G::DEPTH(0) * 8 : 0);   // + one mbarrier per input-ring slot
};

__device__ __forceinline__ void stream_bar() {
#ifndef SB_EXP_NOBAR
    asm volatile("bar.sync 0;\n" ::: "memory");
#endif
}

template <int N>
__device__ __forceinline__ void stream_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// Input rows are fetched PF rows ahead into the input ring; ring slot of row r = (r - r0) mod DEPTH(0).
// TMA form (default): ONE elected thread of level 1 (the lightest level) arms the slot's mbarrier with the row's
// byte count and issues cp.async.bulk.tensor.2d (NBOX boxes of BOXW x 1 floats) on the plane's tensor map; columns
// left / right of the image come back as zeros and are never read (BORDER strips clamp their window columns).  The
// threads of level 1 wait on the slot's mbarrier phase before they read a row.
// cp.async form (kept for comparison, SIFT_B200_STREAM_TMA=0): every level-1 lane copies 16-byte chunks, one commit
// group per step, cp.async.wait_group before the CTA barrier.
__device__ __forceinline__ void stream_mbar_init(unsigned bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void stream_mbar_expect(unsigned bar, int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void stream_mbar_wait(unsigned bar, unsigned parity) {
    for (unsigned spin = 0;; ++spin) {
        unsigned done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"   // non-blocking: the bound below is a time bound
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
        if (spin > (1u << 22)) __trap();   // a protocol bug traps within a fraction of a second instead of hanging the GPU
    }
}
__device__ __forceinline__ void stream_tma_row(unsigned dst, const CUtensorMap* map, int x, int y, unsigned bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(bar) : "memory");
}

template <class G, bool TMA>
struct StreamFetch {
    static constexpr int W0 = G::W(0), W0S = StreamLayout<G, TMA>::W0S, D0 = G::DEPTH(0), NT = 32 * G::WARPS(1), CH = W0 / 4;
    static constexpr int PER = (CH + NT - 1) / NT;
    const float* src;     // cp.async: next row to fetch, at this thread's first chunk
    unsigned dst;         // shared-memory byte address of the ring slot of the next row (cp.async: + this thread's chunk)
    unsigned dst_end;
    unsigned bar, bar0;   // TMA: mbarrier of that slot / of slot 0
    int rows_left, row, gx;
    size_t pitch;
    bool ok[PER];
    bool leader;

    __device__ __forceinline__ void init(const CascadeArgs& a, float* smem, int rt, int r0, int rlast, int sx0) {
        gx = sx0 - G::HO(0);
        rows_left = rlast - r0 + 1;
        row = r0;
        leader = rt == 0;
        const unsigned base = (unsigned)__cvta_generic_to_shared(smem);
        if (TMA) {
            dst = base;
            dst_end = base + (unsigned)(D0 * W0S * 4);
            bar0 = bar = base + (unsigned)StreamLayout<G, TMA>::kRingBytes;
        } else {
            src = a.in + (size_t)r0 * a.pitch + gx + 4 * rt;
            pitch = a.pitch;
            dst = base + 16u * rt;
            dst_end = dst + (unsigned)(D0 * W0S * 4);
#pragma unroll
            for (int k = 0; k < PER; ++k) {
                const int c = rt + k * NT, x = gx + 4 * rt + 4 * k * NT;
                ok[k] = c < CH && x >= 0 && x < a.pitch;
            }
        }
    }
    // map: the address of the kernel's __grid_constant__ tensor map, handed down as a plain argument (kept in a
    // struct member, ptxas lost track of it and fed UTMALDG the wrong uniform register)
    __device__ __forceinline__ void next(const CUtensorMap* map) {
#pragma unroll
        for (int kk = 0; kk < G::K; ++kk)
        if (rows_left > 0) {
            if (TMA) {
                if (leader) {
                    stream_mbar_expect(bar, W0S * 4);
#pragma unroll
                    for (int b = 0; b < G::NBOX; ++b)
                        stream_tma_row(dst + (unsigned)(b * G::BOXW * 4), map, gx + b * G::BOXW, row, bar);
                }
                ++row;
                bar += 8;
            } else {
#pragma unroll
                for (int k = 0; k < PER; ++k)
                    if (ok[k])
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst + 16u * k * NT),
                                     "l"(src + 4 * k * NT));
                src += pitch;
            }
            dst += W0S * 4;
            if (dst == dst_end) { dst -= D0 * W0S * 4; bar = bar0; }
            --rows_left;
        }
        if (!TMA) asm volatile("cp.async.commit_group;\n" ::);
    }
};

// One level of the pipeline, run by the warps of that level.  rt = thread index inside the level.
// Everything that moves with the row (ring slots, plane offsets) is a running counter: no division, no
// 64-bit multiply inside the step.
template <class G, int L, bool BORDER, bool STORE_G, bool TMA>
__device__ __forceinline__ void stream_level(const CascadeArgs& a, const CUtensorMap* tmap, float* smem, const int rt,
                                             const StreamSched& sc, const int sx0) {
    constexpr int R = G::R(L), RA = G::RA(L), C = G::C, C2 = C / 2;
    constexpr int WP = G::W(L - 1), DP = G::DEPTH(L - 1);
    using LY = StreamLayout<G, TMA>;
    constexpr int WPS = LY::RS(L - 1);   // row stride of the previous level's ring (padded for the TMA input ring)
    constexpr int WL = G::W(L), DL = L < G::NL ? G::DEPTH(L) : 1;
    constexpr bool LAST = L == G::NL;
    const float* ringP = smem + LY::OFF(L - 1);
    float* ringL = smem + (LAST ? 0 : LY::OFF(L));
    const BlurTaps& tp = a.taps[L - 1];
    const int w = a.w, h = a.h;

    const bool dup = rt >= G::NTH(L);                     // spare lanes of the last warp repeat the last thread
    const int x = (dup ? G::NTH(L) - 1 : rt) * C;         // first column inside the level's region
    const int gx = sx0 - G::HO(L) + x;                    // ... in the image
    const bool store_ok = !dup && x >= G::HO(L) && x < G::HO(L) + G::WS && gx < w;
    // BORDER: clamp window columns to the image (ring column of image column 0 / w-1)
    const int clo = max(0, -(sx0 - G::HO(L - 1)));
    const int chi = min(WP - 1, w - 1 - (sx0 - G::HO(L - 1)));

    float* gout = a.g[L - 1];
    float* dout = a.d[L - 1];
    float* decp = (LAST && G::NL == 3) ? a.dec : nullptr;   // only the three-level kernel (G0 -> G3) feeds the next octave

    // acc[p] = partial sum of output row (i - R + 1 + p) after input row i: 2R rows are in flight
    float2 acc[2 * R][C2];
#pragma unroll
    for (int p = 0; p < 2 * R; ++p)
#pragma unroll
        for (int c = 0; c < C2; ++c) acc[p][c] = make_float2(0.f, 0.f);
    const int T = sc.T[L], Tend = sc.Tend[L], steps = sc.steps;
    const int fL = sc.f[L], eL = sc.e[L];
    const int i0 = sc.i0[L];
    auto mod = [](int v, int m) { return ((v % m) + m) % m; };
    // running state of the step loop
    constexpr int K = G::K;
    int i = i0;                                            // virtual input row of the first of the K rows of a step
    // ring slot (floats) of the clamped input row / of the previous level's row y (DoG centre); the input ring (read
    // by level 1) counts its rows from the first row fetched, the level rings from row 0
    const int rbase = L == 1 ? sc.r0 : 0;
    int in_off = ((min(max(i0, 0), h - 1) - rbase) % DP) * WPS;
    int cen_off = mod(i0 - R - rbase, DP) * WPS;
    unsigned in_par = 0;                                   // (TMA) mbarrier phase of the slot at in_off
    const unsigned bar0 = (unsigned)__cvta_generic_to_shared(smem) + (unsigned)LY::kRingBytes;
    int out_off = mod(i0 - R, DL) * WL;                    // ... of this level's row y
    unsigned e_off = (unsigned)((long long)(i0 - R) * a.pitch + gx);   // plane offset of (y, gx); wraps while y < 0

    StreamFetch<G, TMA> fetch;
    if (L == 1) {
        fetch.init(a, smem, rt, sc.r0, sc.rlast, sx0);
#pragma unroll 1
        for (int g = 0; g < G::PF; ++g) fetch.next(tmap);
    }

    // Steady steps: the level is active, every row it reads is inside the image (no clamping) and every row it
    // completes is stored; the step body is instantiated without those tests for them.
    int ts_lo, ts_hi;
    {
        auto cdiv = [](int n, int d) { return n >= 0 ? (n + d - 1) / d : -((-n) / d); };
        auto fdiv = [](int n, int d) { return n >= 0 ? n / d : -((-n + d - 1) / d); };
        const int ylo = max(fL, sc.y0), yhi = min(eL, sc.y1 - 1);
        ts_lo = T + max(max(cdiv(-i0, K), cdiv(ylo + R - i0, K)), 0);
        ts_hi = min(T + min(fdiv(h - 2 - (K - 1) - i0, K), fdiv(yhi + R - (K - 1) - i0, K)), Tend);
    }
    auto step = [&](const int t, auto steady_tag) {
        constexpr bool STEADY = decltype(steady_tag)::value;
        if (!STEADY && (t < T || t > Tend)) return;   // pipeline fill / drain: this level has nothing to do
        // ---- horizontal pass of K rows: C outputs each from a register window ----
        float2 hv[K][C2];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const float* rowp = ringP + in_off;
            if (TMA && L == 1) stream_mbar_wait(bar0 + 8u * (unsigned)(in_off / WPS), in_par);   // the row has landed
            float v[C + 2 * RA];
            if (!BORDER) {
                if (C % 4 == 0) {
                    const float4* src = reinterpret_cast<const float4*>(rowp + x);
#pragma unroll
                    for (int q4 = 0; q4 < (C + 2 * RA) / 4; ++q4) {
                        if (4 * q4 + 3 < RA - R || 4 * q4 >= RA + C + R) continue;   // outside the taps
                        const float4 q = src[q4];
                        v[4 * q4] = q.x; v[4 * q4 + 1] = q.y; v[4 * q4 + 2] = q.z; v[4 * q4 + 3] = q.w;
                    }
                } else {   // two columns per thread: x is only 8-byte aligned
                    const float2* src = reinterpret_cast<const float2*>(rowp + x);
#pragma unroll
                    for (int q2 = 0; q2 < (C + 2 * RA) / 2; ++q2) {
                        if (2 * q2 + 1 < RA - R || 2 * q2 >= RA + C + R) continue;
                        const float2 q = src[q2];
                        v[2 * q2] = q.x; v[2 * q2 + 1] = q.y;
                    }
                }
            } else {
#pragma unroll
                for (int j = RA - R; j < RA + C + R; ++j) v[j] = rowp[min(max(x + j, clo), chi)];
            }
#pragma unroll
            for (int c = 0; c < C; ++c) {
                float sacc = 0.f;
#pragma unroll
                for (int u = R; u >= 1; --u) sacc = fmaf(tp.w[u], v[RA + c - u] + v[RA + c + u], sacc);
                const float o = fmaf(tp.w[0], v[RA + c], sacc);
                if (c & 1) hv[k][c / 2].y = o; else hv[k][c / 2].x = o;
            }
            if (STEADY || (unsigned)(i + k) < (unsigned)(h - 1)) {   // rows above / below the image re-read the edge row
                in_off += WPS;
                if (in_off == DP * WPS) { in_off = 0; in_par ^= 1u; }
            }
        }
        // ---- vertical pass, scatter form: row i adds w[|i - y|] * h to every output row y in [i-R, i+R].  The
        // accumulators move down one slot per row THROUGH the FMA (d = a * b + c with c = slot p+1, d = slot p):
        // the rotation costs nothing, every index is static, and each row still receives its terms in ascending
        // input-row order starting from fma(w[R], h, 0) ----
        float2 out[K][C2];
#pragma unroll
        for (int k = 0; k < K; ++k) {
#pragma unroll
            for (int c = 0; c < C2; ++c) {
                if (G::PACK) {
                    out[k][c] = __ffma2_rn(make_float2(tp.w[R], tp.w[R]), hv[k][c], acc[0][c]);
                } else {
                    out[k][c].x = fmaf(tp.w[R], hv[k][c].x, acc[0][c].x);
                    out[k][c].y = fmaf(tp.w[R], hv[k][c].y, acc[0][c].y);
                }
            }
#pragma unroll
            for (int p = 0; p < 2 * R; ++p) {
                const int d = R - 1 - p < 0 ? p + 1 - R : R - 1 - p;
#pragma unroll
                for (int c = 0; c < C2; ++c) {
                    const float2 prev = p + 1 < 2 * R ? acc[p + 1][c] : make_float2(0.f, 0.f);
                    if (G::PACK) {
                        acc[p][c] = __ffma2_rn(make_float2(tp.w[d], tp.w[d]), hv[k][c], prev);
                    } else {
                        acc[p][c].x = fmaf(tp.w[d], hv[k][c].x, prev.x);
                        acc[p][c].y = fmaf(tp.w[d], hv[k][c].y, prev.y);
                    }
                }
            }
        }
        // ---- the rows that just received their last term ----
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int y = i + k - R;
            if (STEADY || (y >= fL && y <= eL)) {
                if (!LAST) {
                    float2* dst = reinterpret_cast<float2*>(ringL + out_off + x);
#pragma unroll
                    for (int c = 0; c < C2; ++c) dst[c] = out[k][c];   // (pairs of STS.64 merge into STS.128)
                }
                if (store_ok && (STEADY || (y >= sc.y0 && y < sc.y1))) {
                    const float2* cen = reinterpret_cast<const float2*>(ringP + cen_off + x + RA);
                    if (C % 4 == 0) {
#pragma unroll
                        for (int q4 = 0; q4 < C / 4; ++q4) {
                            if (gx + 4 * q4 >= w) break;
                            const float4 ov = make_float4(out[k][2 * q4].x, out[k][2 * q4].y, out[k][2 * q4 + 1].x,
                                                          out[k][2 * q4 + 1].y);
#ifdef SB_EXP_NOSTG
                            if (ov.x == 123.456f)
#endif
                            if (STORE_G) *reinterpret_cast<float4*>(gout + e_off + 4 * q4) = ov;
                            {
                                const float4 cv = *reinterpret_cast<const float4*>(cen + 2 * q4);
#ifdef SB_EXP_NOSTG
                                if (cv.x == 123.456f)
#endif
                                *reinterpret_cast<float4*>(dout + e_off + 4 * q4) =
                                    make_float4(ov.x - cv.x, ov.y - cv.y, ov.z - cv.z, ov.w - cv.w);
                            }
                            if (LAST && decp != nullptr && !(y & 1)) {
                                const int dy = y >> 1, dx = (gx + 4 * q4) >> 1;
                                if (dy < a.dec_h) {
                                    if (dx + 1 < a.dec_w)
                                        *reinterpret_cast<float2*>(decp + (size_t)dy * a.dec_pitch + dx) =
                                            make_float2(ov.x, ov.z);
                                    else if (dx < a.dec_w)
                                        decp[(size_t)dy * a.dec_pitch + dx] = ov.x;
                                }
                            }
                        }
                    } else {
#pragma unroll
                        for (int c = 0; c < C2; ++c) {
                            if (gx + 2 * c >= w) break;
                            const float2 ov = out[k][c];
                            if (STORE_G) *reinterpret_cast<float2*>(gout + e_off + 2 * c) = ov;
                            {
                                const float2 cv = cen[c];
                                *reinterpret_cast<float2*>(dout + e_off + 2 * c) = make_float2(ov.x - cv.x, ov.y - cv.y);
                            }
                            if (LAST && decp != nullptr && !(y & 1)) {
                                const int dy = y >> 1, dx = (gx + 2 * c) >> 1;
                                if (dy < a.dec_h && dx < a.dec_w) decp[(size_t)dy * a.dec_pitch + dx] = ov.x;
                            }
                        }
                    }
                }
            }
            // ---- advance the per-row running state ----
            cen_off += WPS;
            if (cen_off == DP * WPS) cen_off = 0;
            if (!LAST) {
                out_off += WL;
                if (out_off == DL * WL) out_off = 0;
            }
            e_off += (unsigned)a.pitch;
        }
        i += K;
    };
    auto sync_step = [&]() {
        if (L == 1) {
            fetch.next(tmap);
            if (!TMA) stream_wait<G::PF>();
        }
        stream_bar();
    };
    int t = 0;
#pragma unroll 1
    for (; t < min(ts_lo, steps); ++t) {
        sync_step();
        step(t, std::false_type{});
    }
#pragma unroll 1
    for (; t <= min(ts_hi, steps - 1); ++t) {
        sync_step();
        step(t, std::true_type{});
    }
#pragma unroll 1
    for (; t < steps; ++t) {
        sync_step();
        step(t, std::false_type{});
    }
}

template <class G, bool BORDER, bool STORE_G, bool TMA>
__device__ __forceinline__ void stream_body(const CascadeArgs& a, const CUtensorMap* tmap, float* smem,
                                            const StreamSched& sc, int sx0) {
    const int warp = threadIdx.x >> 5;
    if (G::NL >= 3 && warp >= G::FIRSTWARP(3))
        stream_level<G, (G::NL >= 3 ? 3 : 1), BORDER, STORE_G, TMA>(a, tmap, smem, threadIdx.x - 32 * G::FIRSTWARP(3), sc, sx0);
    else if (warp >= G::FIRSTWARP(2))
        stream_level<G, 2, BORDER, STORE_G, TMA>(a, tmap, smem, threadIdx.x - 32 * G::FIRSTWARP(2), sc, sx0);
    else
        stream_level<G, 1, BORDER, STORE_G, TMA>(a, tmap, smem, threadIdx.x, sc, sx0);
}

// Schedule (per CTA, uniform), K rows per step.  Level l consumes the virtual input rows
// i0[l] + K (t - T[l]) + k, k < K, at step t and completes the rows R_l above them in the same step.  Level l
// starts the step after level l-1 completed the last row of level l's first step, so from then on the rows it
// needs were always completed at least one step earlier.  At the top of the image i0[l] = -R_l: the level
// re-reads ring row 0 while its producer runs ahead, hence ring depth 2 R_l + 3K - 1 (window rows .. centre row
// of the DoG .. rows being written); the input ring adds the prefetch distance (K PF rows).
// Work split, two forms.  nseg == 0 (any image size): the (strip, row) space is flattened strip-major and cut into
// gridDim.x equal-cost ranges (a row of a strip on the left / right image edge costs 3 units, an interior one 2:
// clamped window loads), so every CTA gets the same amount of work; a range that crosses a strip boundary is
// run as two (or more) passes of the pipeline.  nseg > 0 (when strips x nseg fills the GPU): every strip is cut
// into the same nseg row bands (2 nseg half-height ones for an edge strip), one CTA each: the CTAs of a band walk
// down the same rows at the same time, so a plane row is written as one run across the strips (DRAM pages,
// and the x-halo reads of the neighbours hit L2).
// STORE_G: the level planes are stored (always for G1..G3; G4, G5 only for the debug planes); the DoG planes
// always are.
template <class G, bool STORE_G, bool TMA>
__global__ void __launch_bounds__(G::THREADS, G::MINB)
k_stream(const CascadeArgs a, const __grid_constant__ CUtensorMap tmap, const int nseg) {
    extern __shared__ __align__(128) float smem[];
    const int strips = (a.w + G::WS - 1) / G::WS;
    auto is_border = [&](int s) { return s * G::WS - G::HO(0) < 0 || s * G::WS + G::WS + G::HO(0) > a.w; };
    long long lo, hi;
    if (nseg == 0) {
        long long total = 0;
        for (int s = 0; s < strips; ++s) total += (long long)a.h * (is_border(s) ? 3 : 2);
        lo = total * blockIdx.x / gridDim.x;
        hi = total * (blockIdx.x + 1) / gridDim.x;
    } else {
        int k = blockIdx.x, s = 0;
        long long off = 0;
        for (; s < strips; ++s) {
            const int n = is_border(s) ? 2 * nseg : nseg;
            if (k < n) break;
            k -= n;
            off += (long long)a.h * (is_border(s) ? 3 : 2);
        }
        if (s == strips) return;
        const int c = is_border(s) ? 3 : 2, n = is_border(s) ? 2 * nseg : nseg;
        const int hs = (a.h + n - 1) / n;
        const int y0 = k * hs, y1 = min(y0 + hs, a.h);
        if (y0 >= y1) return;
        lo = off + (long long)y0 * c;
        hi = off + (long long)y1 * c;
    }
    long long beg = 0;
    bool first = true;
    for (int s = 0; s < strips; ++s) {
        const bool border = is_border(s);
        const int c = border ? 3 : 2;
        const long long end = beg + (long long)a.h * c;
        const long long ra = lo > beg ? lo : beg, rb = hi < end ? hi : end;
        const long long cur = beg;
        beg = end;
        if (ra >= rb) continue;
        StreamSched sc;
        sc.y0 = (int)((ra - cur + c - 1) / c);
        sc.y1 = (int)((rb - cur + c - 1) / c);
        if (sc.y0 >= sc.y1) continue;
        if (!first) __syncthreads();   // the rings are reused: the previous pass must be done reading them
        first = false;
        if (TMA) {   // one mbarrier per input-ring slot, (re)initialised per pass: every slot starts at phase 0
            if (threadIdx.x == 0) {
                const unsigned bar0 = (unsigned)__cvta_generic_to_shared(smem) + (unsigned)StreamLayout<G, TMA>::kRingBytes;
                for (int q = 0; q < G::DEPTH(0); ++q) stream_mbar_init(bar0 + 8u * q, 1);
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            }
            __syncthreads();
        }
        sc.f[G::NL] = sc.y0;
        sc.e[G::NL] = sc.y1 - 1;
#pragma unroll
        for (int l = G::NL; l >= 1; --l) {
            sc.i0[l] = sc.f[l] - G::R(l);
            sc.f[l - 1] = max(sc.i0[l], 0);
            sc.e[l - 1] = min(sc.e[l] + G::R(l), a.h - 1);
        }
        sc.r0 = sc.f[0];
        sc.rlast = sc.e[0];
        sc.T[1] = 0;
#pragma unroll
        for (int l = 1; l <= G::NL; ++l) {
            sc.Tend[l] = sc.T[l] + (sc.e[l] + G::R(l) - sc.i0[l]) / G::K;
            if (l < G::NL) sc.T[l + 1] = sc.T[l] + 1 + (sc.f[l] + G::K - 1 + G::R(l) - sc.i0[l]) / G::K;
        }
        sc.steps = sc.Tend[G::NL] + 1;
        // level 1 consumes K rows per step: with K > 1 its last step may touch one row beyond the last row it needs.
        // The cp.async form reads whatever the slot holds (the value is never used); the TMA form WAITS for the row
        // to land, so the row has to be requested as well.
        sc.rlast = min(max(sc.rlast, sc.i0[1] + (sc.Tend[1] - sc.T[1] + 1) * G::K - 1), a.h - 1);
        if (border)
            stream_body<G, true, STORE_G, TMA>(a, &tmap, smem, sc, s * G::WS);
        else
            stream_body<G, false, STORE_G, TMA>(a, &tmap, smem, sc, s * G::WS);
    }
}

// default scale space (radii 4,5,6 | 8,10)
// A: one warp per level (32 / 28 / 24 threads of 4 columns), one row per step, 5 CTAs per SM (6 would cap the
// registers at 96 and spill); B: two warps per level (64 / 58 threads), two rows per step (half the barriers,
// two independent rows in the horizontal pass), 3 CTAs per SM.  Measured alternatives in DESIGN.md.
using StreamA = StreamGeom<3, 4, 5, 6, 4, 96, 12, 5, true, 1>;    // G0 -> G1,G2,G3, D0,D1,D2, next base
// wider strips with two warps per level -- less x halo (level 1 computes 1.14x instead of 1.33x the strip), fuller
// last warps -- at fewer resident warps: used on octaves of >= 16 Mpx.  Measured and not instantiated any more (4K
// batch images/s, the geometry forced on every streaming octave): <.., 224, 12, 3 CTAs/SM> 743 (same as A1),
// <.., 160, 12, 3> 711 (half-empty second warps) against 731 for A everywhere and 743 for A1 everywhere; a narrow
// second kernel <2, 8, 10, 0, 4, 104, 6, 6, true, 2> on octaves below 3 / 10 Mpx: 797 / 795 against 802.
using StreamA1 = StreamGeom<3, 4, 5, 6, 4, 224, 12, 2, true, 1>;
using StreamB = StreamGeom<2, 8, 10, 0, 4, 232, 6, 3, true, 2>;   // G3 -> (G4,G5) -> D3,D4

// 2-D tensor map of one FP32 plane (w x h, `pitch` floats per row) with a BOXW x 1 box, no swizzle, zero fill
cudaError_t stream_make_map(CUtensorMap* map, const float* base, int w, int h, int pitch, int boxw);
bool stream_use_tma();

template <class G>
cudaError_t launch_stream_t(const CascadeArgs& a, int sm_count, cudaStream_t s, int force = 0) {
    const int strips = (a.w + G::WS - 1) / G::WS;
    const int slots = sm_count * G::MINB;   // one wave of CTAs
    int nb = 0;                             // strips on an image edge
    for (int t = 0; t < strips; ++t) nb += t * G::WS - G::HO(0) < 0 || t * G::WS + G::WS + G::HO(0) > a.w;
    // aligned row bands when they fill >= 90 % of the wave with bands of >= 96 rows, else the flattened split
    // (fewer CTAs when the image is too small to give each of them min_rows rows)
    const int min_rows = 48;
    int nseg = slots / (strips + nb), ctas;
    if (nseg < 1 || (strips + nb) * nseg * 10 < slots * 9 || a.h / nseg < 96) nseg = 0;
    if (force < 0) nseg = -force;   // experiments: 1 = flattened, > 1 = flattened with that many CTAs, < 0 = -force bands
    if (force > 0) nseg = 0;
    if (nseg > 0) {
        ctas = (strips + nb) * nseg;
    } else {
        const long long units = ((long long)strips * a.h + min_rows - 1) / min_rows;
        ctas = units < slots ? (int)units : slots;
        if (force > 1) ctas = force;
    }
    for (int l = 0; l < G::NL; ++l)
        if (a.d[l] == nullptr || (a.g[l] == nullptr) != (a.g[0] == nullptr)) return cudaErrorInvalidValue;
    CUtensorMap map;
    memset(&map, 0, sizeof map);
    const bool tma = stream_use_tma();
    if (tma) {
        const cudaError_t e = stream_make_map(&map, a.in, a.w, a.h, a.pitch, G::BOXW);
        if (e != cudaSuccess) return e;
    }
    if (a.g[0] != nullptr) {
        if (tma) k_stream<G, true, true><<<ctas, G::THREADS, StreamLayout<G, true>::kSmem, s>>>(a, map, nseg);
        else k_stream<G, true, false><<<ctas, G::THREADS, StreamLayout<G, false>::kSmem, s>>>(a, map, nseg);
    } else {
        if (tma) k_stream<G, false, true><<<ctas, G::THREADS, StreamLayout<G, true>::kSmem, s>>>(a, map, nseg);
        else k_stream<G, false, false><<<ctas, G::THREADS, StreamLayout<G, false>::kSmem, s>>>(a, map, nseg);
    }
    return cudaGetLastError();
}

// C ABI of the B200 SIFT engine (include/sift_b200.h): context, workspace arena, stage
// orchestration on one CUDA stream.  Everything between the input copy and the result copy is
// enqueued without host synchronisation; list sizes live in device counters.
//
// Stage order follows detect_keypoints_and_descriptors (sift.cpp:712-776):
//   compute_initial_image :113-126 -> compute_gaussian_images :181-202 (+ DoG :209-225 fused)
//   -> detect_extrema :300-319 -> compute_keypoints :330-436 -> compute_orientations :447-533
//   -> clean_keypoints :20-24 -> compute_descriptors :610-682.
#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <string>
#include <vector>

#include "../../include/sift_b200.h"
#include "common.cuh"
#include "kernels.h"

using namespace sb;

static thread_local std::string g_create_error;

struct sift_b200_collection;

struct sift_b200_ctx {
    sift_b200_collection* coll = nullptr;   // communicator + collection-matching state (api_collection.inc)
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    int max_w = 0, max_h = 0;
    std::string err;
    long launches = 0;

    // ---- detect workspace ----
    float* arena = nullptr;       // all pyramid planes
    size_t arena_floats = 0;
    uint8_t* d_input = nullptr;   // staged input pixels (u8 or f32, up to 3 channels)
    size_t input_bytes = 0;
    PyramidDesc pyr;              // host copy (device pointers inside)
    PyramidDesc* d_pyr = nullptr;
    Counters* d_counters = nullptr;
    Counters* h_counters = nullptr;  // pinned
    float* d_range = nullptr;        // (min, max) of a float input
    Cand* d_cands = nullptr;
    CandCube* d_cubes = nullptr;   // the fit's first neighbourhood of candidate slot < cube_cap, written by the scan
    int cube_cap = 0;
    bool use_cubes = true;         // SIFT_B200_CUBES=0: the refinement loads every neighbourhood itself
    KpCore* d_raw = nullptr;
    KpCore* d_oriented = nullptr;
    uint8_t* d_records = nullptr;  // cap_final x 168
    uint8_t* d_desc = nullptr;     // cap_final x 128
    int cap_extrema = 0, cap_raw = 0, cap_oriented = 0;
    SortScratch ss{};
    int ss_nb_cap = 0;
    StageParams sp{};
    bool detect_pending = false;
    bool have_result = false;
    int base_w = 0, base_h = 0;
    int layers = kLayers, dogs = kDogs;   // of the last detect (intervals + 3, intervals + 2)
    sift_b200_stats stats{};
    bool keep_planes = false;    // also store G[4], G[5] (debug plane access)
    bool force_unfused = false;  // per-level kernels instead of the fused octave cascade
    int fused_mode = 0;          // launch_octave_fused mode: 0 auto, 2 tile kernels only, 3 streaming kernels only
    bool centred = true;         // scale space stored relative to the input's mid level (finer FP32 steps, same results)
    float centre_u8 = 128.f;
    bool last_float_input = false;
    bool pyramid_valid = false;  // the planes of the last detect are intact (debug stage calls need them)
    // ---- launch plan: everything after the input stage is replayed from a CUDA graph; inside it the octave
    // chain (first cascade kernel of every octave) runs on the main stream and the second cascade kernel + the
    // extrema scan of each octave on a side stream, so the small octaves overlap the large ones ----
    bool use_graph = true;
    bool three_branches = false; // experiments (SIFT_B200_GRAPH=3): extrema scans on a third graph branch
    bool use_tail = true;        // the latency-bound small octaves in one launch (k_tail); SIFT_B200_TAIL=0: one by one
    long long fork_min_px = 300000;   // octaves of at least this many pixels run their second half on the side branch
    int extrema_form = 0;        // 0 = four columns per lane (default), 1 = one column per lane
    cudaStream_t side = nullptr, side2 = nullptr;
    std::vector<cudaEvent_t> fork_ev;   // [0 .. kMaxOctaves): "octave chain reached o"; [kMaxOctaves]: join of the
                                        // cascade branch; then [kMaxOctaves + 1 + o]: "D3, D4 of octave o written";
                                        // last: join of the extrema branch
    cudaGraphExec_t graph_exec = nullptr;
    std::vector<uint8_t> graph_key;
    int graph_launches = 0;
    int graph_stage_launches[SIFT_B200_STAGE_COUNT] = {0};
    long graphs_built = 0;
    PyramidDesc pyr_uploaded;    // what d_pyr holds
    bool pyr_uploaded_valid = false;
    bool profiling = false;
    std::vector<cudaEvent_t> ev_pool;
    std::vector<int> ev_stage;   // stage of the interval ending at event i (event 0: -1)
    int ev_used = 0;
    int stage_launches[SIFT_B200_STAGE_COUNT] = {0};

    // ---- match workspace (grown on demand) ----
    MatchScratch ms{};
    size_t ms_rows_a = 0, ms_rows_b = 0;
    uint8_t* d_ma = nullptr; size_t ma_cap = 0;   // staged host descriptors
    uint8_t* d_mb = nullptr; size_t mb_cap = 0;
    int* d_best_idx = nullptr; int* d_best_d2 = nullptr; int* d_second_d2 = nullptr; size_t best_cap = 0;
    int* d_out_ia = nullptr; int* d_out_ib = nullptr; double* d_out_dist = nullptr; int* d_out_count = nullptr;
    size_t out_cap = 0;
};

namespace {

void collection_free(sift_b200_ctx* c);   // api_collection.inc

int fail(sift_b200_ctx* c, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf; else g_create_error = buf;
    return code;
}

#define CU(c, call)                                                                           \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess)                                                               \
            return fail((c), SIFT_B200_E_CUDA, "%s failed: %s (%s:%d)", #call,                \
                        cudaGetErrorString(e__), __FILE__, __LINE__);                         \
    } while (0)

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

// compute_octaves_count, sift.cpp:132-137: floor(log2(min(w,h) / 3)) with integer division.
int octave_count(int w, int h) {
    const int m = std::min(w, h) / 3;
    if (m < 1) return 0;
    return (int)floor(log2((double)m));
}

// apply_gaussian_blur_fast, image.cpp:226-235: ceil(3 sigma)+1 half-kernel taps; the per-pixel
// division by the accumulated weight total (image.cpp:185) is folded into the taps.
BlurTaps make_taps(double sigma) {
    BlurTaps t;
    memset(&t, 0, sizeof t);
    const int n = (int)ceil(3 * sigma) + 1;
    double k[64];
    const double denom = 2 * sigma * sigma, coef = 1 / (sqrt(2 * kPi) * sigma);
    double total = 0.0;
    for (int i = 0; i < n && i < 64; ++i) {
        k[i] = exp(-i * i / denom) * coef;
        total += (i == 0) ? k[i] : 2.0 * k[i];
    }
    t.radius = n - 1;
    for (int i = 0; i < n && i <= kMaxRadius; ++i) t.w[i] = (float)(k[i] / total);
    return t;
}

bool is_device_ptr(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

size_t plane_floats(int w, int h) { return (size_t)round_up(w, 32) * h; }

// Carve the arena into planes for a base image of bw x bh with `octaves` octaves.
void layout_pyramid(sift_b200_ctx* c, int bw, int bh, int octaves, int layers, int dogs) {
    float* p = c->arena;
    c->pyr.octaves = octaves;
    int w = bw, h = bh;
    for (int o = 0; o < octaves; ++o) {
        OctaveDesc& od = c->pyr.oct[o];
        od.w = w; od.h = h; od.pitch = round_up(w, 32);
        const size_t n = plane_floats(w, h);
        for (int i = 0; i < kMaxLayers; ++i) { od.G[i] = i < layers ? p : nullptr; if (i < layers) p += n; }
        for (int i = 0; i < kMaxDogs; ++i) { od.D[i] = i < dogs ? p : nullptr; if (i < dogs) p += n; }
        w /= 2; h /= 2;
    }
}

size_t arena_need(int bw, int bh, int planes = kLayers + kDogs) {
    size_t total = 0;
    int w = bw, h = bh;
    for (int o = 0; o < kMaxOctaves && w >= 1 && h >= 1; ++o) {
        total += (size_t)planes * plane_floats(w, h);
        w /= 2; h /= 2;
    }
    return total + 64;
}

int check_params(sift_b200_ctx* c, const sift_b200_params& p) {
    const int bins = (int)p.num_bins;  // declared double, used as int (sift.cpp:450)
    if (p.intervals < 2 || p.intervals > kMaxIntervals || p.window_size < 3 || p.window_size > 7 ||
        !(p.window_size & 1) || 2 * (p.window_size / 2) >= p.intervals + 2 || bins < 4 || bins > 128)
        return fail(c, SIFT_B200_E_UNSUPPORTED,
                    "this build implements intervals in 2..%d, odd window_size in 3..7 that leaves a layer to "
                    "test, num_bins in 4..128 (got %d, %d, %g)",
                    kMaxIntervals, p.intervals, p.window_size, p.num_bins);
    if (!(p.init_sigma > 1.0) || p.init_sigma > 3.0)
        return fail(c, SIFT_B200_E_UNSUPPORTED, "init_sigma must be in (1, 3] (got %g)", p.init_sigma);
    if (p.max_octaves < 0) return fail(c, SIFT_B200_E_INVALID, "max_octaves < 0");
    return SIFT_B200_OK;
}

int grow_i32(sift_b200_ctx* c, int** p, size_t n) {
    if (*p) {
        CU(c, cudaStreamSynchronize(c->stream));   // earlier matcher launches may still read the old buffer
        cudaFree(*p);
    }
    *p = nullptr;
    CU(c, cudaMalloc(p, n * sizeof(int)));
    return SIFT_B200_OK;
}

int ensure_match_scratch(sift_b200_ctx* c, int na, int nb) {
    const int max_splits = 64;
    if ((size_t)na > c->ms_rows_a) {
        const size_t rows = std::max<size_t>((size_t)na, 1024) * 5 / 4;
        int rc;
        if ((rc = grow_i32(c, &c->ms.part_idx, rows * max_splits))) return rc;
        if ((rc = grow_i32(c, &c->ms.part_d1, rows * max_splits))) return rc;
        if ((rc = grow_i32(c, &c->ms.part_d2, rows * max_splits))) return rc;
        if ((rc = grow_i32(c, &c->ms.norms_a, rows))) return rc;
        c->ms_rows_a = rows;
        c->ms.cap_rows = rows;
        c->ms.max_splits = max_splits;
    }
    if ((size_t)nb > c->ms_rows_b) {
        const size_t rows = std::max<size_t>((size_t)nb, 1024) * 5 / 4;
        int rc;  // + room for the tile padding and the per-tile minima of the tensor-core path
        if ((rc = grow_i32(c, &c->ms.norms_b, rows + 1024 + rows / 64))) return rc;
        c->ms_rows_b = rows;
    }
    return SIFT_B200_OK;
}

int enqueue_match(sift_b200_ctx* c, const uint8_t* d_a, int na, const uint8_t* d_b, int nb, int* d_idx,
                  int* d_d1, int* d_d2) {
    int rc = ensure_match_scratch(c, na, nb);
    if (rc) return rc;
    int launches = 0;
    CU(c, launch_match(d_a, na, d_b, nb, d_idx, d_d1, d_d2, c->ms, c->sm_count, c->stream, &launches));
    c->launches += launches;
    return SIFT_B200_OK;
}

// Profiling: mark(stage) closes an interval [previous event, now) attributed to `stage`.
void prof_mark(sift_b200_ctx* c, int stage, int launches = 0) {
    if (stage >= 0) c->stage_launches[stage] += launches;
    c->launches += launches;
    if (!c->profiling) return;
    if (c->ev_used == (int)c->ev_pool.size()) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        c->ev_pool.push_back(e);
        c->ev_stage.push_back(-1);
    }
    cudaEventRecord(c->ev_pool[c->ev_used], c->stream);
    c->ev_stage[c->ev_used] = stage;
    c->ev_used++;
}

// Everything of one detect call that follows the input stage: cascade + DoG + extrema per octave, refine,
// orientation, sort / dedup, descriptors, counters back to the host.  `forked`: the first cascade kernel of every
// octave (which also writes the next octave's base) stays on the main stream, the second cascade kernel and the
// extrema scan of each octave go to the side stream, joined before the refinement -- the octave chain is the only
// true dependency (sift.cpp:187-199), so the many small launches of the low octaves overlap the large ones.
// Enqueued either directly (profiling, debugging) or once under stream capture and replayed as a CUDA graph.
struct DetectPlan {
    int octaves, layers, dogs;
    bool fused;
    BlurTaps taps[kMaxLayers];
};
constexpr size_t kCountersBytes = sizeof(Counters) + (2 + kMaxOctaves) * sizeof(int);

int enqueue_body(sift_b200_ctx* c, const DetectPlan& pl, bool forked, int* n_launches, int* stage_launches) {
    const StageParams& sp = c->sp;
    // Branches of the graph: the octave chain (first cascade kernel of every octave) on the main stream; the second
    // cascade kernel + extrema scan of the LARGE octaves on the side stream; those of the small octaves follow their
    // octave on the main stream, which would otherwise idle once the chain is through -- the two branches then
    // carry about the same time (measured at 4K, latency by threshold: every octave on the side 1.596 ms, >= 4 Mpx 1.615, >= 1 Mpx 1.566, >= 0.3 Mpx 1.547).  (A third branch for the extrema scans was measured and
    // dropped: 4K latency 1.64 ms against 1.60 ms, batch throughput -4 %: the scans then compete with the next
    // octave's cascade kernels for the SMs.)
    const bool three = forked && c->three_branches;
    cudaStream_t s = c->stream;
    int total = 0;
    bool side_forked = false, s3_forked = false;
    auto mark = [&](int stage, int launches) {
        total += launches;
        if (stage_launches) stage_launches[stage] += launches;
        if (!forked) prof_mark(c, stage, 0);
    };
    // the tail: the run of last octaves that are at most one 32 x 32 tile per SM each -- one launch for their
    // cascades, one for their extrema scans (default plan only: the kernel-form switches keep one launch per octave,
    // which is also what the tests compare the tail against)
    const int cube_cap = (c->use_cubes && extrema_hands_cubes(sp.border, c->extrema_form)) ? c->cube_cap : 0;
    int tail_first = pl.octaves;
    if (pl.fused && c->use_tail && c->fused_mode == 0) {
        while (tail_first > 0 && tail_eligible(c->pyr.oct[tail_first - 1], c->sm_count)) --tail_first;
        if (pl.octaves - tail_first < 2 || pl.octaves - tail_first > 12) tail_first = pl.octaves;
    }
    for (int o = 0; o < pl.octaves; ++o) {
        OctaveDesc& od = c->pyr.oct[o];
        if (o == tail_first) {
            const int n = pl.octaves - o;
            CU(c, launch_tail(c->pyr.oct, o, pl.octaves, pl.taps, c->keep_planes, (int*)(c->d_counters + 1), c->sm_count, s));
            mark(SIFT_B200_STAGE_PYRAMID, 1);
            if (extrema_multi_supported(sp.border, c->extrema_form)) {
                CU(c, launch_extrema_multi(c->pyr.oct, o, n, pl.dogs, sp.dog_threshold, c->d_cands, c->cap_extrema, c->d_counters, c->d_cubes, cube_cap, s));
                mark(SIFT_B200_STAGE_EXTREMA, 1);
            } else {
                for (int q = o; q < pl.octaves; ++q) {
                    const OctaveDesc& oq = c->pyr.oct[q];
                    if (oq.w < 2 * sp.border + 1 || oq.h < 2 * sp.border + 1) continue;
                    CU(c, launch_extrema(oq, q, pl.dogs, sp.border, sp.dog_threshold, c->d_cands, c->cap_extrema, c->d_counters, c->d_cubes, cube_cap, c->extrema_form, s));
                    mark(SIFT_B200_STAGE_EXTREMA, 1);
                }
            }
            break;
        }
        const bool on_side = forked && (long long)od.w * od.h >= c->fork_min_px;
        cudaStream_t s2 = on_side ? c->side : s, s3 = (three && on_side) ? c->side2 : s2;
        float* dec = nullptr;
        int dw = 0, dh = 0, dp = 0;
        if (o + 1 < pl.octaves) {  // next base = G[layers-3] decimated, sift.cpp:195-196
            OctaveDesc& nx = c->pyr.oct[o + 1];
            dec = nx.G[0]; dw = nx.w; dh = nx.h; dp = nx.pitch;
        }
        if (pl.fused) {
            CU(c, launch_octave_fused(od, pl.taps, dec, dw, dh, dp, c->keep_planes, c->sm_count, c->fused_mode, 1, s));
            if (on_side) {
                CU(c, cudaEventRecord(c->fork_ev[o], s));
                CU(c, cudaStreamWaitEvent(s2, c->fork_ev[o], 0));
                side_forked = true;
            }
            if (o == 0) mark(SIFT_B200_STAGE_PYRAMID, 1);   // the two largest launches are timed one by one
            CU(c, launch_octave_fused(od, pl.taps, dec, dw, dh, dp, c->keep_planes, c->sm_count, c->fused_mode, 2, s2));
            mark(SIFT_B200_STAGE_PYRAMID, o == 0 ? 1 : 2);
        } else {
            for (int i = 1; i < pl.layers; ++i) {
                const bool last = i == pl.layers - 3;
                CU(c, launch_blur(od.G[i - 1], od.G[i], od.D[i - 1], last ? dec : nullptr, od.w, od.h, od.pitch,
                                  last ? dw : 0, last ? dh : 0, last ? dp : 0, pl.taps[i], s));
            }
            if (on_side) {
                CU(c, cudaEventRecord(c->fork_ev[o], s));
                CU(c, cudaStreamWaitEvent(s2, c->fork_ev[o], 0));
                side_forked = true;
            }
            mark(SIFT_B200_STAGE_PYRAMID, pl.layers - 1);
        }
        if (od.w >= 2 * sp.border + 1 && od.h >= 2 * sp.border + 1) {
            if (three && on_side) {   // experiments: the extrema scans of the large octaves as a third branch
                CU(c, cudaEventRecord(c->fork_ev[kMaxOctaves + 1 + o], s2));
                CU(c, cudaStreamWaitEvent(s3, c->fork_ev[kMaxOctaves + 1 + o], 0));
                s3_forked = true;
            }
            CU(c, launch_extrema(od, o, pl.dogs, sp.border, sp.dog_threshold, c->d_cands, c->cap_extrema, c->d_counters, c->d_cubes, cube_cap, c->extrema_form, s3));
            mark(SIFT_B200_STAGE_EXTREMA, 1);
        }
    }
    if (side_forked) {
        CU(c, cudaEventRecord(c->fork_ev[kMaxOctaves], c->side));
        CU(c, cudaStreamWaitEvent(s, c->fork_ev[kMaxOctaves], 0));
    }
    if (s3_forked) {
        CU(c, cudaEventRecord(c->fork_ev[2 * kMaxOctaves + 1], c->side2));
        CU(c, cudaStreamWaitEvent(s, c->fork_ev[2 * kMaxOctaves + 1], 0));
    }
    CU(c, launch_refine(c->d_pyr, c->d_cands, c->d_raw, c->d_counters, c->d_cubes, cube_cap, sp, c->sm_count, s));
    mark(SIFT_B200_STAGE_REFINE, 1);
    CU(c, launch_orient(c->d_pyr, c->d_raw, c->d_oriented, c->d_counters, sp, c->sm_count, s));
    mark(SIFT_B200_STAGE_ORIENT, 1);
    int l = 0;
    CU(c, launch_sort_dedup(c->d_oriented, c->d_counters, c->ss, sp, c->sm_count, s, &l));
    mark(SIFT_B200_STAGE_SORT, l);
    CU(c, launch_describe(c->d_pyr, c->d_oriented, c->ss.final_order, c->d_counters, c->d_records, c->d_desc,
                          c->cap_oriented, sp, c->sm_count, s));
    mark(SIFT_B200_STAGE_DESCRIBE, 1);
    CU(c, cudaMemcpyAsync(c->h_counters, c->d_counters, sizeof(Counters), cudaMemcpyDeviceToHost, s));
    if (n_launches) *n_launches = total;
    return SIFT_B200_OK;
}

template <typename T>
int enqueue_detect(sift_b200_ctx* c, const T* d_pixels, int width, int height, int channels,
                   const sift_b200_params& p) {
    int rc = check_params(c, p);
    if (rc) return rc;
    if (width < 2 || height < 2) return fail(c, SIFT_B200_E_INVALID, "image %dx%d is too small", width, height);
    if (channels != 1 && channels != 3) return fail(c, SIFT_B200_E_INVALID, "channels must be 1 or 3");
    if (width > c->max_w || height > c->max_h)
        return fail(c, SIFT_B200_E_TOO_LARGE, "image %dx%d exceeds the context's %dx%d", width, height,
                    c->max_w, c->max_h);
    const int doubled = p.double_image_size ? 1 : 0;
    const int bw = doubled ? 2 * width : width, bh = doubled ? 2 * height : height;
    const int layers = p.intervals + 3, dogs = p.intervals + 2;
    // 32-bit plane offsets in the scale-space kernels (also checked for the context's maximum at creation)
    if ((long long)round_up(bw, 32) * bh >= (1ll << 31) || arena_need(bw, bh, layers + dogs) > c->arena_floats)
        return fail(c, SIFT_B200_E_TOO_LARGE, "image %dx%d exceeds the context's workspace", width, height);
    int octaves = octave_count(bw, bh);
    if (p.max_octaves > 0) octaves = std::min(octaves, p.max_octaves);
    octaves = std::min(octaves, kMaxOctaves);
    if (octaves < 0) octaves = 0;
    // resize_inter_nearest throws below 2x2 (image.cpp:42-44); octave_count never gets there
    layout_pyramid(c, bw, bh, octaves, layers, dogs);
    c->layers = layers; c->dogs = dogs;
    c->base_w = bw; c->base_h = bh;
    cudaStream_t s = c->stream;
    if (!c->pyr_uploaded_valid || memcmp(&c->pyr, &c->pyr_uploaded, sizeof(PyramidDesc)) != 0) {
        // pageable source: the copy is staged before the call returns, so c->pyr may change afterwards
        CU(c, cudaMemcpyAsync(c->d_pyr, &c->pyr, sizeof(PyramidDesc), cudaMemcpyHostToDevice, s));
        memcpy(&c->pyr_uploaded, &c->pyr, sizeof(PyramidDesc));
        c->pyr_uploaded_valid = true;
    }
    CU(c, cudaMemsetAsync(c->d_counters, 0, kCountersBytes, s));

    StageParams& sp = c->sp;
    memset(&sp, 0, sizeof sp);   // (it is part of the graph key: no indeterminate padding)
    sp.doubled = doubled;
    sp.intervals = p.intervals;
    sp.dogs = dogs;
    sp.num_bins = (int)p.num_bins;
    sp.border = p.window_size / 2;
    sp.mag_bound = 361.0;   // > sqrt(2) * 255
    sp.range = nullptr;
    c->last_float_input = sizeof(T) != 1;
    if (sizeof(T) != 1) {   // float input: any range -- reduce it on the device
        CU(c, launch_range((const float*)d_pixels, (size_t)width * height * channels, c->d_range, c->sm_count, s));
        sp.range = c->d_range;
        c->launches += 1;
    }
    sp.dog_threshold = (int)floor(0.5 * p.contrast_threshold / p.intervals * 255.0);  // sift.cpp:305-307
    sp.init_sigma = p.init_sigma;
    sp.contrast_threshold = p.contrast_threshold;
    sp.eigen_ratio = p.eigen_ratio;
    sp.peak_ratio = p.peak_ratio;
    sp.ori_sigma_factor = p.ori_sigma_factor;
    sp.desc_scale_factor = p.desc_scale_factor;
    sp.cap_extrema = c->cap_extrema;
    sp.cap_raw = c->cap_raw;
    sp.cap_oriented = c->cap_oriented;

    c->stats = sift_b200_stats{};
    c->stats.octaves = octaves;
    c->stats.base_width = bw;
    c->stats.base_height = bh;
    c->have_result = false;
    c->pyramid_valid = false;
    c->ev_used = 0;
    memset(c->stage_launches, 0, sizeof c->stage_launches);
    prof_mark(c, -1);
    if (octaves == 0) {
        c->detect_pending = true;
        CU(c, cudaMemcpyAsync(c->h_counters, c->d_counters, sizeof(Counters), cudaMemcpyDeviceToHost, s));
        return SIFT_B200_OK;
    }

    // Scale-space sigmas, compute_gaussian_kernels sift.cpp:143-155.
    DetectPlan pl;
    memset(&pl, 0, sizeof pl);
    pl.octaves = octaves; pl.layers = layers; pl.dogs = dogs;
    double sig[kMaxLayers];
    const double k = pow(2.0, 1.0 / p.intervals);
    sig[0] = p.init_sigma;
    for (int i = 1; i < layers; ++i) sig[i] = pow(k, i - 1) * p.init_sigma * sqrt(k * k - 1);
    BlurTaps* taps = pl.taps;
    taps[0] = make_taps(sqrt(p.init_sigma * p.init_sigma - 1.0));  // sift.cpp:124: "-1" in both modes
    for (int i = 1; i < layers; ++i) taps[i] = make_taps(sig[i]);
    for (int i = 0; i < layers; ++i)
        if (taps[i].radius > kMaxRadius)
            return fail(c, SIFT_B200_E_UNSUPPORTED, "blur radius %d > %d (init_sigma / intervals combination)",
                        taps[i].radius, kMaxRadius);

    // Stage 0: gray (+2x) into a scratch plane (octave 0's last G slot, dead until the cascade
    // reaches it), then the initial blur into G[0].  Not part of the graph: the input pointer changes per call.
    OctaveDesc& o0 = c->pyr.oct[0];
    float* scratch = o0.G[layers - 1];
    const float centre = c->centred ? c->centre_u8 : 0.f;
    if (sizeof(T) == 1 && input_fused_supported(channels, taps[0]) && !c->force_unfused) {
        CU(c, launch_input_u8((const uint8_t*)d_pixels, width, height, channels, o0.G[0], bw, bh, o0.pitch, doubled, taps[0], centre, s));
        prof_mark(c, SIFT_B200_STAGE_INPUT, 1);
    } else {
        if (sizeof(T) == 1)
            CU(c, launch_prepare_u8((const uint8_t*)d_pixels, width, height, channels, scratch, bw, bh, o0.pitch,
                                    doubled, centre, s));
        else
            CU(c, launch_prepare_f32((const float*)d_pixels, width, height, channels, scratch, bw, bh, o0.pitch,
                                     doubled, c->centred ? c->d_range : nullptr, s));
        CU(c, launch_blur(scratch, o0.G[0], nullptr, nullptr, bw, bh, o0.pitch, 0, 0, 0, taps[0], s));
        prof_mark(c, SIFT_B200_STAGE_INPUT, 2);
    }

    pl.fused = p.intervals == 3 && cascade_supported(taps) && !c->force_unfused;
    c->ss.nb = std::min(width + 2, c->ss_nb_cap);
    c->detect_pending = true;
    if (!c->use_graph || c->profiling) {
        int n = 0;
        rc = enqueue_body(c, pl, false, &n, c->stage_launches);
        c->launches += n;
        return rc;
    }
    // graph key: everything the captured launches depend on (geometry, taps, stage parameters, debug switches)
    std::vector<uint8_t> key;
    auto put = [&](const void* q, size_t n) { key.insert(key.end(), (const uint8_t*)q, (const uint8_t*)q + n); };
    put(&pl, sizeof pl); put(&sp, sizeof sp); put(&c->pyr, sizeof c->pyr); put(&c->ss.nb, sizeof(int));
    const int dbg[8] = {c->keep_planes, c->force_unfused, c->fused_mode, c->extrema_form, c->three_branches,
                        (int)std::min<long long>(c->fork_min_px, 1ll << 30), c->use_tail, c->use_cubes};
    put(dbg, sizeof dbg);
    if (!c->graph_exec || key != c->graph_key) {
        if (c->graph_exec) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; }
        memset(c->graph_stage_launches, 0, sizeof c->graph_stage_launches);
        CU(c, cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        rc = enqueue_body(c, pl, true, &c->graph_launches, c->graph_stage_launches);
        cudaGraph_t g = nullptr;
        const cudaError_t e = cudaStreamEndCapture(s, &g);
        if (rc) { if (g) cudaGraphDestroy(g); c->detect_pending = false; return rc; }
        if (e != cudaSuccess) {
            c->detect_pending = false;
            return fail(c, SIFT_B200_E_CUDA, "stream capture failed: %s", cudaGetErrorString(e));
        }
        const cudaError_t e2 = cudaGraphInstantiate(&c->graph_exec, g, 0);
        cudaGraphDestroy(g);
        if (e2 != cudaSuccess) {
            c->graph_exec = nullptr;
            c->detect_pending = false;
            return fail(c, SIFT_B200_E_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e2));
        }
        c->graph_key.swap(key);
        c->graphs_built++;
    }
    CU(c, cudaGraphLaunch(c->graph_exec, s));
    c->launches += c->graph_launches;
    for (int i = 0; i < SIFT_B200_STAGE_COUNT; ++i) c->stage_launches[i] += c->graph_stage_launches[i];
    return SIFT_B200_OK;
}

int finish_detect(sift_b200_ctx* c, int* count) {
    if (!c->detect_pending && !c->have_result) return fail(c, SIFT_B200_E_INVALID, "no detect call is pending");
    if (c->detect_pending) {
        CU(c, cudaStreamSynchronize(c->stream));
        c->detect_pending = false;
        const Counters& k = *c->h_counters;
        c->stats.extrema = k.n_extrema;
        c->stats.raw_keypoints = k.n_raw;
        c->stats.oriented_keypoints = k.n_oriented;
        c->stats.final_keypoints = k.n_final;
        if (k.n_extrema > c->cap_extrema || k.n_raw > c->cap_raw || k.n_oriented > c->cap_oriented) {
            c->have_result = false;
            return fail(c, SIFT_B200_E_CAPACITY,
                        "internal list overflow: extrema %d/%d raw %d/%d oriented %d/%d -- create the "
                        "context for a larger image",
                        k.n_extrema, c->cap_extrema, k.n_raw, c->cap_raw, k.n_oriented, c->cap_oriented);
        }
        c->have_result = true;
        c->pyramid_valid = true;
    }
    if (count) *count = c->stats.final_keypoints;
    return SIFT_B200_OK;
}

template <typename T>
int detect_sync(sift_b200_ctx* c, const T* pixels, int width, int height, int channels,
                const sift_b200_params* params, sift_b200_keypoint* out, int capacity, int* count) {
    if (!c) return SIFT_B200_E_INVALID;
    if (!pixels || !count || (capacity > 0 && !out)) return fail(c, SIFT_B200_E_INVALID, "null argument");
    if (width < 2 || height < 2) return fail(c, SIFT_B200_E_INVALID, "image %dx%d is too small", width, height);
    if (channels != 1 && channels != 3) return fail(c, SIFT_B200_E_INVALID, "channels must be 1 or 3");
    CU(c, cudaSetDevice(c->device));
    sift_b200_params p;
    if (params) p = *params; else sift_b200_default_params(&p);
    const T* d_px = pixels;
    if (!is_device_ptr(pixels)) {
        const size_t bytes = (size_t)width * height * channels * sizeof(T);
        if (bytes > c->input_bytes)
            return fail(c, SIFT_B200_E_TOO_LARGE, "image %dx%dx%d exceeds the context's staging buffer", width,
                        height, channels);
        CU(c, cudaMemcpyAsync(c->d_input, pixels, bytes, cudaMemcpyHostToDevice, c->stream));
        d_px = (const T*)c->d_input;
    }
    int rc = enqueue_detect<T>(c, d_px, width, height, channels, p);
    if (rc) return rc;
    int n = 0;
    rc = finish_detect(c, &n);
    *count = n;
    if (rc) return rc;
    const int ncopy = std::min(n, capacity);
    if (ncopy > 0)
        CU(c, cudaMemcpy(out, c->d_records, (size_t)ncopy * sizeof(sift_b200_keypoint), cudaMemcpyDeviceToHost));
    if (n > capacity) return fail(c, SIFT_B200_E_CAPACITY, "%d keypoints found, output capacity %d", n, capacity);
    return SIFT_B200_OK;
}

}  // namespace

extern "C" {

void sift_b200_default_params(sift_b200_params* p) {
    if (!p) return;
    p->double_image_size = 1;
    p->init_sigma = 1.6;
    p->intervals = 3;
    p->window_size = 3;
    p->contrast_threshold = 0.04;
    p->eigen_ratio = 10.0;
    p->num_bins = 36;
    p->peak_ratio = 0.8;
    p->ori_sigma_factor = 1.5;
    p->desc_scale_factor = 3.0;
    p->max_octaves = 0;
}

const char* sift_b200_last_error(const sift_b200_ctx* ctx) {
    return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

int sift_b200_create(int device, int max_width, int max_height, sift_b200_ctx** out) {
    if (!out) return fail(nullptr, SIFT_B200_E_INVALID, "out is null");
    *out = nullptr;
    if (max_width < 2 || max_height < 2) return fail(nullptr, SIFT_B200_E_INVALID, "bad maximum image size");
    // plane offsets are 32-bit in the scale-space kernels: the doubled base plane (pitch padded to 32) must hold
    // fewer than 2^31 pixels (that is 32 768 x 16 384 input pixels; its arena would be ~100 GB anyway)
    if ((2ll * max_width + 32) * (2ll * max_height) >= (1ll << 31))
        return fail(nullptr, SIFT_B200_E_UNSUPPORTED, "maximum image size %d x %d is too large", max_width, max_height);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(nullptr, SIFT_B200_E_NO_DEVICE, "no CUDA device is visible; this library has no CPU path");
    }
    if (device < 0 || device >= ndev) return fail(nullptr, SIFT_B200_E_NO_DEVICE, "device %d of %d", device, ndev);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess)
        return fail(nullptr, SIFT_B200_E_NO_DEVICE, "cannot query device %d", device);
    if (prop.major != 10)
        return fail(nullptr, SIFT_B200_E_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only",
                    device, prop.major, prop.minor);
    sift_b200_ctx* c = new (std::nothrow) sift_b200_ctx();
    if (!c) return fail(nullptr, SIFT_B200_E_INVALID, "out of host memory");
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    if (const char* m = getenv("SIFT_B200_PYRAMID_MODE")) {   // experiments: same values as sift_b200_debug_options
        const int v = atoi(m);
        c->force_unfused = v == 1;
        c->fused_mode = v == 2 || v == 3 ? v : 0;
    }
    if (const char* m = getenv("SIFT_B200_CENTER")) c->centred = atoi(m) != 0;   // experiments
    if (const char* m = getenv("SIFT_B200_GRAPH")) { c->use_graph = atoi(m) != 0; c->three_branches = atoi(m) == 3; }
    if (const char* m = getenv("SIFT_B200_EXTREMA")) c->extrema_form = atoi(m);
    if (const char* m = getenv("SIFT_B200_FORK_MIN_PX")) c->fork_min_px = atoll(m);
    if (const char* m = getenv("SIFT_B200_TAIL")) c->use_tail = atoi(m) != 0;
    int cube_cap_env = 0;   // SIFT_B200_CUBES: 0 = off, 1 = on, n > 1 = hand over the first n candidates only (tests)
    if (const char* m = getenv("SIFT_B200_CUBES")) { c->use_cubes = atoi(m) != 0; cube_cap_env = atoi(m) > 1 ? atoi(m) : 0; }
    c->max_w = max_width;
    c->max_h = max_height;
#define CRT(call)                                                                                   \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            fail(nullptr, SIFT_B200_E_CUDA, "%s failed: %s", #call, cudaGetErrorString(e__));       \
            sift_b200_destroy(c);                                                                   \
            return SIFT_B200_E_CUDA;                                                                \
        }                                                                                           \
    } while (0)
    CRT(cudaSetDevice(device));
    CRT(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CRT(cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking));
    CRT(cudaStreamCreateWithFlags(&c->side2, cudaStreamNonBlocking));
    for (int i = 0; i <= 2 * kMaxOctaves + 1; ++i) {
        cudaEvent_t e;
        CRT(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->fork_ev.push_back(e);
    }
    CRT(pyramid_init());
    CRT(match_init());
    c->arena_floats = arena_need(2 * max_width, 2 * max_height, kMaxLayers + kMaxDogs);
    CRT(cudaMalloc(&c->arena, c->arena_floats * sizeof(float)));
    c->input_bytes = (size_t)max_width * max_height * 3 * sizeof(float);
    CRT(cudaMalloc(&c->d_input, c->input_bytes));
    CRT(cudaMalloc(&c->d_pyr, sizeof(PyramidDesc)));
    CRT(cudaMalloc(&c->d_counters, kCountersBytes));   // the counters + k_tail's ticket / hand-over counters
    CRT(cudaMalloc(&c->d_range, 2 * sizeof(float)));
    CRT(cudaMallocHost(&c->h_counters, sizeof(Counters)));
    const size_t px = (size_t)max_width * max_height;
    c->cap_extrema = (int)std::max<size_t>(1 << 16, px / 2);
    c->cap_raw = (int)std::max<size_t>(1 << 15, px / 8);
    c->cap_oriented = (int)std::max<size_t>(1 << 15, px / 8);
    CRT(cudaMalloc(&c->d_cands, (size_t)c->cap_extrema * sizeof(Cand)));
    c->cube_cap = std::min(c->cap_extrema, 1 << 20);   // 80 B each; candidates beyond it are fetched by the refinement
    if (cube_cap_env) c->cube_cap = std::min(c->cube_cap, cube_cap_env);
    CRT(cudaMalloc(&c->d_cubes, (size_t)c->cube_cap * sizeof(CandCube)));
    CRT(cudaMalloc(&c->d_raw, (size_t)c->cap_raw * sizeof(KpCore)));
    CRT(cudaMalloc(&c->d_oriented, (size_t)c->cap_oriented * sizeof(KpCore)));
    CRT(cudaMalloc(&c->d_records, (size_t)c->cap_oriented * 168));
    CRT(cudaMalloc(&c->d_desc, (size_t)c->cap_oriented * 128));
    c->ss_nb_cap = 2 * max_width + 70;
    const size_t nbp = (size_t)c->ss_nb_cap + 1;
    CRT(cudaMalloc(&c->ss.bucket_cnt, nbp * sizeof(int)));
    CRT(cudaMalloc(&c->ss.bucket_off, nbp * sizeof(int)));
    CRT(cudaMalloc(&c->ss.bucket_fill, nbp * sizeof(int)));
    CRT(cudaMalloc(&c->ss.uniq_cnt, nbp * sizeof(int)));
    CRT(cudaMalloc(&c->ss.uniq_off, nbp * sizeof(int)));
    CRT(cudaMalloc(&c->ss.perm, (size_t)c->cap_oriented * sizeof(int)));
    CRT(cudaMalloc(&c->ss.tmp_sorted, (size_t)c->cap_oriented * sizeof(int)));
    CRT(cudaMalloc(&c->ss.sorted, (size_t)c->cap_oriented * sizeof(int)));
    CRT(cudaMalloc(&c->ss.final_order, (size_t)c->cap_oriented * sizeof(int)));
    CRT(cudaMalloc(&c->d_out_count, sizeof(int)));
#undef CRT
    *out = c;
    return SIFT_B200_OK;
}

void sift_b200_destroy(sift_b200_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->side) cudaStreamSynchronize(c->side);
    collection_free(c);
    void* ptrs[] = {c->arena, c->d_input, c->d_pyr, c->d_counters, c->d_range, c->d_cands, c->d_cubes, c->d_raw, c->d_oriented,
                    c->d_records, c->d_desc, c->ss.bucket_cnt, c->ss.bucket_off, c->ss.bucket_fill,
                    c->ss.uniq_cnt, c->ss.uniq_off, c->ss.perm, c->ss.tmp_sorted, c->ss.sorted,
                    c->ss.final_order, c->ms.part_idx, c->ms.part_d1, c->ms.part_d2, c->ms.norms_a,
                    c->ms.norms_b, c->d_ma, c->d_mb, c->d_best_idx, c->d_best_d2, c->d_second_d2,
                    c->d_out_ia, c->d_out_ib, c->d_out_dist, c->d_out_count};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
    for (cudaEvent_t e : c->fork_ev) cudaEventDestroy(e);
    if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
    if (c->h_counters) cudaFreeHost(c->h_counters);
    if (c->side) cudaStreamDestroy(c->side);
    if (c->side2) cudaStreamDestroy(c->side2);
    if (c->stream) cudaStreamDestroy(c->stream);
    cudaGetLastError();
    delete c;
}

int sift_b200_detect_u8(sift_b200_ctx* c, const uint8_t* pixels, int width, int height, int channels,
                        const sift_b200_params* params, sift_b200_keypoint* out, int capacity, int* count) {
    return detect_sync<uint8_t>(c, pixels, width, height, channels, params, out, capacity, count);
}

int sift_b200_detect_f32(sift_b200_ctx* c, const float* pixels, int width, int height, int channels,
                         const sift_b200_params* params, sift_b200_keypoint* out, int capacity, int* count) {
    return detect_sync<float>(c, pixels, width, height, channels, params, out, capacity, count);
}

int sift_b200_host_alloc(size_t bytes, void** out) {
    if (!out || bytes == 0) return SIFT_B200_E_INVALID;
    *out = nullptr;
    const cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocPortable);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(nullptr, SIFT_B200_E_CUDA, "cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    }
    return SIFT_B200_OK;
}

void sift_b200_host_free(void* p) {
    if (p) cudaFreeHost(p);
    cudaGetLastError();
}

int sift_b200_detect_enqueue_u8(sift_b200_ctx* c, const uint8_t* pixels, int width, int height, int channels,
                                const sift_b200_params* params) {
    if (!c) return SIFT_B200_E_INVALID;
    if (!pixels) return fail(c, SIFT_B200_E_INVALID, "null argument");
    if (width < 2 || height < 2) return fail(c, SIFT_B200_E_INVALID, "image %dx%d is too small", width, height);
    if (channels != 1 && channels != 3) return fail(c, SIFT_B200_E_INVALID, "channels must be 1 or 3");
    CU(c, cudaSetDevice(c->device));
    sift_b200_params p;
    if (params) p = *params; else sift_b200_default_params(&p);
    const uint8_t* d_px = pixels;
    if (!is_device_ptr(pixels)) {
        const size_t bytes = (size_t)width * height * channels;
        if (bytes > c->input_bytes)
            return fail(c, SIFT_B200_E_TOO_LARGE, "image %dx%dx%d exceeds the context's staging buffer", width,
                        height, channels);
        CU(c, cudaMemcpyAsync(c->d_input, pixels, bytes, cudaMemcpyHostToDevice, c->stream));
        d_px = c->d_input;
    }
    return enqueue_detect<uint8_t>(c, d_px, width, height, channels, p);
}

int sift_b200_result_copy(sift_b200_ctx* c, sift_b200_keypoint* out, int capacity, int* count) {
    if (!c || !count || (capacity > 0 && !out)) return SIFT_B200_E_INVALID;
    CU(c, cudaSetDevice(c->device));
    int n = 0;
    int rc = finish_detect(c, &n);
    *count = n;
    if (rc) return rc;
    const int ncopy = std::min(n, capacity);
    if (ncopy > 0) {
        CU(c, cudaMemcpyAsync(out, c->d_records, (size_t)ncopy * sizeof(sift_b200_keypoint),
                              cudaMemcpyDeviceToHost, c->stream));
        CU(c, cudaStreamSynchronize(c->stream));
    }
    if (n > capacity) return fail(c, SIFT_B200_E_CAPACITY, "%d keypoints found, output capacity %d", n, capacity);
    return SIFT_B200_OK;
}

int sift_b200_set_profiling(sift_b200_ctx* c, int on) {
    if (!c) return SIFT_B200_E_INVALID;
    c->profiling = on != 0;
    c->ev_used = 0;
    return SIFT_B200_OK;
}

int sift_b200_get_profile(sift_b200_ctx* c, float* stage_ms, int32_t* stage_launches) {
    if (!c || !stage_ms) return SIFT_B200_E_INVALID;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaStreamSynchronize(c->stream));
    for (int i = 0; i < SIFT_B200_STAGE_COUNT; ++i) {
        stage_ms[i] = 0.f;
        if (stage_launches) stage_launches[i] = c->stage_launches[i];
    }
    for (int i = 1; i < c->ev_used; ++i) {
        float ms = 0.f;
        CU(c, cudaEventElapsedTime(&ms, c->ev_pool[i - 1], c->ev_pool[i]));
        const int st = c->ev_stage[i];
        if (st >= 0 && st < SIFT_B200_STAGE_COUNT) stage_ms[st] += ms;
    }
    return SIFT_B200_OK;
}

int sift_b200_get_profile_marks(sift_b200_ctx* c, int32_t* stage, float* ms, int capacity, int* count) {
    if (!c || !count) return SIFT_B200_E_INVALID;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaStreamSynchronize(c->stream));
    *count = c->ev_used > 0 ? c->ev_used - 1 : 0;
    for (int i = 1; i < c->ev_used && i - 1 < capacity; ++i) {
        float t = 0.f;
        CU(c, cudaEventElapsedTime(&t, c->ev_pool[i - 1], c->ev_pool[i]));
        if (stage) stage[i - 1] = c->ev_stage[i];
        if (ms) ms[i - 1] = t;
    }
    return SIFT_B200_OK;
}

int sift_b200_detect_finish(sift_b200_ctx* c, int* count) {
    if (!c) return SIFT_B200_E_INVALID;
    CU(c, cudaSetDevice(c->device));
    return finish_detect(c, count);
}

int sift_b200_result_device(sift_b200_ctx* c, const sift_b200_keypoint** d_records, const uint8_t** d_descriptors,
                            int* n) {
    if (!c) return SIFT_B200_E_INVALID;
    int count = 0;
    int rc = finish_detect(c, &count);
    if (rc) return rc;
    if (d_records) *d_records = (const sift_b200_keypoint*)c->d_records;
    if (d_descriptors) *d_descriptors = c->d_desc;
    if (n) *n = count;
    return SIFT_B200_OK;
}

int sift_b200_get_stats(sift_b200_ctx* c, sift_b200_stats* stats) {
    if (!c || !stats) return SIFT_B200_E_INVALID;
    if (c->detect_pending) {
        int rc = finish_detect(c, nullptr);
        if (rc && rc != SIFT_B200_E_CAPACITY) return rc;
    }
    *stats = c->stats;
    return SIFT_B200_OK;
}

int sift_b200_match_enqueue(sift_b200_ctx* c, const uint8_t* d_a, int na, const uint8_t* d_b, int nb,
                            int32_t* d_best_idx, int32_t* d_best_d2, int32_t* d_second_d2) {
    if (!c) return SIFT_B200_E_INVALID;
    if (na < 0 || nb < 0 || (na > 0 && (!d_a || !d_best_idx || !d_best_d2 || !d_second_d2)) || (nb > 0 && !d_b))
        return fail(c, SIFT_B200_E_INVALID, "bad match arguments");
    if (((uintptr_t)d_a | (uintptr_t)d_b) & 15)
        return fail(c, SIFT_B200_E_INVALID, "descriptor matrices must be 16-byte aligned");
    CU(c, cudaSetDevice(c->device));
    if (na == 0) return SIFT_B200_OK;
    return enqueue_match(c, d_a, na, d_b, nb, d_best_idx, d_best_d2, d_second_d2);
}

int sift_b200_match(sift_b200_ctx* c, const uint8_t* desc_a, int na, const uint8_t* desc_b, int nb, double ratio,
                    int32_t* idx_a, int32_t* idx_b, double* dist, int capacity, int* count) {
    if (!c) return SIFT_B200_E_INVALID;
    if (!count || na < 0 || nb < 0 || (na > 0 && !desc_a) || (nb > 0 && !desc_b) ||
        (capacity > 0 && (!idx_a || !idx_b || !dist)))
        return fail(c, SIFT_B200_E_INVALID, "bad match arguments");
    *count = 0;
    CU(c, cudaSetDevice(c->device));
    if (na == 0 || nb == 0) return SIFT_B200_OK;  // sift.cpp:789-812: nothing to emit
    cudaStream_t s = c->stream;
    const uint8_t* d_a = desc_a;
    const uint8_t* d_b = desc_b;
    const bool dev_a = is_device_ptr(desc_a), dev_b = is_device_ptr(desc_b);
    if ((dev_a && ((uintptr_t)desc_a & 15)) || (dev_b && ((uintptr_t)desc_b & 15)))
        return fail(c, SIFT_B200_E_INVALID, "device descriptor matrices must be 16-byte aligned");
    if (!dev_a) {
        if ((size_t)na > c->ma_cap) {
            CU(c, cudaStreamSynchronize(s));
            if (c->d_ma) cudaFree(c->d_ma);
            c->d_ma = nullptr;
            c->ma_cap = (size_t)na * 5 / 4 + 1024;
            CU(c, cudaMalloc(&c->d_ma, c->ma_cap * 128));
        }
        CU(c, cudaMemcpyAsync(c->d_ma, desc_a, (size_t)na * 128, cudaMemcpyHostToDevice, s));
        d_a = c->d_ma;
    }
    if (!dev_b) {
        if ((size_t)nb > c->mb_cap) {
            CU(c, cudaStreamSynchronize(s));
            if (c->d_mb) cudaFree(c->d_mb);
            c->d_mb = nullptr;
            c->mb_cap = (size_t)nb * 5 / 4 + 1024;
            CU(c, cudaMalloc(&c->d_mb, c->mb_cap * 128));
        }
        CU(c, cudaMemcpyAsync(c->d_mb, desc_b, (size_t)nb * 128, cudaMemcpyHostToDevice, s));
        d_b = c->d_mb;
    }
    if ((size_t)na > c->best_cap) {
        const size_t rows = (size_t)na * 5 / 4 + 1024;
        int rc;
        if ((rc = grow_i32(c, &c->d_best_idx, rows))) return rc;
        if ((rc = grow_i32(c, &c->d_best_d2, rows))) return rc;
        if ((rc = grow_i32(c, &c->d_second_d2, rows))) return rc;
        if ((rc = grow_i32(c, &c->d_out_ia, rows))) return rc;
        if ((rc = grow_i32(c, &c->d_out_ib, rows))) return rc;
        if (c->d_out_dist) {
            CU(c, cudaStreamSynchronize(s));
            cudaFree(c->d_out_dist);
        }
        c->d_out_dist = nullptr;
        CU(c, cudaMalloc(&c->d_out_dist, rows * sizeof(double)));
        c->best_cap = rows;
    }
    int rc = enqueue_match(c, d_a, na, d_b, nb, c->d_best_idx, c->d_best_d2, c->d_second_d2);
    if (rc) return rc;
    CU(c, launch_match_emit(c->d_best_idx, c->d_best_d2, c->d_second_d2, na, nb, ratio, c->d_out_ia, c->d_out_ib,
                            c->d_out_dist, na, c->d_out_count, s));
    c->launches += 1;
    int n = 0;
    CU(c, cudaMemcpyAsync(&n, c->d_out_count, sizeof(int), cudaMemcpyDeviceToHost, s));
    CU(c, cudaStreamSynchronize(s));
    *count = n;
    const int ncopy = std::min(n, capacity);
    if (ncopy > 0) {
        CU(c, cudaMemcpyAsync(idx_a, c->d_out_ia, (size_t)ncopy * sizeof(int), cudaMemcpyDeviceToHost, s));
        CU(c, cudaMemcpyAsync(idx_b, c->d_out_ib, (size_t)ncopy * sizeof(int), cudaMemcpyDeviceToHost, s));
        CU(c, cudaMemcpyAsync(dist, c->d_out_dist, (size_t)ncopy * sizeof(double), cudaMemcpyDeviceToHost, s));
        CU(c, cudaStreamSynchronize(s));
    }
    if (n > capacity) return fail(c, SIFT_B200_E_CAPACITY, "%d matches, output capacity %d", n, capacity);
    return SIFT_B200_OK;
}

int sift_b200_sync(sift_b200_ctx* c) {
    if (!c) return SIFT_B200_E_INVALID;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaStreamSynchronize(c->stream));
    return SIFT_B200_OK;
}

void* sift_b200_stream(sift_b200_ctx* c) { return c ? (void*)c->stream : nullptr; }

int sift_b200_match_path(int na, int nb) { return match_uses_tensor_cores(na, nb) ? 1 : 0; }

int sift_b200_debug_options(sift_b200_ctx* c, int keep_all_planes, int unfused_pyramid) {
    if (!c) return SIFT_B200_E_INVALID;
    c->keep_planes = keep_all_planes != 0;
    c->force_unfused = unfused_pyramid == 1;
    c->fused_mode = unfused_pyramid == 2 || unfused_pyramid == 3 ? unfused_pyramid : 0;
    return SIFT_B200_OK;
}

int sift_b200_debug_plane_dims(sift_b200_ctx* c, int octave, int* width, int* height) {
    if (!c || octave < 0 || octave >= c->pyr.octaves) return SIFT_B200_E_INVALID;
    if (width) *width = c->pyr.oct[octave].w;
    if (height) *height = c->pyr.oct[octave].h;
    return SIFT_B200_OK;
}

int sift_b200_debug_plane(sift_b200_ctx* c, int kind, int octave, int layer, float* host_out) {
    if (!c || !host_out || octave < 0 || octave >= c->pyr.octaves) return SIFT_B200_E_INVALID;
    const OctaveDesc& od = c->pyr.oct[octave];
    const float* src = nullptr;
    if (kind == SIFT_B200_PLANE_GAUSSIAN && layer >= 0 && layer < c->layers) src = od.G[layer];
    if (kind == SIFT_B200_PLANE_DOG && layer >= 0 && layer < c->dogs) src = od.D[layer];
    if (!src) return fail(c, SIFT_B200_E_INVALID, "bad plane kind/layer");
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaStreamSynchronize(c->stream));
    CU(c, cudaMemcpy2D(host_out, (size_t)od.w * sizeof(float), src, (size_t)od.pitch * sizeof(float),
                       (size_t)od.w * sizeof(float), od.h, cudaMemcpyDeviceToHost));
    if (kind == SIFT_B200_PLANE_GAUSSIAN && c->centred) {   // the planes are stored relative to the input's mid level
        float centre = c->centre_u8;
        if (c->last_float_input) {
            float r[2];
            CU(c, cudaMemcpy(r, c->d_range, sizeof r, cudaMemcpyDeviceToHost));
            centre = (float)(0.5 * ((double)r[0] + (double)r[1]));
        }
        for (size_t i = 0, n = (size_t)od.w * od.h; i < n; ++i) host_out[i] += centre;
    }
    return SIFT_B200_OK;
}

int sift_b200_debug_extrema(sift_b200_ctx* c, int32_t* host_out, int capacity, int* count) {
    if (!c || !count) return SIFT_B200_E_INVALID;
    int rc = finish_detect(c, nullptr);
    if (rc) return rc;
    const int n = c->stats.extrema;
    *count = n;
    const int ncopy = std::min(n, capacity);
    if (ncopy > 0 && host_out)
        CU(c, cudaMemcpy(host_out, c->d_cands, (size_t)ncopy * sizeof(Cand), cudaMemcpyDeviceToHost));
    return SIFT_B200_OK;
}

int sift_b200_debug_keypoints(sift_b200_ctx* c, int stage, sift_b200_keypoint* host_out, int capacity, int* count) {
    if (!c || !count || (stage != 0 && stage != 1)) return SIFT_B200_E_INVALID;
    int rc = finish_detect(c, nullptr);
    if (rc) return rc;
    const int n = stage == 0 ? c->stats.raw_keypoints : c->stats.oriented_keypoints;
    *count = n;
    const int ncopy = std::min(n, capacity);
    if (ncopy > 0 && host_out) {
        std::vector<KpCore> tmp(ncopy);
        CU(c, cudaMemcpy(tmp.data(), stage == 0 ? c->d_raw : c->d_oriented, (size_t)ncopy * sizeof(KpCore),
                         cudaMemcpyDeviceToHost));
        for (int i = 0; i < ncopy; ++i) {
            memset(&host_out[i], 0, sizeof(sift_b200_keypoint));
            host_out[i].x = tmp[i].x; host_out[i].y = tmp[i].y;
            host_out[i].octave = tmp[i].octave; host_out[i].layer = tmp[i].layer;
            host_out[i].size = tmp[i].size; host_out[i].pori = tmp[i].pori;
        }
    }
    return SIFT_B200_OK;
}

// ---- single stages on caller-supplied keypoints over the last detect's scale space (parity attribution) ----
static int debug_stage_ready(sift_b200_ctx* c) {
    if (c->detect_pending) {
        int rc = finish_detect(c, nullptr);
        if (rc) return rc;
    }
    if (!c->pyramid_valid || c->stats.octaves == 0)
        return fail(c, SIFT_B200_E_INVALID, "no scale space: run a detect call first");
    return SIFT_B200_OK;
}

int sift_b200_debug_orient(sift_b200_ctx* c, const sift_b200_keypoint* raw_in, int n, sift_b200_keypoint* out,
                           int capacity, int* count) {
    if (!c || !count || n < 0 || (n > 0 && !raw_in) || (capacity > 0 && !out)) return SIFT_B200_E_INVALID;
    CU(c, cudaSetDevice(c->device));
    int rc = debug_stage_ready(c);
    if (rc) return rc;
    if (n > c->cap_raw) return fail(c, SIFT_B200_E_CAPACITY, "%d keypoints, capacity %d", n, c->cap_raw);
    for (int i = 0; i < n; ++i)
        if (raw_in[i].octave < 0 || raw_in[i].octave >= c->stats.octaves || raw_in[i].layer < 1 ||
            raw_in[i].layer > c->layers - 3)
            return fail(c, SIFT_B200_E_INVALID, "keypoint %d: octave / layer outside the scale space", i);
    std::vector<KpCore> tmp(std::max(n, 1));
    for (int i = 0; i < n; ++i)
        tmp[i] = KpCore{raw_in[i].x, raw_in[i].y, raw_in[i].octave, raw_in[i].layer, raw_in[i].size, raw_in[i].pori};
    Counters k;
    memset(&k, 0, sizeof k);
    k.n_raw = n;
    cudaStream_t s = c->stream;
    c->have_result = false;   // the lists of the detect call are overwritten
    CU(c, cudaMemcpyAsync(c->d_raw, tmp.data(), (size_t)n * sizeof(KpCore), cudaMemcpyHostToDevice, s));
    CU(c, cudaMemcpyAsync(c->d_counters, &k, sizeof k, cudaMemcpyHostToDevice, s));
    CU(c, launch_orient(c->d_pyr, c->d_raw, c->d_oriented, c->d_counters, c->sp, c->sm_count, s));
    CU(c, cudaMemcpyAsync(&k, c->d_counters, sizeof k, cudaMemcpyDeviceToHost, s));
    CU(c, cudaStreamSynchronize(s));
    c->launches += 1;
    *count = k.n_oriented;
    if (k.n_oriented > c->cap_oriented) return fail(c, SIFT_B200_E_CAPACITY, "oriented list overflow");
    const int ncopy = std::min(k.n_oriented, capacity);
    if (ncopy > 0) {
        tmp.resize(ncopy);
        CU(c, cudaMemcpy(tmp.data(), c->d_oriented, (size_t)ncopy * sizeof(KpCore), cudaMemcpyDeviceToHost));
        for (int i = 0; i < ncopy; ++i) {
            memset(&out[i], 0, sizeof(sift_b200_keypoint));
            out[i].x = tmp[i].x; out[i].y = tmp[i].y; out[i].octave = tmp[i].octave; out[i].layer = tmp[i].layer;
            out[i].size = tmp[i].size; out[i].pori = tmp[i].pori;
        }
    }
    if (k.n_oriented > capacity) return fail(c, SIFT_B200_E_CAPACITY, "%d keypoints, output capacity %d", k.n_oriented, capacity);
    return SIFT_B200_OK;
}

int sift_b200_debug_describe(sift_b200_ctx* c, sift_b200_keypoint* inout, int n) {
    if (!c || n < 0 || (n > 0 && !inout)) return SIFT_B200_E_INVALID;
    CU(c, cudaSetDevice(c->device));
    int rc = debug_stage_ready(c);
    if (rc) return rc;
    if (n > c->cap_oriented) return fail(c, SIFT_B200_E_CAPACITY, "%d keypoints, capacity %d", n, c->cap_oriented);
    for (int i = 0; i < n; ++i)
        if (inout[i].octave < 0 || inout[i].octave >= c->stats.octaves || inout[i].layer < 0 ||
            inout[i].layer >= c->layers)
            return fail(c, SIFT_B200_E_INVALID, "keypoint %d: octave / layer outside the scale space", i);
    if (n == 0) return SIFT_B200_OK;
    std::vector<KpCore> tmp(n);
    std::vector<int> order(n);
    for (int i = 0; i < n; ++i) {
        tmp[i] = KpCore{inout[i].x, inout[i].y, inout[i].octave, inout[i].layer, inout[i].size, inout[i].pori};
        order[i] = i;
    }
    Counters k;
    memset(&k, 0, sizeof k);
    k.n_final = n;
    cudaStream_t s = c->stream;
    c->have_result = false;
    CU(c, cudaMemcpyAsync(c->d_oriented, tmp.data(), (size_t)n * sizeof(KpCore), cudaMemcpyHostToDevice, s));
    CU(c, cudaMemcpyAsync(c->ss.final_order, order.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, s));
    CU(c, cudaMemcpyAsync(c->d_counters, &k, sizeof k, cudaMemcpyHostToDevice, s));
    CU(c, launch_describe(c->d_pyr, c->d_oriented, c->ss.final_order, c->d_counters, c->d_records, c->d_desc,
                          c->cap_oriented, c->sp, c->sm_count, s));
    CU(c, cudaMemcpyAsync(inout, c->d_records, (size_t)n * sizeof(sift_b200_keypoint), cudaMemcpyDeviceToHost, s));
    CU(c, cudaStreamSynchronize(s));
    c->launches += 1;
    return SIFT_B200_OK;
}

// Canary audit: sift_b200_debug_canary_arm fills the whole scale-space arena with a NaN pattern; after the next
// detect call sift_b200_debug_canary_check counts (a) elements the pipeline had no business writing -- row padding
// (w <= x < pitch) of every plane, planes that stay on chip (G[4], G[5] unless kept), the arena behind the last plane
// -- that no longer hold the pattern, and (b) elements inside the planes it must write that still hold it.
static const unsigned kCanary = 0x7fc0beefu;   // a quiet NaN no kernel produces

int sift_b200_debug_canary_arm(sift_b200_ctx* c) {
    if (!c) return SIFT_B200_E_INVALID;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaStreamSynchronize(c->stream));
    CU(c, launch_canary_fill(c->arena, c->arena_floats, kCanary, c->sm_count, c->stream));
    c->pyramid_valid = false;
    c->launches += 1;
    return SIFT_B200_OK;
}

int sift_b200_debug_canary_check(sift_b200_ctx* c, int64_t* stray_writes, int64_t* missing_writes) {
    if (!c || !stray_writes || !missing_writes) return SIFT_B200_E_INVALID;
    CU(c, cudaSetDevice(c->device));
    int rc = debug_stage_ready(c);
    if (rc) return rc;
    unsigned long long* d_counts = nullptr;
    CU(c, cudaMalloc(&d_counts, 2 * sizeof(unsigned long long)));
    cudaStream_t s = c->stream;
    cudaMemsetAsync(d_counts, 0, 2 * sizeof(unsigned long long), s);
    const float* end = c->arena;
    const bool fused = c->sp.intervals == 3 && !c->force_unfused;
    for (int o = 0; o < c->pyr.octaves; ++o) {
        const OctaveDesc& od = c->pyr.oct[o];
        for (int i = 0; i < c->layers; ++i) {
            // the fused cascades keep the last two Gaussian levels on chip (unless the debug planes are on); the
            // per-level path writes every level; octave 0's last level is the input stage's scratch plane there
            const bool on_chip = fused && !c->keep_planes && i >= c->layers - 2;
            const bool scratch = o == 0 && i == c->layers - 1 && !fused;
            launch_canary_plane(od.G[i], od.w, od.h, od.pitch, kCanary, (on_chip || scratch) ? (scratch ? 1 : 0) : 1, d_counts, c->sm_count, s);
            end = std::max(end, (const float*)(od.G[i] + (size_t)od.pitch * od.h));
        }
        for (int i = 0; i < c->dogs; ++i) {
            launch_canary_plane(od.D[i], od.w, od.h, od.pitch, kCanary, 1, d_counts, c->sm_count, s);
            end = std::max(end, (const float*)(od.D[i] + (size_t)od.pitch * od.h));
        }
    }
    const size_t used = (size_t)(end - c->arena);
    if (used < c->arena_floats)
        launch_canary_tail(end, c->arena_floats - used, kCanary, d_counts, c->sm_count, s);
    unsigned long long h[2] = {0, 0};
    cudaError_t e = cudaMemcpyAsync(h, d_counts, sizeof h, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(d_counts);
    if (e != cudaSuccess) return fail(c, SIFT_B200_E_CUDA, "canary check failed: %s", cudaGetErrorString(e));
    *stray_writes = (int64_t)h[0];
    *missing_writes = (int64_t)h[1];
    return SIFT_B200_OK;
}

int sift_b200_debug_launch_plan(sift_b200_ctx* c, int use_graph, int centred, int extrema_form) {
    if (!c) return SIFT_B200_E_INVALID;
    if (use_graph >= 0) c->use_graph = use_graph != 0;
    if (centred >= 0) c->centred = centred != 0;
    if (extrema_form >= 0) c->extrema_form = extrema_form;
    return SIFT_B200_OK;
}

int sift_b200_debug_tail(sift_b200_ctx* c, int one_launch) {
    if (!c) return SIFT_B200_E_INVALID;
    c->use_tail = one_launch != 0;
    return SIFT_B200_OK;
}

long sift_b200_graphs_built(const sift_b200_ctx* c) { return c ? c->graphs_built : 0; }

long sift_b200_launch_count(const sift_b200_ctx* c) { return c ? c->launches : 0; }

}  // extern "C"

#include "api_collection.inc"

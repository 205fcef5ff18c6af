/* sift_b200.h -- C ABI of the B200-native SIFT engine.
 *
 * This is the drop-in boundary for the reference's hot path, the two free functions declared
 * in the reference's src/sift.hh:
 *
 *   detect_keypoints_and_descriptors(const Image&, ...)      sift.hh:65-71, sift.cpp:712-776
 *   match_keypoints(const vector<Keypoint>&, ..., ratio)      sift.hh:73-75, sift.cpp:783-815
 *
 * The reference has no FFI of its own (it is one C++ executable); a C++ shim with the exact
 * sift.hh signatures (sift_project_b200/shim/sift_shim.cpp) sits on top of this ABI so that the
 * reference's main.cpp relinks unchanged -- see INTEGRATION.md.
 *
 * Conventions: plain pointers and sizes, no C++/torch types; every call returns a status code
 * and never throws; sift_b200_last_error() gives the text.
 *
 * Parity with the reference is NEAR-exact, not bit-exact: the reference computes in FP64 (image_io.hh:26 stores
 * pixels as double), this build keeps the scale space in FP32 (an HBM-bound stage) and uses FP64 for every
 * per-keypoint scalar.  What the tests enforce (tests/test_gpu_parity.py): keypoint recall / precision >= 99.5 % at
 * 0.01 px / 1e-3 relative size; the descriptor STAGE within 1 quantisation level of the reference's on identical
 * keypoints; end-to-end descriptor differences above 1 level only where the keypoint's own orientation / offset
 * differs within those tolerances (attributed case by case); matcher output identical (integer-exact kernels).
 *
 * Streams: every context enqueues on its own non-blocking CUDA stream (sift_b200_stream()).  DEVICE pointers handed
 * to a call are read / written in that stream's order: the caller must make the stream wait for whatever produced
 * them (cudaStreamWaitEvent on sift_b200_stream(), or a synchronisation) and must not touch outputs before
 * sift_b200_sync() / the stream's completion.  All compute runs in hand-written
 * sm_100a CUDA kernels; there is NO CPU fallback: without a usable CUDA device
 * sift_b200_create() fails with SIFT_B200_E_NO_DEVICE.
 */
#ifndef SIFT_B200_H
#define SIFT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SIFT_B200_OK 0
#define SIFT_B200_E_INVALID 1     /* bad argument (null pointer, non 1/3 channels, tiny image) */
#define SIFT_B200_E_NO_DEVICE 2   /* no CUDA device / wrong architecture */
#define SIFT_B200_E_CUDA 3        /* a CUDA runtime call or kernel failed */
#define SIFT_B200_E_CAPACITY 4    /* an output or internal list overflowed its capacity */
#define SIFT_B200_E_UNSUPPORTED 5 /* parameter combination outside this build (see params) */
#define SIFT_B200_E_TOO_LARGE 6   /* image larger than the context was created for */

/* Byte-for-byte the reference's struct Keypoint (sift.hh:15-23): 168 bytes.
 * x, y, size are in INPUT-image pixels when double_image_size is on (sift.cpp:520-527);
 * octave is the pyramid index (octave 0 = the doubled image); pori in [0, 2 pi). */
typedef struct sift_b200_keypoint {
    double x, y;
    int32_t octave, layer;
    double size, pori;
    uint8_t desc[128];
} sift_b200_keypoint;

/* The arguments of detect_keypoints_and_descriptors (sift.hh:65-71), same meaning and defaults.
 * This build implements intervals in 2..5, odd window_size in 3..7 (as long as a DoG layer is left
 * to test; the fit itself stays 3x3x3 like the reference's get_pixel_cube, sift.cpp:35-37) and
 * num_bins in 4..128; other values, and init_sigma / intervals combinations whose blur radius
 * exceeds 16, return SIFT_B200_E_UNSUPPORTED.  The fused per-octave kernels serve the default sigmas; other settings
 * run one blur kernel per level (same arithmetic).  max_octaves is an extension: 0 = derive as the reference does
 * (sift.cpp:132-137), n > 0 = stop after n octaves. */
typedef struct sift_b200_params {
    int32_t double_image_size;  /* 1 */
    double init_sigma;          /* 1.6 */
    int32_t intervals;          /* 3 */
    int32_t window_size;        /* 3 */
    double contrast_threshold;  /* 0.04 */
    double eigen_ratio;         /* 10 */
    double num_bins;            /* 36 */
    double peak_ratio;          /* 0.8 */
    double ori_sigma_factor;    /* 1.5 */
    double desc_scale_factor;   /* 3.0 */
    int32_t max_octaves;        /* 0 (extension) */
} sift_b200_params;

typedef struct sift_b200_ctx sift_b200_ctx;

/* Per-stage counts of the last detect call, mirroring the counts the reference prints
 * (sift.cpp:746, 752, 758, 763). */
typedef struct sift_b200_stats {
    int32_t octaves;
    int32_t extrema;
    int32_t raw_keypoints;
    int32_t oriented_keypoints;
    int32_t final_keypoints;
    int32_t base_width, base_height;
} sift_b200_stats;

void sift_b200_default_params(sift_b200_params* p);

/* One context = one GPU + one stream + one workspace sized for images up to max_width x
 * max_height (before doubling).  Not thread-safe; use one context per host thread. */
int sift_b200_create(int device, int max_width, int max_height, sift_b200_ctx** out);
void sift_b200_destroy(sift_b200_ctx* ctx);
const char* sift_b200_last_error(const sift_b200_ctx* ctx); /* ctx may be NULL: create errors */

/* detect_keypoints_and_descriptors (sift.cpp:712-776).
 * pixels: row-major, interleaved, channels = 1 (gray) or 3 (RGB), values 0..255; HOST or DEVICE
 * memory (detected).  The u8 entry point loses nothing on the way in for file-loaded images (image_io.cpp:27-33);
 * the f32 one accepts arbitrary finite values of any range (its min / max are reduced on the GPU to
 * scale the fixed-point histograms; note that contrast_threshold assumes a 0..255 scale, sift.cpp:305).  out: HOST array of `capacity` records, filled in the
 * reference's order (sorted by Keypoint::operator<, sift.hh:31-41, duplicates removed).
 * *count receives the number found even when it exceeds capacity (status E_CAPACITY then). */
int sift_b200_detect_u8(sift_b200_ctx* ctx, const uint8_t* pixels, int width, int height,
                        int channels, const sift_b200_params* params, sift_b200_keypoint* out,
                        int capacity, int* count);
int sift_b200_detect_f32(sift_b200_ctx* ctx, const float* pixels, int width, int height,
                         int channels, const sift_b200_params* params, sift_b200_keypoint* out,
                         int capacity, int* count);

/* Page-locked host memory for inputs and outputs (cudaHostAlloc): copies from / to it are truly asynchronous and
 * overlap the kernels of other contexts (image_io.cpp:20-35 decodes into pageable memory; a caller that wants the
 * input copy off the critical path decodes or converts into one of these instead).  Usable from any context. */
int sift_b200_host_alloc(size_t bytes, void** out);
void sift_b200_host_free(void* p);

/* Asynchronous variant: enqueue the whole pipeline on the context's stream and leave the
 * results on the GPU; no host synchronisation happens.  pixels may be DEVICE memory (used in
 * place) or HOST memory (copied with cudaMemcpyAsync first -- pin it for a truly asynchronous
 * copy; it must stay valid until the stream has consumed it). */
int sift_b200_detect_enqueue_u8(sift_b200_ctx* ctx, const uint8_t* pixels, int width, int height,
                                int channels, const sift_b200_params* params);
/* Wait for the enqueued work; returns the number of final keypoints in *count. */
int sift_b200_detect_finish(sift_b200_ctx* ctx, int* count);
/* Wait, then copy the records of the last detect to a HOST array (same contract as
 * sift_b200_detect_u8's out / capacity / count). */
int sift_b200_result_copy(sift_b200_ctx* ctx, sift_b200_keypoint* out, int capacity, int* count);
/* Device pointers to the results of the last detect: `n` 168-byte records and the dense
 * n x 128 u8 descriptor matrix (row i = record i's desc), valid until the next detect. */
int sift_b200_result_device(sift_b200_ctx* ctx, const sift_b200_keypoint** d_records,
                            const uint8_t** d_descriptors, int* n);
int sift_b200_get_stats(sift_b200_ctx* ctx, sift_b200_stats* stats);

/* match_keypoints (sift.cpp:783-815) on dense descriptor matrices (n x 128 u8, row-major; HOST
 * or DEVICE, detected).  For every row i of A in ascending order: best and second-best
 * Euclidean distance over all rows of B (lowest j wins ties), emitted iff
 * best < ratio * second.  Outputs are HOST arrays of `capacity` entries (dist = sqrt of the
 * integer squared distance, as the reference stores it). */
int sift_b200_match(sift_b200_ctx* ctx, const uint8_t* desc_a, int na, const uint8_t* desc_b,
                    int nb, double ratio, int32_t* idx_a, int32_t* idx_b, double* dist,
                    int capacity, int* count);
/* Device-resident variant: fills, for every row of A, the index of its nearest row of B and the
 * two smallest SQUARED distances (exact integers).  All pointers are DEVICE memory; enqueued on
 * the context's stream, no host synchronisation.  nb == 0 leaves best_idx = -1. */
int sift_b200_match_enqueue(sift_b200_ctx* ctx, const uint8_t* d_desc_a, int na,
                            const uint8_t* d_desc_b, int nb, int32_t* d_best_idx,
                            int32_t* d_best_d2, int32_t* d_second_d2);
int sift_b200_sync(sift_b200_ctx* ctx);
/* The CUDA stream (cudaStream_t) the context enqueues on, for event timing by the caller. */
void* sift_b200_stream(sift_b200_ctx* ctx);
/* Which matcher the next match call will use for these sizes: 1 = tcgen05 tensor-core kernel,
 * 0 = SIMT dp4a kernel (small problems). */
int sift_b200_match_path(int na, int nb);

/* ---- batch and multi-GPU --------------------------------------------------------------
 * The reference drives one image pair from one thread (main.cpp:12-18).  A batch or a stitching collection repeats
 * detect per image and match_keypoints per image pair; images are independent (no collective), matching needs one
 * exchange step.  Everything below is driven by the host thread(s) of the caller: one process with several
 * contexts (one per GPU, or several per GPU), or one process per GPU (MPI / torchrun style). */

/* Batch detect: image k goes to ctxs[k % n_ctx] (contexts of one GPU = images in flight on it; contexts of
 * different GPUs = the batch sharded by image, config 3).  images[k]: width x height x channels u8, HOST (pinned for
 * asynchronous copies) or DEVICE memory of that context's GPU.  outs[k] / capacities[k] / counts[k]: as in
 * sift_b200_detect_u8.  Returns the first non-OK status (all images are still attempted). */
int sift_b200_detect_batch_u8(sift_b200_ctx* const* ctxs, int n_ctx, const uint8_t* const* images, int n_images,
                              int width, int height, int channels, const sift_b200_params* params,
                              sift_b200_keypoint* const* outs, const int32_t* capacities, int32_t* counts);

/* NCCL communicator (libnccl.so.2 is loaded on first use).  One process per GPU: rank 0 calls comm_unique_id, ships
 * the bytes to every rank by any means, and every rank calls comm_attach (collective).  One process driving several
 * GPUs: comm_attach_all on its contexts (rank = position in the array).  world == 1 needs no NCCL. */
#define SIFT_B200_UNIQUE_ID_BYTES 128
int sift_b200_comm_unique_id(uint8_t* id_out /* SIFT_B200_UNIQUE_ID_BYTES */);
int sift_b200_comm_attach(sift_b200_ctx* ctx, const uint8_t* id, int world, int rank);
int sift_b200_comm_attach_all(sift_b200_ctx* const* ctxs, int n);
int sift_b200_comm_info(const sift_b200_ctx* ctx, int* world, int* rank);

/* Collection matching (config 5): match_keypoints (sift.cpp:783-815) for every unordered image pair i < j (and
 * j -> i as well when both_directions), image i owned by rank i % world.  local_desc / local_counts: the n x 128 u8
 * descriptor matrices (HOST or DEVICE) and row counts of the images THIS rank owns, in ascending image index.
 * Collective: an all-reduce of the counts, then ONE ncclAllGather of the packed descriptor blocks on the context's
 * side stream while the pairs whose images are both local already run; pairs are dealt to ranks by n_i * n_j
 * weight (sift_b200_partition_pairs), never split.  Results (per row of image i: nearest row of image j and the two
 * smallest squared distances) stay on the GPU; *n_pairs = pairs this rank owns.  Enqueued, not synchronised. */
int sift_b200_collection_match(sift_b200_ctx* ctx, int n_images, const uint8_t* const* local_desc,
                               const int32_t* local_counts, int both_directions, int* n_pairs);
/* The same for one process driving every rank: desc[i] / counts[i] for ALL images, desc[i] reachable from the GPU
 * of ctxs[i % n] (or HOST). */
int sift_b200_collection_match_all(sift_b200_ctx* const* ctxs, int n, int n_images, const uint8_t* const* desc,
                                   const int32_t* counts, int both_directions);
/* The pairs this rank owns, in result order: (i, j, rows of image i). */
int sift_b200_collection_pairs(sift_b200_ctx* ctx, int32_t* pair_i, int32_t* pair_j, int32_t* rows, int capacity,
                               int* count);
/* Matches of owned pair `pair` after the ratio test (HOST outputs, same contract as sift_b200_match). */
int sift_b200_collection_fetch(sift_b200_ctx* ctx, int pair, double ratio, int32_t* idx_a, int32_t* idx_b,
                               double* dist, int capacity, int* count);
/* Device pointers to owned pair `pair`'s per-row results (valid until the next collection call). */
int sift_b200_collection_device(sift_b200_ctx* ctx, int pair, const int32_t** d_best_idx, const int32_t** d_best_d2,
                                const int32_t** d_second_d2, int* rows);
/* Digest of this rank's results -- number of matches under `ratio` and an order-free 64-bit checksum over
 * (pair, row, nearest index, d1^2, d2^2) -- summed over every rank when all_ranks (collective then): any number of
 * GPUs must report the same two numbers for the same collection. */
int sift_b200_collection_digest(sift_b200_ctx* ctx, double ratio, int all_ranks, int64_t* n_matches, uint64_t* hash);
/* The deal itself (pure host code, no GPU): pairs of `rank` among `world`, longest-processing-time on n_i * n_j. */
int sift_b200_partition_pairs(const int32_t* counts, int n_images, int world, int both_directions, int rank,
                              int32_t* pair_i, int32_t* pair_j, int capacity, int* count);

/* ---- introspection used by the parity tests (stage-by-stage comparison with the oracle) ---- */
#define SIFT_B200_PLANE_GAUSSIAN 0 /* layer 0..5 */
#define SIFT_B200_PLANE_DOG 1      /* layer 0..4 */
/* keep_all_planes: also store Gaussian layers 4 and 5 (the fused octave kernel keeps them on chip
 * otherwise; layers 0..3 and all DoG layers are always stored).  unfused_pyramid selects the scale-space
 * kernels (all give bit-identical planes): 0 = default (fused per-octave cascade: streaming kernels on large
 * octaves, tile kernels on small ones), 1 = one kernel per level, 2 = fused tile kernels on every octave,
 * 3 = fused streaming kernels on every octave. */
int sift_b200_debug_options(sift_b200_ctx* ctx, int keep_all_planes, int unfused_pyramid);
int sift_b200_debug_plane_dims(sift_b200_ctx* ctx, int octave, int* width, int* height);
int sift_b200_debug_plane(sift_b200_ctx* ctx, int kind, int octave, int layer, float* host_out);
/* rows of (x, y, layer, octave) int32; order is NOT the reference's emission order */
int sift_b200_debug_extrema(sift_b200_ctx* ctx, int32_t* host_out, int capacity, int* count);
/* stage 0 = raw (after refine, doubled-image frame), 1 = oriented (before sort/dedup) */
int sift_b200_debug_keypoints(sift_b200_ctx* ctx, int stage, sift_b200_keypoint* host_out,
                              int capacity, int* count);
/* Single stages on CALLER-SUPPLIED keypoints (HOST arrays) over the scale space of the last detect call, so that a
 * parity test can attribute a difference to the stage that produced it.  debug_orient: compute_orientations
 * (sift.cpp:447-533) on raw keypoints in the frame debug_keypoints(stage 0) returns them in; one output per
 * histogram peak, unordered.  debug_describe: compute_descriptors (sift.cpp:610-682) on oriented keypoints (output
 * frame), in place.  Both overwrite the lists of the last detect (its records can no longer be fetched). */
int sift_b200_debug_orient(sift_b200_ctx* ctx, const sift_b200_keypoint* raw_in, int n, sift_b200_keypoint* out,
                           int capacity, int* count);
int sift_b200_debug_describe(sift_b200_ctx* ctx, sift_b200_keypoint* inout, int n);
/* Write audit of the scale-space kernels (there is no sanitizer on the GPU pool): arm fills the whole arena with a
 * NaN pattern; after the NEXT detect call, check returns how many elements the pipeline must not write (row padding
 * of every plane, planes kept on chip, the arena behind the last plane) lost the pattern, and how many elements it
 * must write still hold it.  Both must be 0. */
int sift_b200_debug_canary_arm(sift_b200_ctx* ctx);
int sift_b200_debug_canary_check(sift_b200_ctx* ctx, int64_t* stray_writes, int64_t* missing_writes);
/* Launch plan switches (each: 0 / 1, or -1 to leave unchanged).  use_graph (default 1; env SIFT_B200_GRAPH): the
 * stages after the input kernel are captured once per (image size, parameters) into a CUDA graph -- octave chain on
 * one branch, second cascade kernel + extrema of each octave on another -- and replayed with one launch per image;
 * 0 = plain launches on one stream.  centred (default 1; env SIFT_B200_CENTER): the scale space is stored relative
 * to the input's mid level (u8: 128; float: the midpoint of its range), which makes the FP32 rounding steps 2-4x
 * finer; every consumer takes differences, so results change only by less rounding.  extrema_form (default 0;
 * env SIFT_B200_EXTREMA): 0 = the four-columns-per-lane extrema kernel, 1 = the one-column form (same set). */
int sift_b200_debug_launch_plan(sift_b200_ctx* ctx, int use_graph, int centred, int extrema_form);
/* one_launch = 1 (default; env SIFT_B200_TAIL): the octaves of at most one 32 x 32 tile per SM -- seven of the ten
 * octaves of a 4K image, latency-bound one by one -- run their cascades in ONE ticket-ordered launch and their extrema
 * scans in another; 0 = two cascade launches + one scan per octave as for the large octaves.  Same bytes. */
int sift_b200_debug_tail(sift_b200_ctx* ctx, int one_launch);
/* how many CUDA graphs this context has captured so far (one per distinct image size / parameter set in a row) */
long sift_b200_graphs_built(const sift_b200_ctx* ctx);
/* number of kernel launches issued by this context since creation (bench's gpu_launches) */
long sift_b200_launch_count(const sift_b200_ctx* ctx);

/* ---- per-stage device timing (CUDA events on the context's stream) ---- */
#define SIFT_B200_STAGE_INPUT 0    /* gray + 2x upsample + initial blur (sift.cpp:113-126) */
#define SIFT_B200_STAGE_PYRAMID 1  /* Gaussian cascade + DoG + decimation (sift.cpp:181-225) */
#define SIFT_B200_STAGE_EXTREMA 2  /* sift.cpp:264-319 */
#define SIFT_B200_STAGE_REFINE 3   /* sift.cpp:330-436 */
#define SIFT_B200_STAGE_ORIENT 4   /* sift.cpp:447-533 */
#define SIFT_B200_STAGE_SORT 5     /* sift.cpp:20-24 */
#define SIFT_B200_STAGE_DESCRIBE 6 /* sift.cpp:610-682 */
#define SIFT_B200_STAGE_COUNT 7
/* When on, every detect call brackets its stages with events (a few microseconds each). */
int sift_b200_set_profiling(sift_b200_ctx* ctx, int on);
/* Milliseconds and kernel-launch counts per stage of the last profiled detect (arrays of
 * SIFT_B200_STAGE_COUNT); waits for the stream. */
int sift_b200_get_profile(sift_b200_ctx* ctx, float* stage_ms, int32_t* stage_launches);
/* The same profile unaggregated: one (stage, milliseconds) entry per bracketed group of launches, in issue
 * order (input; octave 0 first pyramid kernel, octave 0 second pyramid kernel, octave 0 extrema; octave 1
 * pyramid (both kernels), octave 1 extrema; ...; refine; ...).  *count = entries available; at most `capacity` are written. */
int sift_b200_get_profile_marks(sift_b200_ctx* ctx, int32_t* stage, float* ms, int capacity, int* count);

#ifdef __cplusplus
}
#endif
#endif /* SIFT_B200_H */

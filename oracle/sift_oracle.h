/* TEST INFRASTRUCTURE ONLY -- CPU oracle for the SIFT hot path.
 *
 * A plain FP64, single-threaded restatement of the algorithm in the reference's
 * src/sift.cpp + src/image.cpp, written from the algorithm description, with
 * every function citing the reference file:line it follows.  It is PINNED: tests/test_oracle.py
 * checks it bit-for-bit against the real reference compiled into oracle/_ref/libsift_ref.so and
 * against the golden vectors under tests/golden/ (which were produced by that real reference).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  libsift_b200.so never links or calls it; the product has no CPU path.
 */
#ifndef SIFT_ORACLE_H
#define SIFT_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same 168-byte layout as the reference's struct Keypoint (sift.hh:15-23). */
typedef struct OracleKeypoint {
    double x, y;
    int32_t octave, layer;
    double size, pori;
    uint8_t desc[128];
} OracleKeypoint;

/* The arguments of detect_keypoints_and_descriptors (sift.hh:65-71). */
typedef struct OracleParams {
    int32_t double_image_size;
    double init_sigma;
    int32_t intervals;
    double contrast_threshold, eigen_ratio, peak_ratio, ori_sigma_factor, desc_scale_factor;
    int32_t window_size;
    double num_bins;
} OracleParams;
void oracle_default_params(OracleParams* p);

typedef struct OracleRun OracleRun;

/* Runs every stage of detect_keypoints_and_descriptors (sift.cpp:712-776) with the reference's
 * default parameters (sift.hh:65-71) and keeps the intermediate results.
 * pixels: interleaved row-major doubles 0..255, c = 1 or 3 (image_io.cpp:81-83). */
OracleRun* oracle_run_create(const double* pixels, int w, int h, int c, int double_image_size,
                             int keep_pyramid);
OracleRun* oracle_run_create_ex(const double* pixels, int w, int h, int c, const OracleParams* params,
                                int keep_pyramid);
void oracle_run_destroy(OracleRun* r);
int oracle_run_octaves(const OracleRun* r);
int oracle_run_sigmas(const OracleRun* r, double* out, int cap);
int oracle_run_layer_dims(const OracleRun* r, int octave, int* w, int* h);
const double* oracle_run_gaussian(const OracleRun* r, int octave, int layer); /* layer 0..5 */
const double* oracle_run_dog(const OracleRun* r, int octave, int layer);      /* layer 0..4 */
int oracle_run_extrema(const OracleRun* r, double* out_xyzo, int cap);
/* stage 0 = raw (after refine), 1 = oriented, 2 = final (sorted, deduplicated, described) */
int oracle_run_keypoints(const OracleRun* r, int stage, OracleKeypoint* out, int cap);

/* The per-keypoint stages on caller-supplied keypoints over this run's pyramid (keep_pyramid must be on):
 * compute_orientations (sift.cpp:447-533) on raw keypoints, compute_descriptors (sift.cpp:610-682) on
 * oriented keypoints (in place).  Return the output count, -1 without a pyramid. */
int oracle_run_orient_given(OracleRun* r, const OracleKeypoint* in, int n, OracleKeypoint* out, int cap);
int oracle_run_describe_given(OracleRun* r, OracleKeypoint* inout, int n);

/* match_keypoints (sift.cpp:783-815) on raw 128-byte descriptors; returns the match count. */
int oracle_match(const uint8_t* desc_a, int na, const uint8_t* desc_b, int nb, double ratio,
                 int* idx_a, int* idx_b, double* dist, int cap);

/* Stand-alone pieces, exposed for unit tests. */
int oracle_gaussian_taps(double sigma, double* taps, int cap); /* image.cpp:226-235 */
void oracle_blur(const double* in, int w, int h, double sigma, double* out); /* image.cpp:156-238 */

#ifdef __cplusplus
}
#endif
#endif

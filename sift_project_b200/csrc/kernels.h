// Host-callable launchers of the sm_100a kernels (internal to libsift_b200.so).
#pragma once
#include "common.cuh"

namespace sb {

// pyramid.cu
cudaError_t pyramid_init();
// centre: the value subtracted from every pixel on the way in (the scale space is linear and DC-preserving, and every
// consumer takes differences, so the result is the same with 2-4x smaller FP32 rounding steps).  u8: a host
// constant; f32: the midpoint of the device-side (min, max) in `range` when centred, else 0.
cudaError_t launch_prepare_u8(const uint8_t* src, int sw, int sh, int ch, float* dst, int dw, int dh,
                              int dpitch, int doubled, float centre, cudaStream_t s);
cudaError_t launch_prepare_f32(const float* src, int sw, int sh, int ch, float* dst, int dw, int dh,
                               int dpitch, int doubled, const float* centre_range, cudaStream_t s);
cudaError_t launch_blur(const float* in, float* out, float* dog, float* dec, int w, int h, int pitch,
                        int dec_w, int dec_h, int dec_pitch, const BlurTaps& taps, cudaStream_t s);

bool input_fused_supported(int channels, const BlurTaps& taps);
cudaError_t launch_input_u8(const uint8_t* src, int sw, int sh, int channels, float* dst, int w, int h, int pitch,
                            int doubled, const BlurTaps& taps, float centre, cudaStream_t s);
bool cascade_supported(const BlurTaps* taps);
cudaError_t launch_octave_fused(const OctaveDesc& od, const BlurTaps* taps, float* dec, int dec_w, int dec_h,
                                int dec_pitch, bool keep_all, int sm_count, int mode, int part, cudaStream_t s);
// the latency-bound small octaves (at most one 32 x 32 tile per SM) in one launch; sync = 1 + kMaxOctaves zeroed ints
bool tail_eligible(const OctaveDesc& od, int sm_count);
cudaError_t launch_tail(const OctaveDesc* octs, int first, int octaves, const BlurTaps* taps, bool keep_all, int* sync,
                        int sm_count, cudaStream_t s);

// detect.cu
struct SortScratch {
    int nb;             // number of x buckets (= output-frame image width)
    int* bucket_cnt;    // [nb + 1]
    int* bucket_off;    // [nb + 1]
    int* bucket_fill;   // [nb]
    int* uniq_cnt;      // [nb + 1]
    int* uniq_off;      // [nb + 1]
    int* perm;          // [cap_oriented]
    int* tmp_sorted;    // [cap_oriented]
    int* sorted;        // [cap_oriented]
    int* final_order;   // [cap_oriented]
};
cudaError_t launch_range(const float* px, size_t n, float* range, int sm_count, cudaStream_t s);
// canary audit (debug): fill / verify a NaN pattern around what the scale-space kernels are allowed to write
cudaError_t launch_canary_fill(void* p, size_t words, unsigned pattern, int sm_count, cudaStream_t s);
cudaError_t launch_canary_plane(const float* plane, int w, int h, int pitch, unsigned pattern, int expect_written,
                                unsigned long long* counts, int sm_count, cudaStream_t s);
cudaError_t launch_canary_tail(const void* p, size_t words, unsigned pattern, unsigned long long* counts, int sm_count,
                               cudaStream_t s);
// cubes / cube_cap: the scan also writes the fit's first 3x3x3 neighbourhood of candidate slot < cube_cap (only the
// kernels for which extrema_hands_cubes() is true; pass the same cube_cap to launch_refine, 0 otherwise)
bool extrema_hands_cubes(int border, int form);
cudaError_t launch_extrema(const OctaveDesc& oct, int octave, int dogs, int border, int threshold, Cand* cands,
                           int cap, Counters* counters, CandCube* cubes, int cube_cap, int form, cudaStream_t s);
bool extrema_multi_supported(int border, int form);
cudaError_t launch_extrema_multi(const OctaveDesc* octs, int first, int n, int dogs, int threshold, Cand* cands, int cap,
                                 Counters* counters, CandCube* cubes, int cube_cap, cudaStream_t s);
cudaError_t launch_refine(const PyramidDesc* d_pyr, const Cand* cands, KpCore* raw, Counters* counters,
                          const CandCube* cubes, int cube_cap, const StageParams& sp, int sm_count, cudaStream_t s);
cudaError_t launch_orient(const PyramidDesc* d_pyr, const KpCore* raw, KpCore* oriented,
                          Counters* counters, const StageParams& sp, int sm_count, cudaStream_t s);
cudaError_t launch_sort_dedup(const KpCore* oriented, Counters* counters, const SortScratch& ss,
                              const StageParams& sp, int sm_count, cudaStream_t s, int* launches);
cudaError_t launch_describe(const PyramidDesc* d_pyr, const KpCore* oriented, const int* final_order,
                            Counters* counters, uint8_t* records, uint8_t* desc, int cap_final,
                            const StageParams& sp, int sm_count, cudaStream_t s);

// match_simt.cu
struct MatchScratch {
    int* part_idx;   // [splits][na]
    int* part_d1;
    int* part_d2;
    int* norms_a;    // [na]
    int* norms_b;    // [nb]
    size_t cap_rows; // rows of A the partial buffers were sized for (x max splits)
    int max_splits;
};
cudaError_t launch_norms(const uint8_t* d, int n, int* norms, cudaStream_t s);
cudaError_t launch_match_simt(const uint8_t* a, int na, const uint8_t* b, int nb, int* best_idx,
                              int* best_d2, int* second_d2, const MatchScratch& ms, int sm_count,
                              cudaStream_t s, int* launches);
cudaError_t launch_match_merge(const MatchScratch& ms, int splits, int na, int* best_idx, int* best_d2,
                               int* second_d2, cudaStream_t s);
// match_tc.cu: tcgen05 / TMA / TMEM path
cudaError_t match_tc_init();
int match_tc_padded_rows(int nb);
cudaError_t launch_match_tc(const uint8_t* a, int na, const uint8_t* b, int nb, int* best_idx, int* best_d2,
                            int* second_d2, const MatchScratch& ms, int sm_count, cudaStream_t s, int* launches);
// match.cu: picks the tcgen05 kernel (match_tc.cu) for large problems, else the SIMT kernel
cudaError_t match_init();
bool match_uses_tensor_cores(int na, int nb);
cudaError_t launch_match(const uint8_t* a, int na, const uint8_t* b, int nb, int* best_idx, int* best_d2,
                         int* second_d2, const MatchScratch& ms, int sm_count, cudaStream_t s, int* launches);
cudaError_t launch_match_emit(const int* best_idx, const int* best_d2, const int* second_d2, int na,
                              int nb, double ratio, int* out_ia, int* out_ib, double* out_dist,
                              int cap, int* out_count, cudaStream_t s);

}  // namespace sb

// Shared device/host definitions for the B200 SIFT engine (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace sb {

constexpr int kLayers = 6;      // Gaussian images per octave at the default intervals = 3 (intervals + 3, sift.cpp:144)
constexpr int kDogs = 5;        // DoG images per octave at the default intervals (intervals + 2, sift.cpp:212)
constexpr int kMaxIntervals = 5;
constexpr int kMaxLayers = kMaxIntervals + 3;
constexpr int kMaxDogs = kMaxIntervals + 2;
constexpr int kMaxOctaves = 16;
constexpr int kMaxRadius = 16;  // largest blur half-width this build instantiates
constexpr int kOriBins = 36;    // sift.hh:69
constexpr double kTwoPi = 6.283185307179586;  // M_PI2, sift.hh:5
constexpr double kPi = 3.14159265358979323846;
constexpr float kFix = 4294967296.0f;           // 2^32 fixed-point scale for deterministic sums
constexpr double kUnfix = 1.0 / 4294967296.0;

// One octave of the scale space in HBM.  Every plane is FP32, row-major, `pitch` floats per row
// (pitch is a multiple of 32 floats = 128 B so that rows are 16-byte aligned for vector loads).
struct OctaveDesc {
    int w, h, pitch;
    float* G[kMaxLayers];  // G[0] is the octave base (sift.cpp:165); intervals + 3 are in use
    float* D[kMaxDogs];    // D[i] = G[i+1] - G[i] (sift.cpp:217); intervals + 2 are in use
};

struct PyramidDesc {
    int octaves;
    OctaveDesc oct[kMaxOctaves];
};

// A scale-space extremum candidate (sift.cpp:14 "Extrema" tuple).
struct Cand {
    int x, y, z, o;
};

// The 19 cells of a candidate's 3x3x3 DoG neighbourhood that the quadratic fit reads (sift.cpp:49-80: centre, 6 face
// and 12 edge neighbours, no corners), handed over by the extrema scan -- which has just read them -- so that the first
// fit of a candidate needs no scattered loads.  Order: see emit_cube in detect.cu.
struct CandCube {
    float v[20];
};

// The non-descriptor part of the reference's Keypoint (sift.hh:15-21), 40 bytes.
struct KpCore {
    double x, y;
    int octave, layer;
    double size, pori;
};

struct BlurTaps {
    int radius;
    float w[kMaxRadius + 1];  // normalised half kernel: w[0] centre, w[u] = tap at distance u
};

// Device-side counters of one detect call.
struct Counters {
    int n_extrema;
    int n_raw;
    int n_oriented;
    int n_final;
    int overflow;  // bit 0 extrema, bit 1 raw, bit 2 oriented
    int next_orient;    // work-stealing cursors of the warp-per-keypoint kernels
    int next_describe;
    int pad[1];
};

// Parameters of the per-keypoint stages, all FP64 like the reference's scalars.
struct StageParams {
    int doubled;
    int intervals;
    int dogs;                 // intervals + 2
    int num_bins;             // orientation histogram bins (declared double, used as int, sift.cpp:450)
    int border;               // window_size / 2 (sift.cpp:272, :336)
    // Bound of the gradient magnitude for the fixed-point histograms: sqrt(2) * (max - min) of the
    // input (blurs and the bilinear up-sampling stay inside the input's range).  8-bit input: 361
    // on the host; float input: range[0..1] = (min, max), reduced on the device before the pyramid.
    double mag_bound;
    const float* range;
    int dog_threshold;        // floor(0.5*ct/intervals*255) squeezed into an int (sift.cpp:266,305)
    double init_sigma;
    double contrast_threshold;
    double eigen_ratio;
    double peak_ratio;
    double ori_sigma_factor;
    double desc_scale_factor;
    int cap_extrema, cap_raw, cap_oriented;
};

}  // namespace sb

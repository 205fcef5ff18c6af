// Batch driver above the C ABI for callers that start from image FILES.
//
// The reference's driver loads each file with the Image(path) constructor (main.cpp:12-13 -> image_io.cpp:20-35:
// stb decode into pageable doubles) and only then detects, one image after the other.  At B200 speed the decode
// (tens of milliseconds per photo) and the host -> device copy are the critical path, not the GPU.  This driver
// keeps the reference's decoder (pixel-identical input) but pipelines it:
//
//   decode threads (Image(path), any number)  ->  page-locked u8 staging slots (sift_b200_host_alloc)
//       ->  sift_b200_detect_enqueue_u8 on a ring of contexts (asynchronous DMA + the whole GPU pipeline)
//       ->  sift_b200_result_copy when the context is needed again
//
// so that file k+1.. are being decoded and file k's pixels are crossing PCIe while the GPU works on file k-1.
// Results are identical to calling detect_keypoints_and_descriptors per image.
#pragma once
#include <string>
#include <vector>

#include "sift.hh"   // the REFERENCE's header: Keypoint, Image (found through -I)

struct SiftBatchOptions {
    std::vector<int> devices = {0};   // CUDA devices to use; images are dealt to contexts round-robin
    int contexts_per_device = 3;      // images in flight per GPU
    int decode_threads = 4;           // host threads running the reference's decoder
    bool double_image_size = true;    // the arguments of sift.hh:65-71
    double init_sigma = 1.6;
    int intervals = 3;
    int window_size = 3;
    double contrast_threshold = 0.04;
    double eigen_ratio = 10.0;
    double num_bins = 36;
    double peak_ratio = 0.8;
    double ori_sigma_factor = 1.5;
    double desc_scale_factor = 3.0;
};

struct SiftBatchTimes {
    double wall_s = 0.0;          // whole call
    double decode_cpu_s = 0.0;    // summed over the decode threads
};

// One std::vector<Keypoint> per path, in the order of `paths`.  Throws std::runtime_error like the reference
// (unreadable file: image_io.cpp:23-26; CUDA / capacity errors carry sift_b200_last_error()).
std::vector<std::vector<Keypoint>> detect_image_files(const std::vector<std::string>& paths,
                                                      const SiftBatchOptions& options = SiftBatchOptions(),
                                                      SiftBatchTimes* times = nullptr);

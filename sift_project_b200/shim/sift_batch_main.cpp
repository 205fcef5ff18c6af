// sift_batch <image files...>: detect every file through the overlapped decode -> pinned ring -> GPU pipeline
// (sift_batch.hh) and print, per file, the keypoint count and an FNV-1a hash of the 168-byte records, then the
// timing.  Environment: SIFT_BATCH_DEVICES=0,1,..  SIFT_BATCH_CONTEXTS=3  SIFT_BATCH_DECODERS=4
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <sstream>

#include "sift_batch.hh"

int main(int argc, char** argv) {
    if (argc < 2) {
        std::cerr << "usage: sift_batch <image> [<image> ...]" << std::endl;
        return 1;
    }
    SiftBatchOptions opt;
    if (const char* d = std::getenv("SIFT_BATCH_DEVICES")) {
        opt.devices.clear();
        std::stringstream ss(d);
        for (std::string tok; std::getline(ss, tok, ',');) opt.devices.push_back(std::atoi(tok.c_str()));
    }
    if (const char* v = std::getenv("SIFT_BATCH_CONTEXTS")) opt.contexts_per_device = std::atoi(v);
    if (const char* v = std::getenv("SIFT_BATCH_DECODERS")) opt.decode_threads = std::atoi(v);
    std::vector<std::string> paths(argv + 1, argv + argc);
    try {
        SiftBatchTimes t;
        const auto res = detect_image_files(paths, opt, &t);
        for (size_t k = 0; k < res.size(); ++k) {
            unsigned long long h = 1469598103934665603ull;
            const unsigned char* p = reinterpret_cast<const unsigned char*>(res[k].data());
            for (size_t i = 0; i < res[k].size() * sizeof(Keypoint); ++i) h = (h ^ p[i]) * 1099511628211ull;
            std::printf("%s: %zu keypoints, records %016llx\n", paths[k].c_str(), res[k].size(), h);
        }
        std::printf("batch: %zu files in %.3f s wall (%.1f files/s), decode %.3f s of CPU over %d threads, %d contexts on %zu GPU(s)\n",
                    res.size(), t.wall_s, res.size() / t.wall_s, t.decode_cpu_s, opt.decode_threads,
                    opt.contexts_per_device, opt.devices.size());
    } catch (const std::exception& e) {
        std::cerr << "sift_batch: " << e.what() << std::endl;
        return 2;
    }
    return 0;
}

// Drop-in replacement for the reference's src/sift.cpp: this translation unit defines the four
// functions declared in the reference's src/sift.hh (:65-81) on top of libsift_b200's C ABI, so
// that the reference's main.cpp, image.cpp and image_io.cpp link against it unchanged:
//
//   g++ -std=c++17 -I<reference>/src -Iinclude <reference>/src/{main,image,image_io}.cpp \
//       sift_project_b200/shim/sift_shim.cpp -Lsift_project_b200 -lsift_b200 -o sift
//
// "sift.hh" below is the REFERENCE's own header (found through -I); none of its types are
// re-declared here.  Keypoint is layout-compatible with sift_b200_keypoint (checked below), so
// records move with memcpy.  Errors surface as std::runtime_error, like the reference's.
//
// Fidelity switches (environment):
//   SIFT_B200_QUIET=1         skip the per-stage std::cout lines (sift.cpp:719-773 prints them)
//   SIFT_B200_NO_KEYPOINTS_PNG=1  skip the ./keypoints.png side effect (sift.cpp:765-768)
//   SIFT_B200_DEVICE=n        CUDA device (default 0)
#include <algorithm>
#include <array>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "sift.hh"
#include "sift_b200.h"

static_assert(sizeof(Keypoint) == sizeof(sift_b200_keypoint), "Keypoint must be the 168-byte record");
static_assert(offsetof(Keypoint, octave) == offsetof(sift_b200_keypoint, octave), "layout");
static_assert(offsetof(Keypoint, size) == offsetof(sift_b200_keypoint, size), "layout");
static_assert(offsetof(Keypoint, desc) == offsetof(sift_b200_keypoint, desc), "layout");

namespace {

bool env_on(const char* name) {
    const char* v = std::getenv(name);
    return v != nullptr && v[0] != '\0' && v[0] != '0';
}

int env_device() {
    const char* v = std::getenv("SIFT_B200_DEVICE");
    return v ? std::atoi(v) : 0;
}

struct CtxDeleter {
    void operator()(sift_b200_ctx* c) const { sift_b200_destroy(c); }
};

// One context per host thread, re-created only when a larger image arrives.
sift_b200_ctx* context_for(int w, int h) {
    thread_local std::unique_ptr<sift_b200_ctx, CtxDeleter> ctx;
    thread_local int cap_w = 0, cap_h = 0;
    if (!ctx || w > cap_w || h > cap_h) {
        ctx.reset();
        cap_w = std::max(w, cap_w);
        cap_h = std::max(h, cap_h);
        sift_b200_ctx* raw = nullptr;
        const int rc = sift_b200_create(env_device(), cap_w, cap_h, &raw);
        if (rc != SIFT_B200_OK)
            throw std::runtime_error(std::string("sift_b200_create: ") + sift_b200_last_error(nullptr));
        ctx.reset(raw);
    }
    return ctx.get();
}

// Page-locked staging buffer for the pixels (grown on demand, one per host thread): the host -> device copy of
// sift_b200_detect_* is then a true asynchronous DMA instead of a staged copy out of pageable memory.
struct PinnedBuf {
    void* p = nullptr;
    size_t cap = 0;
    ~PinnedBuf() { sift_b200_host_free(p); }
    void* get(size_t bytes) {
        if (bytes > cap) {
            sift_b200_host_free(p);
            p = nullptr; cap = 0;
            if (sift_b200_host_alloc(bytes, &p) != SIFT_B200_OK)
                throw std::runtime_error(std::string("sift_b200_host_alloc: ") + sift_b200_last_error(nullptr));
            cap = bytes;
        }
        return p;
    }
};

// f(begin, end) -> bool over [0, n) in up to 8 chunks on as many threads (one per 2 M elements); AND of the results
template <typename F>
bool parallel_chunks(size_t n, F f) {
    const size_t hw = std::max(1u, std::thread::hardware_concurrency());
    const size_t parts = std::max<size_t>(1, std::min<size_t>({(size_t)8, hw, n >> 21}));
    if (parts == 1) return f((size_t)0, n);
    std::vector<char> ok(parts, 1);
    std::vector<std::thread> pool;
    for (size_t t = 1; t < parts; ++t)
        pool.emplace_back([&, t] { ok[t] = f(n * t / parts, n * (t + 1) / parts) ? 1 : 0; });
    ok[0] = f((size_t)0, n / parts) ? 1 : 0;
    for (std::thread& th : pool) th.join();
    bool all = true;
    for (char c : ok) all = all && c;
    return all;
}

[[noreturn]] void fail(sift_b200_ctx* c, const char* what) {
    throw std::runtime_error(std::string(what) + ": " + sift_b200_last_error(c));
}

}  // namespace

std::vector<Keypoint> detect_keypoints_and_descriptors(const Image& img, const bool double_image_size,
                                                       const double init_sigma, const int intervals,
                                                       const int window_size, const double contrast_threshold,
                                                       const double eigen_ratio, const double num_bins,
                                                       const double peak_ratio, const double ori_sigma_factor,
                                                       const double desc_scale_factor) {
    const bool chatty = !env_on("SIFT_B200_QUIET");
    if (img.channels != 1 && img.channels != 3)
        throw std::runtime_error("detect_keypoints_and_descriptors: image must have 1 or 3 channels");
    sift_b200_ctx* ctx = context_for(img.width, img.height);
    sift_b200_params p;
    sift_b200_default_params(&p);
    p.double_image_size = double_image_size ? 1 : 0;
    p.init_sigma = init_sigma;
    p.intervals = intervals;
    p.window_size = window_size;
    p.contrast_threshold = contrast_threshold;
    p.eigen_ratio = eigen_ratio;
    p.num_bins = num_bins;
    p.peak_ratio = peak_ratio;
    p.ori_sigma_factor = ori_sigma_factor;
    p.desc_scale_factor = desc_scale_factor;

    // Image::data holds doubles 0..255 that came from 8-bit files (image_io.cpp:27-33): ship them
    // as bytes when that is lossless, as floats otherwise.
    // One pass over the doubles (66 MB for a 4K gray image: the conversion, not the GPU, is what a caller waits for,
    // so it is split over a few host threads), a second one only when some pixel is not a whole number in 0..255.
    const size_t n = img.data.size();
    std::vector<Keypoint> out(std::max<size_t>(4096, (size_t)img.width * img.height / 16));
    int count = 0, rc;
    thread_local PinnedBuf staging;
    uint8_t* bytes = static_cast<uint8_t*>(staging.get(n * sizeof(float)));   // room for the float form as well
    const double* src = img.data.data();
    const bool bytes_ok = parallel_chunks(n, [&](size_t b, size_t e) {
        bool ok = true;
        for (size_t i = b; i < e; ++i) {
            const double v = src[i];
            const double c = (v >= 0.0 && v <= 255.0) ? v : 0.0;
            const uint8_t q = (uint8_t)c;
            bytes[i] = q;
            ok &= (double)q == v;
        }
        return ok;
    });
    if (!bytes_ok) {
        float* px = static_cast<float*>(staging.p);
        parallel_chunks(n, [&](size_t b, size_t e) {
            for (size_t i = b; i < e; ++i) px[i] = (float)src[i];
            return true;
        });
    }
    for (int attempt = 0; attempt < 2; ++attempt) {
        if (bytes_ok)
            rc = sift_b200_detect_u8(ctx, static_cast<const uint8_t*>(staging.p), img.width, img.height, img.channels, &p,
                                     reinterpret_cast<sift_b200_keypoint*>(out.data()), (int)out.size(), &count);
        else
            rc = sift_b200_detect_f32(ctx, static_cast<const float*>(staging.p), img.width, img.height, img.channels, &p,
                                      reinterpret_cast<sift_b200_keypoint*>(out.data()), (int)out.size(), &count);
        if (rc == SIFT_B200_E_CAPACITY && count > (int)out.size()) {
            out.resize(count);
            continue;
        }
        break;
    }
    if (rc != SIFT_B200_OK) fail(ctx, "sift_b200_detect");
    out.resize(count);

    if (chatty) {  // the counts the reference prints (sift.cpp:719-763)
        sift_b200_stats st;
        sift_b200_get_stats(ctx, &st);
        std::cout << "Initial image computed: " << st.base_width << "x" << st.base_height << std::endl;
        std::cout << "Octaves count: " << st.octaves << std::endl;
        std::cout << "Extrema points detected: " << st.extrema << std::endl;
        std::cout << "Raw keypoints computed: " << st.raw_keypoints << std::endl;
        std::cout << "Oriented keypoints computed: " << st.oriented_keypoints << std::endl;
        std::cout << "Final keypoints: " << st.final_keypoints << std::endl;
    }
    if (!env_on("SIFT_B200_NO_KEYPOINTS_PNG")) {  // sift.cpp:765-768
        Image canvas(img);
        draw_keypoints(canvas, out, intervals + 3);
        canvas.save("keypoints.png");
    }
    if (chatty) std::cout << "Descriptors computed!" << std::endl;
    return out;
}

std::vector<KeypointMatch> match_keypoints(const std::vector<Keypoint>& keypoints1,
                                           const std::vector<Keypoint>& keypoints2, double ratio_threshold) {
    std::vector<KeypointMatch> matches;
    const int na = (int)keypoints1.size(), nb = (int)keypoints2.size();
    if (na == 0 || nb == 0) return matches;  // sift.cpp:789-812 emits nothing
    sift_b200_ctx* ctx = context_for(64, 64);
    std::vector<uint8_t> a((size_t)na * 128), b((size_t)nb * 128);
    for (int i = 0; i < na; ++i) std::memcpy(&a[(size_t)i * 128], keypoints1[i].desc, 128);
    for (int j = 0; j < nb; ++j) std::memcpy(&b[(size_t)j * 128], keypoints2[j].desc, 128);
    std::vector<int32_t> ia(na), ib(na);
    std::vector<double> dist(na);
    int count = 0;
    if (sift_b200_match(ctx, a.data(), na, b.data(), nb, ratio_threshold, ia.data(), ib.data(), dist.data(), na,
                        &count) != SIFT_B200_OK)
        fail(ctx, "sift_b200_match");
    matches.reserve(count);
    for (int k = 0; k < count; ++k) matches.emplace_back(keypoints1[ia[k]], keypoints2[ib[k]], dist[k]);
    return matches;
}

// Host-side cosmetics (sift.cpp:821-876 in the reference): a ring sized by the layer and a spoke
// along the principal orientation per keypoint; both images side by side with one line per match.
void draw_keypoints(Image& img, const std::vector<Keypoint>& keypoints, double scales_count) {
    static const std::array<Color, 7> palette = {Color::RED,     Color::GREEN, Color::BLUE, Color::YELLOW,
                                                 Color::MAGENTA, Color::CYAN,  Color::BLACK};
    const double r_small = 5.0, r_large = 110.0;
    for (const Keypoint& k : keypoints) {
        const int cx = (int)k.x, cy = (int)k.y;
        const int r = (int)(r_small * std::exp(k.layer / (scales_count - 1) * std::log(r_large / r_small)));
        const int colour = palette[k.layer % palette.size()];
        img.draw_circle(cx, cy, r, colour);
        img.draw_line(cx, cy, (int)(cx + r * std::cos(k.pori)), (int)(cy + r * std::sin(k.pori)), colour);
    }
}

void draw_matches(const Image& a, const Image& b, std::vector<KeypointMatch> matches) {
    Image canvas(a.width + b.width, std::max(a.height, b.height), 3);
    const Channel rgb[3] = {R, G, B};
    auto paste = [&](const Image& src, int x_off) {
        for (int x = 0; x < src.width; ++x)
            for (int y = 0; y < src.height; ++y)
                for (int c = 0; c < 3; ++c)
                    canvas.set_pixel(x_off + x, y, rgb[c], src.get_pixel(x, y, src.channels == 3 ? rgb[c] : R));
    };
    paste(a, 0);
    paste(b, a.width);
    for (const KeypointMatch& m : matches)
        canvas.draw_line((int)m.kp1.x, (int)m.kp1.y, (int)(a.width + m.kp2.x), (int)m.kp2.y);
    canvas.save("matches.png");
}

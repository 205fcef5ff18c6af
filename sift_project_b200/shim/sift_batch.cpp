// detect_image_files: decode -> page-locked ring -> asynchronous H2D -> detect, overlapped.  See sift_batch.hh.
// Links against the reference's image_io.cpp (Image(path) = its vendored stb decoder) and libsift_b200.so.
#include "sift_batch.hh"

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <deque>
#include <exception>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <thread>

#include "sift_b200.h"

static_assert(sizeof(Keypoint) == sizeof(sift_b200_keypoint), "Keypoint must be the 168-byte record");

namespace {

using Clock = std::chrono::steady_clock;

// a decoded image waiting for (or occupying) a GPU context: u8 pixels in page-locked memory
struct Slot {
    void* pixels = nullptr;
    size_t cap = 0;
    int index = -1, w = 0, h = 0, c = 0;
    ~Slot() { sift_b200_host_free(pixels); }
    void reserve(size_t bytes) {
        if (bytes <= cap) return;
        sift_b200_host_free(pixels);
        pixels = nullptr;
        cap = 0;
        if (sift_b200_host_alloc(bytes, &pixels) != SIFT_B200_OK)
            throw std::runtime_error(std::string("sift_b200_host_alloc: ") + sift_b200_last_error(nullptr));
        cap = bytes;
    }
};

struct Lane {   // one GPU context and what it is working on
    sift_b200_ctx* ctx = nullptr;
    int device = 0, max_w = 0, max_h = 0;
    Slot* busy = nullptr;
    ~Lane() { sift_b200_destroy(ctx); }
};

}  // namespace

std::vector<std::vector<Keypoint>> detect_image_files(const std::vector<std::string>& paths,
                                                      const SiftBatchOptions& opt, SiftBatchTimes* times) {
    const auto t_begin = Clock::now();
    const int n = (int)paths.size();
    std::vector<std::vector<Keypoint>> out(n);
    if (n == 0) return out;
    if (opt.devices.empty() || opt.contexts_per_device < 1 || opt.decode_threads < 1)
        throw std::runtime_error("detect_image_files: need at least one device, context and decode thread");

    sift_b200_params prm;
    sift_b200_default_params(&prm);
    prm.double_image_size = opt.double_image_size ? 1 : 0;
    prm.init_sigma = opt.init_sigma;
    prm.intervals = opt.intervals;
    prm.window_size = opt.window_size;
    prm.contrast_threshold = opt.contrast_threshold;
    prm.eigen_ratio = opt.eigen_ratio;
    prm.num_bins = opt.num_bins;
    prm.peak_ratio = opt.peak_ratio;
    prm.ori_sigma_factor = opt.ori_sigma_factor;
    prm.desc_scale_factor = opt.desc_scale_factor;

    const int n_lanes = (int)opt.devices.size() * opt.contexts_per_device;
    const int n_slots = n_lanes + opt.decode_threads;   // every lane busy + every decoder holding one
    std::vector<std::unique_ptr<Slot>> slots;
    for (int i = 0; i < n_slots; ++i) slots.emplace_back(new Slot());
    std::vector<Lane> lanes(n_lanes);
    for (int i = 0; i < n_lanes; ++i) lanes[i].device = opt.devices[i % opt.devices.size()];

    std::mutex mu;
    std::condition_variable cv;
    std::deque<Slot*> free_slots, ready;
    for (auto& s : slots) free_slots.push_back(s.get());
    std::atomic<int> next{0};
    std::exception_ptr failure;
    double decode_cpu = 0.0;
    int decoders_done = 0;

    // ---- decode threads: the reference's own decoder, then doubles -> bytes into page-locked memory ----
    auto decode = [&]() {
        double mine = 0.0;
        try {
            for (;;) {
                const int k = next.fetch_add(1);
                if (k >= n) break;
                Slot* s;
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&] { return !free_slots.empty() || failure; });
                    if (failure) break;
                    s = free_slots.front();
                    free_slots.pop_front();
                }
                const auto t0 = Clock::now();
                Image img(paths[k]);   // image_io.cpp:20-35 (throws on an unreadable file)
                if (img.channels != 1 && img.channels != 3)
                    throw std::runtime_error(paths[k] + ": 1 or 3 channels expected");
                const size_t px = img.data.size();
                s->reserve(px);
                uint8_t* dst = static_cast<uint8_t*>(s->pixels);
                for (size_t i = 0; i < px; ++i) dst[i] = (uint8_t)img.data[i];   // file pixels are integers 0..255
                s->index = k; s->w = img.width; s->h = img.height; s->c = img.channels;
                mine += std::chrono::duration<double>(Clock::now() - t0).count();
                {
                    std::lock_guard<std::mutex> lk(mu);
                    ready.push_back(s);
                }
                cv.notify_all();
            }
        } catch (...) {
            std::lock_guard<std::mutex> lk(mu);
            if (!failure) failure = std::current_exception();
        }
        {
            std::lock_guard<std::mutex> lk(mu);
            decode_cpu += mine;
            ++decoders_done;
        }
        cv.notify_all();
    };
    std::vector<std::thread> pool;
    for (int t = 0; t < opt.decode_threads; ++t) pool.emplace_back(decode);

    // ---- this thread drives the GPUs ----
    auto collect = [&](Lane& ln) {
        if (!ln.busy) return;
        Slot* s = ln.busy;
        std::vector<Keypoint>& kp = out[s->index];
        kp.resize(std::max<size_t>(4096, (size_t)s->w * s->h / 16));
        int count = 0;
        int rc = sift_b200_result_copy(ln.ctx, reinterpret_cast<sift_b200_keypoint*>(kp.data()), (int)kp.size(), &count);
        if (rc == SIFT_B200_E_CAPACITY && count > (int)kp.size()) {
            kp.resize(count);
            rc = sift_b200_result_copy(ln.ctx, reinterpret_cast<sift_b200_keypoint*>(kp.data()), (int)kp.size(), &count);
        }
        if (rc != SIFT_B200_OK) throw std::runtime_error(std::string("sift_b200_result_copy: ") + sift_b200_last_error(ln.ctx));
        kp.resize(count);
        ln.busy = nullptr;
        {
            std::lock_guard<std::mutex> lk(mu);
            free_slots.push_back(s);
        }
        cv.notify_all();
    };
    try {
        int done = 0, turn = 0;
        while (done < n) {
            Slot* s = nullptr;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return !ready.empty() || failure || decoders_done == (int)pool.size(); });
                if (failure) std::rethrow_exception(failure);
                if (ready.empty()) break;   // decoders finished and everything was handed over
                s = ready.front();
                ready.pop_front();
            }
            Lane& ln = lanes[turn++ % n_lanes];
            collect(ln);                                      // the lane's previous image, if any
            if (!ln.ctx || s->w > ln.max_w || s->h > ln.max_h) {   // first use, or a larger image than before
                sift_b200_destroy(ln.ctx);
                ln.ctx = nullptr;
                ln.max_w = std::max(ln.max_w, s->w);
                ln.max_h = std::max(ln.max_h, s->h);
                if (sift_b200_create(ln.device, ln.max_w, ln.max_h, &ln.ctx) != SIFT_B200_OK)
                    throw std::runtime_error(std::string("sift_b200_create: ") + sift_b200_last_error(nullptr));
            }
            if (sift_b200_detect_enqueue_u8(ln.ctx, static_cast<const uint8_t*>(s->pixels), s->w, s->h, s->c, &prm) !=
                SIFT_B200_OK)
                throw std::runtime_error(std::string("sift_b200_detect_enqueue_u8: ") + sift_b200_last_error(ln.ctx));
            ln.busy = s;
            ++done;
        }
        for (Lane& ln : lanes) collect(ln);
    } catch (...) {
        {
            std::lock_guard<std::mutex> lk(mu);
            if (!failure) failure = std::current_exception();
        }
        cv.notify_all();
    }
    for (std::thread& t : pool) t.join();
    if (failure) std::rethrow_exception(failure);
    if (times) {
        times->wall_s = std::chrono::duration<double>(Clock::now() - t_begin).count();
        times->decode_cpu_s = decode_cpu;
    }
    return out;
}

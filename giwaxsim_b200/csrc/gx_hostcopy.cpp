// Host side of the device -> host boundary: the bulk host pass every call of voxelgridmaker_fitting pays.
//   gx_host_widen_f32_f64  page-locked staging buffer (fp32 voxel grid as it crossed PCIe) -> the float64 array
//                          the reference's interface returns (comparison.py:769-786 hands out float64)
// Memory-bandwidth bound.  It runs on a small persistent thread pool (torchrun pins OMP_NUM_THREADS=1, and a
// thread per call costs more than a chunk) and writes with non-temporal stores: the 0.5 GB destination is not
// read back by the CPU before the caller gets it, so the read-for-ownership of every destination line - 40 %
// of the traffic of the pass - is saved.  (The opposite direction, pageable user array -> 8 MB page-locked
// staging halves, stays with torch's threaded copy: its destination is cache-resident and streaming stores
// measured slower there.)  SSE2 only (baseline x86-64); any other host takes the plain loop.
#include <stddef.h>
#include <stdint.h>

#include "../../include/giwaxs_b200.h"

#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#if defined(__x86_64__) || defined(_M_X64)
#include <emmintrin.h>
#define GX_HOST_SSE2 1
#else
#define GX_HOST_SSE2 0
#endif

void gx_set_error(const char *fmt, ...);           // gx_api.cu

namespace {

class HostPool {
public:
    // runs f(0) .. f(parts - 1) on the workers and the calling thread; returns when all are done
    void run(int parts, int threads, const std::function<void(int)> &f)
    {
        std::lock_guard<std::mutex> one_job(call_);
        {
            std::lock_guard<std::mutex> lk(m_);
            while ((int)workers_.size() < threads - 1) {
                workers_.emplace_back([this] { work(); });
                workers_.back().detach();
            }
            job_ = &f; parts_ = parts; next_ = 0; done_ = 0; ++gen_;
        }
        wake_.notify_all();
        std::unique_lock<std::mutex> lk(m_);
        drain(lk);
        finished_.wait(lk, [this] { return done_ == parts_; });
        job_ = nullptr;
    }

private:
    void drain(std::unique_lock<std::mutex> &lk)
    {
        while (job_ && next_ < parts_) {
            const int i = next_++;
            const std::function<void(int)> *f = job_;
            lk.unlock();
            (*f)(i);
            lk.lock();
            if (++done_ == parts_) finished_.notify_all();
        }
    }
    void work()
    {
        uint64_t seen = 0;
        std::unique_lock<std::mutex> lk(m_);
        for (;;) {
            wake_.wait(lk, [&] { return gen_ != seen; });
            seen = gen_;
            drain(lk);
        }
    }
    std::mutex call_, m_;
    std::condition_variable wake_, finished_;
    std::vector<std::thread> workers_;
    const std::function<void(int)> *job_ = nullptr;
    int parts_ = 0, next_ = 0, done_ = 0;
    uint64_t gen_ = 0;
};

HostPool &pool()
{
    static HostPool *p = new HostPool();          // never destroyed: its detached workers outlive static teardown
    return *p;
}

void widen_part(double *dst, const float *src, size_t n)
{
#if GX_HOST_SSE2
    size_t i = 0;
    while (i < n && ((uintptr_t)(dst + i) & 15)) { dst[i] = (double)src[i]; ++i; }
    for (; i + 8 <= n; i += 8) {
        const __m128 a = _mm_loadu_ps(src + i), b = _mm_loadu_ps(src + i + 4);
        _mm_stream_pd(dst + i, _mm_cvtps_pd(a));
        _mm_stream_pd(dst + i + 2, _mm_cvtps_pd(_mm_movehl_ps(a, a)));
        _mm_stream_pd(dst + i + 4, _mm_cvtps_pd(b));
        _mm_stream_pd(dst + i + 6, _mm_cvtps_pd(_mm_movehl_ps(b, b)));
    }
    for (; i < n; ++i) dst[i] = (double)src[i];
    _mm_sfence();
#else
    for (size_t i = 0; i < n; ++i) dst[i] = (double)src[i];
#endif
}

// [0, n) cut into `parts` ranges whose inner boundaries are multiples of `grain` items
void part_range(size_t n, int parts, int i, size_t grain, size_t &lo, size_t &hi)
{
    const size_t per = ((n + parts - 1) / parts + grain - 1) / grain * grain;
    lo = per * (size_t)i < n ? per * (size_t)i : n;
    hi = lo + per < n ? lo + per : n;
}

int clamp_threads(int threads, size_t bytes)
{
    if (threads < 1) threads = 1;
    if (threads > 64) threads = 64;
    const size_t by_size = bytes / (256u << 10) + 1;      // a part below 256 KB is not worth a wake-up
    return (size_t)threads < by_size ? threads : (int)by_size;
}

}  // namespace

extern "C" int gx_host_widen_f32_f64(const float *h_src, double *h_dst, int64_t n, int threads)
{
    if (n == 0) return GX_OK;
    if (!h_dst || !h_src || n < 0) {
        gx_set_error("gx_host_widen_f32_f64: %s", n < 0 ? "negative length" : "NULL pointer");
        return GX_ERR_INVALID;
    }
    const float *src = h_src;
    double *dst = h_dst;
    threads = clamp_threads(threads, (size_t)n * sizeof(double));
    if (threads == 1) { widen_part(dst, src, (size_t)n); return GX_OK; }
    pool().run(threads, threads, [&](int i) {
        size_t lo, hi;
        part_range((size_t)n, threads, i, 1024, lo, hi);
        if (hi > lo) widen_part(dst + lo, src + lo, hi - lo);
    });
    return GX_OK;
}

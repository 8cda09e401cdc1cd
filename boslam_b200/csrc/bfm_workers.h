// bfm_workers.h - a small host thread pool for the BFM_MEM_HOST path with pageable caller arrays.
// (included by bfm_api.cu before the handle definition)
//
// Pageable memory cannot be read by the GPU, and one CPU thread copies at ~16 GB/s on this pool's hosts
// (tools/memcpy_bw_probe.py; 8 threads: ~44 GB/s).  The pool stages the caller's arrays into a pinned
// buffer slice by slice while the matching kernel is already running: its feeder CTAs pick every slice up
// as soon as the host has published it (bfm_pipeline.cuh).
#pragma once
#include <atomic>
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#include <cstring>
#if defined(__SSE2__) || defined(_M_X64)
#include <emmintrin.h>
#define BFM_HAVE_SSE2 1
#endif

// memcpy into pinned staging memory with non-temporal stores: the data is about to be read by the GPU over
// PCIe, not by this CPU, and lines left modified in the CPU caches make those device reads several times
// slower (measured: 2.4 ms instead of 1.2 ms per 32.8 MB batch with a cache-allocating memcpy).
inline void stream_copy(void *dst, const void *src, size_t n) {
#ifndef BFM_HAVE_SSE2
    std::memcpy(dst, src, n);   // non-x86 host: no portable non-temporal store; plain copy + full fence
    std::atomic_thread_fence(std::memory_order_seq_cst);
#else
    char *d = static_cast<char *>(dst);
    const char *s = static_cast<const char *>(src);
    while (n && (reinterpret_cast<uintptr_t>(d) & 15)) { *d++ = *s++; --n; }
    size_t blocks = n / 64;
    while (blocks--) {
        const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s));
        const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s + 16));
        const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s + 32));
        const __m128i e = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s + 48));
        _mm_stream_si128(reinterpret_cast<__m128i *>(d), a);
        _mm_stream_si128(reinterpret_cast<__m128i *>(d + 16), b);
        _mm_stream_si128(reinterpret_cast<__m128i *>(d + 32), c);
        _mm_stream_si128(reinterpret_cast<__m128i *>(d + 48), e);
        s += 64;
        d += 64;
    }
    n &= 63;
    if (n) std::memcpy(d, s, n);
    _mm_sfence();
#endif
}

class WorkerPool {
public:
    explicit WorkerPool(int n) {
        for (int i = 0; i < n; ++i) threads_.emplace_back([this] { loop(); });
    }
    ~WorkerPool() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &t : threads_) t.join();
    }
    int size() const { return (int)threads_.size(); }
    void submit(std::function<void()> job) {
        {
            std::lock_guard<std::mutex> lk(m_);
            q_.push_back(std::move(job));
            ++pending_;
        }
        cv_.notify_one();
    }
    // many jobs under ONE lock and one wake-up of every worker: a lock + notify per job costs ~1.5 us each, and a
    // staged call queues two dozen of them before it may launch its kernel
    void submit_batch(std::vector<std::function<void()>> &&jobs) {
        if (jobs.empty()) return;
        {
            std::lock_guard<std::mutex> lk(m_);
            pending_ += (int)jobs.size();
            for (auto &j : jobs) q_.push_back(std::move(j));
        }
        cv_.notify_all();
    }
    // the calling thread works through the queue too, then waits for the jobs other threads still run
    void help_and_wait() {
        for (;;) {
            std::function<void()> job;
            {
                std::lock_guard<std::mutex> lk(m_);
                if (q_.empty()) break;
                job = std::move(q_.front());
                q_.pop_front();
            }
            job();
            finish_one();
        }
        std::unique_lock<std::mutex> lk(m_);
        done_cv_.wait(lk, [this] { return pending_ == 0; });
    }

private:
    void finish_one() {
        std::lock_guard<std::mutex> lk(m_);
        if (--pending_ == 0) done_cv_.notify_all();
    }
    void loop() {
        for (;;) {
            std::function<void()> job;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [this] { return stop_ || !q_.empty(); });
                if (stop_ && q_.empty()) return;
                job = std::move(q_.front());
                q_.pop_front();
            }
            job();
            finish_one();
        }
    }
    std::vector<std::thread> threads_;
    std::deque<std::function<void()>> q_;
    std::mutex m_;
    std::condition_variable cv_, done_cv_;
    int pending_ = 0;
    bool stop_ = false;
};

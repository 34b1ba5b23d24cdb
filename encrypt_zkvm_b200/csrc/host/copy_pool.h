// Fork-join memcpy over a few persistent host threads.
//
// Used by the trace upload when the caller's columns are ordinary pageable memory (what a `Vec<BaseElement>` of the
// reference's `TraceTable` is, vm/src/lib.rs:18): the driver stages such copies through its own buffer with one
// thread (measured: 448 MiB in 41 ms, i.e. longer than the 27 ms the whole proof needs on the device), so the prover
// copies the columns into its own page-locked ring with several threads and lets the DMA engine run from there.
//
// A copy is cut into pieces of kPiece bytes that the threads (the caller included) claim from a shared counter, so a
// thread that the host schedules late or on a slow core takes fewer pieces instead of holding the others up; between
// copies the workers spin for a short while before they sleep, because the next slot of the ring follows within
// microseconds while a column is travelling.
#pragma once
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace ezk {

class CopyPool {
public:
    // `threads` includes the calling thread: threads - 1 workers are started
    explicit CopyPool(unsigned threads) : parts_(std::max(1u, threads)) {
        for (unsigned i = 1; i < parts_; i++) workers_.emplace_back([this] { run(); });
    }
    ~CopyPool() {
        stopping_.store(true, std::memory_order_release);
        {
            std::lock_guard<std::mutex> lock(mu_);
            stop_ = true;
        }
        wake_.notify_all();
        for (auto& t : workers_) t.join();
    }
    CopyPool(const CopyPool&) = delete;
    CopyPool& operator=(const CopyPool&) = delete;

    unsigned threads() const { return parts_; }

    // dst[0, bytes) = src[0, bytes); returns when every piece is written.  One caller at a time.
    void copy(void* dst, const void* src, size_t bytes) {
        if (parts_ == 1 || bytes < kMinParallel) {
            memcpy(dst, src, bytes);
            return;
        }
        Job job;
        {
            std::lock_guard<std::mutex> lock(mu_);  // workers take their snapshot of the job under this mutex
            job_.dst = (uint8_t*)dst, job_.src = (const uint8_t*)src, job_.bytes = bytes;
            job_.gen = (uint32_t)(job_.gen + 1);
            left_.store((bytes + kPiece - 1) / kPiece, std::memory_order_relaxed);
            next_.store((uint64_t)job_.gen << 32, std::memory_order_release);
            job = job_;
        }
        wake_.notify_all();
        work(job);
        // pieces claimed by workers may still be in flight
        for (unsigned spins = 0; left_.load(std::memory_order_acquire) != 0; spins++)
            if (spins > 64) std::this_thread::yield();
    }

private:
    static constexpr size_t kMinParallel = 256 << 10, kPiece = 512 << 10;
    static constexpr int kSpinMicros = 200;

    struct Job {
        uint8_t* dst = nullptr;
        const uint8_t* src = nullptr;
        size_t bytes = 0;
        uint32_t gen = 0;
    };

    // Claims pieces of `job` until none is left.  next_ = (generation << 32) | next piece: a claim only succeeds while
    // the counter still belongs to the job of the snapshot, and a job cannot end (nor the next one begin) before every
    // claimed piece has been written and counted off in left_.
    void work(const Job& job) {
        const uint64_t pieces = (job.bytes + kPiece - 1) / kPiece;
        uint64_t cur = next_.load(std::memory_order_acquire);
        for (;;) {
            if ((uint32_t)(cur >> 32) != job.gen) return;
            const uint64_t i = cur & 0xFFFFFFFFull;
            if (i >= pieces) return;
            if (!next_.compare_exchange_weak(cur, cur + 1, std::memory_order_acq_rel, std::memory_order_acquire)) continue;
            const size_t lo = (size_t)i * kPiece, hi = std::min(job.bytes, lo + kPiece);
            memcpy(job.dst + lo, job.src + lo, hi - lo);
            left_.fetch_sub(1, std::memory_order_release);
            cur = next_.load(std::memory_order_acquire);
        }
    }

    void run() {
        uint32_t seen = 0;
        for (;;) {
            // a new job shows in the upper half of next_: spin briefly (the next ring slot usually follows at once), then sleep
            const auto until = std::chrono::steady_clock::now() + std::chrono::microseconds(kSpinMicros);
            while ((uint32_t)(next_.load(std::memory_order_acquire) >> 32) == seen && !stopping_.load(std::memory_order_acquire) &&
                   std::chrono::steady_clock::now() < until) {
            }
            Job job;
            {
                std::unique_lock<std::mutex> lock(mu_);
                wake_.wait(lock, [&] { return stop_ || job_.gen != seen; });
                if (stop_) return;
                job = job_;
            }
            seen = job.gen;
            work(job);
        }
    }

    const unsigned parts_;
    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable wake_;
    std::atomic<uint64_t> next_{0}, left_{0};
    std::atomic<bool> stopping_{false};
    bool stop_ = false;
    Job job_;
};

}  // namespace ezk

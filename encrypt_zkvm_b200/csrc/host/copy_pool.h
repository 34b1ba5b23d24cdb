// Fork-join memcpy over a few persistent host threads.
//
// Used by the trace upload when the caller's columns are ordinary pageable memory (what a `Vec<BaseElement>` of the
// reference's `TraceTable` is, vm/src/lib.rs:18): the driver stages such copies through its own buffer with one
// thread (measured: 448 MiB in 41 ms, i.e. longer than the 27 ms the whole proof needs on the device), so the prover
// copies the columns into its own page-locked ring with several threads and lets the DMA engine run from there.
#pragma once
#include <algorithm>
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
        for (unsigned i = 1; i < parts_; i++) workers_.emplace_back([this, i] { run(i); });
    }
    ~CopyPool() {
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

    // dst[0, bytes) = src[0, bytes); returns when every part is written.  One caller at a time.
    void copy(void* dst, const void* src, size_t bytes) {
        if (parts_ == 1 || bytes < kMinParallel) {
            memcpy(dst, src, bytes);
            return;
        }
        {
            std::lock_guard<std::mutex> lock(mu_);
            dst_ = (uint8_t*)dst, src_ = (const uint8_t*)src, bytes_ = bytes;
            pending_ = parts_ - 1;
            generation_++;
        }
        wake_.notify_all();
        copy_part(0, (uint8_t*)dst, (const uint8_t*)src, bytes);
        std::unique_lock<std::mutex> lock(mu_);
        done_.wait(lock, [this] { return pending_ == 0; });
    }

private:
    static constexpr size_t kMinParallel = 256 << 10, kAlign = 4096;

    void copy_part(unsigned part, uint8_t* dst, const uint8_t* src, size_t bytes) const {
        const size_t per = ((bytes + parts_ - 1) / parts_ + kAlign - 1) / kAlign * kAlign;  // parts_ * per >= bytes
        const size_t lo = std::min(bytes, (size_t)part * per), hi = std::min(bytes, lo + per);
        if (hi > lo) memcpy(dst + lo, src + lo, hi - lo);
    }

    void run(unsigned part) {
        uint64_t seen = 0;
        for (;;) {
            uint8_t* dst;
            const uint8_t* src;
            size_t bytes;
            {
                std::unique_lock<std::mutex> lock(mu_);
                wake_.wait(lock, [&] { return stop_ || generation_ != seen; });
                if (stop_) return;
                seen = generation_;
                dst = dst_, src = src_, bytes = bytes_;
            }
            copy_part(part, dst, src, bytes);
            bool last;
            {
                std::lock_guard<std::mutex> lock(mu_);
                last = --pending_ == 0;
            }
            if (last) done_.notify_one();
        }
    }

    const unsigned parts_;
    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable wake_, done_;
    uint64_t generation_ = 0;
    unsigned pending_ = 0;
    bool stop_ = false;
    uint8_t* dst_ = nullptr;
    const uint8_t* src_ = nullptr;
    size_t bytes_ = 0;
};

}  // namespace ezk

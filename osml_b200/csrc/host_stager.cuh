// Pageable host memory -> device, faster than the driver does it.
//
// cudaMemcpyAsync from pageable memory is staged by the driver through its own pinned buffer on the calling thread:
// ~19 GB/s on the B200 hosts against ~54 GB/s from pinned memory, and the call blocks for the whole copy.  A pandas
// frame (`DataFrame.values`, osml/predict.py:787, 807) is pageable, so the reference-facing call
// `model.predict(DataFrame)` spent more time uploading a 10 000 x 784 float32 frame (1.6 ms) than searching it.
// HostStager cuts a copy into chunks of about a megabyte, lets a few worker threads gather the chunks into a ring of
// pinned slots in parallel and issues one cudaMemcpyAsync per chunk from the calling thread, so gathering, DMA and --
// from the second block of a pipelined call on -- the kernels overlap.  Plain host code: no arithmetic.
#pragma once
#include <cuda_runtime.h>
#include <unistd.h>
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

namespace osml {

class HostStager {
public:
    static constexpr size_t kSlotBytes = (size_t)1 << 20;
    static constexpr int kSlots = 16;
    static constexpr int kAhead = 8;                 // slots the workers may fill ahead of the chunk being issued
    // gather threads ($OSML_STAGER_THREADS for experiments); the caller issues every copy and gathers a chunk itself when
    // it would otherwise wait.  Measured on a 16-core B200 host, 31 MB column-major, whole predict call, workers only:
    // 2 threads 1.60 ms, 3: 1.67, 4: 1.96, 6: 2.07, 8: 3.05 (driver's pageable path: 2.68)
    static constexpr int kWorkers = 2;

    // One chunk: `fill(dst)` gathers `bytes` (<= kSlotBytes) of host data into the pinned slot `dst`; the slot then goes
    // to `device_dst` on `stream`.
    struct Chunk {
        void* device_dst;
        size_t bytes;
        std::function<void(char*)> fill;
    };

    static HostStager& instance() {
        static HostStager* s = new HostStager();     // never destroyed: worker threads may outlive static destructors
        return *s;
    }

    // Copies all chunks; returns once every chunk's DMA is queued on `stream` (the host data is no longer needed).
    // Worker threads only gather (chunk c goes to slot c % kSlots once the DMA of chunk c - kSlots has left it); every
    // CUDA call comes from the calling thread, in chunk order: several threads issuing copies contend for the context
    // lock and were slower than the driver's own pageable path.
    cudaError_t copy(int device, const std::vector<Chunk>& chunks, cudaStream_t stream) {
        std::lock_guard<std::mutex> job_guard(job_mutex_);          // one copy at a time
        cudaError_t e = ensure(device);
        if (e != cudaSuccess) return e;
        const int n = (int)chunks.size();
        if ((int)filled_size_ < n) { filled_.reset(new std::atomic<unsigned char>[n]); filled_size_ = n; }
        for (int c = 0; c < n; ++c) filled_[c].store(0, std::memory_order_relaxed);
        job_chunks_ = &chunks;
        next_.store(0);
        confirmed_.store(-kSlots);                   // chunks -kSlots .. -1: whatever earlier copies left in the slots
        start_workers(0);
        for (int c = 0; c < n && e == cudaSuccess; ++c) {
            // confirm old DMAs so that the workers can run kAhead chunks ahead (they finished long ago: the DMA is
            // faster than the gathering, so these waits return at once)
            while (e == cudaSuccess && confirmed_.load(std::memory_order_relaxed) < std::min(c, c - kSlots + 1 + kAhead)) {
                const int o = confirmed_.load(std::memory_order_relaxed);
                e = cudaEventSynchronize(events_[((o % kSlots) + kSlots) % kSlots]);
                confirmed_.store(o + 1, std::memory_order_release);
            }
            if (e != cudaSuccess) break;             // (a worker waiting for this confirmation is released below)
            // while chunk c is still being gathered the calling thread gathers one itself (it is one more core at work;
            // the DMA engine is the faster side and has the ring's slots to draw from)
            while (e == cudaSuccess && !filled_[c].load(std::memory_order_acquire)) {
                int mine = next_.load();
                // only a chunk whose slot was last used by an ISSUED chunk (mine - kSlots < c): that DMA can be waited for here
                if (mine >= n || mine - kSlots >= c || !next_.compare_exchange_strong(mine, mine + 1)) {
                    __builtin_ia32_pause();
                    continue;
                }
                while (e == cudaSuccess && confirmed_.load(std::memory_order_relaxed) < mine - kSlots + 1) {
                    const int o = confirmed_.load(std::memory_order_relaxed);
                    e = cudaEventSynchronize(events_[((o % kSlots) + kSlots) % kSlots]);
                    confirmed_.store(o + 1, std::memory_order_release);
                }
                if (e != cudaSuccess) break;
                chunks[mine].fill(static_cast<char*>(ring_) + (size_t)(mine % kSlots) * kSlotBytes);
                filled_[mine].store(1, std::memory_order_release);
            }
            if (e != cudaSuccess) break;
            char* src = static_cast<char*>(ring_) + (size_t)(c % kSlots) * kSlotBytes;
            if (e == cudaSuccess) e = cudaMemcpyAsync(chunks[c].device_dst, src, chunks[c].bytes, cudaMemcpyHostToDevice, stream);
            if (e == cudaSuccess) e = cudaEventRecord(events_[c % kSlots], stream);
        }
        if (e != cudaSuccess) {                      // let the workers run out: nothing is copied any more, slots are free
            abort_.store(true);
            confirmed_.store(n + kSlots, std::memory_order_release);
        }
        wait_workers();
        abort_.store(false);
        return e;
    }

    // The way back: device -> pageable host memory.  The calling thread issues one DMA per chunk into the ring and
    // waits for each in order; the workers copy finished slots out to `drain`'s destination, so the page faults of a
    // freshly allocated result array (most of the cost of a plain cudaMemcpy into one) are spread over threads and run
    // under the DMA of the following chunks.
    struct OutChunk {
        const void* device_src;
        size_t bytes;
        std::function<void(const char*)> drain;
    };

    cudaError_t copy_out(int device, const std::vector<OutChunk>& chunks, cudaStream_t stream) {
        std::lock_guard<std::mutex> job_guard(job_mutex_);
        cudaError_t e = ensure(device);
        if (e != cudaSuccess) return e;
        const int n = (int)chunks.size();
        if ((int)filled_size_ < n) { filled_.reset(new std::atomic<unsigned char>[n]); filled_size_ = n; }
        if ((int)drained_size_ < n) { drained_.reset(new std::atomic<unsigned char>[n]); drained_size_ = n; }
        for (int c = 0; c < n; ++c) { filled_[c].store(0, std::memory_order_relaxed); drained_[c].store(0, std::memory_order_relaxed); }
        for (int s = 0; s < kSlots && e == cudaSuccess; ++s) e = cudaEventSynchronize(events_[s]);   // earlier copies have left the ring
        if (e != cudaSuccess) return e;
        out_chunks_ = &chunks;
        next_.store(0);
        start_workers(1);
        int issued = 0, readied = 0;
        while (readied < n) {
            while (e == cudaSuccess && issued < n &&
                   (issued < kSlots || drained_[issued - kSlots].load(std::memory_order_acquire))) {
                char* dst = static_cast<char*>(ring_) + (size_t)(issued % kSlots) * kSlotBytes;
                e = cudaMemcpyAsync(dst, chunks[issued].device_src, chunks[issued].bytes, cudaMemcpyDeviceToHost, stream);
                if (e == cudaSuccess) e = cudaEventRecord(events_[issued % kSlots], stream);
                ++issued;
            }
            if (e != cudaSuccess) break;
            if (readied < issued) {
                e = cudaEventSynchronize(events_[readied % kSlots]);
                if (e != cudaSuccess) break;
                filled_[readied].store(1, std::memory_order_release);      // the slot holds chunk `readied`
                ++readied;
            } else __builtin_ia32_pause();
        }
        if (e != cudaSuccess) {                      // let the workers run out without touching anything
            abort_.store(true);
            for (int c = 0; c < n; ++c) filled_[c].store(1, std::memory_order_release);
        }
        wait_workers();
        abort_.store(false);
        return e;
    }

private:
    HostStager() = default;

    void start_workers(int kind) {
        std::lock_guard<std::mutex> g(m_);
        job_kind_ = kind;
        pending_workers_ = nworkers_;
        ++generation_;
        cv_work_.notify_all();
    }
    void wait_workers() {
        std::unique_lock<std::mutex> g(m_);
        cv_done_.wait(g, [&] { return pending_workers_ == 0; });
    }

    cudaError_t ensure(int device) {
        if (pid_ != getpid()) {                      // first use, or a forked child (threads do not survive a fork)
            spawned_ = false;
            ring_ = nullptr;
            events_device_ = -1;
            for (int s = 0; s < kSlots; ++s) events_[s] = nullptr;
            pid_ = getpid();
        }
        if (!ring_) {
            cudaError_t e = cudaHostAlloc(&ring_, kSlotBytes * kSlots, cudaHostAllocPortable);
            if (e != cudaSuccess) { ring_ = nullptr; return e; }
        }
        if (events_device_ != device) {              // the slot events belong to one device's context
            for (int s = 0; s < kSlots; ++s) {
                if (events_[s]) { cudaEventSynchronize(events_[s]); cudaEventDestroy(events_[s]); }
                cudaError_t e = cudaEventCreateWithFlags(&events_[s], cudaEventDisableTiming);
                if (e != cudaSuccess) return e;
            }
            events_device_ = device;
        }
        if (!spawned_) {
            nworkers_ = kWorkers;
            if (const char* env = getenv("OSML_STAGER_THREADS")) nworkers_ = std::max(1, std::min(32, atoi(env)));
            for (int w = 0; w < nworkers_; ++w) std::thread([this] { worker_loop(); }).detach();
            spawned_ = true;
        }
        return cudaSuccess;
    }

    void worker_loop() {
        long seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> g(m_);
                cv_work_.wait(g, [&] { return generation_ != seen; });
                seen = generation_;
            }
            if (job_kind_ == 0) {                    // host -> device: gather chunks into free slots
                const std::vector<Chunk>& chunks = *job_chunks_;
                const int n = (int)chunks.size();
                for (;;) {
                    const int c = next_.fetch_add(1);
                    if (c >= n) break;
                    while (confirmed_.load(std::memory_order_acquire) < c - kSlots + 1) __builtin_ia32_pause();
                    if (!abort_.load()) chunks[c].fill(static_cast<char*>(ring_) + (size_t)(c % kSlots) * kSlotBytes);
                    filled_[c].store(1, std::memory_order_release);
                }
            } else {                                 // device -> host: copy landed slots out
                const std::vector<OutChunk>& chunks = *out_chunks_;
                const int n = (int)chunks.size();
                for (;;) {
                    const int c = next_.fetch_add(1);
                    if (c >= n) break;
                    while (!filled_[c].load(std::memory_order_acquire)) __builtin_ia32_pause();
                    if (!abort_.load()) chunks[c].drain(static_cast<const char*>(ring_) + (size_t)(c % kSlots) * kSlotBytes);
                    drained_[c].store(1, std::memory_order_release);
                }
            }
            {
                std::lock_guard<std::mutex> g(m_);
                --pending_workers_;
            }
            cv_done_.notify_all();
        }
    }

    std::mutex job_mutex_, m_;
    std::condition_variable cv_work_, cv_done_;
    bool spawned_ = false;
    int nworkers_ = 0;
    long generation_ = 0;
    int pending_workers_ = 0;
    pid_t pid_ = -1;
    void* ring_ = nullptr;
    cudaEvent_t events_[kSlots] = {};
    int events_device_ = -1;
    std::unique_ptr<std::atomic<unsigned char>[]> filled_, drained_;
    size_t filled_size_ = 0, drained_size_ = 0;
    int job_kind_ = 0;
    const std::vector<OutChunk>* out_chunks_ = nullptr;
    std::atomic<int> next_{0}, confirmed_{0};
    std::atomic<bool> abort_{false};
    const std::vector<Chunk>* job_chunks_ = nullptr;
};

// true when `p` is ordinary pageable host memory (not pinned / registered / managed)
inline bool is_pageable_host(const void* p) {
    cudaPointerAttributes attr;
    const cudaError_t e = cudaPointerGetAttributes(&attr, p);
    if (e != cudaSuccess) { cudaGetLastError(); return true; }
    return attr.type == cudaMemoryTypeUnregistered;
}

}  // namespace osml

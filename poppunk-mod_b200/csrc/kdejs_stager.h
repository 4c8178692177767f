// kdejs_stager.h -- host-to-device upload of PAGEABLE caller memory at PCIe speed.
//
// The reference's callers hand over ordinary numpy arrays (np.loadtxt / pandas, poppunk_mod.py:360,
// scripts/pairwise_2D_distance.py:38-45).  cudaMemcpyAsync from such memory is staged by the driver through a
// single bounce buffer on the calling thread (~11 GB/s measured here, and synchronous for the caller); pinned
// memory reaches ~48 GB/s.  The stager closes most of that gap without asking the caller for pinned arrays:
// a few worker threads memcpy chunks of the caller's ranges into a ring of pinned slots, running ahead of the
// thread that owns the CUDA stream, which enqueues one asynchronous H2D per staged chunk and records an event
// so that the slot is refilled only after its copy has completed.  Chunks are issued strictly in plan order, so
// stream-ordering arguments made by the caller (wave k's kernels wait for wave k's copies) stay valid.
#pragma once

#include <cuda_runtime.h>

#include <algorithm>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace kdejs {

class Stager {
public:
    struct Range { void *dst; const void *src; size_t bytes; };

    static constexpr size_t kChunk = (size_t)2 << 20;     // 2 MB: one 100k-pair set is a single chunk
    static constexpr int kSlots = 16;

    // One stager per device, created on first use and never destroyed (its threads are detached: a process
    // may exit, or dlclose the library, without joining them).  Returns nullptr when staging is disabled
    // (KDEJS_STAGE_THREADS=0) or the pinned ring cannot be allocated.
    static Stager *get(int device) {
        static std::mutex mu;
        static Stager *inst[64] = {nullptr};
        static bool tried[64] = {false};
        if (device < 0 || device >= 64) return nullptr;
        std::lock_guard<std::mutex> lk(mu);
        if (!tried[device]) {
            tried[device] = true;
            // measured on the 16-vCPU B200 host: 2 threads 13.7k, 3 17.6k, 4 19-23k, 6 16-21k, 8 15-20k evaluations/s
            int threads = (int)std::min<unsigned>(4u, std::max<unsigned>(2u, std::thread::hardware_concurrency() / 2u));
            if (const char *e = getenv("KDEJS_STAGE_THREADS")) threads = std::max(0, std::min(32, atoi(e)));
            if (threads > 0) {
                Stager *s = new Stager(device);
                if (s->start(threads)) inst[device] = s;     // on failure the object is leaked, deliberately small
            }
        }
        return inst[device];
    }

    static bool pageable(const void *p) {
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
        return a.type == cudaMemoryTypeUnregistered;
    }

    // Hand the workers every range of this call, in the order in which issue_range() will be called.
    void plan(const std::vector<Range> &ranges) {
        std::unique_lock<std::mutex> lk(mu_);
        next_stage_ = chunks_.size();                        // nothing of an earlier plan may still be picked up
        cv_.wait(lk, [&] { return busy_ == 0; });
        for (int i = 0; i < kSlots; ++i)                     // nor may one of its copies still read a slot
            if (ev_used_[i]) { cudaEventSynchronize(ev_[i]); ev_used_[i] = false; }
        chunks_.clear();
        range_end_.clear();
        for (const Range &r : ranges) {
            for (size_t o = 0; o < r.bytes; o += kChunk)
                chunks_.push_back(Chunk{(char *)r.dst + o, (const char *)r.src + o, std::min(kChunk, r.bytes - o)});
            range_end_.push_back(chunks_.size());
        }
        staged_.assign(chunks_.size(), 0);
        next_stage_ = 0;
        issued_ = 0;
        next_range_ = 0;
        generation_++;
        cv_.notify_all();
    }

    // Enqueue on `st` the H2D copies of the next planned range (blocks until its chunks are staged).
    cudaError_t issue_range(cudaStream_t st) {
        size_t end;
        {
            std::lock_guard<std::mutex> lk(mu_);
            if (next_range_ >= range_end_.size()) return cudaErrorInvalidValue;
            end = range_end_[next_range_++];
        }
        for (;;) {
            size_t k;
            {
                std::unique_lock<std::mutex> lk(mu_);
                k = issued_;
                if (k >= end) break;
                cv_.wait(lk, [&] { return staged_[k] != 0; });
            }
            const Chunk &c = chunks_[k];
            cudaError_t e = cudaMemcpyAsync(c.dst, slot(k), c.bytes, cudaMemcpyHostToDevice, st);
            if (e == cudaSuccess) e = cudaEventRecord(ev_[k % kSlots], st);
            {
                std::lock_guard<std::mutex> lk(mu_);
                ev_used_[k % kSlots] = true;
                issued_ = k + 1;                 // slot k % kSlots may be refilled once ev_ has completed
                cv_.notify_all();
            }
            if (e != cudaSuccess) { abandon(); return e; }
        }
        return cudaSuccess;
    }

    // Forget the rest of the plan (error paths).  Workers finish the chunk they hold and go idle.
    void abandon() {
        std::unique_lock<std::mutex> lk(mu_);
        next_stage_ = chunks_.size();
        cv_.wait(lk, [&] { return busy_ == 0; });
        chunks_.clear();
        range_end_.clear();
        issued_ = 0;
        next_range_ = 0;
    }

private:
    struct Chunk { char *dst; const char *src; size_t bytes; };

    explicit Stager(int device) : dev_(device) {}

    bool start(int threads) {
        if (cudaSetDevice(dev_) != cudaSuccess) return false;
        if (cudaHostAlloc((void **)&ring_, kChunk * kSlots, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return false; }
        for (int i = 0; i < kSlots; ++i)
            if (cudaEventCreateWithFlags(&ev_[i], cudaEventDisableTiming) != cudaSuccess) return false;
        for (int t = 0; t < threads; ++t) std::thread([this] { work(); }).detach();
        return true;
    }

    char *slot(size_t k) const { return ring_ + (k % kSlots) * kChunk; }

    void work() {
        cudaSetDevice(dev_);
        std::unique_lock<std::mutex> lk(mu_);
        for (;;) {
            // a chunk may be staged once the slot's previous occupant (chunk k - kSlots) has been issued
            cv_.wait(lk, [&] { return next_stage_ < chunks_.size() && next_stage_ < issued_ + kSlots; });
            const size_t k = next_stage_++;
            const Chunk c = chunks_[k];
            const uint64_t gen = generation_;
            busy_++;
            lk.unlock();
            if (k >= (size_t)kSlots) cudaEventSynchronize(ev_[k % kSlots]);    // that occupant's H2D has completed
            memcpy(slot(k), c.src, c.bytes);
            lk.lock();
            busy_--;
            if (gen == generation_ && k < staged_.size()) staged_[k] = 1;
            cv_.notify_all();
        }
    }

    const int dev_;
    char *ring_ = nullptr;
    cudaEvent_t ev_[kSlots] = {};
    bool ev_used_[kSlots] = {};
    std::mutex mu_;
    std::condition_variable cv_;
    std::vector<Chunk> chunks_;
    std::vector<size_t> range_end_;
    std::vector<char> staged_;
    size_t next_stage_ = 0, issued_ = 0, next_range_ = 0;
    int busy_ = 0;
    uint64_t generation_ = 0;
};

}  // namespace kdejs

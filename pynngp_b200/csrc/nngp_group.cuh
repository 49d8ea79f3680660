// nngp_group.cuh -- the host-side runtime of a multi-device handle (nngp_create_multi): one sub-handle per
// device and one persistent host thread per device beyond the first.
//
// Why threads: a kernel launch costs the calling thread 2-4 us, so eight launches issued back to back from one
// thread start up to ~25 us apart -- half the run time of the n = 1e6 evaluation on eight GPUs (55 us), and the
// exchange at the kernels' tails waits for the last one.  Here every device has its own launcher: the caller's
// thread serves device 0, workers 1..N-1 serve the others, all released by one atomic generation counter, so the
// launches leave within about a microsecond of each other.  Workers spin while calls keep coming (an MCMC loop
// evaluates back to back) and go to sleep on a condition variable after ~2 ms of idleness.
#pragma once
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "nngp_common.cuh"

#if defined(__x86_64__) || defined(__i386__)
#include <immintrin.h>
static inline void nngp_cpu_relax() { _mm_pause(); }
#else
static inline void nngp_cpu_relax() { std::this_thread::yield(); }
#endif

struct nngp_group {
    std::vector<nngp_handle *> subs;  // subs[r]->px.rank == r

    // runs fn(r) for every r: r = 0 on the calling thread, the others on their workers; returns the first
    // nonzero result (in rank order)
    template <class F>
    int run(F &&fn)
    {
        const int world = int(subs.size());
        if (world == 1 || workers.empty()) {
            int rc = 0;
            for (int r = 0; r < world; ++r) {
                const int v = fn(r);
                if (v && !rc) rc = v;
            }
            return rc;
        }
        using Fn = typename std::remove_reference<F>::type;
        ctx = const_cast<void *>(static_cast<const void *>(&fn));
        tramp = [](void *c, int r) -> int { return (*static_cast<Fn *>(c))(r); };
        pending.store(world - 1, std::memory_order_relaxed);
        gen.fetch_add(1, std::memory_order_seq_cst);  // releases the spinning workers
        if (sleepers.load(std::memory_order_seq_cst) > 0) {
            std::lock_guard<std::mutex> lk(mu);
            cv.notify_all();
        }
        rcs[0] = fn(0);
        while (pending.load(std::memory_order_acquire) != 0) nngp_cpu_relax();
        for (int r = 0; r < world; ++r)
            if (rcs[r]) return rcs[r];
        return 0;
    }

    void start()
    {
        const int world = int(subs.size());
        rcs.assign(world, 0);
        for (int r = 1; r < world; ++r) workers.emplace_back([this, r] { loop(r); });
    }

    void stop()
    {
        if (workers.empty()) return;
        {
            std::lock_guard<std::mutex> lk(mu);
            quit.store(true, std::memory_order_seq_cst);
            cv.notify_all();
        }
        for (std::thread &t : workers) t.join();
        workers.clear();
    }

private:
    void loop(int r)
    {
        cudaSetDevice(subs[r]->device);  // this thread only ever drives this device
        unsigned long long seen = 0;
        for (;;) {
            auto idle0 = std::chrono::steady_clock::now();
            unsigned int spins = 0;
            while (gen.load(std::memory_order_acquire) == seen) {
                if (quit.load(std::memory_order_acquire)) return;
                nngp_cpu_relax();
                if ((++spins & 0x3ffu) == 0 &&
                    std::chrono::steady_clock::now() - idle0 > std::chrono::milliseconds(2)) {
                    std::unique_lock<std::mutex> lk(mu);
                    sleepers.fetch_add(1, std::memory_order_seq_cst);
                    cv.wait(lk, [&] { return gen.load(std::memory_order_seq_cst) != seen || quit.load(std::memory_order_seq_cst); });
                    sleepers.fetch_sub(1, std::memory_order_seq_cst);
                    idle0 = std::chrono::steady_clock::now();
                }
            }
            ++seen;
            rcs[r] = tramp(ctx, r);
            pending.fetch_sub(1, std::memory_order_release);
        }
    }

    std::vector<std::thread> workers;
    std::vector<int> rcs;
    void *ctx = nullptr;
    int (*tramp)(void *, int) = nullptr;
    std::atomic<unsigned long long> gen{0};
    std::atomic<int> pending{0};
    std::atomic<int> sleepers{0};
    std::atomic<bool> quit{false};
    std::mutex mu;
    std::condition_variable cv;
};

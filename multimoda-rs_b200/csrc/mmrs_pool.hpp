// mmrs_pool.hpp — host-side parallelism of the library (host code only; shared by mmrs_host.cpp and mmrs_sweep.cu).
#pragma once
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <deque>
#include <exception>
#include <mutex>
#include <thread>
#include <type_traits>
#include <vector>

namespace mmrs {

using std::size_t;

// Host-side parallelism over independent geometries (the reference runs its 4 pullbacks in a
// crossbeam scope, binding/entry.rs:140-203). The first exception wins and is rethrown.
// Host threads this process may use: MMRS_HOST_THREADS, else the hardware concurrency divided by the ranks that share
// the host (LOCAL_WORLD_SIZE / WORLD_SIZE under torchrun: one process per GPU must not oversubscribe the cores N-fold).
inline size_t host_thread_budget() {
    static const size_t budget = [] {
        if (const char* e = std::getenv("MMRS_HOST_THREADS"))
            if (std::atoi(e) > 0) return (size_t)std::atoi(e);
        size_t hw = std::max(1u, std::thread::hardware_concurrency());
        const char* w = std::getenv("LOCAL_WORLD_SIZE");
        if (!w) w = std::getenv("WORLD_SIZE");
        const int ranks = w ? std::atoi(w) : 1;
        return ranks > 1 ? std::max<size_t>(2, hw / (size_t)ranks) : hw;
    }();
    return budget;
}
// A persistent pool instead of a thread per call: the small clinical cases (config 1: 14 ms per call) issue dozens of
// parallel_for over a few hundred microseconds of work each, and spawning + joining up to 16 threads per call cost more
// than the work. The SUBMITTER always works on its own job and only then waits for the helpers that joined it, so nested
// calls (pullbacks -> frames) and concurrent submitters (process_cases_pipelined) cannot dead-lock; the pool is leaked on
// purpose (no static-destruction order to get wrong inside a Python process), and after a fork() a job simply finds no
// helpers and runs on its submitter.
class HostPool {
  public:
    struct Job {
        void (*call)(void*, size_t) = nullptr;
        void* fn = nullptr;
        size_t n = 0;
        std::atomic<size_t> next{0};
        int wanted = 0;   // helpers that may still join (guarded by the pool mutex)
        int active = 0;   // helpers inside the job (guarded by the pool mutex)
        std::exception_ptr err;   // of the LOWEST failing index: what a serial loop would have thrown first
        size_t err_index = ~(size_t)0;
        std::mutex err_mu;
    };
    static HostPool& get() {
        static HostPool* pool = new HostPool(host_thread_budget() > 1 ? host_thread_budget() - 1 : 0);
        return *pool;
    }
    static void work(Job& j) {
        for (;;) {
            const size_t i = j.next.fetch_add(1);
            if (i >= j.n) return;
            try {
                j.call(j.fn, i);
            } catch (...) {
                // Indices are handed out in increasing order, so every index below a failing one has been claimed and
                // runs to its end: keeping the exception of the lowest failing index reproduces the serial loop's error
                // (the reference reports the first bad frame). No new indices are handed out after a failure.
                std::lock_guard<std::mutex> lk(j.err_mu);
                if (i < j.err_index) j.err_index = i, j.err = std::current_exception();
                j.next.store(j.n);
                return;
            }
        }
    }
    void run(Job& j, int helpers) {
        if (helpers > 0 && !workers_.empty()) {
            int wanted;
            {
                std::lock_guard<std::mutex> lk(mu_);
                wanted = j.wanted = std::min<int>(helpers, (int)workers_.size());
                open_.push_back(&j);
            }
            if (wanted == 1) cv_.notify_one();
            else cv_.notify_all();
        }
        work(j);
        if (helpers > 0 && !workers_.empty()) {
            std::unique_lock<std::mutex> lk(mu_);
            for (auto it = open_.begin(); it != open_.end(); ++it)
                if (*it == &j) {
                    open_.erase(it);
                    break;
                }
            j.wanted = 0;
            done_.wait(lk, [&] { return j.active == 0; });
        }
        if (j.err) std::rethrow_exception(j.err);
    }

  private:
    explicit HostPool(size_t n) {
        for (size_t t = 0; t < n; ++t) workers_.emplace_back([this] { loop(); });
        for (auto& w : workers_) w.detach();
    }
    void loop() {
        std::unique_lock<std::mutex> lk(mu_);
        for (;;) {
            cv_.wait(lk, [&] { return !open_.empty(); });
            Job* j = open_.front();
            if (--j->wanted <= 0) open_.pop_front();
            ++j->active;
            lk.unlock();
            work(*j);
            lk.lock();
            if (--j->active == 0) done_.notify_all();
        }
    }
    std::mutex mu_;
    std::condition_variable cv_, done_;
    std::deque<Job*> open_;
    std::vector<std::thread> workers_;
};

template <class F>
void parallel_for(size_t n, F&& f, size_t max_threads = ~(size_t)0) {
    const size_t nt = std::min(max_threads, std::min<size_t>(n, host_thread_budget()));
    if (nt <= 1) {
        for (size_t i = 0; i < n; ++i) f(i);
        return;
    }
    using Fn = std::remove_reference_t<F>;
    HostPool::Job job;
    job.call = [](void* p, size_t i) { (*static_cast<Fn*>(p))(i); };
    job.fn = const_cast<void*>(static_cast<const void*>(&f));
    job.n = n;
    HostPool::get().run(job, (int)nt - 1);
}

}  // namespace mmrs

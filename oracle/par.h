// ORACLE (test infrastructure): tiny fork-join helper standing in for the reference's Rayon
// data-parallel loops (reference prover/src/prover.rs:564,700-703,785; proof.rs:312,318; and
// plonky2's per-polynomial FFT / per-leaf hashing loops).  Thread count: orc::set_threads(n),
// default = ORC_THREADS env var or hardware_concurrency.
#pragma once
#include <thread>
#include <vector>
#include <atomic>
#include <functional>
#include <cstdlib>
#include <algorithm>
#include <exception>
#include <mutex>

namespace orc {

inline int& threads_ref() {
    static int n = [] {
        const char* e = std::getenv("ORC_THREADS");
        int v = e ? std::atoi(e) : (int)std::thread::hardware_concurrency();
        return v > 0 ? v : 1;
    }();
    return n;
}
inline void set_threads(int n) { threads_ref() = n > 0 ? n : 1; }
inline int get_threads() { return threads_ref(); }

// Calls f(i) for i in [0, n) with dynamic chunking.
template <class F>
inline void parallel_for(size_t n, F&& f, size_t chunk = 0) {
    int nt = get_threads();
    if (nt <= 1 || n <= 1) {
        for (size_t i = 0; i < n; i++) f(i);
        return;
    }
    if (chunk == 0) chunk = std::max<size_t>(1, n / ((size_t)nt * 8));
    std::atomic<size_t> next(0);
    std::exception_ptr err;
    std::mutex err_mu;
    auto worker = [&] {
        try {
            for (;;) {
                size_t b = next.fetch_add(chunk);
                if (b >= n) break;
                size_t e = std::min(n, b + chunk);
                for (size_t i = b; i < e; i++) f(i);
            }
        } catch (...) {
            std::lock_guard<std::mutex> g(err_mu);
            if (!err) err = std::current_exception();
            next.store(n);
        }
    };
    std::vector<std::thread> ts;
    int spawn = (int)std::min<size_t>((size_t)nt, (n + chunk - 1) / chunk);
    for (int t = 1; t < spawn; t++) ts.emplace_back(worker);
    worker();
    for (auto& t : ts) t.join();
    if (err) std::rethrow_exception(err);
}

}  // namespace orc

// host/compat: minimal stand-in for the oneTBB subset that the TWILIGHT host sources use.
// TEST INFRASTRUCTURE ONLY — lets the *unmodified* reference sources under /root/reference compile in an
// image that has no TBB. Not part of the product path.
//
// Provided: tbb::blocked_range<T>, tbb::parallel_for (range and index overloads),
// tbb::this_task_arena::{isolate,max_concurrency}. Work is distributed over std::thread workers with a
// shared atomic cursor (dynamic schedule); nested calls run inline on the calling worker.
#pragma once
#include <algorithm>
#include <atomic>
#include <climits>
#include <cmath>
#include <cstddef>
#include <thread>
#include <vector>

namespace tbb {

namespace compat_detail {
inline int &parallelism_cap() {
    static int cap = static_cast<int>(std::max(1u, std::thread::hardware_concurrency()));
    return cap;
}
inline bool &inside_worker() {
    thread_local bool flag = false;
    return flag;
}
template <typename Index, typename Body>
void run_chunks(Index first, Index last, Index chunk, const Body &body) {
    if (last <= first) return;
    const long long total = static_cast<long long>(last) - static_cast<long long>(first);
    long long nChunks = (total + chunk - 1) / chunk;
    int workers = static_cast<int>(std::min<long long>(parallelism_cap(), nChunks));
    if (inside_worker() || workers <= 1) {
        body(first, last);
        return;
    }
    std::atomic<long long> cursor{0};
    auto loop = [&]() {
        inside_worker() = true;
        for (;;) {
            long long c = cursor.fetch_add(1);
            if (c >= nChunks) break;
            Index b = static_cast<Index>(first + c * chunk);
            Index e = static_cast<Index>(std::min<long long>(static_cast<long long>(last), static_cast<long long>(b) + chunk));
            body(b, e);
        }
        inside_worker() = false;
    };
    std::vector<std::thread> pool;
    pool.reserve(workers - 1);
    for (int w = 1; w < workers; ++w) pool.emplace_back(loop);
    loop();
    for (auto &t : pool) t.join();
}
} // namespace compat_detail

template <typename T>
class blocked_range {
  public:
    using const_iterator = T;
    blocked_range(T b, T e, std::size_t grain = 1) : b_(b), e_(e), grain_(grain) {}
    T begin() const { return b_; }
    T end() const { return e_; }
    std::size_t size() const { return static_cast<std::size_t>(e_ - b_); }
    std::size_t grainsize() const { return grain_; }
    bool empty() const { return !(b_ < e_); }

  private:
    T b_, e_;
    std::size_t grain_;
};

template <typename T, typename Body>
void parallel_for(const blocked_range<T> &range, const Body &body) {
    const long long total = static_cast<long long>(range.end()) - static_cast<long long>(range.begin());
    if (total <= 0) return;
    // Small chunks keep uneven per-item cost balanced; never below 1.
    long long chunk = std::max<long long>(1, total / (8LL * compat_detail::parallelism_cap()));
    compat_detail::run_chunks<T>(range.begin(), range.end(), static_cast<T>(chunk),
                                 [&](T b, T e) { body(blocked_range<T>(b, e)); });
}

template <typename Index, typename Body>
void parallel_for(Index first, Index last, const Body &body) {
    const long long total = static_cast<long long>(last) - static_cast<long long>(first);
    if (total <= 0) return;
    long long chunk = std::max<long long>(1, total / (8LL * compat_detail::parallelism_cap()));
    compat_detail::run_chunks<Index>(first, last, static_cast<Index>(chunk), [&](Index b, Index e) {
        for (Index i = b; i < e; ++i) body(i);
    });
}

namespace this_task_arena {
template <typename F>
void isolate(const F &f) {
    f();
}
inline int max_concurrency() { return static_cast<int>(std::max(1u, std::thread::hardware_concurrency())); }
} // namespace this_task_arena

} // namespace tbb

#pragma once
#include <mutex>
#include <queue>
namespace tbb {
template <class T>
class concurrent_queue {
    std::queue<T> q;
    std::mutex mu;
public:
    void push(const T &v) { std::lock_guard<std::mutex> g(mu); q.push(v); }
    bool try_pop(T &v) { std::lock_guard<std::mutex> g(mu); if (q.empty()) return false; v = q.front(); q.pop(); return true; }
    bool empty() const { return q.empty(); }
    size_t unsafe_size() const { return q.size(); }
};
}

#pragma once
#include <mutex>
#include <unordered_set>
namespace tbb {
template <class K, class H = std::hash<K>, class E = std::equal_to<K>>
class concurrent_unordered_set {
    std::unordered_set<K, H, E> s;
    std::mutex mu;
public:
    typedef typename std::unordered_set<K, H, E>::iterator iterator;
    typedef typename std::unordered_set<K, H, E>::const_iterator const_iterator;
    concurrent_unordered_set() {}
    concurrent_unordered_set(const concurrent_unordered_set &o) : s(o.s) {}
    concurrent_unordered_set &operator=(const concurrent_unordered_set &o) { s = o.s; return *this; }
    std::pair<iterator, bool> insert(const K &k) { std::lock_guard<std::mutex> g(mu); return s.insert(k); }
    size_t size() const { return s.size(); }
    size_t count(const K &k) const { return s.count(k); }
    iterator find(const K &k) { return s.find(k); }
    iterator begin() { return s.begin(); }
    iterator end() { return s.end(); }
    const_iterator begin() const { return s.begin(); }
    const_iterator end() const { return s.end(); }
    void clear() { s.clear(); }
    void swap(concurrent_unordered_set &o) { s.swap(o.s); }
};
}

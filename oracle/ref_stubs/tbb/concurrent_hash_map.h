// mutex-protected std::unordered_map with the accessor interface the reference uses (see ../README.md)
#pragma once
#include <functional>
#include <mutex>
#include <unordered_map>
#include <utility>
namespace tbb {
template <class K> struct tbb_hash_compare {
    static size_t hash(const K &k) { return std::hash<K>()(k); }
    static bool equal(const K &a, const K &b) { return a == b; }
};
template <class K, class V, class HC = tbb_hash_compare<K>>
class concurrent_hash_map {
    struct H { size_t operator()(const K &k) const { return HC::hash(k); } };
    struct E { bool operator()(const K &a, const K &b) const { return HC::equal(a, b); } };
    typedef std::unordered_map<K, V, H, E> map_t;
    map_t m;
    std::recursive_mutex mu;
public:
    typedef typename map_t::value_type value_type;
    typedef typename map_t::iterator iterator;
    typedef typename map_t::const_iterator const_iterator;
    // like TBB's accessor, it keeps the element locked for its lifetime (coarsely: the whole map), so that the
    // reference's "find / insert, then modify through the accessor" sequences stay safe under its OpenMP loops
    class accessor {
        friend class concurrent_hash_map;
        value_type *p = nullptr;
        std::unique_lock<std::recursive_mutex> lk;
    public:
        accessor() {}
        accessor(const accessor &) = delete;
        value_type &operator*() const { return *p; }
        value_type *operator->() const { return p; }
        bool empty() const { return p == nullptr; }
        void release() { p = nullptr; if (lk.owns_lock()) lk.unlock(); }
    };
    typedef accessor const_accessor;
    bool insert(accessor &a, const K &k) {
        a.release();
        a.lk = std::unique_lock<std::recursive_mutex>(mu);
        auto r = m.emplace(k, V());
        a.p = &*r.first;
        return r.second;
    }
    bool insert(accessor &a, const value_type &v) {
        a.release();
        a.lk = std::unique_lock<std::recursive_mutex>(mu);
        auto r = m.insert(v);
        a.p = &*r.first;
        return r.second;
    }
    bool insert(const value_type &v) { std::lock_guard<std::recursive_mutex> g(mu); return m.insert(v).second; }
    bool find(accessor &a, const K &k) {
        a.release();
        a.lk = std::unique_lock<std::recursive_mutex>(mu);
        auto it = m.find(k);
        if (it == m.end()) { a.release(); return false; }
        a.p = &*it;
        return true;
    }
    bool erase(const K &k) { std::lock_guard<std::recursive_mutex> g(mu); return m.erase(k) > 0; }
    size_t count(const K &k) { std::lock_guard<std::recursive_mutex> g(mu); return m.count(k); }
    size_t size() const { return m.size(); }
    bool empty() const { return m.empty(); }
    void clear() { m.clear(); }
    void swap(concurrent_hash_map &o) { m.swap(o.m); }
    iterator begin() { return m.begin(); }
    iterator end() { return m.end(); }
    const_iterator begin() const { return m.begin(); }
    const_iterator end() const { return m.end(); }
};
}  // namespace tbb

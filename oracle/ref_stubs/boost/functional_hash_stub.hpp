#pragma once
#include <functional>
#include <utility>
namespace boost {
template <class T> struct hash : std::hash<T> {};
template <class A, class B> struct hash<std::pair<A, B>> {
    size_t operator()(const std::pair<A, B> &p) const {
        size_t h = std::hash<A>()(p.first);
        return h ^ (std::hash<B>()(p.second) + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2));
    }
};
}

#pragma once
#include <mutex>
#include <unordered_map>
namespace tbb {
template <class K, class V, class H = std::hash<K>, class E = std::equal_to<K>>
class concurrent_unordered_map : public std::unordered_map<K, V, H, E> {
public:
    using std::unordered_map<K, V, H, E>::unordered_map;
};
}

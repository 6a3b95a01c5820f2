#pragma once
#include <unordered_map>
#include "functional_hash_stub.hpp"
namespace boost { template <class K, class V, class H = boost::hash<K>, class E = std::equal_to<K>> using unordered_map = std::unordered_map<K, V, H, E>; }

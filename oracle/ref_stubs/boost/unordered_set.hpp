#pragma once
#include <unordered_set>
#include "functional_hash_stub.hpp"
namespace boost { template <class K, class H = boost::hash<K>, class E = std::equal_to<K>> using unordered_set = std::unordered_set<K, H, E>; }

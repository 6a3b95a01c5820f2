// Shadow of core/bind.hpp (hwloc core binding): not on the path, nothing to bind in a test shim.
#pragma once
#include <vector>
static inline void load_node_topo() {}
static inline void bind_to_core(int) {}
extern std::vector<std::vector<int>> core_bindings;

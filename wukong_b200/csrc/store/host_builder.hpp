// Product-side builder of the cluster-hash graph store (host C++, OpenMP).
//
// Produces the same data structure the reference's loader + StaticGStore produce
// (core/loader/base_loader.hpp:308-378, core/store/static_gstore.hpp:383-454,
// core/store/gstore.hpp:428-472, 529-888): 16-byte key/pointer slots in 8-way buckets with the
// last slot as chain pointer, a flat 4-byte edge array, and one metadata record per
// (index, pid, dir) segment.  Only what a probe can observe is contractual (key -> edge run, run
// order = sorted ids); slot placement is free.  The implementation is sort-based and parallel over
// segments rather than the reference's per-thread triple lists + TBB maps.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "wukong_b200.h"

namespace wkhost {

struct StoreBuildOptions {
    int num_servers = 1;           // Global::num_servers: vid % num_servers sharding (utils/math.hpp:51-55)
    int sid = 0;
    int num_normal_preds = 0;      // (#lines of str_index) - 1, base_loader.hpp:409-424
    uint64_t kvstore_bytes = 0;    // Global::memstore_size_gb worth of bytes; 0 = size from the data
    int est_load_factor = 55;      // Global::est_load_factor, used when kvstore_bytes == 0
    bool gpu_ext_extents = true;   // one 15% extent per segment (USE_GPU build, store/meta.hpp:38-40)
};

struct HostStore {
    std::vector<wk_vertex_t> vertices;   // num_slots
    std::vector<wk_sid_t> edges;         // used entries only
    std::vector<wk_segmeta_t> segs;
    uint64_t num_buckets = 0, num_buckets_ext = 0, used_ext = 0;
    uint64_t num_keys = 0, num_triples_out = 0, num_triples_in = 0;
    std::string error;
    bool ok() const { return error.empty(); }
    // host-side probe (debug / checks); returns pointer into edges or nullptr
    const wk_sid_t *get_edges(wk_sid_t vid, wk_sid_t pid, int dir, uint64_t &size) const;
};

// triples: n x (s, p, o) uint32
void build_store(const wk_sid_t *triples, uint64_t n, const StoreBuildOptions &opt, HostStore &out);

}  // namespace wkhost

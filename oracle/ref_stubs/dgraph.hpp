// Shadow of core/dgraph.hpp for oracle/_ref (see README.md): the real header also pulls the POSIX / HDFS loaders (Boost.MPI,
// libhdfs).  The engine only calls the two one-line forwards below (dgraph.hpp:106-112), restated over the REAL GStore.
#pragma once
#include "store/static_gstore.hpp"
#include "string_server.hpp"
class DGraph {
public:
    int sid;
    GStore *gstore;
    DGraph(int sid, GStore *g) : sid(sid), gstore(g) {}
    edge_t *get_triples(int tid, sid_t vid, sid_t pid, dir_t d, uint64_t &sz) { return gstore->get_edges(tid, vid, pid, d, sz); }
    edge_t *get_index(int tid, sid_t pid, dir_t d, uint64_t &sz) { return gstore->get_edges(tid, 0, pid, d, sz); }
    // attribute values are not on this path
    attr_t get_attr(int tid, sid_t vid, sid_t pid, dir_t d, bool &has_value) { has_value = false; return attr_t(); }
    int gstore_check(bool, bool) { return 0; }
    int dynamic_load_data(std::string, bool) { return 0; }
};

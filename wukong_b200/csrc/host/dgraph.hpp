// Dataset loading + graph store of one server (the surface of reference core/dgraph.hpp:55-112 and
// core/loader/base_loader.hpp:386-469): reads an ID-triple directory (id_*.nt, str_index), builds the
// cluster-hash store with the product builder and uploads it to the GPU through the C ABI.
#pragma once
#include <dirent.h>

#include <algorithm>
#include <cstdio>
#include <fstream>
#include <string>
#include <vector>

#include "global.hpp"
#include "store/host_builder.hpp"
#include "string_server.hpp"
#include "type.hpp"
#include "wukong_b200.h"

namespace wukong {

class DGraph {
public:
    int sid;
    wkhost::HostStore store;
    wk_store_t *gstore = nullptr;
    uint64_t num_triples = 0;
    int num_normal_preds = 0;
    std::string error;

    // files whose name starts with `prefix`, sorted (base_loader.hpp: list_files + sort)
    static std::vector<std::string> list_files(const std::string &dname, const std::string &prefix) {
        std::vector<std::string> out;
        if (DIR *d = opendir(dname.c_str())) {
            while (dirent *e = readdir(d)) {
                std::string n = e->d_name;
                if (n.compare(0, prefix.size(), prefix) == 0) out.push_back(dname + n);
            }
            closedir(d);
        }
        std::sort(out.begin(), out.end());
        return out;
    }
    static int count_lines(const std::string &fname) {
        std::ifstream f(fname.c_str());
        std::string a;
        int n = 0;
        while (std::getline(f, a))
            if (!a.empty()) n++;
        return n;
    }

    DGraph(int sid, const Global &g, int device) : sid(sid) {
        std::string dir = g.input_folder;
        std::vector<sid_t> triples;
        for (const std::string &fn : list_files(dir, "id_")) {
            FILE *f = fopen(fn.c_str(), "r");
            if (!f) continue;
            unsigned s, p, o;
            while (fscanf(f, "%u %u %u", &s, &p, &o) == 3) { triples.push_back(s); triples.push_back(p); triples.push_back(o); }
            fclose(f);
        }
        num_triples = triples.size() / 3;
        const int lines = count_lines(dir + "str_index");
        if (lines == 0) { error = "Encoding file of predicates should be named as \"str_index\"."; return; }
        num_normal_preds = lines - 1;   // skip PREDICATE_ID
        wkhost::StoreBuildOptions opt;
        opt.num_servers = g.num_servers;
        opt.sid = sid;
        opt.num_normal_preds = num_normal_preds;
        opt.est_load_factor = g.est_load_factor;
        opt.kvstore_bytes = 0;   // sized from the data; Global::memstore_size_gb is an upper bound on real deployments
        wkhost::build_store(triples.data(), num_triples, opt, store);
        if (!store.ok()) { error = store.error; return; }
        if (device >= 0) {
            int rc = wk_store_create(device, store.vertices.data(), store.vertices.size(), store.edges.data(), store.edges.size(),
                                     store.segs.data(), (int)store.segs.size(), &gstore);
            if (rc) error = std::string("wk_store_create: ") + wk_strerror(rc);
        }
    }
    ~DGraph() { if (gstore) wk_store_destroy(gstore); }
    bool ok() const { return error.empty(); }
};

}  // namespace wukong

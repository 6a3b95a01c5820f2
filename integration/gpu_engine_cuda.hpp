// Replacement for core/gpu/gpu_engine_cuda.hpp of SJTU-IPADS/wukong (the class at gpu_engine_cuda.hpp:45-409): same name,
// same constructor, same member functions -- core/gpu/gpu_engine.hpp and gpu_agent.hpp compile against it unchanged -- but
// every body is a call into libwukong_b200.so (include/wukong_b200.h).  core/gpu/gpu_hash.{hpp,cu} are no longer needed.
//
// One line has to be added to the reference: `friend class GPUEngineCuda;` inside class GPUCache (gpu_cache.hpp:46), so that
// this class can take the host store arrays and the segment table GPUCache was constructed with (gpu_cache.hpp:425-427).
// GPUCache's block cache itself, GPUMem's result buffers and GPUStreamPool are not used by the query path any more: the store
// lives in flat HBM arrays inside wk_store_t, the double result buffer inside wk_engine_t.
//
// oracle/ref_gpu_engine_shim.cpp compiles this file against the reference's real headers (tests/test_integration_binding.py).
#pragma once

#ifdef USE_GPU

#include <vector>
#include <utility>

#include "global.hpp"
#include "assertion.hpp"
#include "errors.hpp"
#include "query.hpp"
#include "unit.hpp"

#include "gpu_cache.hpp"
#include "gpu_stream.hpp"

#include "wukong_b200.h"

class GPUEngineCuda final {
private:
    int sid;
    wk_store_t *store = nullptr;
    wk_engine_t *engine = nullptr;
    std::vector<sid_t> pending;      // a history table received as bytes (load_result_buf(const char *, size)): its column
                                     // count is only known when the next pattern arrives

    // error convention of the engine code: utils/assertion.hpp:91-98 (caught at sparql.hpp:1663-1667)
    static void check(int rc) {
        if (rc == WK_SUCCESS) return;
        logstream(LOG_ERROR) << "wukong_b200: " << wk_strerror(rc) << LOG_endl;
        throw WukongException(rc >= 0 && rc <= 12 ? rc : UNKNOWN_ERROR);
    }

    bool has_next_pattern(const SPARQLQuery &req) {
        return req.pattern_step + 1 < req.pattern_group.patterns.size();
    }

    void upload_pending(SPARQLQuery &req) {
        if (pending.empty()) return;
        const int cols = req.result.get_col_num();
        ASSERT(cols > 0 && pending.size() % cols == 0);
        check(wk_table_upload(engine, pending.data(), pending.size() / cols, cols));
        pending.clear();
    }

    // row bookkeeping of Result::gpu (query.hpp:257-308); the table itself stays inside the engine until the last pattern
    void finish_step(SPARQLQuery &req, uint64_t rows, int cols, std::vector<sid_t> &new_table) {
        req.result.gpu.set_rbuf((char *)engine, rows * (uint64_t)cols);
        if (has_next_pattern(req)) return;
        new_table.resize(rows * (uint64_t)cols);
        if (new_table.empty()) return;
        int c = 0;
        check(wk_table_download(engine, new_table.data(), new_table.size(), &rows, &c));
        ASSERT(c == cols);
    }

public:
    GPUEngineCuda(int sid, GPUCache *gcache, GPUMem *gmem, GPUStreamPool *stream_pool) : sid(sid) {
        // flat copy of the segment table (gstore.hpp:174): one wk_segmeta_t per (segid_t, rdf_seg_meta_t) pair
        std::vector<wk_segmeta_t> segs;
        uint64_t nbuckets = 0, nentries = 0;
        for (auto &kv : gcache->rdf_metas) {
            wk_segmeta_t m;
            memset(&m, 0, sizeof(m));
            m.index = kv.first.index; m.dir = kv.first.dir; m.pid = kv.first.pid;
            m.num_keys = kv.second.num_keys; m.num_buckets = kv.second.num_buckets;
            m.bucket_start = kv.second.bucket_start;
            m.num_edges = kv.second.num_edges; m.edge_start = kv.second.edge_start;
            nbuckets = std::max(nbuckets, m.bucket_start + m.num_buckets);
            if (kv.second.get_ext_bucket_list_size() > 0) {      // the probe follows chain pointers; the extents only bound the array
                m.ext_start = kv.second.ext_bucket_list[0].start;
                m.ext_num = kv.second.ext_bucket_list[0].num_ext_buckets;
                for (size_t i = 0; i < kv.second.get_ext_bucket_list_size(); i++)
                    nbuckets = std::max(nbuckets, (uint64_t)kv.second.ext_bucket_list[i].start + kv.second.ext_bucket_list[i].num_ext_buckets);
            }
            nentries = std::max(nentries, m.edge_start + m.num_edges);
            segs.push_back(m);
        }
        int dev = 0;
        cudaGetDevice(&dev);
        // vertex_t == wk_vertex_t (16 bytes: ikey_t | iptr_t), edge_t == wk_sid_t
        check(wk_store_create(dev, (const wk_vertex_t *)gcache->vertex_addr, nbuckets * GStore::ASSOCIATIVITY,
                              (const wk_sid_t *)gcache->edge_addr, nentries, segs.data(), (int)segs.size(), &store));
        check(wk_engine_create(store, MiB2B(Global::gpu_rbuf_size_mb), &engine));
    }

    ~GPUEngineCuda() {
        wk_engine_destroy(engine);
        wk_store_destroy(store);
    }

    // the token returned stands for "the table is inside the engine": Result::gpu only tests it against nullptr
    char *load_result_buf(const SPARQLQuery::Result &r) {
        const int cols = const_cast<SPARQLQuery::Result &>(r).get_col_num();
        pending.clear();
        if (cols > 0 && !r.result_table.empty())      // an empty table: GPUEngine skips the backend (Result::gpu.is_rbuf_empty)
            check(wk_table_upload(engine, r.result_table.data(), r.result_table.size() / cols, cols));
        return (char *)engine;
    }

    char *load_result_buf(const char *rbuf, uint64_t size) {
        pending.assign((const sid_t *)rbuf, (const sid_t *)(rbuf + size));
        return (char *)engine;
    }

    vector<sid_t> index_to_unknown(SPARQLQuery &req, sid_t tpid, dir_t d) {
        upload_pending(req);
        uint64_t rows = 0;
        check(wk_index_to_unknown(engine, tpid, d, req.mt_tid % req.mt_factor, req.mt_factor, &rows));
        vector<sid_t> out(rows);
        int c = 0;
        if (rows) check(wk_table_download(engine, out.data(), out.size(), &rows, &c));
        return out;
    }

    void known_to_unknown(SPARQLQuery &req, ssid_t start, ssid_t pid, dir_t d, vector<sid_t> &new_table) {
        upload_pending(req);
        uint64_t rows = 0;
        check(wk_known_to_unknown(engine, req.result.var2col(start), pid, d, &rows));
        finish_step(req, rows, req.result.get_col_num() + 1, new_table);
    }

    void known_to_known(SPARQLQuery &req, ssid_t start, sid_t pid, ssid_t end, dir_t d, vector<sid_t> &new_table) {
        upload_pending(req);
        uint64_t rows = 0;
        check(wk_known_to_known(engine, req.result.var2col(start), pid, d, req.result.var2col(end), &rows));
        finish_step(req, rows, req.result.get_col_num(), new_table);
    }

    void known_to_const(SPARQLQuery &req, ssid_t start, ssid_t pid, ssid_t end, dir_t d, vector<sid_t> &new_table) {
        upload_pending(req);
        uint64_t rows = 0;
        check(wk_known_to_const(engine, req.result.var2col(start), pid, d, end, &rows));
        finish_step(req, rows, req.result.get_col_num(), new_table);
    }

    // fork-join split (gpu_engine_cuda.hpp:364-407): rows bucketised by row[var2col(start)] % num_jobs; the device run of every
    // destination stays valid until the next call into the engine (send_dev2host copies it out, gpu_agent.hpp:94-110)
    void generate_sub_query(SPARQLQuery &req, sid_t start, int num_jobs, vector<sid_t *> &buf_ptrs, vector<int> &buf_sizes) {
        upload_pending(req);
        std::vector<uint64_t> part_rows(num_jobs, 0);
        check(wk_partition(engine, req.result.var2col(start), num_jobs, part_rows.data()));
        const int cols = req.result.get_col_num();
        for (int i = 0; i < num_jobs; i++) {
            const wk_sid_t *p = nullptr;
            uint64_t rows = 0;
            check(wk_partition_ptr(engine, i, &p, &rows));
            buf_ptrs[i] = (sid_t *)p;
            buf_sizes[i] = (int)(rows * (uint64_t)cols);
        }
    }
};

#endif  // USE_GPU

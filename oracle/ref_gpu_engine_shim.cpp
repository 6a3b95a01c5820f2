// TEST INFRASTRUCTURE ONLY.  Compile-and-link check of the binding INTEGRATION.md describes: the reference's OWN GPU engine
// (core/gpu/gpu_engine.hpp: GPUEngine with its dispatch switch, index_to_unknown / const_to_unknown on the CPU and the
// known_to_* patterns handed to `GPUEngineCuda backend`), its GPUMem / GPUCache / GPUStreamPool and query.hpp under -DUSE_GPU,
// all unmodified, over integration/gpu_engine_cuda.hpp -- the replacement GPUEngineCuda whose bodies call libwukong_b200.so.
//
// How the replacement gets in without editing the reference: gpu_engine.hpp includes "gpu_engine_cuda.hpp" from its own
// directory, so the reference's file is included first with USE_GPU switched off (its whole body sits inside #ifdef USE_GPU:
// nothing is compiled, and #pragma once marks it as seen), then the replacement is included, then gpu_engine.hpp.
// `friend class GPUEngineCuda;` in GPUCache (the one line a maintainer adds) is granted here by ref_shim.h's access macro.
//
// refg_query drives GPUEngine::execute_one_pattern the way GPUAgent::execute_sparql_query does (gpu_agent.hpp:203-217); it
// needs a GPU (GPUMem allocates device memory) and is exercised by an opt-in test only -- this round's GPU budget was spent
// before it could be run there, so what the CPU suite holds is: this file compiles against the reference's headers, links
// against libwukong_b200.so and exports refg_query.
#define USE_GPU
#define WK_REF_WITH_ENGINE 1
#include "ref_shim.h"
#include "engine/msgr.hpp"
#include <list>
#include <cuda_runtime.h>
#define private public            // = `friend class GPUEngineCuda;` in GPUCache
#include "gpu/gpu_mem.hpp"
#include "gpu/gpu_cache.hpp"
#include "gpu/gpu_stream.hpp"
#undef private
#undef USE_GPU
#include "gpu/gpu_engine_cuda.hpp"      // the reference's class: compiled out, marked as seen
#define USE_GPU
#include "../integration/gpu_engine_cuda.hpp"
#include "gpu/gpu_engine.hpp"

extern "C" {

// planned patterns through the reference's GPUEngine over the replacement backend; blind: row count only.
// Returns the reference's status code (utils/errors.hpp), -1 when `out` is too small, -2 for a CUDA failure at start-up.
int refg_query(void *h, const int32_t *pats, int npat, int nvars, const int32_t *required, int nreq, int blind, int rbuf_mb,
               uint32_t *out, uint64_t cap_words, uint64_t *rows, int *cols) {
    RefStore *r = (RefStore *)h;
    Global::num_servers = 1;
    Global::num_gpus = 1;
    Global::gpu_kvcache_size_gb = 1;              // GPUCache's block cache is constructed but never loaded
    Global::gpu_rbuf_size_mb = rbuf_mb > 0 ? rbuf_mb : 64;
    Global::gpu_rdma_buf_size_mb = 0;
    Global::gpu_enable_pipeline = false;
    *rows = 0;
    *cols = 0;
    try {
        GPUMem gmem(0, 1, 1);
        GPUCache gcache(&gmem, r->g->vertices, r->g->edges, r->g->rdf_seg_meta_map);
        GPUStreamPool pool(4);
        DGraph graph(0, r->g);
        GPUEngine eng(0, 0, &gmem, &gcache, &pool, &graph);
        SPARQLQuery::PatternGroup pg;
        for (int i = 0; i < npat; i++)
            pg.patterns.push_back(SPARQLQuery::Pattern((ssid_t)pats[4 * i], (ssid_t)pats[4 * i + 1], (ssid_t)pats[4 * i + 2], (ssid_t)pats[4 * i + 3]));
        std::vector<ssid_t> req(required, required + nreq);
        SPARQLQuery q(pg, nvars, req);
        q.result.blind = blind != 0;
        q.dev_type = SPARQLQuery::DeviceType::GPU;
        if (npat == 0) throw WukongException(SYNTAX_ERROR);
        while (!q.done(SPARQLQuery::SQState::SQ_PATTERN)) {
            if (!eng.result_buf_ready(q)) eng.load_result_buf(q);
            eng.execute_one_pattern(q);
        }
        q.result.update_nrows();
        *rows = (uint64_t)q.result.get_row_num();
        *cols = q.result.get_col_num();
        const uint64_t words = q.result.result_table.size();
        if (!q.result.blind && out) {
            if (words > cap_words) return -1;
            memcpy(out, q.result.result_table.data(), words * sizeof(uint32_t));
        }
    } catch (WukongException &ex) {
        return ex.code();
    }
    return SUCCESS;
}

}  // extern "C"

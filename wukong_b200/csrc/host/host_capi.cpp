// C API of the host-side layer (libwukong_host.so) for ctypes: store builder, dataset reader,
// and the Wukong-surface mirror (query / planner / engine / proxy).  Everything GPU goes through
// the C ABI in include/wukong_b200.h.
#include <chrono>
#include <cstdint>
#include <cstring>
#include <string>

#include "store/host_builder.hpp"
#include "wukong_b200.h"

using namespace wkhost;

extern "C" {

void *wkh_store_build(const uint32_t *triples, uint64_t n, int num_servers, int sid, int num_normal_preds,
                      uint64_t kvstore_bytes, int est_load_factor, int gpu_ext_extents) {
    StoreBuildOptions opt;
    opt.num_servers = num_servers;
    opt.sid = sid;
    opt.num_normal_preds = num_normal_preds;
    opt.kvstore_bytes = kvstore_bytes;
    opt.est_load_factor = est_load_factor > 0 ? est_load_factor : 55;
    opt.gpu_ext_extents = gpu_ext_extents != 0;
    HostStore *st = new HostStore();
    build_store(triples, n, opt, *st);
    return st;
}
void wkh_store_free(void *h) { delete (HostStore *)h; }
int wkh_store_ok(void *h) { return ((HostStore *)h)->ok() ? 1 : 0; }
const char *wkh_store_error(void *h) { return ((HostStore *)h)->error.c_str(); }
const void *wkh_store_vertices(void *h) { return ((HostStore *)h)->vertices.data(); }
uint64_t wkh_store_num_slots(void *h) { return ((HostStore *)h)->vertices.size(); }
const uint32_t *wkh_store_edges(void *h) { return ((HostStore *)h)->edges.data(); }
uint64_t wkh_store_num_edges(void *h) { return ((HostStore *)h)->edges.size(); }
uint64_t wkh_store_num_keys(void *h) { return ((HostStore *)h)->num_keys; }
uint64_t wkh_store_num_buckets(void *h) { return ((HostStore *)h)->num_buckets; }
uint64_t wkh_store_used_ext(void *h) { return ((HostStore *)h)->used_ext; }
int wkh_store_num_segs(void *h) { return (int)((HostStore *)h)->segs.size(); }
void wkh_store_segs(void *h, wk_segmeta_t *out) {
    HostStore *st = (HostStore *)h;
    memcpy(out, st->segs.data(), st->segs.size() * sizeof(wk_segmeta_t));
}
uint64_t wkh_store_get_edges(void *h, uint32_t vid, uint32_t pid, int dir, const uint32_t **out) {
    uint64_t sz = 0;
    const uint32_t *p = ((HostStore *)h)->get_edges(vid, pid, dir, sz);
    if (out) *out = p;
    return sz;
}
// upload to the GPU through the C ABI (wk_store_create)
int wkh_store_upload(void *h, int device, wk_store_t **out) {
    HostStore *st = (HostStore *)h;
    if (!st->ok()) return WK_ERR_BAD_ARG;
    return wk_store_create(device, st->vertices.data(), st->vertices.size(), st->edges.data(), st->edges.size(),
                           st->segs.data(), (int)st->segs.size(), out);
}

// Timed loop around the public C-ABI call, with no interpreter between iterations.
//   wall_us[i]  host wall clock around wk_query_execute (host buffers, copies inside the region)
//   dev_us[i]   CUDA-event time of the same call on the engine's stream (needs profiling >= 1)
// flush != 0 evicts the L2 before every iteration (outside the timed region).
int wkh_time_query(wk_engine_t *e, const wk_pattern_t *pats, int npat, int nvars, const int32_t *req, int nreq,
                   int mt_tid, int mt_factor, int blind, wk_sid_t *table, uint64_t cap_words, int reps, int flush,
                   double *wall_us, float *dev_us, uint64_t *rows, int *cols) {
    for (int i = 0; i < reps; i++) {
        if (flush) {
            int rc = wk_engine_flush_l2(e);
            if (rc) return rc;
        }
        const auto t0 = std::chrono::steady_clock::now();
        int rc = wk_query_execute(e, pats, npat, nvars, req, nreq, mt_tid, mt_factor, blind, table, cap_words, rows, cols);
        const auto t1 = std::chrono::steady_clock::now();
        if (rc) return rc;
        if (wall_us) wall_us[i] = std::chrono::duration<double, std::micro>(t1 - t0).count();
        if (dev_us) {
            rc = wk_engine_last_query_device_us(e, &dev_us[i]);
            if (rc) return rc;
        }
    }
    return WK_SUCCESS;
}

}  // extern "C"

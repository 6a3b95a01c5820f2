// C API of the host-side layer (libwukong_host.so) for ctypes: store builder, dataset reader,
// and the Wukong-surface mirror (query / planner / engine / proxy).  Everything GPU goes through
// the C ABI in include/wukong_b200.h.
#include <chrono>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

#include <sstream>

#include "host/bundle.hpp"
#include "host/dgraph.hpp"
#include "host/global.hpp"
#include "host/gpu_engine.hpp"
#include "host/proxy.hpp"
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

// One timed collective query of a sharded group, called by every rank of the group at the same time.  L2 flush and stream
// sync first, then a spin barrier on `slots` (shared memory, one generation counter per rank, `stride` int64 apart; release skew
// well below a microsecond), then the clock around wk_query_execute_sharded.  resident: the call was answered by the resident
// servers (no launch to bracket with events: dev_us = wall_us then); server_ns: in-kernel span of this rank's server request.
int wkh_time_query_sharded(wk_engine_t *e, const wk_pattern_t *pats, int npat, int nvars, const int32_t *req, int nreq, int blind,
                           wk_sid_t *table, uint64_t cap_words, int flush, volatile int64_t *slots, int stride, int rank, int world,
                           int64_t gen, double *wall_us, float *dev_us, uint64_t *rows, int *cols, int *resident, uint64_t *server_ns) {
    if (flush) {
        int rc = wk_engine_flush_l2(e);
        if (rc) return rc;
    }
    int rc = wk_engine_sync(e);
    if (rc) return rc;
    if (slots) {
        slots[(size_t)rank * stride] = gen;
        __atomic_thread_fence(__ATOMIC_SEQ_CST);
        const auto tb = std::chrono::steady_clock::now();
        uint64_t spins = 0;
        for (int r = 0; r < world; r++) {
            while (slots[(size_t)r * stride] < gen) {
                if ((++spins & 0xFFFFF) == 0 && std::chrono::duration<double>(std::chrono::steady_clock::now() - tb).count() > 120.0) return WK_ERR_COMM;
            }
        }
    }
    const auto t0 = std::chrono::steady_clock::now();
    rc = wk_query_execute_sharded(e, pats, npat, nvars, req, nreq, 0, 1, blind, table, cap_words, rows, cols);
    const auto t1 = std::chrono::steady_clock::now();
    if (rc) return rc;
    const double w = std::chrono::duration<double, std::micro>(t1 - t0).count();
    if (wall_us) *wall_us = w;
    int64_t res = 0, ns = 0;
    wk_engine_get_option(e, WK_INFO_LAST_RESIDENT, &res);
    wk_engine_get_option(e, WK_INFO_LAST_RESIDENT_NS, &ns);
    if (resident) *resident = (int)res;
    if (server_ns) *server_ns = res ? (uint64_t)ns : 0;
    if (dev_us) {
        if (res) *dev_us = (float)w;
        else if (wk_engine_last_query_device_us(e, dev_us) != WK_SUCCESS) *dev_us = -1.0f;   // profiling is off: no event pair
    }
    return WK_SUCCESS;
}

// ---- Wukong-surface environment: config + string server + graph + engine + proxy of one server ----
struct HostEnv {
    wukong::Global global;
    wukong::StringServer str_server;
    wukong::DGraph *graph = nullptr;
    wukong::GPUEngine *engine = nullptr;
    wukong::Proxy *proxy = nullptr;
    std::string error;
    ~HostEnv() { delete proxy; delete engine; delete graph; }
};

// config_text: "global_* value" lines (must contain global_input_folder).  device < 0: host-only
// (config / string server / parser / planner / store build are usable without a GPU).
void *wkh_env_create(const char *config_text, int device) {
    HostEnv *env = new HostEnv();
    std::string bad;
    if (!env->global.load_str(config_text ? config_text : "", &bad)) { env->error = "bad config item: " + bad; return env; }
    if (env->global.input_folder.empty()) { env->error = "global_input_folder is not set"; return env; }
    if (!env->str_server.load(env->global.input_folder)) { env->error = "cannot read str_index in " + env->global.input_folder; return env; }
    env->graph = new wukong::DGraph(0, env->global, device);
    if (!env->graph->ok()) { env->error = env->graph->error; return env; }
    if (device >= 0) {
        env->engine = new wukong::GPUEngine(0, env->graph, env->global);
        if (!env->engine->ok()) { env->error = env->engine->error; return env; }
    }
    env->proxy = new wukong::Proxy(0, 0, &env->str_server, env->engine, &env->global);
    return env;
}
void wkh_env_destroy(void *h) { delete (HostEnv *)h; }
const char *wkh_env_error(void *h) { return ((HostEnv *)h)->error.c_str(); }
uint64_t wkh_env_num_triples(void *h) { HostEnv *e = (HostEnv *)h; return e->graph ? e->graph->num_triples : 0; }
int wkh_env_num_normal_preds(void *h) { HostEnv *e = (HostEnv *)h; return e->graph ? e->graph->num_normal_preds : 0; }
uint64_t wkh_env_num_keys(void *h) { HostEnv *e = (HostEnv *)h; return e->graph ? e->graph->store.num_keys : 0; }
int wkh_env_config_int(void *h, const char *key) {
    wukong::Global &g = ((HostEnv *)h)->global;
    std::string k(key);
    if (k == "global_num_engines") return g.num_engines;
    if (k == "global_mt_threshold") return g.mt_threshold;
    if (k == "global_silent") return g.silent;
    if (k == "global_est_load_factor") return g.est_load_factor;
    if (k == "global_gpu_rbuf_size_mb") return g.gpu_rbuf_size_mb;
    if (k == "global_enable_planner") return g.enable_planner;
    if (k == "global_num_threads") return g.num_threads;
    return -1;
}

// ---- wire codec (csrc/host/bundle.hpp) through a flat int description of a query, for tests ------------------------------
// flat = 19 scalars (qid .. distinct in save() order), group, norders, (id, descending)*, col_num, row_num, attr_col_num,
//        status_code, blind, nvars, nreq, req*, nv2c, v2c*, nmatched, matched*, ntable, table*, gpu_nelems, gpu_col_num
// group = npat, (s, p, d, o, pred_type)*, nnewvars, newvars*, nunions, group*, noptional, group*
static bool flat_read_group(const int64_t *&p, const int64_t *end, wukong::WireQuery::Group &g, int depth = 0) {
    if (depth > 16 || p >= end) return false;
    int64_t n = *p++;
    if (n < 0 || p + 5 * n > end) return false;
    for (int64_t i = 0; i < n; i++, p += 5) {
        wukong::SPARQLQuery::Pattern pt((wukong::ssid_t)p[0], (wukong::ssid_t)p[1], (wukong::dir_t)p[2], (wukong::ssid_t)p[3]);
        pt.pred_type = (char)p[4];
        g.patterns.push_back(pt);
    }
    if (p >= end) return false;
    n = *p++;
    if (n < 0 || p + n > end) return false;
    for (int64_t i = 0; i < n; i++) g.optional_new_vars.insert((wukong::ssid_t)*p++);
    for (int kind = 0; kind < 2; kind++) {
        if (p >= end) return false;
        n = *p++;
        if (n < 0 || n > 64) return false;
        for (int64_t i = 0; i < n; i++) {
            wukong::WireQuery::Group sub;
            if (!flat_read_group(p, end, sub, depth + 1)) return false;
            (kind == 0 ? g.unions : g.optional).push_back(sub);
        }
    }
    return true;
}
static void flat_write_group(const wukong::WireQuery::Group &g, std::vector<int64_t> &o) {
    o.push_back((int64_t)g.patterns.size());
    for (const auto &pt : g.patterns) { o.push_back(pt.subject); o.push_back(pt.predicate); o.push_back(pt.direction); o.push_back(pt.object); o.push_back(pt.pred_type); }
    o.push_back((int64_t)g.optional_new_vars.size());
    for (auto v : g.optional_new_vars) o.push_back(v);
    o.push_back((int64_t)g.unions.size());
    for (const auto &u : g.unions) flat_write_group(u, o);
    o.push_back((int64_t)g.optional.size());
    for (const auto &x : g.optional) flat_write_group(x, o);
}
static bool flat_to_wire(const int64_t *flat, int n, wukong::WireQuery &w) {
    const int64_t *p = flat, *end = flat + n;
    if (n < 19) return false;
    w.qid = (int)p[0]; w.pqid = (int)p[1]; w.pg_type = (int)p[2]; w.state = (int)p[3]; w.dev_type = (int)p[4]; w.job_type = (int)p[5];
    w.priority = (int)p[6]; w.mt_factor = (int)p[7]; w.mt_tid = (int)p[8]; w.pattern_step = (int)p[9]; w.local_var = (wukong::ssid_t)p[10];
    w.corun_enabled = p[11] != 0; w.corun_step = (int)p[12]; w.fetch_step = (int)p[13]; w.union_done = p[14] != 0; w.optional_step = (int)p[15];
    w.limit = (int)p[16]; w.offset = (unsigned)p[17]; w.distinct = p[18] != 0;
    p += 19;
    if (!flat_read_group(p, end, w.pattern_group)) return false;
    auto need = [&](int64_t k) { return p + k <= end; };
    if (!need(1)) return false;
    int64_t k = *p++;
    if (k < 0 || !need(2 * k)) return false;
    for (int64_t i = 0; i < k; i++, p += 2) { wukong::WireQuery::Order o; o.id = (wukong::ssid_t)p[0]; o.descending = p[1] != 0; w.orders.push_back(o); }
    if (!need(6)) return false;
    w.col_num = (int)p[0]; w.row_num = (int)p[1]; w.attr_col_num = (int)p[2]; w.status_code = (int)p[3]; w.blind = p[4] != 0; w.nvars = (int)p[5];
    p += 6;
    auto vec = [&](auto &dst) { if (!need(1)) return false; int64_t m = *p++; if (m < 0 || !need(m)) return false; for (int64_t i = 0; i < m; i++) dst.push_back((typename std::remove_reference<decltype(dst)>::type::value_type)*p++); return true; };
    if (!vec(w.required_vars) || !vec(w.v2c_map)) return false;
    { if (!need(1)) return false; int64_t m = *p++; if (m < 0 || !need(m)) return false; for (int64_t i = 0; i < m; i++) w.optional_matched_rows.push_back(*p++ != 0); }
    if (!vec(w.result_table)) return false;
    if (!need(2)) return false;
    w.gpu_result_buf_nelems = (uint64_t)p[0]; w.gpu_col_num = (int)p[1];
    p += 2;
    return p == end;
}
static void wire_to_flat(const wukong::WireQuery &w, std::vector<int64_t> &o) {
    const int64_t sc[] = {w.qid, w.pqid, w.pg_type, w.state, w.dev_type, w.job_type, w.priority, w.mt_factor, w.mt_tid, w.pattern_step, w.local_var,
                          w.corun_enabled, w.corun_step, w.fetch_step, w.union_done, w.optional_step, w.limit, w.offset, w.distinct};
    o.assign(sc, sc + 19);
    flat_write_group(w.pattern_group, o);
    o.push_back((int64_t)w.orders.size());
    for (const auto &x : w.orders) { o.push_back(x.id); o.push_back(x.descending); }
    const int64_t rs[] = {w.col_num, w.row_num, w.attr_col_num, w.status_code, w.blind, w.nvars};
    o.insert(o.end(), rs, rs + 6);
    o.push_back((int64_t)w.required_vars.size()); for (auto v : w.required_vars) o.push_back(v);
    o.push_back((int64_t)w.v2c_map.size()); for (auto v : w.v2c_map) o.push_back(v);
    o.push_back((int64_t)w.optional_matched_rows.size()); for (bool v : w.optional_matched_rows) o.push_back(v);
    o.push_back((int64_t)w.result_table.size()); for (auto v : w.result_table) o.push_back(v);
    o.push_back((int64_t)w.gpu_result_buf_nelems); o.push_back(w.gpu_col_num);
}
// flat -> Bundle::to_str() bytes (req_type SPARQL_QUERY + archive).  Returns the length, -1 malformed, -2 buffer too small.
int64_t wkh_bundle_encode(const int64_t *flat, int n, int gpu_build, uint8_t *out, int64_t cap) {
    wukong::WireQuery w;
    if (!flat || !flat_to_wire(flat, n, w)) return -1;
    const std::string s = wukong::bundle_to_str(wukong::SPARQL_QUERY, wukong::encode_query(w, gpu_build != 0));
    if ((int64_t)s.size() > cap) return -2;
    memcpy(out, s.data(), s.size());
    return (int64_t)s.size();
}
// Bundle bytes -> flat.  Returns the number of ints, -1 not a SPARQL_QUERY bundle / malformed archive, -2 buffer too small.
int64_t wkh_bundle_decode(const uint8_t *bytes, int64_t len, int gpu_build, int64_t *flat, int64_t cap) {
    wukong::req_type t;
    std::string data;
    if (!bytes || !wukong::bundle_from_str(std::string((const char *)bytes, (size_t)len), t, data) || t != wukong::SPARQL_QUERY) return -1;
    wukong::WireQuery w;
    if (!wukong::decode_query(data, w, gpu_build != 0)) return -1;
    std::vector<int64_t> o;
    wire_to_flat(w, o);
    if ((int64_t)o.size() > cap) return -2;
    memcpy(flat, o.data(), o.size() * sizeof(int64_t));
    return (int64_t)o.size();
}
// a host-mirror SPARQLQuery through the wire and back (to_wire -> encode -> decode -> from_wire): planned patterns, required
// variables and the result metadata survive.  Returns 0 when every carried field is equal.
int wkh_bundle_roundtrip_query(const int32_t *pats, int npat, int nvars, const int32_t *req, int nreq, int blind, const uint32_t *table,
                               int rows, int cols) {
    wukong::SPARQLQuery::PatternGroup pg;
    for (int i = 0; i < npat; i++) pg.patterns.push_back(wukong::SPARQLQuery::Pattern(pats[4 * i], pats[4 * i + 1], (wukong::dir_t)pats[4 * i + 2], pats[4 * i + 3]));
    std::vector<wukong::ssid_t> rq(req, req + nreq);
    wukong::SPARQLQuery q(pg, nvars, rq), r;
    q.result.blind = blind != 0;
    q.result.col_num = cols;
    q.result.row_num = rows;
    if (table && !blind) q.result.result_table.assign(table, table + (size_t)rows * cols);
    q.qid = 77; q.pqid = 5; q.mt_factor = 3; q.mt_tid = 2; q.pattern_step = npat; q.limit = 10; q.offset = 4; q.distinct = true;
    const std::string s = wukong::bundle_to_str(wukong::SPARQL_QUERY, wukong::encode_query(wukong::to_wire(q)));
    wukong::req_type t;
    std::string data;
    wukong::WireQuery w;
    if (!wukong::bundle_from_str(s, t, data) || t != wukong::SPARQL_QUERY || !wukong::decode_query(data, w)) return 1;
    wukong::from_wire(w, r);
    if (r.qid != q.qid || r.pqid != q.pqid || r.mt_factor != q.mt_factor || r.mt_tid != q.mt_tid || r.pattern_step != q.pattern_step ||
        r.limit != q.limit || r.offset != q.offset || r.distinct != q.distinct) return 2;
    if (r.pattern_group.patterns.size() != q.pattern_group.patterns.size()) return 3;
    for (size_t i = 0; i < r.pattern_group.patterns.size(); i++) {
        const auto &a = r.pattern_group.patterns[i], &b = q.pattern_group.patterns[i];
        if (a.subject != b.subject || a.predicate != b.predicate || a.object != b.object || a.direction != b.direction || a.pred_type != b.pred_type) return 4;
    }
    if (r.result.required_vars != q.result.required_vars || r.result.v2c_map != q.result.v2c_map || r.result.nvars != q.result.nvars ||
        r.result.blind != q.result.blind || r.result.col_num != q.result.col_num || r.result.row_num != q.result.row_num) return 5;
    if (r.result.result_table != q.result.result_table) return 6;
    return 0;
}

// Planner::set_plan on a pattern-group TREE (UNION / OPTIONAL sub-groups), no string server needed.
// tree = [npat, (subject, predicate, direction, object) x npat, nunions, tree ..., noptional, tree ...]; the planned tree comes
// back in the same encoding.  Returns the number of ints written, -1 when set_plan refuses the plan, -2 on a malformed tree.
static bool tree_read(const int32_t *&p, const int32_t *end, wukong::SPARQLQuery::PatternGroup &g, int depth = 0) {
    if (depth > 16 || p >= end) return false;
    const int npat = *p++;
    if (npat < 0 || p + 4 * (size_t)npat > end) return false;
    for (int i = 0; i < npat; i++, p += 4)
        g.patterns.push_back(wukong::SPARQLQuery::Pattern(p[0], p[1], (wukong::dir_t)p[2], p[3]));
    for (int kind = 0; kind < 2; kind++) {
        if (p >= end) return false;
        const int n = *p++;
        if (n < 0 || n > 64) return false;
        for (int i = 0; i < n; i++) {
            wukong::SPARQLQuery::PatternGroup sub;
            if (!tree_read(p, end, sub, depth + 1)) return false;
            (kind == 0 ? g.unions : g.optional).push_back(sub);
        }
    }
    return true;
}
static void tree_write(const wukong::SPARQLQuery::PatternGroup &g, std::vector<int32_t> &out) {
    out.push_back((int32_t)g.patterns.size());
    for (const auto &pt : g.patterns) { out.push_back(pt.subject); out.push_back(pt.predicate); out.push_back(pt.direction); out.push_back(pt.object); }
    out.push_back((int32_t)g.unions.size());
    for (const auto &u : g.unions) tree_write(u, out);
    out.push_back((int32_t)g.optional.size());
    for (const auto &o : g.optional) tree_write(o, out);
}
int wkh_set_plan_tree(const int32_t *tree, int n, const char *fmt, int32_t *out, int cap) {
    wukong::SPARQLQuery::PatternGroup g;
    const int32_t *p = tree;
    if (!tree || n <= 0 || !tree_read(p, tree + n, g)) return -2;
    wukong::Planner planner;
    std::istringstream fs(std::string(fmt ? fmt : ""));
    if (!planner.set_plan(g, fs)) return -1;
    std::vector<int32_t> v;
    tree_write(g, v);
    if ((int)v.size() > cap) return -2;
    memcpy(out, v.data(), v.size() * sizeof(int32_t));
    return (int)v.size();
}

// load_config(fname, nsrvs) [+ reload_config(reload)] of the reference (core/config.hpp:160-230) on a fresh Global; the
// integer items come back in the order of CONFIG_ITEMS (wukong_b200/host.py), the input folder as a string.
int wkh_config_load(const char *fname, int nsrvs, int gpu_build, const char *reload, int32_t *out, int cap, char *folder, int folder_cap) {
    wukong::Global g;
    g.reference_defaults();
    if (!g.load_config(fname ? fname : "", nsrvs, gpu_build != 0)) return -1;
    if (reload && reload[0]) g.reload_config(reload);
    const int32_t v[] = {g.num_servers, g.num_threads, g.num_proxies, g.num_engines, g.data_port_base, g.ctrl_port_base,
                         g.rdma_buf_size_mb, g.rdma_rbf_size_mb, g.use_rdma, g.rdma_threshold, g.mt_threshold, g.enable_caching,
                         g.enable_workstealing, g.stealing_pattern, g.silent, g.enable_planner, g.generate_statistics, g.enable_vattr,
                         g.memstore_size_gb, g.est_load_factor, g.num_gpus, g.gpu_kvcache_size_gb, g.gpu_rbuf_size_mb,
                         g.gpu_rdma_buf_size_mb, g.gpu_key_blk_size_mb, g.gpu_value_blk_size_mb, g.gpu_enable_pipeline};
    const int n = (int)(sizeof(v) / sizeof(v[0]));
    if (cap < n) return -2;
    for (int i = 0; i < n; i++) out[i] = v[i];
    if (folder && folder_cap > 0) snprintf(folder, (size_t)folder_cap, "%s", g.input_folder.c_str());
    return n;
}

// parse + apply a user-defined plan; patterns come back as (subject, predicate, direction, object)
int wkh_parse_plan(void *h, const char *query, const char *fmt, int32_t *pats, int max_pats, int *npats, int *nvars,
                   int32_t *required, int max_req, int *nreq) {
    HostEnv *env = (HostEnv *)h;
    if (!env->proxy) return wukong::UNKNOWN_ERROR;
    std::istringstream is(query), fs(fmt);
    wukong::SPARQLQuery q;
    int rc = env->proxy->prepare(is, fs, q);
    if (rc) return rc;
    if ((int)q.pattern_group.patterns.size() > max_pats || (int)q.result.required_vars.size() > max_req) return wukong::UNKNOWN_ERROR;
    *npats = (int)q.pattern_group.patterns.size();
    for (int i = 0; i < *npats; i++) {
        const auto &p = q.pattern_group.patterns[i];
        pats[4 * i] = p.subject; pats[4 * i + 1] = p.predicate; pats[4 * i + 2] = p.direction; pats[4 * i + 3] = p.object;
    }
    *nvars = q.result.nvars;
    *nreq = (int)q.result.required_vars.size();
    for (int i = 0; i < *nreq; i++) required[i] = q.result.required_vars[i];
    return wukong::SUCCESS;
}

// Proxy::run_single_query.  Returns the reply's status code; the (last) reply's table is copied out.
int wkh_run_single_query(void *h, const char *query, const char *fmt, int mt_factor, int cnt, int per_pattern,
                         uint32_t *table, uint64_t cap_words, uint64_t *rows, int *cols, double *latency_us,
                         uint64_t *table_words) {
    HostEnv *env = (HostEnv *)h;
    if (!env->proxy || !env->engine) return WK_ERR_NO_DEVICE;
    std::istringstream is(query), fs(fmt);
    wukong::SPARQLQuery reply;
    wukong::Monitor mon;
    int rc = env->proxy->run_single_query(is, fs, mt_factor, cnt, per_pattern != 0, reply, mon);
    if (rows) *rows = (uint64_t)reply.result.row_num;
    if (cols) *cols = reply.result.col_num;
    if (latency_us) *latency_us = mon.latency_usec();
    if (table_words) *table_words = reply.result.result_table.size();   // 0 for blind / silent replies
    if (rc == 0 && table && reply.result.result_table.size() <= cap_words)
        memcpy(table, reply.result.result_table.data(), reply.result.result_table.size() * sizeof(uint32_t));
    return rc;
}

}  // extern "C"

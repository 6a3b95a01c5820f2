// TEST INFRASTRUCTURE ONLY.  The reference's OWN query engine -- core/engine/sparql.hpp (SPARQLEngine: the dispatch switch,
// index_to_unknown / const_to_unknown / known_to_unknown / known_to_known / known_to_const / index_to_known /
// const_to_known, final_process) with core/query.hpp, rmap.hpp, msgr.hpp, coder.hpp, and core/planner.hpp (set_plan /
// set_direction only) -- compiled over the store of
// ref_store_shim.cpp.  Shadowed because they only reach code off this path (ref_stubs/): dgraph.hpp (two one-line forwards to
// the real GStore), bind.hpp, comm/adaptor.hpp, string_server.hpp.  This pins SURVEY.md §8 rows a6-a14 of the oracle against
// compiled reference code, on a single server; the fork-join transport and the proxy are not exercised.
//
// Built by `make -C oracle ref` into oracle/_ref/ (git-ignored); accommodations for current GCC are listed in the Makefile
// and in ref_shim.h.
#define WK_REF_WITH_ENGINE 1
#include <thread>
#include "ref_shim.h"

std::vector<std::vector<int>> core_bindings;   // bind.hpp

extern "C" {
// ---- the reference's engine over that store ---------------------------------------------------------------------
// Runs planned patterns through SPARQLEngine::execute_one_pattern (the reference's own dispatch switch and pattern
// functions, sparql.hpp:80-1061) until the pattern phase is done, then SPARQLEngine::final_process (:1424-1551).
// mt_factor > 1: the index start is run once per slice (mt_tid = 0 .. mt_factor - 1) and the replies are concatenated
// the way RMap merges them (query.hpp:536-557 append_result); the fork-join transport itself is not exercised.
// Returns the reference's status code (utils/errors.hpp).
int refe_query(void *h, const int32_t *pats, int npat, int nvars, const int32_t *required, int nreq, int blind, int mt_factor,
               int distinct, int64_t offset, int64_t limit, uint32_t *out, uint64_t cap_words, uint64_t *rows, int *cols) {
    RefStore *r = (RefStore *)h;
    Global::num_servers = 1;
    StringServer strs;
    DGraph graph(0, r->g);
    Coder coder(0, 0);
    Adaptor adaptor(0);
    Messenger msgr(0, 0, &adaptor);
    SPARQLEngine eng(0, 0, &strs, &graph, &coder, &msgr);
    SPARQLQuery::PatternGroup pg;
    for (int i = 0; i < npat; i++)
        pg.patterns.push_back(SPARQLQuery::Pattern((ssid_t)pats[4 * i], (ssid_t)pats[4 * i + 1], (ssid_t)pats[4 * i + 2], (ssid_t)pats[4 * i + 3]));
    std::vector<ssid_t> req(required, required + nreq);
    SPARQLQuery fin(pg, nvars, req);
    fin.result.blind = blind != 0;
    fin.distinct = distinct != 0;
    fin.offset = (unsigned)(offset < 0 ? 0 : offset);
    fin.limit = (int)limit;
    *rows = 0;
    *cols = 0;
    try {
        if (npat == 0) throw WukongException(SYNTAX_ERROR);
        const int slices = (fin.start_from_index() && mt_factor > 1) ? mt_factor : 1;
        bool first = true;
        for (int t = 0; t < slices; t++) {
            SPARQLQuery q(pg, nvars, req);
            q.result.blind = blind != 0;
            q.mt_factor = slices;
            q.mt_tid = t;
            while (!q.done(SPARQLQuery::SQState::SQ_PATTERN)) eng.execute_one_pattern(q);
            q.result.update_nrows();
            if (first) {
                fin.result.v2c_map = q.result.v2c_map;
                fin.result.col_num = q.result.col_num;
                fin.result.result_table.swap(q.result.result_table);
                fin.result.row_num = q.result.row_num;
                first = false;
            } else {
                fin.result.row_num += q.result.row_num;
                fin.result.result_table.insert(fin.result.result_table.end(), q.result.result_table.begin(), q.result.result_table.end());
            }
        }
        fin.pattern_step = npat;
        if (!fin.result.blind) {
            fin.result.update_nrows();
            eng.final_process(fin);
        }
    } catch (WukongException &ex) {
        return ex.code();
    }
    *rows = fin.result.blind ? (uint64_t)fin.result.row_num : (uint64_t)fin.result.get_row_num();
    *cols = fin.result.get_col_num();
    const uint64_t words = fin.result.result_table.size();
    if (!fin.result.blind && out && words <= cap_words) memcpy(out, fin.result.result_table.data(), words * sizeof(uint32_t));
    if (!fin.result.blind && words > cap_words) return -1;
    return SUCCESS;
}


// ---- timing: the reference's own engine as the CPU arm -----------------------------------------------------------------------
// One query, `reps` times: the pattern phase runs as mt_factor slices of the index start (sparql.hpp:1064-1089 dispatch /
// :211-221 slicing), each slice on its own host thread with its own SPARQLEngine when `threaded`; the replies are
// concatenated and final_process runs on the merged result, as the proxy-side engine does.  usec[i] = wall time of rep i
// (pattern phase + merge + final_process; no Bundle serialisation, no proxy hop).  Returns the reference's status code.
// digest != NULL: an order-independent digest of the final table of the last repetition (sum over rows of a 64-bit mix of the
// row's words, computed outside the timed region) -- bench.py compares it with the same digest of the GPU engine's table
static uint64_t row_digest(const sid_t *row, int cols) {
    uint64_t hh = 0x9E3779B97F4A7C15ull;
    for (int c = 0; c < cols; c++) {
        hh = (hh ^ (uint64_t)row[c]) * 0xBF58476D1CE4E5B9ull;
        hh ^= hh >> 29;
    }
    hh *= 0x94D049BB133111EBull;
    hh ^= hh >> 32;
    return hh;
}
int refe_time_query_digest(void *h, const int32_t *pats, int npat, int nvars, const int32_t *required, int nreq, int blind, int mt_factor,
                           int threaded, int reps, double *usec, uint64_t *rows_out, uint64_t *digest);
int refe_time_query(void *h, const int32_t *pats, int npat, int nvars, const int32_t *required, int nreq, int blind, int mt_factor,
                    int threaded, int reps, double *usec, uint64_t *rows_out) {
    return refe_time_query_digest(h, pats, npat, nvars, required, nreq, blind, mt_factor, threaded, reps, usec, rows_out, nullptr);
}
int refe_time_query_digest(void *h, const int32_t *pats, int npat, int nvars, const int32_t *required, int nreq, int blind, int mt_factor,
                           int threaded, int reps, double *usec, uint64_t *rows_out, uint64_t *digest) {
    RefStore *r = (RefStore *)h;
    Global::num_servers = 1;
    SPARQLQuery::PatternGroup pg;
    for (int i = 0; i < npat; i++)
        pg.patterns.push_back(SPARQLQuery::Pattern((ssid_t)pats[4 * i], (ssid_t)pats[4 * i + 1], (ssid_t)pats[4 * i + 2], (ssid_t)pats[4 * i + 3]));
    std::vector<ssid_t> req(required, required + nreq);
    StringServer strs;
    DGraph graph(0, r->g);
    int status = SUCCESS;
    for (int rep = 0; rep < reps; rep++) {
        const uint64_t t0 = timer::get_usec();
        SPARQLQuery fin(pg, nvars, req);
        fin.result.blind = blind != 0;
        const int slices = (fin.start_from_index() && mt_factor > 1) ? mt_factor : 1;
        std::vector<SPARQLQuery> parts(slices, SPARQLQuery(pg, nvars, req));
        std::vector<int> codes(slices, SUCCESS);
        auto run = [&](int t) {
            Coder coder(0, t);
            Adaptor adaptor(t);
            Messenger msgr(0, t, &adaptor);
            SPARQLEngine eng(0, t, &strs, &graph, &coder, &msgr);
            SPARQLQuery &q = parts[t];
            q.result.blind = blind != 0;
            q.mt_factor = slices;
            q.mt_tid = t;
            try {
                while (!q.done(SPARQLQuery::SQState::SQ_PATTERN)) eng.execute_one_pattern(q);
                q.result.update_nrows();
            } catch (WukongException &ex) {
                codes[t] = ex.code();
            }
        };
        if (threaded && slices > 1) {
            std::vector<std::thread> th;
            for (int t = 0; t < slices; t++) th.emplace_back(run, t);
            for (auto &x : th) x.join();
        } else {
            for (int t = 0; t < slices; t++) run(t);
        }
        for (int t = 0; t < slices; t++)
            if (codes[t] != SUCCESS) status = codes[t];
        if (status != SUCCESS) return status;
        fin.result.v2c_map = parts[0].result.v2c_map;
        fin.result.col_num = parts[0].result.col_num;
        fin.result.row_num = 0;
        for (int t = 0; t < slices; t++) {   // append_result, query.hpp:536-557
            fin.result.row_num += parts[t].result.row_num;
            if (!fin.result.blind) {
                if (t == 0) fin.result.result_table.swap(parts[t].result.result_table);
                else fin.result.result_table.insert(fin.result.result_table.end(), parts[t].result.result_table.begin(), parts[t].result.result_table.end());
            }
        }
        fin.pattern_step = npat;
        try {
            if (!fin.result.blind) {
                Coder coder(0, 0);
                Adaptor adaptor(0);
                Messenger msgr(0, 0, &adaptor);
                SPARQLEngine eng(0, 0, &strs, &graph, &coder, &msgr);
                fin.result.update_nrows();
                eng.final_process(fin);
            }
        } catch (WukongException &ex) {
            return ex.code();
        }
        usec[rep] = (double)(timer::get_usec() - t0);
        *rows_out = fin.result.blind ? (uint64_t)fin.result.row_num : (uint64_t)fin.result.get_row_num();
        if (digest && rep == reps - 1) {
            uint64_t d = 0;
            if (!fin.result.blind && fin.result.col_num > 0) {
                const int C = fin.result.col_num;
                const uint64_t n = fin.result.result_table.size() / (uint64_t)C;
                for (uint64_t i = 0; i < n; i++) d += row_digest(&fin.result.result_table[i * C], C);
            }
            *digest = d;
        }
    }
    return status;
}

// ---- a simulated cluster: n shard stores in one process, the reference's own per-server functions ----------------------------
// Every server i has its own GStore (shard i of n, built by refs_build) and its own SPARQLEngine.  The control flow of
// SPARQLEngine::execute_patterns (sparql.hpp:1113-1154) is followed with the transport replaced by an in-process work list:
// run execute_one_pattern; when the plan is not done, either replicate the table to every server (the rule of
// dispatch(r, false), :1091-1110), or split it with generate_sub_query when need_fork_join says so (:802-814; without RDMA:
// before every step), and hand every part to its server.  Index starts go to every server (dispatch, :1064-1089), constant
// starts to the owner (proxy.hpp:205).  Finished parts are concatenated (RMap / append_result) and final_process runs once.
// What is the reference's: every per-server decision and every row.  What is this shim's: the work list.
int refe_cluster_query(void **stores, int n, const int32_t *pats, int npat, int nvars, const int32_t *required, int nreq, int blind,
                       uint32_t *out, uint64_t cap_words, uint64_t *rows, int *cols) {
    std::vector<RefStore *> rs(n);
    for (int i = 0; i < n; i++) { rs[i] = (RefStore *)stores[i]; rs[i]->g->sid = i; }
    StringServer strs;
    std::vector<DGraph *> graphs;
    std::vector<Coder *> coders;
    std::vector<Adaptor *> adaptors;
    std::vector<Messenger *> msgrs;
    std::vector<SPARQLEngine *> engs;
    for (int i = 0; i < n; i++) {
        graphs.push_back(new DGraph(i, rs[i]->g));
        coders.push_back(new Coder(i, 0));
        adaptors.push_back(new Adaptor(0));
        msgrs.push_back(new Messenger(i, 0, adaptors[i]));
        engs.push_back(new SPARQLEngine(i, 0, &strs, graphs[i], coders[i], msgrs[i]));
    }
    SPARQLQuery::PatternGroup pg;
    for (int i = 0; i < npat; i++)
        pg.patterns.push_back(SPARQLQuery::Pattern((ssid_t)pats[4 * i], (ssid_t)pats[4 * i + 1], (ssid_t)pats[4 * i + 2], (ssid_t)pats[4 * i + 3]));
    std::vector<ssid_t> req(required, required + nreq);
    SPARQLQuery fin(pg, nvars, req);
    fin.result.blind = blind != 0;
    const bool rdma0 = Global::use_rdma;
    Global::num_servers = n;
    Global::use_rdma = false;
    int status = SUCCESS;
    *rows = 0;
    *cols = 0;
    try {
        if (npat == 0) throw WukongException(SYNTAX_ERROR);
        std::vector<std::pair<int, SPARQLQuery>> work;
        SPARQLQuery proto(pg, nvars, req);
        proto.result.blind = false;   // the tables travel between the simulated servers
        if (proto.start_from_index()) {
            for (int i = 0; i < n; i++) work.push_back(std::make_pair(i, proto));
        } else {
            work.push_back(std::make_pair(wukong::math::hash_mod(pg.patterns[0].subject, n), proto));
        }
        bool first = true;
        while (!work.empty()) {
            const int sid = work.back().first;
            SPARQLQuery q = work.back().second;
            work.pop_back();
            while (true) {
                engs[sid]->execute_one_pattern(q);
                if (q.done(SPARQLQuery::SQState::SQ_PATTERN)) {
                    q.result.update_nrows();
                    if (first) { fin.result.v2c_map = q.result.v2c_map; fin.result.col_num = q.result.col_num; first = false; }
                    if (q.result.col_num == fin.result.col_num || q.result.result_table.empty())
                        fin.result.result_table.insert(fin.result.result_table.end(), q.result.result_table.begin(), q.result.result_table.end());
                    break;
                }
                SPARQLQuery::Pattern &pt = q.get_pattern();
                if (pt.predicate == TYPE_ID && pt.direction == IN) {
                    std::vector<SPARQLQuery> subs = engs[sid]->generate_sub_query(q, false);
                    for (int i = 0; i < n; i++) work.push_back(std::make_pair(i, subs[i]));
                    break;
                }
                if (engs[sid]->need_fork_join(q)) {
                    std::vector<SPARQLQuery> subs = engs[sid]->generate_sub_query(q);
                    for (int i = 0; i < n; i++)
                        if (subs[i].result.get_row_num() > 0) work.push_back(std::make_pair(i, subs[i]));
                    break;
                }
            }
        }
        if (first) fin.result.col_num = 0;   // every branch died before the last step: an empty answer
        fin.pattern_step = npat;
        fin.result.update_nrows();
        Global::num_servers = 1;
        if (!fin.result.blind) engs[0]->final_process(fin);
    } catch (WukongException &ex) {
        status = ex.code();
    }
    Global::num_servers = 1;
    Global::use_rdma = rdma0;
    for (int i = 0; i < n; i++) { rs[i]->g->sid = 0; delete engs[i]; delete msgrs[i]; delete adaptors[i]; delete coders[i]; delete graphs[i]; }
    if (status != SUCCESS) return status;
    *rows = (uint64_t)fin.result.get_row_num();
    *cols = fin.result.get_col_num();
    const uint64_t words = fin.result.result_table.size();
    if (!fin.result.blind) {
        if (words > cap_words) return -1;
        if (out && words) memcpy(out, fin.result.result_table.data(), words * sizeof(uint32_t));
    }
    return SUCCESS;
}

// ---- fork-join decisions and row split --------------------------------------------------------------
// Which steps of a plan exchange when the store is sharded over n servers, decided by the reference's own
// SPARQLEngine::need_fork_join (sparql.hpp:802-814, RDMA rule: fork when the next hop is not local; threshold 0) and the
// replicate rule of dispatch(r, false) (sparql.hpp:1091-1110, restated here in one line because dispatch itself sends).
// The plan is walked on a real single-server store so that v2c_map / local_var evolve through the reference's pattern
// functions (index_to_unknown sets local_var, :230; sub-queries inherit local_var = start, :760).
// out[s]: -1 no exchange before step s, -2 replicate, c >= 0 re-shard by column c.  Returns 0 or the reference's status code.
int refe_fork_plan(void *h, const int32_t *pats, int npat, int nvars, int n, int32_t *out) {
    RefStore *r = (RefStore *)h;
    StringServer strs;
    DGraph graph(0, r->g);
    Coder coder(0, 0);
    Adaptor adaptor(0);
    Messenger msgr(0, 0, &adaptor);
    SPARQLEngine eng(0, 0, &strs, &graph, &coder, &msgr);
    SPARQLQuery::PatternGroup pg;
    for (int i = 0; i < npat; i++)
        pg.patterns.push_back(SPARQLQuery::Pattern((ssid_t)pats[4 * i], (ssid_t)pats[4 * i + 1], (ssid_t)pats[4 * i + 2], (ssid_t)pats[4 * i + 3]));
    std::vector<ssid_t> req;
    SPARQLQuery q(pg, nvars, req);
    const bool rdma0 = Global::use_rdma;
    const int thr0 = Global::rdma_threshold;
    int rc = SUCCESS;
    try {
        for (int s = 0; s < npat; s++) {
            out[s] = -1;
            if (s > 0) {
                Global::num_servers = n;
                Global::use_rdma = true;
                Global::rdma_threshold = 0;
                SPARQLQuery::Pattern &pt = q.get_pattern();
                if (pt.predicate == TYPE_ID && pt.direction == IN) {          // dispatch(r, false): replicate to every server
                    out[s] = -2;
                    q.local_var = pt.subject;                                 // generate_sub_query(r, false), :760
                } else if (eng.need_fork_join(q)) {
                    out[s] = q.result.var2col(pt.subject);
                    q.local_var = pt.subject;                                 // generate_sub_query(r), :760
                }
                Global::num_servers = 1;
            }
            eng.execute_one_pattern(q);
        }
    } catch (WukongException &ex) {
        rc = ex.code();
    }
    Global::num_servers = 1;
    Global::use_rdma = rdma0;
    Global::rdma_threshold = thr0;
    return rc;
}

// SPARQLEngine::generate_sub_query (sparql.hpp:746-799): rows of an nrows x ncols table go to server row[col] % n.
// out receives the n sub-tables back to back, counts[i] their row numbers.
int refe_split(void *h, const uint32_t *table, uint64_t nrows, int ncols, int col, int n, uint32_t *out, uint64_t *counts) {
    RefStore *r = (RefStore *)h;
    StringServer strs;
    DGraph graph(0, r->g);
    Coder coder(0, 0);
    Adaptor adaptor(0);
    Messenger msgr(0, 0, &adaptor);
    SPARQLEngine eng(0, 0, &strs, &graph, &coder, &msgr);
    SPARQLQuery::PatternGroup pg;
    pg.patterns.push_back(SPARQLQuery::Pattern((ssid_t)(-(col + 1)), (ssid_t)2, (ssid_t)OUT, (ssid_t)(-(ncols + 1))));
    std::vector<ssid_t> req;
    SPARQLQuery q(pg, ncols + 1, req);
    q.result.col_num = ncols;
    for (int c = 0; c < ncols; c++) q.result.v2c_map[c] = c;
    q.result.result_table.assign(table, table + nrows * ncols);
    q.result.update_nrows();
    Global::num_servers = n;
    std::vector<SPARQLQuery> subs = eng.generate_sub_query(q);
    Global::num_servers = 1;
    uint64_t off = 0;
    for (int i = 0; i < n; i++) {
        const std::vector<sid_t> &t = subs[i].result.result_table;
        counts[i] = t.size() / ncols;
        if (!t.empty()) memcpy(out + off, t.data(), t.size() * sizeof(uint32_t));
        off += t.size();
    }
    return SUCCESS;
}

// ---- the reference's plan application: Planner::set_plan + set_direction (core/planner.hpp:1647-1754) -------------------
// pats: npat x (subject, predicate, direction, object) as parsed (direction ignored); fmt: the text of a .fmt plan file.
// Writes the planned patterns to out (4 ints each) and returns their number, or -1 when set_plan refuses the plan.
int refp_set_plan(const int32_t *pats, int npat, const char *fmt, int32_t *out, int cap) {
    Global::enable_planner = false;
    SPARQLQuery::PatternGroup pg;
    for (int i = 0; i < npat; i++)
        pg.patterns.push_back(SPARQLQuery::Pattern((ssid_t)pats[4 * i], (ssid_t)pats[4 * i + 1], (ssid_t)pats[4 * i + 2], (ssid_t)pats[4 * i + 3]));
    Planner planner(0, nullptr, nullptr);
    std::istringstream is(std::string(fmt ? fmt : ""));
    if (!planner.set_plan(pg, is)) return -1;
    const int n = (int)pg.patterns.size();
    if (n > cap) return -1;
    for (int i = 0; i < n; i++) {
        out[4 * i] = (int32_t)pg.patterns[i].subject;
        out[4 * i + 1] = (int32_t)pg.patterns[i].predicate;
        out[4 * i + 2] = (int32_t)pg.patterns[i].direction;
        out[4 * i + 3] = (int32_t)pg.patterns[i].object;
    }
    return n;
}

// ---- Planner::set_plan on a pattern-group TREE (UNION / OPTIONAL blocks of a .fmt file, core/planner.hpp:1722-1738) ------------
// tree = [npat, (subject, predicate, direction, object) x npat, nunions, tree ..., noptional, tree ...]
static bool ref_tree_read(const int32_t *&p, const int32_t *end, SPARQLQuery::PatternGroup &g, int depth) {
    if (depth > 16 || p >= end) return false;
    const int npat = *p++;
    if (npat < 0 || p + 4 * (size_t)npat > end) return false;
    for (int i = 0; i < npat; i++, p += 4)
        g.patterns.push_back(SPARQLQuery::Pattern((ssid_t)p[0], (ssid_t)p[1], (ssid_t)p[2], (ssid_t)p[3]));
    for (int kind = 0; kind < 2; kind++) {
        if (p >= end) return false;
        const int n = *p++;
        if (n < 0 || n > 64) return false;
        for (int i = 0; i < n; i++) {
            SPARQLQuery::PatternGroup sub;
            if (!ref_tree_read(p, end, sub, depth + 1)) return false;
            (kind == 0 ? g.unions : g.optional).push_back(sub);
        }
    }
    return true;
}
static void ref_tree_write(const SPARQLQuery::PatternGroup &g, std::vector<int32_t> &out) {
    out.push_back((int32_t)g.patterns.size());
    for (const auto &pt : g.patterns) { out.push_back((int32_t)pt.subject); out.push_back((int32_t)pt.predicate); out.push_back((int32_t)pt.direction); out.push_back((int32_t)pt.object); }
    out.push_back((int32_t)g.unions.size());
    for (const auto &u : g.unions) ref_tree_write(u, out);
    out.push_back((int32_t)g.optional.size());
    for (const auto &o : g.optional) ref_tree_write(o, out);
}
int refp_set_plan_tree(const int32_t *tree, int n, const char *fmt, int32_t *out, int cap) {
    Global::enable_planner = false;
    SPARQLQuery::PatternGroup g;
    const int32_t *p = tree;
    if (!tree || n <= 0 || !ref_tree_read(p, tree + n, g, 0)) return -2;
    Planner planner(0, nullptr, nullptr);
    std::istringstream is(std::string(fmt ? fmt : ""));
    if (!planner.set_plan(g, is)) return -1;
    std::vector<int32_t> v;
    ref_tree_write(g, v);
    if ((int)v.size() > cap) return -2;
    memcpy(out, v.data(), v.size() * sizeof(int32_t));
    return (int)v.size();
}

// ---- the reference's config loader: load_config + reload_config (core/config.hpp:160-230), non-GPU build, no RDMA device ----
// Fills the integer items in the order of CONFIG_ITEMS (wukong_b200/host.py) and the input folder; returns their number.
int refc_load_config(const char *fname, int nsrvs, const char *reload, int32_t *out, int cap, char *folder, int folder_cap) {
    // the loader writes process-wide statics: start from the reference's defaults (global.hpp:28-124)
    Global::num_servers = 1; Global::num_threads = 2; Global::num_proxies = 1; Global::num_engines = 1;
    Global::input_folder = ""; Global::data_port_base = 5500; Global::ctrl_port_base = 9576;
    Global::rdma_buf_size_mb = 64; Global::rdma_rbf_size_mb = 16; Global::use_rdma = true; Global::rdma_threshold = 300;
    Global::mt_threshold = 16; Global::enable_caching = true; Global::enable_workstealing = false; Global::stealing_pattern = 0;
    Global::silent = true; Global::enable_planner = true; Global::generate_statistics = true; Global::enable_vattr = false;
    Global::memstore_size_gb = 20; Global::est_load_factor = 55; Global::num_gpus = 0; Global::gpu_kvcache_size_gb = 10;
    Global::gpu_rbuf_size_mb = 32; Global::gpu_rdma_buf_size_mb = 64; Global::gpu_key_blk_size_mb = 16;
    Global::gpu_value_blk_size_mb = 4; Global::gpu_enable_pipeline = true;
    load_config(std::string(fname), nsrvs);
    if (reload && reload[0]) reload_config(std::string(reload));
    const int32_t v[] = {Global::num_servers, Global::num_threads, Global::num_proxies, Global::num_engines, Global::data_port_base,
                         Global::ctrl_port_base, Global::rdma_buf_size_mb, Global::rdma_rbf_size_mb, Global::use_rdma,
                         Global::rdma_threshold, Global::mt_threshold, Global::enable_caching, Global::enable_workstealing,
                         Global::stealing_pattern, Global::silent, Global::enable_planner, Global::generate_statistics,
                         Global::enable_vattr, Global::memstore_size_gb, Global::est_load_factor, Global::num_gpus,
                         Global::gpu_kvcache_size_gb, Global::gpu_rbuf_size_mb, Global::gpu_rdma_buf_size_mb,
                         Global::gpu_key_blk_size_mb, Global::gpu_value_blk_size_mb, Global::gpu_enable_pipeline};
    const int n = (int)(sizeof(v) / sizeof(v[0]));
    if (cap < n) return -2;
    for (int i = 0; i < n; i++) out[i] = v[i];
    if (folder && folder_cap > 0) snprintf(folder, (size_t)folder_cap, "%s", Global::input_folder.c_str());
    return n;
}

}  // extern "C"

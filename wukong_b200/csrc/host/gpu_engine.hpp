// GPU query engine with the reference's GPUEngine surface (core/gpu/gpu_engine.hpp:48-393): the same
// execute_one_pattern dispatch, the same Result bookkeeping after every pattern, the same status codes;
// the device work goes through the C ABI (include/wukong_b200.h).  Two ways to run a query:
//   execute_patterns(): one ABI call per pattern, like the reference's agent loop;
//   execute_sparql_query(): the whole pattern phase + final_process in one wk_query_execute call.
#pragma once
#include "dgraph.hpp"
#include "query.hpp"
#include "wukong_b200.h"

namespace wukong {

class GPUEngine {
    int sid;
    wk_engine_t *eng = nullptr;

    static void check(int rc) {
        if (rc == WK_SUCCESS) return;
        throw WukongException(rc <= UNKNOWN_FILTER ? rc : UNKNOWN_ERROR + 1000 * rc);   // engine codes kept in the upper digits
    }
    void after_expand(SPARQLQuery &req, ssid_t end, uint64_t rows, bool new_col) {
        SPARQLQuery::Result &res = req.result;
        if (new_col) {
            res.add_var2col(end, res.get_col_num());
            res.set_col_num(res.get_col_num() + 1);
        }
        res.row_num = (int)rows;
        req.pattern_step++;
    }

public:
    std::string error;
    GPUEngine(int sid, DGraph *graph, const Global &g) : sid(sid) {
        int rc = wk_engine_create(graph->gstore, (uint64_t)g.gpu_rbuf_size_mb << 20, &eng);
        if (rc) error = std::string("wk_engine_create: ") + wk_strerror(rc);
    }
    ~GPUEngine() { if (eng) wk_engine_destroy(eng); }
    bool ok() const { return error.empty(); }
    wk_engine_t *handle() { return eng; }

    // same switch as the reference (gpu_engine.hpp:263-336 / sparql.hpp:938-1061)
    bool execute_one_pattern(SPARQLQuery &req) {
        SPARQLQuery::Pattern &pt = req.get_pattern();
        const ssid_t start = pt.subject, predicate = pt.predicate, end = pt.object;
        const dir_t d = pt.direction;
        SPARQLQuery::Result &res = req.result;
        uint64_t rows = 0;
        if (req.pattern_step == 0) check(wk_engine_reset(eng));
        if (req.pattern_step == 0 && req.start_from_index()) {
            ASSERT_ERROR_CODE(res.var2col(end) == NO_RESULT, UNKNOWN_PATTERN);   // index_to_known: not on the device path
            ASSERT_ERROR_CODE(res.get_col_num() == 0, FIRST_PATTERN_ERROR);
            check(wk_index_to_unknown(eng, (sid_t)start, d, req.mt_tid, req.mt_factor, &rows));
            after_expand(req, end, rows, true);
            req.local_var = end;
            return true;
        }
        ASSERT_ERROR_CODE(res.var_stat(predicate) == CONST_VAR, UNKNOWN_PATTERN);
        switch (const_pair(res.var_stat(start), res.var_stat(end))) {
        case const_pair(CONST_VAR, UNKNOWN_VAR):
            ASSERT_ERROR_CODE(res.get_col_num() == 0, FIRST_PATTERN_ERROR);
            check(wk_const_to_unknown(eng, (sid_t)start, (sid_t)predicate, d, &rows));
            after_expand(req, end, rows, true);
            break;
        case const_pair(CONST_VAR, KNOWN_VAR):   // const_to_known, sparql.hpp:144-186
            check(wk_const_to_known(eng, (sid_t)start, (sid_t)predicate, d, res.var2col(end), &rows));
            after_expand(req, end, rows, false);
            break;
        case const_pair(KNOWN_VAR, CONST_VAR):
            check(wk_known_to_const(eng, res.var2col(start), (sid_t)predicate, d, (sid_t)end, &rows));
            after_expand(req, end, rows, false);
            break;
        case const_pair(KNOWN_VAR, KNOWN_VAR):
            check(wk_known_to_known(eng, res.var2col(start), (sid_t)predicate, d, res.var2col(end), &rows));
            after_expand(req, end, rows, false);
            break;
        case const_pair(KNOWN_VAR, UNKNOWN_VAR):
            check(wk_known_to_unknown(eng, res.var2col(start), (sid_t)predicate, d, &rows));
            after_expand(req, end, rows, true);
            break;
        case const_pair(UNKNOWN_VAR, CONST_VAR):
        case const_pair(UNKNOWN_VAR, KNOWN_VAR):
        case const_pair(UNKNOWN_VAR, UNKNOWN_VAR):
            ASSERT_ERROR_CODE(false, UNKNOWN_SUB);
        default:
            ASSERT_ERROR_CODE(false, UNKNOWN_PATTERN);
        }
        return true;
    }

    // pattern loop + final_process projection (sparql.hpp:1113-1154, 1424-1551), one ABI call per pattern
    void execute_patterns(SPARQLQuery &r) {
        try {
            r.state = SPARQLQuery::SQ_PATTERN;
            while (!r.done(SPARQLQuery::SQ_PATTERN)) execute_one_pattern(r);
            SPARQLQuery::Result &res = r.result;
            if (!res.blind && res.row_num > 0) {
                ASSERT_ERROR_CODE(!res.required_vars.empty(), NO_REQUIRED_VAR);
                std::vector<int32_t> cols;
                for (ssid_t v : res.required_vars) cols.push_back(res.var2col(v));
                uint64_t rows = 0;
                // final_process: DISTINCT, OFFSET, LIMIT, then the projection (sparql.hpp:1428-1550)
                if (r.distinct) check(wk_table_distinct(eng, cols.data(), (int)cols.size(), &rows));
                if (r.offset > 0 || r.limit >= 0) check(wk_table_slice(eng, r.offset, r.limit, &rows));
                check(wk_project(eng, cols.data(), (int)cols.size(), &rows));
                res.set_col_num((int)cols.size());
                res.result_table.resize((size_t)rows * cols.size());
                int c = 0;
                check(wk_table_download(eng, res.result_table.data(), res.result_table.size(), &rows, &c));
                res.update_nrows();
            }
        } catch (WukongException &ex) {
            r.result.set_status_code(ex.code());
        }
        r.shrink();
        r.state = SPARQLQuery::SQ_REPLY;
    }

    // the whole query in one call (plan executed without a host synchronisation per pattern)
    void execute_sparql_query(SPARQLQuery &r) {
        SPARQLQuery::Result &res = r.result;
        std::vector<wk_pattern_t> pats;
        for (auto &p : r.pattern_group.patterns) pats.push_back(wk_pattern_t{p.subject, p.predicate, (int32_t)p.direction, p.object});
        uint64_t rows = 0;
        int cols = 0;
        if (!res.blind && out_buf.size() < (size_t)1 << 20) out_buf.resize((size_t)1 << 20);
        int rc;
        while (true) {
            wk_query_opts_t o;
            o.mt_tid = r.mt_tid; o.mt_factor = r.mt_factor; o.blind = res.blind ? 1 : 0;
            o.distinct = r.distinct ? 1 : 0; o.offset = (int64_t)r.offset; o.limit = (int64_t)r.limit;   // final_process, sparql.hpp:1428-1499
            rc = wk_query_execute_ex(eng, pats.data(), (int)pats.size(), res.nvars, res.required_vars.data(),
                                     (int)res.required_vars.size(), &o,
                                     res.blind ? nullptr : out_buf.data(), out_buf.size(), &rows, &cols);
            if (rc == WK_ERR_BAD_ARG && !res.blind && rows * (uint64_t)cols > out_buf.size()) { out_buf.resize(rows * (uint64_t)cols); continue; }
            break;
        }
        res.set_status_code(rc);
        res.col_num = cols;
        res.row_num = (int)rows;
        if (rc == WK_SUCCESS && !res.blind) res.result_table.assign(out_buf.begin(), out_buf.begin() + rows * (uint64_t)cols);
        r.pattern_step = (int)pats.size();
        r.shrink();
        r.state = SPARQLQuery::SQ_REPLY;
    }

private:
    std::vector<sid_t> out_buf;
};

}  // namespace wukong

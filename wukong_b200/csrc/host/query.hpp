// SPARQL query / result data model with the reference's field names and semantics
// (core/query.hpp:48-116 Pattern, :170-240 PatternGroup, :251-557 Result, :560-682 SPARQLQuery),
// without boost serialisation.  Only the parts the graph-exploration path touches are modelled;
// filters / unions / optionals are carried as empty containers.
#pragma once
#include <vector>

#include "errors.hpp"
#include "type.hpp"

namespace wukong {

enum vstat { KNOWN_VAR = 0, UNKNOWN_VAR, CONST_VAR };
#define NBITS_COL 16
#define NO_RESULT ((1 << NBITS_COL) - 1)
constexpr int const_pair(int t1, int t2) { return (t1 << 4) | t2; }
static inline int col2ext(int col, int t) { return (t << NBITS_COL) | col; }
static inline int ext2col(int ext) { return ext & ((1 << NBITS_COL) - 1); }
static inline int ext2type(int ext) { return (ext >> NBITS_COL) & ((1 << NBITS_COL) - 1); }

class SPARQLQuery {
public:
    enum SQState { SQ_PATTERN = 0, SQ_UNION, SQ_FILTER, SQ_OPTIONAL, SQ_FINAL, SQ_REPLY };
    enum PGType { BASIC, UNION, OPTIONAL };
    enum DeviceType { CPU, GPU };

    class Pattern {
    public:
        ssid_t subject = 0, predicate = 0, object = 0;
        dir_t direction = OUT;
        char pred_type = (char)SID_t;
        Pattern() {}
        Pattern(ssid_t s, ssid_t p, dir_t d, ssid_t o) : subject(s), predicate(p), object(o), direction(d) {}
    };

    class PatternGroup {
    public:
        std::vector<Pattern> patterns;
        std::vector<PatternGroup> unions, optional;   // out of scope on this path (kept empty)
        // id of the first pattern's subject: decides the start server (proxy.hpp:205)
        ssid_t get_start() const { return patterns.empty() ? 0 : patterns[0].subject; }
    };

    class Result {
    public:
        int col_num = 0, row_num = 0, attr_col_num = 0;
        int status_code = SUCCESS;
        bool blind = false;
        int nvars = 0;
        std::vector<ssid_t> required_vars;
        std::vector<int> v2c_map;
        std::vector<sid_t> result_table;

        void clear() { result_table.clear(); required_vars.clear(); }

        int var2col(ssid_t vid) {
            ASSERT_ERROR_CODE(vid < 0, VERTEX_INVALID);
            if (v2c_map.empty()) v2c_map.resize(nvars, NO_RESULT);
            int idx = -(vid + 1);
            ASSERT_ERROR_CODE(idx < nvars && idx >= 0, VERTEX_INVALID);
            return ext2col(v2c_map[idx]);
        }
        vstat var_stat(ssid_t vid) {
            if (vid >= 0) return CONST_VAR;
            return var2col(vid) == NO_RESULT ? UNKNOWN_VAR : KNOWN_VAR;
        }
        void add_var2col(ssid_t vid, int col, int t = SID_t) {
            if (v2c_map.empty()) v2c_map.resize(nvars, NO_RESULT);
            int idx = -(vid + 1);
            ASSERT_ERROR_CODE(vid < 0 && idx < nvars && v2c_map[idx] == NO_RESULT, VERTEX_INVALID);
            v2c_map[idx] = col2ext(col, t);
        }
        void set_col_num(int n) { col_num = n; }
        int get_col_num() const { return col_num; }
        int get_row_num() const { return col_num == 0 ? 0 : (int)(result_table.size() / col_num); }
        void update_nrows() { row_num = get_row_num(); }
        sid_t get_row_col(int r, int c) const { return result_table[(size_t)col_num * r + c]; }
        void append_row_to(int r, std::vector<sid_t> &update) const {
            for (int c = 0; c < col_num; c++) update.push_back(get_row_col(r, c));
        }
        void set_status_code(int code) { status_code = code; }
        int get_status_code() const { return status_code; }
        // replies of mt_factor / fork-join sub-queries are concatenated (query.hpp:536-557)
        void append_result(Result &r) {
            v2c_map = r.v2c_map;
            col_num = r.col_num;
            row_num += r.row_num;
            if (r.blind) return;
            result_table.insert(result_table.end(), r.result_table.begin(), r.result_table.end());
        }
    };

    int qid = -1, pqid = -1;
    PGType pg_type = BASIC;
    SQState state = SQ_PATTERN;
    DeviceType dev_type = GPU;
    int priority = 0;
    int mt_factor = 1, mt_tid = 0;
    int pattern_step = 0;
    ssid_t local_var = 0;
    int limit = -1;
    unsigned offset = 0;
    bool distinct = false;
    PatternGroup pattern_group;
    Result result;

    SPARQLQuery() {}
    SPARQLQuery(PatternGroup g, int nvars, std::vector<ssid_t> &required_vars) : pattern_group(g) {
        result.nvars = nvars;
        result.required_vars = required_vars;
        result.v2c_map.resize(nvars, NO_RESULT);
    }
    Pattern &get_pattern() { return pattern_group.patterns.at(pattern_step); }
    Pattern &get_pattern(int step) { return pattern_group.patterns.at(step); }
    bool has_pattern() const { return !pattern_group.patterns.empty(); }
    bool done(SQState s) const { return s == SQ_PATTERN ? pattern_step >= (int)pattern_group.patterns.size() : true; }
    // the planner hints an index start with a dummy first pattern whose subject is a predicate / type id
    bool start_from_index() const {
        if (pattern_group.patterns.empty()) return false;
        if (is_tpid(pattern_group.patterns[0].subject)) {
            ASSERT_ERROR_CODE(pattern_group.patterns[0].predicate == PREDICATE_ID || pattern_group.patterns[0].predicate == TYPE_ID,
                              OBJ_ERROR);
            return true;
        }
        return false;
    }
    // drop what a reply does not need (blind replies carry only the metadata)
    void shrink() {
        pattern_group.patterns.clear();
        if (result.blind) result.clear();
    }
};

}  // namespace wukong

// Wire format of a SPARQL request / reply: the reference's Bundle (core/query.hpp:1130-1232) = 4-byte req_type followed by a
// boost::archive::binary_oarchive of the SPARQLQuery, written field by field by the free save() functions of
// core/query.hpp:917-1075.  Boost is not available here (and not wanted on this path), so the archive layout is restated:
//
//   header   std::size_t 22 + "serialization::archive" + uint16 archive version (16 = Boost 1.66/1.67) + the native-size
//            bytes {sizeof(int), sizeof(long), sizeof(float), sizeof(double)} + int 1 (endianness probe)
//            (basic_binary_oarchive::init + basic_binary_oprimitive::init)
//   classes  every class on this path is object_serializable + track_never (query.hpp:1096-1119): no class id, no version,
//            no tracking words -- just the members in save() order
//   scalars  native little-endian: int / enum 4 bytes, bool / char 1 byte, unsigned 4, ssid_t 4 (DTYPE_64BIT off), size_t 8
//   vector<T>, T arithmetic (ssid_t, int, sid_t): std::size_t count + the raw array (array optimisation, no item version)
//   vector<T>, T a class: std::size_t count + uint32 item version (0) + the items; vector<bool>: count + one byte each;
//   std::set<ssid_t>: count + uint32 item version + the items in order
// NOT PINNED against a real Boost build (none in this image): the layout above is what Boost 1.67 documents and its headers
// implement; tests/test_bundle.py holds an independent Python statement of the same rules and checks every byte.
// FILTER expression trees (pointer graphs, which bring Boost's pointer tracking into the archive) and attribute result
// columns are refused (UNSUPPORT) rather than approximated.
#pragma once
#include <cstring>
#include <set>
#include <string>
#include <vector>

#include "query.hpp"

namespace wukong {

enum req_type { SPARQL_QUERY = 0, DYNAMIC_LOAD = 1, GSTORE_CHECK = 2, SPARQL_HISTORY = 3 };   // query.hpp:1122

// everything the reference serialises of a SPARQLQuery; the members this path does not use keep the reference's defaults
struct WireQuery {
    struct Order { ssid_t id = 0; bool descending = false; };
    struct Group {
        std::vector<SPARQLQuery::Pattern> patterns;
        std::vector<Group> unions, optional;
        std::set<ssid_t> optional_new_vars;
    };
    int qid = -1, pqid = -1;
    int pg_type = 0, state = 0, dev_type = 0, job_type = 0;
    int priority = 0, mt_factor = 1, mt_tid = 0, pattern_step = 0;
    ssid_t local_var = 0;
    bool corun_enabled = false;
    int corun_step = 0, fetch_step = 0;
    bool union_done = false;
    int optional_step = 0;
    int limit = -1;
    unsigned offset = 0;
    bool distinct = false;
    Group pattern_group;
    std::vector<Order> orders;
    // Result
    int col_num = 0, row_num = 0, attr_col_num = 0, status_code = 0;
    bool blind = false;
    int nvars = 0;
    std::vector<ssid_t> required_vars;
    std::vector<int> v2c_map;
    std::vector<bool> optional_matched_rows;
    std::vector<sid_t> result_table;
    // GPUResult (only in the archive of a -DUSE_GPU build, query.hpp:1000-1002)
    uint64_t gpu_result_buf_nelems = 0;
    int gpu_col_num = 0;
};

class BinaryOArchive {
    std::string out;
    template <class T> void raw(const T &v) { out.append(reinterpret_cast<const char *>(&v), sizeof(T)); }
public:
    explicit BinaryOArchive(bool header = true) {
        if (!header) return;
        const std::string sig = "serialization::archive";
        raw<uint64_t>(sig.size());
        out += sig;
        raw<uint16_t>(16);                       // BOOST_ARCHIVE_VERSION of Boost 1.66 / 1.67
        raw<uint8_t>(sizeof(int)); raw<uint8_t>(sizeof(long)); raw<uint8_t>(sizeof(float)); raw<uint8_t>(sizeof(double));
        raw<int32_t>(1);
    }
    void i32(int32_t v) { raw(v); }
    void u32(uint32_t v) { raw(v); }
    void u64(uint64_t v) { raw(v); }
    void b(bool v) { raw<uint8_t>(v ? 1 : 0); }
    void c(char v) { raw(v); }
    void count(size_t n) { raw<uint64_t>(n); }
    void item_version() { raw<uint32_t>(0); }
    template <class T> void array(const std::vector<T> &v) { count(v.size()); if (!v.empty()) out.append(reinterpret_cast<const char *>(v.data()), v.size() * sizeof(T)); }
    const std::string &str() const { return out; }
};

class BinaryIArchive {
    const std::string &in;
    size_t pos = 0;
public:
    bool ok = true;
    explicit BinaryIArchive(const std::string &s, bool header = true) : in(s) {
        if (!header) return;
        const uint64_t n = u64();
        if (!ok || n != 22 || pos + 22 > in.size() || in.compare(pos, 22, "serialization::archive") != 0) { ok = false; return; }
        pos += 22;
        const uint16_t ver = rd<uint16_t>();
        const uint8_t si = rd<uint8_t>(), sl = rd<uint8_t>(), sf = rd<uint8_t>(), sd = rd<uint8_t>();
        const int32_t one = i32();
        if (ver < 9 || si != sizeof(int) || sl != sizeof(long) || sf != sizeof(float) || sd != sizeof(double) || one != 1) ok = false;
    }
    template <class T> T rd() {
        T v{};
        if (!ok || pos + sizeof(T) > in.size()) { ok = false; return v; }
        memcpy(&v, in.data() + pos, sizeof(T));
        pos += sizeof(T);
        return v;
    }
    int32_t i32() { return rd<int32_t>(); }
    uint32_t u32() { return rd<uint32_t>(); }
    uint64_t u64() { return rd<uint64_t>(); }
    bool b() { return rd<uint8_t>() != 0; }
    char c() { return rd<char>(); }
    size_t count() { const uint64_t n = u64(); if (n > in.size()) ok = false; return ok ? (size_t)n : 0; }   // every item takes >= 1 byte
    void item_version() { (void)u32(); }
    template <class T> void array(std::vector<T> &v) {
        const size_t n = count();
        if (!ok || pos + n * sizeof(T) > in.size()) { ok = false; return; }
        v.resize(n);
        if (n) memcpy(v.data(), in.data() + pos, n * sizeof(T));
        pos += n * sizeof(T);
    }
    bool at_end() const { return pos == in.size(); }
};

namespace wire {
static const char OCCUPIED = 0, EMPTY = 1;   // query.hpp:920-921

inline void save(BinaryOArchive &ar, const SPARQLQuery::Pattern &t) {   // query.hpp:924-931
    ar.i32(t.subject); ar.i32(t.predicate); ar.i32(t.object); ar.i32((int)t.direction); ar.c(t.pred_type);
}
inline void load(BinaryIArchive &ar, SPARQLQuery::Pattern &t) {
    t.subject = ar.i32(); t.predicate = ar.i32(); t.object = ar.i32(); t.direction = (dir_t)ar.i32(); t.pred_type = ar.c();
}
inline void save(BinaryOArchive &ar, const WireQuery::Group &g) {        // query.hpp:943-962
    ar.count(g.patterns.size()); ar.item_version();
    for (const auto &p : g.patterns) save(ar, p);
    ar.count(g.optional_new_vars.size()); ar.item_version();
    for (ssid_t v : g.optional_new_vars) ar.i32(v);
    ar.c(EMPTY);                                                          // filters: never present on this path
    if (!g.optional.empty()) { ar.c(OCCUPIED); ar.count(g.optional.size()); ar.item_version(); for (const auto &o : g.optional) save(ar, o); }
    else ar.c(EMPTY);
    if (!g.unions.empty()) { ar.c(OCCUPIED); ar.count(g.unions.size()); ar.item_version(); for (const auto &u : g.unions) save(ar, u); }
    else ar.c(EMPTY);
}
inline bool load(BinaryIArchive &ar, WireQuery::Group &g, int depth = 0) {   // query.hpp:964-984
    if (depth > 32) return false;
    size_t n = ar.count(); ar.item_version();
    g.patterns.resize(ar.ok ? n : 0);
    for (auto &p : g.patterns) load(ar, p);
    n = ar.count(); ar.item_version();
    for (size_t i = 0; i < n && ar.ok; i++) g.optional_new_vars.insert(ar.i32());
    if (ar.c() == OCCUPIED) return false;                                   // a FILTER tree: not carried by this codec
    if (ar.c() == OCCUPIED) { n = ar.count(); ar.item_version(); g.optional.resize(ar.ok ? n : 0); for (auto &o : g.optional) if (!load(ar, o, depth + 1)) return false; }
    if (ar.c() == OCCUPIED) { n = ar.count(); ar.item_version(); g.unions.resize(ar.ok ? n : 0); for (auto &u : g.unions) if (!load(ar, u, depth + 1)) return false; }
    return ar.ok;
}
}  // namespace wire

// SPARQLQuery -> archive bytes (what Bundle(const SPARQLQuery &) keeps in `data`).  gpu_build: the archive of a -DUSE_GPU
// reference carries Result::gpu as well.
inline std::string encode_query(const WireQuery &q, bool gpu_build = false) {
    BinaryOArchive ar;
    ar.i32(q.qid); ar.i32(q.pqid); ar.i32(q.pg_type); ar.i32(q.state); ar.i32(q.dev_type); ar.i32(q.job_type);
    ar.i32(q.priority); ar.i32(q.mt_factor); ar.i32(q.mt_tid); ar.i32(q.pattern_step); ar.i32(q.local_var);
    ar.b(q.corun_enabled); ar.i32(q.corun_step); ar.i32(q.fetch_step); ar.b(q.union_done); ar.i32(q.optional_step);
    ar.i32(q.limit); ar.u32(q.offset); ar.b(q.distinct);
    wire::save(ar, q.pattern_group);
    if (!q.orders.empty()) {
        ar.c(wire::OCCUPIED);
        ar.count(q.orders.size()); ar.item_version();
        for (const auto &o : q.orders) { ar.i32(o.id); ar.b(o.descending); }
    } else ar.c(wire::EMPTY);
    // Result, query.hpp:986-1003
    ar.i32(q.col_num); ar.i32(q.row_num); ar.i32(q.attr_col_num); ar.i32(q.status_code); ar.b(q.blind); ar.i32(q.nvars);
    ar.array(q.required_vars);
    ar.array(q.v2c_map);
    ar.count(q.optional_matched_rows.size());
    for (bool x : q.optional_matched_rows) ar.b(x);
    if (q.row_num > 0) {
        ar.c(wire::OCCUPIED);
        ar.array(q.result_table);
        ar.count(0); ar.item_version();          // attr_res_table: vector<attr_t>, empty on this path
    } else ar.c(wire::EMPTY);
    if (gpu_build) { ar.u64(q.gpu_result_buf_nelems); ar.i32(q.gpu_col_num); }
    return ar.str();
}

// archive bytes -> SPARQLQuery; false on a malformed / truncated archive or one that carries FILTERs or attribute columns
inline bool decode_query(const std::string &data, WireQuery &q, bool gpu_build = false) {
    BinaryIArchive ar(data);
    if (!ar.ok) return false;
    q.qid = ar.i32(); q.pqid = ar.i32(); q.pg_type = ar.i32(); q.state = ar.i32(); q.dev_type = ar.i32(); q.job_type = ar.i32();
    q.priority = ar.i32(); q.mt_factor = ar.i32(); q.mt_tid = ar.i32(); q.pattern_step = ar.i32(); q.local_var = ar.i32();
    q.corun_enabled = ar.b(); q.corun_step = ar.i32(); q.fetch_step = ar.i32(); q.union_done = ar.b(); q.optional_step = ar.i32();
    q.limit = ar.i32(); q.offset = ar.u32(); q.distinct = ar.b();
    if (!wire::load(ar, q.pattern_group)) return false;
    if (ar.c() == wire::OCCUPIED) {
        const size_t n = ar.count(); ar.item_version();
        q.orders.resize(ar.ok ? n : 0);
        for (auto &o : q.orders) { o.id = ar.i32(); o.descending = ar.b(); }
    }
    q.col_num = ar.i32(); q.row_num = ar.i32(); q.attr_col_num = ar.i32(); q.status_code = ar.i32(); q.blind = ar.b(); q.nvars = ar.i32();
    ar.array(q.required_vars);
    ar.array(q.v2c_map);
    {
        const size_t n = ar.count();
        q.optional_matched_rows.resize(ar.ok ? n : 0);
        for (size_t i = 0; i < q.optional_matched_rows.size(); i++) q.optional_matched_rows[i] = ar.b();
    }
    if (ar.c() == wire::OCCUPIED) {
        ar.array(q.result_table);
        const size_t na = ar.count(); ar.item_version();
        if (na != 0) return false;               // attribute values (boost::variant items): not carried
    }
    if (gpu_build) { q.gpu_result_buf_nelems = ar.u64(); q.gpu_col_num = ar.i32(); }
    return ar.ok && ar.at_end();
}

// Bundle::to_str() / Bundle::init(), query.hpp:1160-1165, 1219-1230: 4-byte req_type + the archive
inline std::string bundle_to_str(req_type type, const std::string &data) {
    std::string s(reinterpret_cast<const char *>(&type), sizeof(req_type));
    return s + data;
}
inline bool bundle_from_str(const std::string &s, req_type &type, std::string &data) {
    if (s.size() < sizeof(req_type)) return false;
    memcpy(&type, s.data(), sizeof(req_type));
    data.assign(s, sizeof(req_type), std::string::npos);
    return true;
}

// host mirror <-> wire
inline WireQuery to_wire(const SPARQLQuery &q) {
    WireQuery w;
    w.qid = q.qid; w.pqid = q.pqid; w.pg_type = (int)q.pg_type; w.state = (int)q.state; w.dev_type = (int)q.dev_type;
    w.priority = q.priority; w.mt_factor = q.mt_factor; w.mt_tid = q.mt_tid; w.pattern_step = q.pattern_step; w.local_var = q.local_var;
    w.limit = q.limit; w.offset = q.offset; w.distinct = q.distinct;
    struct Conv { static void run(const SPARQLQuery::PatternGroup &g, WireQuery::Group &o) {
        o.patterns = g.patterns;
        o.unions.resize(g.unions.size()); for (size_t i = 0; i < g.unions.size(); i++) run(g.unions[i], o.unions[i]);
        o.optional.resize(g.optional.size()); for (size_t i = 0; i < g.optional.size(); i++) run(g.optional[i], o.optional[i]);
    } };
    Conv::run(q.pattern_group, w.pattern_group);
    const SPARQLQuery::Result &r = q.result;
    w.col_num = r.col_num; w.row_num = r.row_num; w.attr_col_num = r.attr_col_num; w.status_code = r.status_code; w.blind = r.blind;
    w.nvars = r.nvars; w.required_vars = r.required_vars; w.v2c_map = r.v2c_map; w.result_table = r.result_table;
    return w;
}
inline void from_wire(const WireQuery &w, SPARQLQuery &q) {
    q.qid = w.qid; q.pqid = w.pqid; q.pg_type = (SPARQLQuery::PGType)w.pg_type; q.state = (SPARQLQuery::SQState)w.state;
    q.dev_type = (SPARQLQuery::DeviceType)w.dev_type; q.priority = w.priority; q.mt_factor = w.mt_factor; q.mt_tid = w.mt_tid;
    q.pattern_step = w.pattern_step; q.local_var = w.local_var; q.limit = w.limit; q.offset = w.offset; q.distinct = w.distinct;
    struct Conv { static void run(const WireQuery::Group &g, SPARQLQuery::PatternGroup &o) {
        o.patterns = g.patterns;
        o.unions.resize(g.unions.size()); for (size_t i = 0; i < g.unions.size(); i++) run(g.unions[i], o.unions[i]);
        o.optional.resize(g.optional.size()); for (size_t i = 0; i < g.optional.size(); i++) run(g.optional[i], o.optional[i]);
    } };
    q.pattern_group = SPARQLQuery::PatternGroup();
    Conv::run(w.pattern_group, q.pattern_group);
    SPARQLQuery::Result &r = q.result;
    r.col_num = w.col_num; r.row_num = w.row_num; r.attr_col_num = w.attr_col_num; r.status_code = w.status_code; r.blind = w.blind;
    r.nvars = w.nvars; r.required_vars = w.required_vars; r.v2c_map = w.v2c_map; r.result_table = w.result_table;
}

}  // namespace wukong

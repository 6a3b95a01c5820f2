// See host_builder.hpp.  Data-structure contract: reference core/store/{vertex,meta,gstore,static_gstore}.hpp
// and core/loader/base_loader.hpp (cited per step below).
#include "host_builder.hpp"

#include <algorithm>
#include <atomic>
#include <cstring>
#include <map>
#include <parallel/algorithm>

#include <omp.h>

namespace wkhost {
namespace {

constexpr int ASSOC = WK_ASSOCIATIVITY;
constexpr int KEY_VID_SHIFT = WK_NBITS_IDX + 1;
constexpr int PTR_SIZE_BITS = 28;

inline bool is_tpid(uint64_t id) { return id > 1 && id < (1u << WK_NBITS_IDX); }
inline uint64_t make_key(uint64_t vid, uint64_t pid, uint64_t dir) { return (vid << KEY_VID_SHIFT) | (pid << 1) | dir; }
inline uint64_t make_ptr(uint64_t size, uint64_t off) { return size | (off << PTR_SIZE_BITS); }

inline uint64_t wang64(uint64_t key) {   // ikey_t::hash -> math::hash_u64 (utils/math.hpp:58-67)
    key = (~key) + (key << 21);
    key ^= key >> 24;
    key = (key + (key << 3)) + (key << 8);
    key ^= key >> 14;
    key = (key + (key << 2)) + (key << 4);
    key ^= key >> 28;
    key += key << 31;
    return key;
}

uint64_t prime_at_most(uint64_t upper) {   // math::hash_prime_u64 (utils/math.hpp:105-131)
    static const uint64_t primes[] = {98317ull, 196613ull, 393241ull, 786433ull, 1572869ull, 3145739ull, 6291469ull,
                                      12582917ull, 25165843ull, 50331653ull, 100663319ull, 201326611ull,
                                      402653189ull, 805306457ull, 1610612741ull};
    if (upper >= (1ull << 31)) return upper;
    uint64_t best = upper;   // below the table: returned unchanged
    bool any = false;
    for (uint64_t p : primes)
        if (p <= upper) { best = p; any = true; }
    return any ? best : upper;
}

struct Triple { uint32_t s, p, o; };

struct Seg {   // working copy of one segment
    wk_segmeta_t m;
    uint64_t ext_used = 0;
};

struct Ctx {
    HostStore &st;
    const StoreBuildOptions &opt;
    uint64_t num_slots = 0;
    std::atomic<uint64_t> last_ext{0};
    std::atomic<bool> failed{false};
    Ctx(HostStore &s, const StoreBuildOptions &o) : st(s), opt(o) {}
    void fail(const char *msg) {
        bool exp = false;
        if (failed.compare_exchange_strong(exp, true)) st.error = msg;
    }
    uint64_t ext_len(uint64_t nb) const { return opt.gpu_ext_extents ? (nb * 15 / 100 + 1) : 256; }   // meta.hpp:38-43
    // GStore::alloc_ext_buckets (gstore.hpp:414-426)
    uint64_t alloc_ext(uint64_t n) {
        const uint64_t orig = last_ext.fetch_add(n);
        if (orig + n >= st.num_buckets_ext) { fail("out of indirect-header region (raise memstore size)"); return 0; }
        return st.num_buckets + orig;
    }
};

// GStore::insert_key (gstore.hpp:789-856) for one segment; segments never share buckets, so
// distinct segments can be filled concurrently without the reference's bucket spinlocks.
uint64_t insert_key(Ctx &cx, Seg &seg, std::vector<std::pair<uint64_t, uint64_t>> &extents, uint64_t key) {
    wk_vertex_t *V = cx.st.vertices.data();
    uint64_t bucket = seg.m.bucket_start + wang64(key) % seg.m.num_buckets;
    while (true) {
        uint64_t slot = bucket * ASSOC;
        for (int i = 0; i < ASSOC - 1; i++, slot++) {
            if (V[slot].key == key) { cx.fail("duplicate key"); return slot; }
            if (V[slot].key == 0) { V[slot].key = key; return slot; }
        }
        if (V[slot].key != 0) { bucket = V[slot].key >> KEY_VID_SHIFT; continue; }
        // link a fresh indirect-header bucket
        uint64_t ext = 0;
        for (auto &e : extents)
            if (e.second > 0) { ext = e.first++; e.second--; break; }
        if (ext == 0) {
            if (cx.opt.gpu_ext_extents) { cx.fail("segment exceeded its single ext extent"); return 0; }
            const uint64_t n = cx.ext_len(seg.m.num_buckets);
            const uint64_t start = cx.alloc_ext(n);
            if (cx.failed) return 0;
            extents.emplace_back(start + 1, n - 1);
            ext = start;
        }
        V[slot].key = make_key(ext, 0, 0);
        bucket = ext;
        seg.ext_used++;
    }
}

}  // namespace

const wk_sid_t *HostStore::get_edges(wk_sid_t vid, wk_sid_t pid, int dir, uint64_t &size) const {
    size = 0;
    const wk_segmeta_t *m = nullptr;
    for (auto &s : segs) {
        if (vid == 0 ? (s.index == 1 && s.dir == dir) : (s.index == 0 && s.pid == pid && s.dir == dir)) { m = &s; break; }
    }
    if (!m || m->num_buckets == 0) return nullptr;
    const uint64_t key = make_key(vid, pid, (uint64_t)dir);
    uint64_t bucket = m->bucket_start + wang64(key) % m->num_buckets;
    while (true) {
        const wk_vertex_t *b = &vertices[bucket * ASSOC];
        for (int i = 0; i < ASSOC - 1; i++)
            if (b[i].key == key) {
                size = b[i].ptr & ((1ull << PTR_SIZE_BITS) - 1);
                return &edges[(b[i].ptr >> PTR_SIZE_BITS) & ((1ull << 34) - 1)];
            }
        if (b[ASSOC - 1].key == 0) return nullptr;
        bucket = b[ASSOC - 1].key >> KEY_VID_SHIFT;
    }
}

void build_store(const wk_sid_t *tr, uint64_t n, const StoreBuildOptions &opt, HostStore &st) {
    st = HostStore();
    Ctx cx(st, opt);
    const uint32_t S = (uint32_t)std::max(1, opt.num_servers), sid = (uint32_t)opt.sid;
    const int npreds = opt.num_normal_preds;
    if (npreds <= 0 || npreds >= (1 << WK_NBITS_IDX)) { st.error = "bad num_normal_preds"; return; }

    // ---- 1. partition by owner, order and deduplicate (base_loader.hpp:343-373) -------------------
    // OUT edges live with the subject's owner, IN edges with the object's owner.
    std::vector<Triple> pso, pos;
    pso.reserve(n / S + 16);
    pos.reserve(n / S + 16);
    for (uint64_t i = 0; i < n; i++) {
        const Triple t{tr[3 * i], tr[3 * i + 1], tr[3 * i + 2]};
        if (t.p == 0 || t.p > (uint32_t)npreds) { st.error = "triple with predicate id outside str_index"; return; }
        if (t.s % S == sid) pso.push_back(t);
        if (t.o % S == sid) pos.push_back(t);
    }
    auto by_pso = [](const Triple &a, const Triple &b) {
        return a.p != b.p ? a.p < b.p : (a.s != b.s ? a.s < b.s : a.o < b.o);
    };
    auto by_pos = [](const Triple &a, const Triple &b) {
        return a.p != b.p ? a.p < b.p : (a.o != b.o ? a.o < b.o : a.s < b.s);
    };
    auto same = [](const Triple &a, const Triple &b) { return a.s == b.s && a.p == b.p && a.o == b.o; };
    __gnu_parallel::sort(pso.begin(), pso.end(), by_pso);
    __gnu_parallel::sort(pos.begin(), pos.end(), by_pos);
    pso.erase(std::unique(pso.begin(), pso.end(), same), pso.end());
    pos.erase(std::unique(pos.begin(), pos.end(), same), pos.end());
    // type triples are not indexed by object: the reference skips the leading POS run whose objects
    // are type ids (static_gstore.hpp:127-130) and answers "?x type T" from the type index instead
    {
        size_t skip = 0;
        while (skip < pos.size() && is_tpid(pos[skip].o)) skip++;
        pos.erase(pos.begin(), pos.begin() + skip);
    }
    st.num_triples_out = pso.size();
    st.num_triples_in = pos.size();

    // ---- 2. per-predicate extents, key and edge counts (init_seg_metas, gstore.hpp:530-786) --------
    struct Range { size_t b = 0, e = 0; uint64_t keys = 0; };
    std::vector<Range> out_r(npreds + 1), in_r(npreds + 1);
    std::vector<uint64_t> type_cnt(npreds + 1, 0);   // instances per type id
    auto scan = [&](const std::vector<Triple> &v, std::vector<Range> &r, bool by_subject) {
        size_t i = 0;
        while (i < v.size()) {
            const uint32_t p = v[i].p;
            size_t j = i;
            uint64_t keys = 0;
            uint32_t prev = 0;
            bool first = true;
            for (; j < v.size() && v[j].p == p; j++) {
                const uint32_t k = by_subject ? v[j].s : v[j].o;
                if (first || k != prev) { keys++; prev = k; first = false; }
            }
            r[p].b = i; r[p].e = j; r[p].keys = keys;
            i = j;
        }
    };
    scan(pso, out_r, true);
    scan(pos, in_r, false);
    for (size_t i = out_r[WK_TYPE_ID].b; i < out_r[WK_TYPE_ID].e; i++)
        if (is_tpid(pso[i].o) && pso[i].o <= (uint32_t)npreds) type_cnt[pso[i].o]++;

    std::vector<uint32_t> local_preds;
    uint64_t total_keys = 0, num_typeid = 0;
    for (int p = 1; p <= npreds; p++) {
        const uint64_t ne_out = out_r[p].e - out_r[p].b, ne_in = in_r[p].e - in_r[p].b;
        if (ne_out + ne_in > 0) {
            local_preds.push_back((uint32_t)p);
            total_keys += out_r[p].keys + in_r[p].keys;
        } else if (type_cnt[p] > 0) {
            num_typeid++;
        }
    }
    total_keys += local_preds.size() * 2 + num_typeid;
    st.num_keys = total_keys;

    // ---- 3. region sizes (GStore ctor, gstore.hpp:979-1025) ------------------------------------------
    const uint64_t nsegs = (uint64_t)npreds * 2 + 2;
    uint64_t total_edges = pso.size() + pos.size();
    for (int p = 1; p <= npreds; p++) total_edges += out_r[p].keys + in_r[p].keys + type_cnt[p];
    uint64_t num_entries;
    if (opt.kvstore_bytes) {
        const uint64_t header = opt.kvstore_bytes * (128 * 100 / (128 + 3 * 32)) / 100;   // HD_RATIO
        cx.num_slots = header / sizeof(wk_vertex_t);
        st.num_buckets = prime_at_most((cx.num_slots / ASSOC) * 80 / 100);              // MHD_RATIO
        st.num_buckets_ext = cx.num_slots / ASSOC - st.num_buckets;
        num_entries = (opt.kvstore_bytes - header) / sizeof(wk_sid_t);
    } else {
        // size from the data: #buckets = #keys * 100 / (ASSOCIATIVITY * est_load_factor) (global.hpp:99-104)
        const uint64_t lf = (uint64_t)std::max(1, std::min(100, opt.est_load_factor));
        st.num_buckets = total_keys * 100 / (ASSOC * lf) + nsegs + 8;
        st.num_buckets_ext = (opt.gpu_ext_extents ? st.num_buckets * 15 / 100 : 512 * nsegs + st.num_buckets) + nsegs + 8;
        cx.num_slots = (st.num_buckets + st.num_buckets_ext) * ASSOC;
        num_entries = total_edges + 1;
    }
    if (st.num_buckets <= nsegs) { st.error = "kvstore too small"; return; }
    if (total_edges >= num_entries) { st.error = "out of entry region (raise memstore size)"; return; }

    // ---- 4. segment metadata: edge extents, then buckets in proportion to keys (gstore.hpp:428-472) ---
    std::map<std::tuple<uint32_t, int, int>, Seg> segs;   // ordered like segid_t::operator< (pid, index, dir)
    uint64_t last_entry = 0, main_off = 0;
    const uint64_t num_free = st.num_buckets - nsegs;
    auto alloc_edges = [&](uint64_t k) { uint64_t o = k ? last_entry : 0; last_entry += k; return o; };
    auto alloc_buckets = [&](Seg &sg) {
        uint64_t nb = 0;
        if (sg.m.num_keys != 0) {
            const double ratio = static_cast<double>(sg.m.num_keys) / total_keys;
            nb = (uint64_t)(ratio * num_free);
        }
        sg.m.num_buckets = std::max<uint64_t>(nb, 1);
        sg.m.bucket_start = main_off;
        main_off += sg.m.num_buckets;
        const uint64_t el = cx.ext_len(sg.m.num_buckets);
        sg.m.ext_start = cx.alloc_ext(el);
        sg.m.ext_num = el;
    };
    auto new_seg = [&](int index, uint32_t pid, int dir) -> Seg & {
        Seg &sg = segs[std::make_tuple(pid, index, dir)];
        memset(&sg.m, 0, sizeof(sg.m));
        sg.m.index = index; sg.m.pid = pid; sg.m.dir = dir;
        return sg;
    };
    for (int d = 0; d <= 1; d++) new_seg(0, 0, d);   // [0|PREDICATE_ID|d]: only used by VERSATILE builds, stays empty
    Seg &idx_in = new_seg(1, WK_PREDICATE_ID, WK_DIR_IN), &idx_out = new_seg(1, WK_PREDICATE_ID, WK_DIR_OUT);
    for (int p = 1; p <= npreds; p++) {
        Seg &so = new_seg(0, (uint32_t)p, WK_DIR_OUT), &si = new_seg(0, (uint32_t)p, WK_DIR_IN);
        so.m.num_edges = out_r[p].e - out_r[p].b;
        si.m.num_edges = in_r[p].e - in_r[p].b;
        idx_out.m.num_edges += in_r[p].keys;                    // [0|p|OUT] lists the objects of p
        idx_in.m.num_edges += out_r[p].keys + type_cnt[p];      // [0|p|IN] subjects of p, [0|t|IN] instances of t
        so.m.num_keys = so.m.num_edges ? out_r[p].keys : 0;
        si.m.num_keys = si.m.num_edges ? in_r[p].keys : 0;
        so.m.edge_start = alloc_edges(so.m.num_edges);
        si.m.edge_start = alloc_edges(si.m.num_edges);
        alloc_buckets(so);
        alloc_buckets(si);
    }
    idx_out.m.edge_start = alloc_edges(idx_out.m.num_edges);
    idx_out.m.num_keys = local_preds.size();
    alloc_buckets(idx_out);
    idx_in.m.edge_start = alloc_edges(idx_in.m.num_edges);
    idx_in.m.num_keys = local_preds.size() + num_typeid;
    alloc_buckets(idx_in);
    if (main_off > st.num_buckets) cx.fail("main header overflow");
    if (cx.failed) return;

    st.vertices.assign(cx.num_slots, wk_vertex_t{0, 0});
    st.edges.assign(last_entry + 1, 0);

    // ---- 5. normal segments: one key + one contiguous sorted edge run per (s,p) / (o,p) group
    //         (insert_triples, static_gstore.hpp:64-161); segments are independent -> parallel ------
    std::vector<Seg *> work;
    for (uint32_t p : local_preds) {
        work.push_back(&segs[std::make_tuple(p, 0, WK_DIR_OUT)]);
        work.push_back(&segs[std::make_tuple(p, 0, WK_DIR_IN)]);
    }
#pragma omp parallel for schedule(dynamic, 1)
    for (size_t w = 0; w < work.size(); w++) {
        Seg &sg = *work[w];
        if (sg.m.num_edges == 0 || cx.failed) continue;
        const bool outdir = sg.m.dir == WK_DIR_OUT;
        const std::vector<Triple> &v = outdir ? pso : pos;
        const Range &r = outdir ? out_r[sg.m.pid] : in_r[sg.m.pid];
        std::vector<std::pair<uint64_t, uint64_t>> extents{{sg.m.ext_start, sg.m.ext_num}};
        uint64_t off = sg.m.edge_start;
        size_t i = r.b;
        while (i < r.e && !cx.failed) {
            const uint32_t k = outdir ? v[i].s : v[i].o;
            size_t j = i;
            while (j < r.e && (outdir ? v[j].s : v[j].o) == k) j++;
            const uint64_t slot = insert_key(cx, sg, extents, make_key(k, sg.m.pid, (uint64_t)sg.m.dir));
            if (cx.failed) break;
            st.vertices[slot].ptr = make_ptr(j - i, off);
            for (size_t t = i; t < j; t++) st.edges[off++] = outdir ? v[t].o : v[t].s;
            i = j;
        }
    }
    if (cx.failed) return;

    // ---- 6. index segments (insert_idx, static_gstore.hpp:217-265; collect_idx_info, gstore.hpp:858-888)
    for (int d : {WK_DIR_IN, WK_DIR_OUT}) {
        Seg &sg = (d == WK_DIR_IN) ? idx_in : idx_out;
        std::vector<std::pair<uint64_t, uint64_t>> extents{{sg.m.ext_start, sg.m.ext_num}};
        uint64_t off = sg.m.edge_start;
        for (uint32_t p : local_preds) {
            if (p == WK_TYPE_ID) continue;   // type triples feed the type index, not a predicate index
            const bool subj = (d == WK_DIR_IN);
            const std::vector<Triple> &v = subj ? pso : pos;
            const Range &r = subj ? out_r[p] : in_r[p];
            if (r.keys == 0) continue;
            const uint64_t slot = insert_key(cx, sg, extents, make_key(0, p, (uint64_t)d));
            if (cx.failed) return;
            st.vertices[slot].ptr = make_ptr(r.keys, off);
            uint32_t prev = 0;
            bool first = true;
            for (size_t i = r.b; i < r.e; i++) {
                const uint32_t k = subj ? v[i].s : v[i].o;
                if (first || k != prev) { st.edges[off++] = k; prev = k; first = false; }
            }
        }
        if (d == WK_DIR_IN) {
            // type index [0|t|IN]: instances of t, gathered from the (s, TYPE_ID, t) triples
            std::vector<uint64_t> toff(npreds + 2, 0);
            for (int t = 1; t <= npreds; t++) toff[t + 1] = toff[t] + type_cnt[t];
            std::vector<uint64_t> cur(toff.begin(), toff.end());
            for (size_t i = out_r[WK_TYPE_ID].b; i < out_r[WK_TYPE_ID].e; i++)
                if (is_tpid(pso[i].o) && pso[i].o <= (uint32_t)npreds) st.edges[off + cur[pso[i].o]++] = pso[i].s;
            for (int t = 2; t <= npreds; t++) {
                if (type_cnt[t] == 0) continue;
                const uint64_t slot = insert_key(cx, sg, extents, make_key(0, (uint64_t)t, WK_DIR_IN));
                if (cx.failed) return;
                st.vertices[slot].ptr = make_ptr(type_cnt[t], off + toff[t]);
            }
            off += toff[npreds + 1];
        }
        if (off > sg.m.edge_start + sg.m.num_edges) { cx.fail("index segment overflow"); return; }
    }
    st.used_ext = cx.last_ext.load();
    st.edges.resize(last_entry ? last_entry : 1);
    for (auto &kv : segs) st.segs.push_back(kv.second.m);
}

}  // namespace wkhost

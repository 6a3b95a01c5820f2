// Lean single-CTA interpreter for const-start ("light") plans.
//
// Light queries (LUBM Q4-Q6: a constant, then a handful of steps over tens to hundreds of rows)
// are bound by dependent-load latency, not bandwidth.  This kernel therefore
//   * keeps the whole binding table in shared memory across steps (no global round trips),
//   * issues all probes of a step at once (thread per row, 8 independent 16-byte loads each),
//     then all edge loads at once (thread per OUTPUT row, binary search in the scanned degrees),
//   * keeps every counter in shared memory (no global atomics),
//   * reports through one 32-byte record in mapped pinned memory, validated by a checksum on the
//     host instead of a system-scope fence on the device.
// If a table outgrows shared memory the kernel spills it to the engine's global buffer and tells
// the host at which step to continue on the multi-CTA path.
//
// Step semantics: reference core/engine/sparql.hpp:238-285 (const_to_unknown), :295-407
// (known_to_unknown), :416-476 (known_to_known), :484-549 (known_to_const), :1507-1550 (projection).
#pragma once
#include "wk_device.cuh"

namespace wk {

enum { LIGHT_WORDS = 4096, LIGHT_ROWS = 1024, MAX_LIGHT_STEPS = 24 };
// diagnostics (profiling level 3): [0] entry, [1] control block cleared, [2 + s] step s done, [26] table written, [27] record
// stored, [28] request acquired / [29] decoded (resident server), [32 + 4 s + k] inside step s: k = 0 probes issued and
// consumed, 1 multiplicities scanned, 2 rows materialised
enum { LIGHT_TRACE_WORDS = 32 + 4 * MAX_LIGHT_STEPS };
// thread count of the single-CTA latency path (the interpreter is a template on it)
enum { LIGHT_THREADS = 256, LIGHT_MAX_WARPS = 32 };   // measured (round 1, block-wide steps only): 1024 threads cost ~0.4 us more per step in barriers
// the resident server runs 1024 threads: tables of <= 32 rows are handled by one warp (one barrier per step whatever the CTA
// size), and a table of hundreds of rows is probed in ONE round instead of three dependent ones
enum { LIGHT_SRV_THREADS = 1024 };
enum { STAR_ROWS = 128 };   // tallest table of a star round (4 patterns x 128 rows of multiplicities and pointers in LightSmem)
enum { LKIND_I2U = 0, LKIND_C2U = 1, LKIND_K2U = 2, LKIND_K2K = 3, LKIND_K2C = 4 };

struct LightStep {
    SegParam seg;
    uint64_t key;          // seeds
    int32_t kind, C;
    int32_t col_start, col_end;
    uint32_t end_const, _pad0;
    int32_t mt_tid, mt_factor;
};

// completion record in mapped pinned host memory (32 bytes, written with two 16-byte stores)
struct LightRecord {
    uint64_t seq;
    uint64_t rows;
    uint64_t status_resume;   // status | resume_step << 32
    uint64_t check;           // record_check(seq, rows, status_resume, table checksum)
};

struct LightPlan {
    const uint4 *vertices;
    const uint32_t *edges;
    uint32_t *buf[2];           // engine result buffers (spill target)
    uint64_t *counts;           // CtlBlock::counts
    uint64_t *stats;            // CtlBlock::stats
    uint32_t *status;           // CtlBlock::status
    uint64_t *ctl_words;        // whole CtlBlock as 8-byte words (cleared here)
    int32_t ctl_nwords;
    int32_t nsteps;
    LightRecord *rec;           // device pointer of the mapped record
    uint32_t *host_table;       // device pointer of the mapped staging area
    uint64_t host_table_words;
    uint64_t cap_words;         // per result buffer
    uint64_t seq;
    int32_t do_project, proj_n;
    int32_t collect_stats, _pad1;   // per-step counters cost two block reductions per step: only when profiling
    long long *trace;               // diagnostics (profiling level 3): clock64() of thread 0 at phase boundaries
    int8_t proj_cols[MAX_COLS];
    LightStep steps[MAX_LIGHT_STEPS];
};

__host__ __device__ __forceinline__ uint64_t table_word_mix(uint32_t w, uint64_t i) {
    return ((uint64_t)w + 0x9E3779B97F4A7C15ull) * (2 * i + 1);
}
__host__ __device__ __forceinline__ uint64_t record_check(uint64_t seq, uint64_t rows, uint64_t sr, uint64_t tsum) {
    uint64_t h = seq * 0xD6E8FEB86659FD93ull;
    h ^= (rows + 0x9E3779B97F4A7C15ull) * 0xBF58476D1CE4E5B9ull;
    h = (h << 31) | (h >> 33);
    h ^= (sr + 0x2545F4914F6CDD1Dull) * 0x94D049BB133111EBull;
    h = (h << 29) | (h >> 35);
    return h + tsum * 0xFF51AFD7ED558CCDull + 1;
}

__device__ __forceinline__ void st_sys_u32(uint32_t *p, uint32_t v) {
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_sys_v2u64(uint64_t *p, uint64_t a, uint64_t b) {
    asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}

__device__ __forceinline__ void st_sys_u64(uint64_t *p, uint64_t v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_sys_u64(const uint64_t *p) {
    uint64_t v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// spin until *p >= want; bounded (about 2 s) so that a dead peer cannot hang the GPU
__device__ __forceinline__ bool wait_flag(const uint64_t *p, uint64_t want) {
    for (uint32_t i = 0; i < (1u << 23); i++) {
        if (ld_sys_u64(p) >= want) return true;
        __nanosleep(200);
    }
    return false;
}

struct LightSmem {
    uint32_t tab[2][LIGHT_WORDS];
    uint64_t ptr[LIGHT_ROWS];       // raw iptr_t of each row's key (0 = miss)
    uint32_t pre[LIGHT_ROWS + 4];   // multiplicities, then their exclusive prefix
    uint32_t wsum[LIGHT_MAX_WARPS];
    uint32_t total;
    uint32_t seed_len;
    uint64_t seed_ptr;
    uint64_t red[LIGHT_MAX_WARPS];
};

// Where a key lives.  LocalView: the whole store is on this GPU.  PeerView: the store is sharded by vid % n and the
// shards of the other GPUs are mapped into this address space (CUDA IPC): a probe is a load over NVLink from the
// owner's header and edge arrays -- the reference's in-place execution with one-sided RDMA reads for small tables
// (core/engine/sparql.hpp:802-814 need_fork_join, rdma_threshold), without the RDMA.
enum { LIGHT_PEERS = 8, LIGHT_SHARDED_STEPS = 12 };
struct SegLite { uint64_t bucket_start; FastMod fm; };
struct LocalView {
    const uint4 *v;
    const uint32_t *e;
    __device__ __forceinline__ const uint4 *vertices(uint32_t) const { return v; }
    __device__ __forceinline__ const uint32_t *edges(uint32_t) const { return e; }
    __device__ __forceinline__ uint64_t bucket(const LightStep &ls, int, uint64_t key, uint32_t) const {
        return ls.seg.bucket_start + fastmod(hash_u64(key), ls.seg.fm);
    }
};
struct PeerView {
    const uint4 *const *v;                       // [rank]
    const uint32_t *const *e;                    // [rank]
    const SegLite (*segr)[LIGHT_PEERS];          // [step][rank]: every shard sizes its segments on its own
    uint32_t n;
    __device__ __forceinline__ const uint4 *vertices(uint32_t vid) const { return v[vid % n]; }
    __device__ __forceinline__ const uint32_t *edges(uint32_t vid) const { return e[vid % n]; }
    __device__ __forceinline__ uint64_t bucket(const LightStep &, int s, uint64_t key, uint32_t vid) const {
        const SegLite &sg = segr[s][vid % n];
        return sg.bucket_start + fastmod(hash_u64(key), sg.fm);
    }
};

// probe one key, thread-serial over the bucket chain, 8 independent slot loads per bucket
__device__ __forceinline__ uint64_t probe_thread(const uint4 *__restrict__ vertices, uint64_t key, uint64_t bucket,
                                                 uint32_t &visited) {
    visited = 0;
    while (true) {
        uint4 v[8];
        const uint4 *b = vertices + bucket * 8;
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = ld_slot(b + i);
        visited++;
        uint64_t found = 0;
#pragma unroll
        for (int i = 0; i < 7; i++) {
            const uint64_t kk = (uint64_t)v[i].x | ((uint64_t)v[i].y << 32);
            if (kk == key) found = (uint64_t)v[i].z | ((uint64_t)v[i].w << 32);
        }
        if (found) return found;
        const uint64_t chain = (uint64_t)v[7].x | ((uint64_t)v[7].y << 32);
        if (chain == 0) return 0;
        bucket = chain >> WK_KEY_VID_SHIFT;
    }
}

// first index k < size with edges[k] == target, scanning 4 independent loads at a time
__device__ __forceinline__ bool list_contains(const uint32_t *__restrict__ e, uint32_t size, uint32_t target,
                                              uint32_t &scanned) {
    for (uint32_t k0 = 0; k0 < size; k0 += 4) {
        uint32_t x[4];
#pragma unroll
        for (int j = 0; j < 4; j++) x[j] = (k0 + j < size) ? ld_edge(e + k0 + j) : ~target;
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (k0 + j < size && x[j] == target) { scanned = k0 + j + 1; return true; }
    }
    scanned = size;
    return false;
}

// exclusive scan of sm.pre[0..n) in place (n <= LIGHT_ROWS), total in sm.total
template <int NT>
__device__ __forceinline__ void light_scan(LightSmem &sm, uint32_t n, int tid) {
    constexpr int EPT = LIGHT_ROWS / NT;   // elements per thread
    const int lane = tid & 31, warp = tid >> 5;
    uint32_t v[EPT], s = 0;
#pragma unroll
    for (int j = 0; j < EPT; j++) {
        const uint32_t i = (uint32_t)tid * EPT + j;
        v[j] = (i < n) ? sm.pre[i] : 0;
        s += v[j];
    }
    uint32_t incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += y;
    }
    if (lane == 31) sm.wsum[warp] = incl;
    __syncthreads();
    uint32_t woff = 0, tot = 0;
    if (NT <= 256) {
#pragma unroll
        for (int w = 0; w < NT / 32; w++) {
            const uint32_t x = sm.wsum[w];
            if (w < warp) woff += x;
            tot += x;
        }
    } else {
        // 32 warp sums: one more warp-level scan instead of 32 serial shared-memory reads
        const uint32_t x = sm.wsum[lane];
        uint32_t wi = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, wi, o);
            if (lane >= o) wi += y;
        }
        tot = __shfl_sync(0xFFFFFFFFu, wi, 31);
        woff = __shfl_sync(0xFFFFFFFFu, wi - x, warp);
    }
    uint32_t run = woff + incl - s;
#pragma unroll
    for (int j = 0; j < EPT; j++) {
        const uint32_t i = (uint32_t)tid * EPT + j;
        if (i < n) sm.pre[i] = run;
        run += v[j];
    }
    if (tid == 0) sm.total = tot;
    __syncthreads();
}

template <int NT>
__device__ __forceinline__ uint64_t block_sum_u64(uint64_t x, LightSmem &sm, int tid) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xFFFFFFFFu, x, o);
    if ((tid & 31) == 0) sm.red[tid >> 5] = x;
    __syncthreads();
    uint64_t t = 0;
#pragma unroll
    for (int w = 0; w < NT / 32; w++) t += sm.red[w];
    __syncthreads();
    return t;
}

// The interpreter proper: runs steps[0..nsteps) with the table in shared memory.  On return the table
// (N x C) is in sm.tab[cur]; done = number of steps completed; spilled = the next step's output does
// not fit shared memory (nothing of that step has been written).
struct LightState { uint32_t N; int C, cur, done; bool spilled; };

template <int NT, class SV, bool WARPM = true>
__device__ __forceinline__ LightState light_interpret(const LightStep *steps, int nsteps, const SV &sv, LightSmem &sm, uint64_t *stats,
                                                      uint64_t *counts, long long *trace = nullptr) {
    const int tid = threadIdx.x;
    uint32_t N = 0;          // rows of the current table (in sm.tab[cur])
    int C = 0;
    int cur = 0;
    int s = 0;
    bool spilled = false;
    uint32_t shadowed = 0;   // steps whose lines were already pulled into L2 by a shadow probe
    for (; s < nsteps; s++) {
        const LightStep &ls = steps[s];
        const int nxt = cur ^ 1;
        uint64_t st_visited = 0, st_edges = 0;
        if (ls.kind == LKIND_C2U || ls.kind == LKIND_I2U) {
            if (tid < 32) {
                uint32_t visited;
                const uint32_t seed_vid = (uint32_t)(ls.key >> WK_KEY_VID_SHIFT);
                const uint4 *vertices = sv.vertices(seed_vid);
                const uint64_t bucket = sv.bucket(ls, s, ls.key, seed_vid);
                // warp-cooperative single probe (8 lanes load the bucket)
                uint64_t result = 0, b = bucket;
                visited = 0;
                while (true) {
                    uint4 v = make_uint4(0, 0, 0, 0);
                    if (tid < 8) v = ld_slot(vertices + (b * 8 + tid));
                    const uint64_t kk = (uint64_t)v.x | ((uint64_t)v.y << 32);
                    const uint64_t pp = (uint64_t)v.z | ((uint64_t)v.w << 32);
                    const uint32_t hit = __ballot_sync(0xFFFFFFFFu, tid < 7 && kk == ls.key && kk != 0);
                    const uint64_t chain = __shfl_sync(0xFFFFFFFFu, kk, 7);
                    visited++;
                    if (hit) { result = __shfl_sync(0xFFFFFFFFu, pp, __ffs(hit) - 1); break; }
                    if (chain == 0) break;
                    b = chain >> WK_KEY_VID_SHIFT;
                }
                // the whole seed is this warp's business (the other warps only execute the barrier below: every
                // instruction of the interpreter that all 8..32 warps run costs issue slots of the one that works)
                const uint64_t size = ptr_size(result), off = ptr_off(result);
                uint64_t begin = 0, len = size;
                if (ls.mt_factor > 1) {   // mt slicing (index starts only), sparql.hpp:211-221
                    const uint64_t mtf = (uint64_t)ls.mt_factor, start = (uint64_t)ls.mt_tid % mtf, length = size / mtf;
                    begin = start * length;
                    len = (start == mtf - 1) ? (size - begin) : length;
                }
                if (len <= LIGHT_ROWS) {
                    const uint32_t *seed_edges = sv.edges(seed_vid) + off + begin;
                    for (uint32_t k = tid; k < len; k += 32) sm.tab[nxt][k] = ld_edge(seed_edges + k);
                }
                if (tid == 0) { sm.seed_ptr = result; sm.seed_len = len > LIGHT_ROWS ? 0xFFFFFFFFu : (uint32_t)len; st_visited = visited; st_edges = len; }
            }
            __syncthreads();
            const uint32_t slen = sm.seed_len;
            if (slen == 0xFFFFFFFFu) { spilled = true; break; }   // nothing done yet: resume at this very step
            N = slen;
            C = 1;
        } else {
            const int Cin = ls.C;
            const int Cout = (ls.kind == LKIND_K2U) ? Cin + 1 : Cin;
            const uint32_t *tin = sm.tab[cur];
            // phase 1: all probes of the step in flight at once (thread per row).  Later steps that start
            // from a column which already exists ("star" plans: several patterns on the same variable) are
            // shadow-probed at the same time: their bucket and edge lines are pulled into L2 now, so that
            // those steps later see L2 hits instead of a chain of cold DRAM + page-walk latencies.
            int sh[3], nsh = 0;
            for (int s2 = s + 1; s2 < nsteps && nsh < 3; s2++) {
                const LightStep &l2 = steps[s2];
                if (l2.kind >= LKIND_K2U && l2.col_start < Cin && !((shadowed >> s2) & 1u)) { sh[nsh++] = s2; shadowed |= 1u << s2; }
            }
            if ((uint64_t)N * (uint64_t)(1 + nsh) > 2048) nsh = 0;
            uint32_t *tout = sm.tab[nxt];
            // ---- star round: consecutive steps that only read columns which exist NOW do not depend on each other (a join is
            // commutative): LUBM Q4 is "?X worksFor D . ?X type T . ?X name ?a . ?X email ?b . ?X phone ?c" -- four patterns on ?X.
            // One warp per pattern probes all (<= 32) rows at the same time; a row's fan-out is the product of the patterns'
            // multiplicities and an output row picks one edge per expanding pattern.  Four serial steps (each a chain of
            // dependent instructions and memory round trips) become one.  Not taken when per-step statistics are collected
            // (the round probes rows a serial execution would already have dropped) or when the result would not fit.
            if (WARPM && N <= STAR_ROWS && stats == nullptr && NT >= 128) {
                int L = 1;
                while (s + L < nsteps && L < 4) {
                    const LightStep &x = steps[s + L];
                    if (x.kind < LKIND_K2U || x.col_start >= Cin || (x.kind == LKIND_K2K && x.col_end >= Cin)) break;
                    L++;
                }
                if (L >= 2) {
                    const int w = tid >> 5, lane = tid & 31;
                    if (w < L) {
                        const LightStep &x = steps[s + w];
                        for (uint32_t r = lane; r < N; r += 32) {   // every probe of the pattern in flight before the first is consumed
                            const uint32_t c0 = tin[r * Cin + x.col_start];
                            const uint64_t key = step_key(x.seg, c0);
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(sv.vertices(c0) + sv.bucket(x, s + w, key, c0) * 8));
                        }
                        for (uint32_t r = lane; r < N; r += 32) {
                            const uint32_t c0 = tin[r * Cin + x.col_start];
                            const uint64_t key = step_key(x.seg, c0);
                            uint32_t visited, m;
                            const uint64_t ptr = probe_thread(sv.vertices(c0), key, sv.bucket(x, s + w, key, c0), visited);
                            const uint32_t size = ptr_size(ptr);
                            if (x.kind == LKIND_K2U) {
                                m = size > LIGHT_ROWS ? (uint32_t)LIGHT_ROWS + 1u : size;
                                if (size) asm volatile("prefetch.global.L2 [%0];" ::"l"(sv.edges(c0) + ptr_off(ptr)));
                            } else {
                                const uint32_t target = (x.kind == LKIND_K2K) ? tin[r * Cin + x.col_end] : x.end_const;
                                uint32_t scanned;
                                m = list_contains(sv.edges(c0) + ptr_off(ptr), size, target, scanned) ? 1u : 0u;
                            }
                            sm.pre[w * STAR_ROWS + r] = m;
                            sm.ptr[w * STAR_ROWS + r] = ptr;
                        }
                    }
                    __syncthreads();
                    int nk2u = 0;
                    for (int j = 0; j < L; j++) nk2u += steps[s + j].kind == LKIND_K2U ? 1 : 0;
                    const int CoutR = Cin + nk2u;
                    if (tid < 32) {
                        uint32_t base = 0;
                        for (uint32_t r0 = 0; r0 < N; r0 += 32) {      // fan-out of a row = product over the patterns; prefix over the rows
                            const uint32_t r = r0 + lane;
                            uint64_t M = r < N ? 1u : 0u;
                            for (int j = 0; j < L && r < N; j++) {
                                M *= sm.pre[j * STAR_ROWS + r];
                                if (M > LIGHT_ROWS) M = (uint64_t)LIGHT_ROWS + 1;
                            }
                            uint32_t incl = (uint32_t)M;
#pragma unroll
                            for (int o = 1; o < 32; o <<= 1) {
                                const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                                if (lane >= o) incl += y;
                            }
                            if (r < N) sm.pre[4 * STAR_ROWS + r] = base + incl - (uint32_t)M;
                            const uint32_t tot = __shfl_sync(0xFFFFFFFFu, incl, 31);
                            base = (base + tot > 2u * LIGHT_ROWS) ? 2u * LIGHT_ROWS : base + tot;   // saturate: it will not fit anyway
                        }
                        if (lane == 0) {
                            sm.total = base;
                            sm.wsum[0] = (base <= LIGHT_ROWS && (uint64_t)base * (uint64_t)CoutR <= LIGHT_WORDS) ? 0u : 1u;
                        }
                    }
                    __syncthreads();
                    if (sm.wsum[0] == 0) {
                        const uint32_t total = sm.total;
                        const uint32_t *rpre = sm.pre + 4 * STAR_ROWS;
                        for (uint32_t o = tid; o < total; o += NT) {
                            uint32_t lo = 0, hi = N;   // largest r with rpre[r] <= o (rows without fan-out share their successor's prefix)
                            while (hi - lo > 1) {
                                const uint32_t mid = (lo + hi) >> 1;
                                if (rpre[mid] <= o) lo = mid; else hi = mid;
                            }
                            const uint32_t r = lo;
                            uint32_t idx = o - rpre[r];
                            uint32_t *dst = tout + o * CoutR;
                            for (int c = 0; c < Cin; c++) dst[c] = tin[r * Cin + c];
                            int kc = Cin;
                            for (int j = 0; j < L; j++) {
                                const LightStep &x = steps[s + j];
                                if (x.kind != LKIND_K2U) continue;
                                const uint32_t sz = sm.pre[j * STAR_ROWS + r];
                                const uint32_t k = idx % sz;
                                idx /= sz;
                                dst[kc++] = ld_edge(sv.edges(tin[r * Cin + x.col_start]) + ptr_off(sm.ptr[j * STAR_ROWS + r]) + k);
                            }
                        }
                        if (tid == 0 && counts)
                            for (int j = 0; j < L; j++) counts[s + j + 1] = total;
                        N = total;
                        C = CoutR;
                        s += L - 1;
                        cur = nxt;
                        __syncthreads();
                        continue;
                    }
                    // does not fit shared memory as a whole: take the steps one by one (nothing has been written)
                }
            }
            if (WARPM && N <= 32) {
                // ---- warp mode: a table of at most 32 rows is ONE warp's business: lane = row, multiplicities are scanned
                // with shuffles, rows are compacted straight into the next table, and the step costs one CTA barrier instead
                // of four plus a block-wide scan.  The other warps pull the lines of later steps towards L2 meanwhile.
                const int lane = tid & 31;
                if (tid < 32) {
                    const bool act = (uint32_t)lane < N;
                    uint64_t ptr = 0;
                    uint32_t c0 = 0, mult = 0, size = 0;
                    if (act) {
                        c0 = tin[lane * Cin + ls.col_start];
                        const uint64_t key = step_key(ls.seg, c0);
                        uint32_t visited;
                        ptr = probe_thread(sv.vertices(c0), key, sv.bucket(ls, s, key, c0), visited);
                        st_visited += visited;
                        size = ptr_size(ptr);
                        if (ls.kind == LKIND_K2U) {
                            mult = size > LIGHT_ROWS ? (uint32_t)LIGHT_ROWS + 1u : size;
                            st_edges += size;
                        } else {
                            const uint32_t target = (ls.kind == LKIND_K2K) ? tin[lane * Cin + ls.col_end] : ls.end_const;
                            uint32_t scanned;
                            mult = list_contains(sv.edges(c0) + ptr_off(ptr), size, target, scanned) ? 1u : 0u;
                            st_edges += scanned;
                        }
                    }
                    uint32_t incl = mult;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                        if (lane >= o) incl += y;
                    }
                    const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
                    const uint32_t excl = incl - mult;
                    const bool fits = total <= LIGHT_ROWS && (uint64_t)total * (uint64_t)Cout <= LIGHT_WORDS;
                    // 0: rows written by this warp, 1: spill, 2: known_to_unknown with a large fan-out: every thread of the CTA
                    // takes output rows (prefix and pointers left in shared memory)
                    uint32_t how = fits ? 0u : 1u;
                    if (fits && ls.kind == LKIND_K2U && total > 64) how = 2u;
                    if (how == 0) {
                        if (ls.kind == LKIND_K2U) {
                            const uint32_t *e = sv.edges(c0) + ptr_off(ptr);
                            for (uint32_t k = 0; k < mult; k++) {
                                uint32_t *dst = tout + (excl + k) * Cout;
                                const uint32_t ev = ld_edge(e + k);
                                for (int c = 0; c < Cin; c++) dst[c] = tin[lane * Cin + c];
                                dst[Cin] = ev;
                            }
                        } else if (mult) {
                            for (int c = 0; c < Cin; c++) tout[excl * Cout + c] = tin[lane * Cin + c];
                        }
                    } else if (how == 2 && act) {
                        sm.pre[lane] = excl;
                        sm.ptr[lane] = ptr;
                    }
                    if (lane == 0) { sm.total = total; sm.wsum[0] = how; }
                } else if (nsh && tid < 64) {   // one more warp pulls the lines of later steps towards L2; the rest only meet at the barrier
                    for (uint32_t item = tid - 32; item < N * (uint32_t)nsh; item += 32) {
                        uint32_t r = item, j = 0;
                        while (r >= N) { r -= N; j++; }
                        const LightStep &l2 = steps[sh[j]];
                        const uint32_t c2 = tin[r * Cin + l2.col_start];
                        const uint64_t key2 = step_key(l2.seg, c2);
                        uint32_t v2;
                        const uint64_t ptr2 = probe_thread(sv.vertices(c2), key2, sv.bucket(l2, sh[j], key2, c2), v2);
                        if (ptr2) asm volatile("prefetch.global.L2 [%0];" ::"l"(sv.edges(c2) + ptr_off(ptr2)));
                    }
                }
                if (trace && tid == 0) trace[32 + 4 * s] = clock64();
                __syncthreads();
                const uint32_t total = sm.total, how = sm.wsum[0];
                if (trace && tid == 0) trace[32 + 4 * s + 1] = clock64();
                if (how == 1) { spilled = true; break; }
                if (how == 2) {
                    for (uint32_t o = tid; o < total; o += NT) {
                        uint32_t lo = 0, hi = N;   // largest r with pre[r] <= o
                        while (hi - lo > 1) {
                            const uint32_t mid = (lo + hi) >> 1;
                            if (sm.pre[mid] <= o) lo = mid; else hi = mid;
                        }
                        const uint32_t e = ld_edge(sv.edges(tin[lo * Cin + ls.col_start]) + ptr_off(sm.ptr[lo]) + (o - sm.pre[lo]));
                        for (int c = 0; c < Cin; c++) tout[o * Cout + c] = tin[lo * Cin + c];
                        tout[o * Cout + Cin] = e;
                    }
                }
                N = total;
                C = Cout;
                if (trace && tid == 0) trace[32 + 4 * s + 2] = clock64();
            } else {
            // tables taller than the CTA are walked in several rounds whose loads depend on each other through
            // the loop; pull every later round's bucket line towards L2 first so that only round one pays DRAM
            for (uint32_t r = tid + NT; r < N; r += NT) {
                const uint32_t c0 = tin[r * Cin + ls.col_start];
                const uint64_t key = step_key(ls.seg, c0);
                asm volatile("prefetch.global.L2 [%0];" ::"l"(sv.vertices(c0) + sv.bucket(ls, s, key, c0) * 8));
            }
            // A table taller than the CTA takes several rounds per thread.  A round that probes AND reads its edge list
            // chains two cold accesses per round; instead every round first resolves its key (the bucket line was
            // prefetched above) and asks for the edge line, and the lists are read in a second pass: two cold latencies
            // per step however tall the table is.  (A thread owns the same rows in both passes: no barrier in between.)
            const bool two_pass = N > (uint32_t)NT;
            for (uint32_t item = tid; item < N * (uint32_t)(1 + nsh); item += NT) {
                uint32_t r = item, j = 0;   // item = j * N + r with j <= 3: no integer division on this path
                while (r >= N) { r -= N; j++; }
                if (j != 0) {
                    const LightStep &l2 = steps[sh[j - 1]];
                    const uint32_t c2 = tin[r * Cin + l2.col_start];
                    const uint64_t key2 = step_key(l2.seg, c2);
                    uint32_t v2;
                    const uint64_t ptr2 = probe_thread(sv.vertices(c2), key2, sv.bucket(l2, sh[j - 1], key2, c2), v2);
                    if (ptr2) asm volatile("prefetch.global.L2 [%0];" ::"l"(sv.edges(c2) + ptr_off(ptr2)));
                    continue;
                }
                const uint32_t c0 = tin[r * Cin + ls.col_start];
                const uint64_t key = step_key(ls.seg, c0);
                const uint64_t bucket = sv.bucket(ls, s, key, c0);
                uint32_t visited;
                const uint64_t ptr = probe_thread(sv.vertices(c0), key, bucket, visited);
                st_visited += visited;
                sm.ptr[r] = ptr;
                const uint32_t size = ptr_size(ptr);
                if (ls.kind == LKIND_K2U) {
                    // anything above LIGHT_ROWS spills anyway: clamp so that the 32-bit scan over <= 1024 rows cannot wrap
                    sm.pre[r] = size > LIGHT_ROWS ? (uint32_t)LIGHT_ROWS + 1u : size;
                    st_edges += size;
                    if (size) asm volatile("prefetch.global.L2 [%0];" ::"l"(sv.edges(c0) + ptr_off(ptr)));   // read when the rows are materialised
                } else if (two_pass) {
                    if (size) asm volatile("prefetch.global.L2 [%0];" ::"l"(sv.edges(c0) + ptr_off(ptr)));
                } else {
                    const uint32_t target = (ls.kind == LKIND_K2K) ? tin[r * Cin + ls.col_end] : ls.end_const;
                    uint32_t scanned;
                    const bool hit = list_contains(sv.edges(c0) + ptr_off(ptr), size, target, scanned);
                    st_edges += scanned;
                    sm.pre[r] = hit ? 1u : 0u;
                }
            }
            if (two_pass && ls.kind != LKIND_K2U) {
                for (uint32_t r = tid; r < N; r += NT) {
                    const uint32_t c0 = tin[r * Cin + ls.col_start];
                    const uint64_t ptr = sm.ptr[r];
                    const uint32_t target = (ls.kind == LKIND_K2K) ? tin[r * Cin + ls.col_end] : ls.end_const;
                    uint32_t scanned;
                    const bool hit = list_contains(sv.edges(c0) + ptr_off(ptr), ptr_size(ptr), target, scanned);
                    st_edges += scanned;
                    sm.pre[r] = hit ? 1u : 0u;
                }
            }
            if (trace && tid == 0) trace[32 + 4 * s] = clock64();
            __syncthreads();
            // phase 2: scan multiplicities
            light_scan<NT>(sm, N, tid);
            const uint32_t total = sm.total;
            if (trace && tid == 0) trace[32 + 4 * s + 1] = clock64();
            if (total > LIGHT_ROWS || (uint64_t)total * (uint64_t)Cout > LIGHT_WORDS) { spilled = true; break; }
            // phase 3: materialise (thread per OUTPUT row: every edge load of the step in flight at once)
            if (ls.kind == LKIND_K2U) {
                for (uint32_t o = tid; o < total; o += NT) {
                    uint32_t lo = 0, hi = N;   // largest r with pre[r] <= o
                    while (hi - lo > 1) {
                        const uint32_t mid = (lo + hi) >> 1;
                        if (sm.pre[mid] <= o) lo = mid; else hi = mid;
                    }
                    const uint32_t k = o - sm.pre[lo];
                    const uint32_t e = ld_edge(sv.edges(tin[lo * Cin + ls.col_start]) + ptr_off(sm.ptr[lo]) + k);
                    for (int c = 0; c < Cin; c++) tout[o * Cout + c] = tin[lo * Cin + c];
                    tout[o * Cout + Cin] = e;
                }
            } else {
                for (uint32_t r = tid; r < N; r += NT) {
                    const uint32_t p0 = sm.pre[r];
                    const uint32_t p1 = (r + 1 < N) ? sm.pre[r + 1] : total;
                    if (p1 != p0)
                        for (int c = 0; c < Cin; c++) tout[p0 * Cout + c] = tin[r * Cin + c];
                }
            }
            N = total;
            C = Cout;
            if (trace && tid == 0) trace[32 + 4 * s + 2] = clock64();
            }   // block mode
        }
        // per-step statistics (algorithmic-bytes accounting)
        if ((stats != nullptr)) {
            const uint64_t v = block_sum_u64<NT>(st_visited, sm, tid);
            const uint64_t e = block_sum_u64<NT>(st_edges, sm, tid);
            if (tid == 0) { stats[2 * s] = v; stats[2 * s + 1] = e; }
        }
        if (tid == 0 && counts) counts[s + 1] = N;
        if (trace && tid == 0) trace[2 + s] = clock64();
        cur = nxt;
        __syncthreads();
    }
    LightState st;
    st.N = N; st.C = C; st.cur = cur; st.done = s; st.spilled = spilled;
    return st;
}

// The step descriptors are read many times and by data-dependent index.  Kernel parameters are freshly
// written per launch (every first touch of a parameter line is a miss), so all lines are fetched at once,
// by all threads, into shared memory instead of one by one on the interpreter's critical path.
template <int NT>
__device__ __forceinline__ void stage_steps(LightStep *dst, const LightStep *src, int nsteps, int tid) {
    const uint32_t words = (uint32_t)nsteps * (uint32_t)(sizeof(LightStep) / sizeof(uint32_t));
    const uint32_t *s = reinterpret_cast<const uint32_t *>(src);
    uint32_t *d = reinterpret_cast<uint32_t *>(dst);
    for (uint32_t i = tid; i < words; i += NT) d[i] = s[i];
}

// whole light query: interpret, then project into the mapped staging area (or hand the table over on a spill) and
// store the completion record.  Returns true when the table outgrew shared memory.
// `s_steps` are the step descriptors in shared memory.  clear_ctl: the launch-per-query kernel owns the control block and
// clears it up front; the resident server (wk_server.cuh) only touches it when it has to (spill, per-step statistics).
template <int NT, class SV, bool WARPM = true>
__device__ __forceinline__ bool light_run(const LightPlan &plan, const LightStep *s_steps, const SV &sv, LightSmem &sm, bool clear_ctl,
                                          uint64_t *times = nullptr, uint64_t t_acquired = 0) {
    const int tid = threadIdx.x;
    if (plan.trace && tid == 0) plan.trace[0] = clock64();
    if (clear_ctl) {
        for (int i = tid; i < plan.ctl_nwords; i += NT) plan.ctl_words[i] = 0;
        __syncthreads();
    }
    if (plan.trace && tid == 0) plan.trace[1] = clock64();
    const LightState ls_ = light_interpret<NT, SV, WARPM>(s_steps, plan.nsteps, sv, sm, (clear_ctl && plan.collect_stats) ? plan.stats : nullptr,
                                           clear_ctl ? plan.counts : nullptr, plan.trace);
    const uint32_t N = ls_.N;
    const int C = ls_.C, cur = ls_.cur, s = ls_.done;
    const bool spilled = ls_.spilled;
    uint32_t status = 0;
    const int done_steps = s;   // steps [0, done_steps) ran; the table (N x C) is in sm.tab[cur]
    uint64_t tsum = 0;
    if (spilled) {
        // hand the table over to the multi-CTA path: buf[done_steps & 1], counts[done_steps]
        if (!clear_ctl) {   // the control block may hold a previous query's counters and status
            for (int i = tid; i < plan.ctl_nwords; i += NT) plan.ctl_words[i] = 0;
            __syncthreads();
            if (tid == 0) plan.counts[done_steps] = N;
        }
        if (done_steps > 0) {
            uint32_t *dst = plan.buf[done_steps & 1];
            const uint32_t words = N * (uint32_t)C;
            if ((uint64_t)words <= plan.cap_words) {
                for (uint32_t i = tid; i < words; i += NT) dst[i] = sm.tab[cur][i];
            } else {
                status = 1;
            }
        }
        __threadfence();   // the continuation is launched by the host once it has seen the record
    } else if (plan.do_project && N > 0) {
        // final_process projection straight into the mapped staging area
        const uint32_t words = N * (uint32_t)plan.proj_n;
        uint64_t part = 0;
        if (words <= 128) {
            // a handful of rows: one warp writes and sums them, nobody else executes a reduction
            if (tid < 32) {
                for (uint32_t w = tid; w < words; w += 32) {
                    const uint32_t r = w / (uint32_t)plan.proj_n, j = w - r * (uint32_t)plan.proj_n;
                    const uint32_t val = sm.tab[cur][r * C + plan.proj_cols[j]];
                    st_sys_u32(plan.host_table + w, val);
                    part += table_word_mix(val, w);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) part += __shfl_down_sync(0xFFFFFFFFu, part, o);
                tsum = part;   // lane 0 holds the sum; it also writes the record
            }
        } else {
            if ((uint64_t)words <= plan.host_table_words) {
                for (uint32_t w = tid; w < words; w += NT) {
                    const uint32_t r = w / (uint32_t)plan.proj_n, j = w - r * (uint32_t)plan.proj_n;
                    const uint32_t val = sm.tab[cur][r * C + plan.proj_cols[j]];
                    st_sys_u32(plan.host_table + w, val);
                    part += table_word_mix(val, w);
                }
            } else {
                status = 1;
            }
            tsum = block_sum_u64<NT>(part, sm, tid);
        }
    }
    if (spilled) __syncthreads();
    if (plan.trace && tid == 0) plan.trace[2 + MAX_LIGHT_STEPS] = clock64();
    if (tid == 0) {
        if (status) *plan.status = status;
        const uint64_t rows = N;
        const uint64_t sr = (uint64_t)status | ((uint64_t)(uint32_t)done_steps << 32);
        uint64_t *rec = (uint64_t *)plan.rec;
        if (times) {   // diagnostics of the resident server: in-kernel span of this request (not covered by the checksum)
            uint64_t now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            st_sys_v2u64(times, t_acquired, now);
        }
        st_sys_v2u64(rec + 2, sr, record_check(plan.seq, rows, sr, tsum));
        st_sys_v2u64(rec, plan.seq, rows);
        if (plan.trace) plan.trace[3 + MAX_LIGHT_STEPS] = clock64();
    }
    return spilled;
}

template <int NT, class SV>
__device__ __forceinline__ bool light_query_body(const LightPlan &plan, const SV &sv, LightSmem &sm, LightStep *s_steps) {
    stage_steps<NT>(s_steps, plan.steps, plan.nsteps, threadIdx.x);
    return light_run<NT>(plan, s_steps, sv, sm, true);
}

__global__ void __launch_bounds__(LIGHT_THREADS) light_query_kernel(const __grid_constant__ LightPlan plan) {
    __shared__ LightSmem sm;
    __shared__ LightStep s_steps[MAX_LIGHT_STEPS];
    LocalView sv;
    sv.v = plan.vertices;
    sv.e = plan.edges;
    light_query_body<LIGHT_THREADS>(plan, sv, sm, s_steps);
}

// The same on a sharded store: this GPU owns the query's constant and walks the other shards through peer memory.
// When done it tells every peer whether the answer stands (flag = 2 * epoch) or the table outgrew shared memory and
// the query has to be redone by all shards together (flag = 2 * epoch + 1).
struct LightPlanSharded {
    LightPlan lp;
    const uint4 *pv[LIGHT_PEERS];
    const uint32_t *pe[LIGHT_PEERS];
    uint64_t *peer_flag[LIGHT_PEERS];       // &XchCtl::flagL[owner] inside every peer's control block (nullptr for myself)
    uint64_t epoch;
    uint32_t nranks, _pad;
    SegLite segr[LIGHT_SHARDED_STEPS][LIGHT_PEERS];
};

__device__ __forceinline__ void st_sys_u64_light(uint64_t *p, uint64_t v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// 1024 threads: a remote access costs about a microsecond, so a table taller than the CTA (three dependent rounds of
// probe + list scan with 256 threads) is worth far more than the extra barrier cost measured on the local kernel
enum { LIGHT_SHARDED_THREADS = 1024 };
__global__ void __launch_bounds__(LIGHT_SHARDED_THREADS) light_sharded_kernel(const __grid_constant__ LightPlanSharded plan) {
    __shared__ LightSmem sm;
    __shared__ LightStep s_steps[LIGHT_SHARDED_STEPS];
    PeerView sv;
    sv.v = plan.pv;
    sv.e = plan.pe;
    sv.segr = plan.segr;
    sv.n = plan.nranks;
    const bool spilled = light_query_body<LIGHT_SHARDED_THREADS>(plan.lp, sv, sm, s_steps);
    if (threadIdx.x < plan.nranks && plan.peer_flag[threadIdx.x] != nullptr)
        st_sys_u64_light(plan.peer_flag[threadIdx.x], 2 * plan.epoch + (spilled ? 1 : 0));
}


// ---- throughput path: many independent light plans in one launch, one CTA per query (blind) -------------
enum { BATCH_STEPS = 8 };
struct BatchPlan {
    int32_t nsteps, _pad;
    LightStep steps[BATCH_STEPS];
};
struct BatchResult {
    uint64_t rows;
    uint32_t status;   // 0 ok, 2 = table outgrew shared memory: run this query through wk_query_execute
    uint32_t done;
};

__global__ void __launch_bounds__(CTA_THREADS) light_batch_kernel(const BatchPlan *plans, BatchResult *results, int nqueries,
                                                                  const uint4 *vertices, const uint32_t *edges) {
    __shared__ LightSmem sm;
    __shared__ LightStep s_steps[BATCH_STEPS];
    LocalView sv;
    sv.v = vertices;
    sv.e = edges;
    for (int q = blockIdx.x; q < nqueries; q += gridDim.x) {
        const BatchPlan *bp = plans + q;
        const int nsteps = bp->nsteps;
        stage_steps<CTA_THREADS>(s_steps, bp->steps, nsteps, threadIdx.x);
        __syncthreads();
        const LightState st = light_interpret<CTA_THREADS>(s_steps, nsteps, sv, sm, nullptr, nullptr);
        if (threadIdx.x == 0) {
            results[q].rows = st.N;
            results[q].status = st.spilled ? 2u : 0u;
            results[q].done = (uint32_t)st.done;
        }
        __syncthreads();
    }
}

}  // namespace wk

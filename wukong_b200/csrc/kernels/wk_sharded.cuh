// Sharded execution: the store is partitioned by vid % nranks (reference utils/math.hpp:51-55,
// base_loader.hpp:343-351); before a step whose start variable is not the current sharding column
// the binding table is bucketised by row[col] % nranks and exchanged all-to-all(v) over NCCL
// (replaces generate_sub_query + the Bundle/RDMA hop: sparql.hpp:746-814, gpu_hash.cu:599-760,
// gpu_engine_cuda.hpp:364-407).  Included by engine.cu (needs wk_engine internals).
#pragma once
#include <dlfcn.h>
#include <nccl.h>

// ---- bucketise-by-owner kernels -----------------------------------------------------------------
enum { MAX_PARTS = 64 };

__global__ void __launch_bounds__(CTA_THREADS) part_count_kernel(const uint32_t *in, const uint64_t *in_count, int C, int col,
                                                                 uint32_t nparts, uint64_t *counts) {
    __shared__ uint32_t hist[MAX_PARTS];
    if (threadIdx.x < MAX_PARTS) hist[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t N = ld_count(in_count);
    for (uint64_t r = (uint64_t)blockIdx.x * CTA_THREADS + threadIdx.x; r < N; r += (uint64_t)gridDim.x * CTA_THREADS)
        atomicAdd(&hist[ld_table(in + r * (uint64_t)C + col) % nparts], 1u);
    __syncthreads();
    if (threadIdx.x < nparts && hist[threadIdx.x]) atomicAdd((unsigned long long *)&counts[threadIdx.x], (unsigned long long)hist[threadIdx.x]);
}

__global__ void part_scan_kernel(const uint64_t *counts, uint64_t *cursor, uint32_t nparts) {
    uint64_t acc = 0;
    for (uint32_t d = 0; d < nparts; d++) { cursor[d] = acc; acc += counts[d]; }
}

// each tile reserves a run per destination, then places its rows (order inside a run is free)
__global__ void __launch_bounds__(CTA_THREADS) part_scatter_kernel(const uint32_t *in, const uint64_t *in_count, int C, int col,
                                                                   uint32_t nparts, uint64_t *cursor, uint32_t *out) {
    __shared__ uint32_t hist[MAX_PARTS];
    __shared__ uint64_t base[MAX_PARTS];
    const uint64_t N = ld_count(in_count);
    for (uint64_t t0 = (uint64_t)blockIdx.x * CTA_THREADS; t0 < N; t0 += (uint64_t)gridDim.x * CTA_THREADS) {
        if (threadIdx.x < MAX_PARTS) hist[threadIdx.x] = 0;
        __syncthreads();
        const uint64_t r = t0 + threadIdx.x;
        uint32_t d = 0, local = 0;
        if (r < N) {
            d = ld_table(in + r * (uint64_t)C + col) % nparts;
            local = atomicAdd(&hist[d], 1u);
        }
        __syncthreads();
        if (threadIdx.x < nparts && hist[threadIdx.x])
            base[threadIdx.x] = atomicAdd((unsigned long long *)&cursor[threadIdx.x], (unsigned long long)hist[threadIdx.x]);
        __syncthreads();
        if (r < N) {
            const uint32_t *src = in + r * (uint64_t)C;
            uint32_t *dst = out + (base[d] + local) * (uint64_t)C;
            for (int c = 0; c < C; c++) dst[c] = ld_table(src + c);
        }
        __syncthreads();
    }
}

// ---- NCCL through dlopen (no link-time dependency; shares the copy torch may already have loaded) ----
struct NcclApi {
    void *h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

static NcclApi &nccl_api() {
    static NcclApi api;
    if (api.h) return api;
    api.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!api.h) api.h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!api.h) { fprintf(stderr, "[wukong_b200] cannot load libnccl: %s\n", dlerror()); return api; }
#define WK_NCCL_SYM(field, name) *(void **)(&api.field) = dlsym(api.h, name)
    WK_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
    WK_NCCL_SYM(CommInitRank, "ncclCommInitRank");
    WK_NCCL_SYM(CommDestroy, "ncclCommDestroy");
    WK_NCCL_SYM(GroupStart, "ncclGroupStart");
    WK_NCCL_SYM(GroupEnd, "ncclGroupEnd");
    WK_NCCL_SYM(Send, "ncclSend");
    WK_NCCL_SYM(Recv, "ncclRecv");
    WK_NCCL_SYM(AllGather, "ncclAllGather");
    WK_NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef WK_NCCL_SYM
    api.ok = api.GetUniqueId && api.CommInitRank && api.GroupStart && api.GroupEnd && api.Send && api.Recv && api.AllGather;
    return api;
}

#define NCCL_TRY(x)                                                                                        \
    do {                                                                                                   \
        ncclResult_t _r = (x);                                                                             \
        if (_r != ncclSuccess) {                                                                           \
            fprintf(stderr, "[wukong_b200] NCCL error at %s:%d: %s\n", __FILE__, __LINE__,                 \
                    nccl_api().GetErrorString ? nccl_api().GetErrorString(_r) : "?");                      \
            return WK_ERR_COMM;                                                                            \
        }                                                                                                  \
    } while (0)

struct wk_comm {
    ncclComm_t comm = nullptr;
    int nranks = 1, rank = 0;
    uint64_t *d_counts = nullptr;    // [MAX_PARTS] rows per destination
    uint64_t *d_cursor = nullptr;    // [MAX_PARTS] running offsets
    uint64_t *d_matrix = nullptr;    // [nranks][MAX_PARTS] all-gathered counts
    uint64_t *h_matrix = nullptr;    // pinned copy
    uint64_t part_rows[MAX_PARTS] = {0}, part_off[MAX_PARTS] = {0};
    bool partitioned = false;
    uint64_t exchanges = 0, rows_sent = 0, rows_recv = 0;
    // peer-memory path (CUDA IPC)
    bool p2p_ready = false;
    struct XchCtl *d_xctl = nullptr;
    struct P2PLocal *d_p2p_local = nullptr;
    struct P2PTable *p2p = nullptr;     // host copy of the peer pointer table (passed to kernels by value)
    uint64_t epoch = 0;
    std::vector<void *> ipc_opened;
    // peers' stores mapped through CUDA IPC (in-place light queries): header / edge arrays and segment tables per rank
    bool peer_stores = false;
    const uint4 *peer_v[8] = {nullptr};
    const uint32_t *peer_e[8] = {nullptr};
    std::vector<std::map<std::tuple<int, uint32_t, int>, wk_segmeta_t>> peer_segs;
};

// Which steps need an exchange: out[i] = -1 none, -2 replicate to every rank, c >= 0 re-shard by column c.
// Mirrors need_fork_join()/local_var (sparql.hpp:802-814) and dispatch(r,false) for type-index
// lookups of a known variable (sparql.hpp:1091-1110).
static void plan_exchanges(const std::vector<PlannedStep> &steps, std::vector<int> &out) {
    out.assign(steps.size(), -1);
    int shard_col = -1;
    for (size_t i = 0; i < steps.size(); i++) {
        const PlannedStep &ps = steps[i];
        if (ps.kind == KIND_I2U) { shard_col = 0; continue; }     // local index slice: the new column is local
        if (ps.kind == KIND_C2U) { shard_col = -1; continue; }    // rows only on the owner of the constant
        if (ps.kind == KIND_K2U && ps.pid == WK_TYPE_ID && ps.dir == WK_DIR_IN) {
            out[i] = -2;
            shard_col = ps.in_cols;   // the appended instances are local to the rank that found them
            continue;
        }
        if (ps.col_start != shard_col) { out[i] = ps.col_start; shard_col = ps.col_start; }
    }
}

// =============================================================================================
// Peer-memory exchange (NVLink / NVSwitch, no NCCL, no host synchronisation):
//   count -> publish counts to every peer + wait for theirs (barrier 1: also proves that every peer has
//   finished the previous step, so its next-table buffer may be overwritten) -> scatter rows STRAIGHT INTO
//   the peers' next-table buffers with plain stores over NVLink -> "pushed" flags (barrier 2).
// Buffers and control blocks of the peers are mapped with CUDA IPC (one process per GPU).
// =============================================================================================
struct XchCtl {
    uint64_t counts[MAX_PARTS][MAX_PARTS];   // [src][dst]; row `src` is written by rank src into every peer's copy
    uint64_t flagA[MAX_PARTS];               // epoch of the last counts publication seen from each rank
    uint64_t flagB[MAX_PARTS];               // epoch of the last completed push seen from each rank
    uint64_t flagL[MAX_PARTS];               // light query answered in place by rank r: 2 * epoch (+ 1: redo it collectively)
};

struct P2PTable {
    uint32_t *buf[2][MAX_PARTS / 4];         // peers' result buffers (up to 16 ranks per box)
    XchCtl *ctl[MAX_PARTS / 4];
    int nranks, rank;
};
enum { P2P_MAX_RANKS = MAX_PARTS / 4 };

struct P2PLocal {                            // device scratch of one rank
    uint64_t cursor[MAX_PARTS];              // running row offset into each destination's buffer
    uint32_t done_ctas;
    uint32_t skip;                           // set when the exchange must not push (overflow somewhere)
    uint64_t rows_sent, rows_recv;           // running totals over all exchanges (wk_comm_stats)
};

__device__ __forceinline__ void st_sys_u64(uint64_t *p, uint64_t v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_sys_u64(const uint64_t *p) {
    uint64_t v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// spin until *p == want; bounded (about 2 s) so that a dead peer cannot hang the GPU
__device__ __forceinline__ bool wait_flag(const uint64_t *p, uint64_t want) {
    for (uint32_t i = 0; i < (1u << 23); i++) {
        if (ld_sys_u64(p) >= want) return true;
        __nanosleep(200);
    }
    return false;
}

// barrier 1: publish my per-destination counts to every peer, wait for all of theirs, derive offsets
__global__ void p2p_publish_kernel(P2PTable t, XchCtl *my, P2PLocal *loc, const uint64_t *my_counts, const uint64_t *in_count,
                                   int dup, uint64_t epoch, uint64_t cap_rows, uint64_t *out_count, uint32_t *status) {
    const int p = threadIdx.x;
    __shared__ int ok;
    if (p == 0) ok = 1;
    __syncthreads();
    if (p < t.nranks) {
        XchCtl *peer = t.ctl[p];
        for (int d = 0; d < t.nranks; d++) st_sys_u64(&peer->counts[t.rank][d], dup ? ld_count(in_count) : my_counts[d]);
        __threadfence_system();
        st_sys_u64(&peer->flagA[t.rank], epoch);
        if (!wait_flag(&my->flagA[p], epoch)) atomicExch(&ok, 0);
    }
    __syncthreads();
    if (p == 0) {
        loc->done_ctas = 0;
        loc->skip = 0;
        if (!ok) { atomicOr(status, 2u); loc->skip = 1; *out_count = 0; return; }
        // every rank sees the same matrix: offsets and overflow decisions agree everywhere
        bool overflow = false;
        for (int d = 0; d < t.nranks; d++) {
            uint64_t tot = 0, before = 0;
            for (int src = 0; src < t.nranks; src++) {
                const uint64_t c = ld_sys_u64(&my->counts[src][d]);
                if (src < t.rank) before += c;
                tot += c;
            }
            if (tot > cap_rows) overflow = true;
            loc->cursor[d] = before;
            if (d == t.rank) {
                *out_count = tot;
                loc->rows_recv += tot - ld_sys_u64(&my->counts[t.rank][d]);
            } else {
                loc->rows_sent += ld_sys_u64(&my->counts[t.rank][d]);
            }
        }
        if (overflow) { atomicOr(status, 1u); loc->skip = 1; }
    }
}

// scatter every row into its owner's next-table buffer; the last CTA to finish publishes the "pushed" flag to every
// peer after a system-scope fence.  A tile of RPT * 256 rows is loaded with coalesced reads, grouped by destination in
// shared memory, and every destination's run leaves as consecutive words from consecutive threads: the stores that cross
// NVLink are full, contiguous segments instead of 4-byte pieces of 12-byte rows (2.4x on 16 M rows, 2 GPUs).
template <int RPT>
__global__ void __launch_bounds__(CTA_THREADS) p2p_scatter_kernel(P2PTable t, P2PLocal *loc, const uint32_t *in, const uint64_t *in_count,
                                                                  int C, int col, int dup, int dst_buf, uint64_t epoch) {
    extern __shared__ uint32_t p2p_dyn[];
    constexpr uint32_t TILE = CTA_THREADS * RPT;
    uint32_t *rows = p2p_dyn, *stage = p2p_dyn + (size_t)TILE * C;
    __shared__ uint32_t hist[P2P_MAX_RANKS], off[P2P_MAX_RANKS + 1];
    __shared__ uint64_t base[P2P_MAX_RANKS];
    __shared__ uint32_t last;
    const uint32_t n = (uint32_t)t.nranks;
    const uint64_t N = ld_count(in_count);
    const uint32_t tid = threadIdx.x;
    if (!__ldcg(&loc->skip)) {
        for (uint64_t t0 = (uint64_t)blockIdx.x * TILE; t0 < N; t0 += (uint64_t)gridDim.x * TILE) {
            const uint32_t nrows = (uint32_t)((N - t0 < TILE) ? (N - t0) : TILE);
            const uint32_t words = nrows * (uint32_t)C;
            if (tid < P2P_MAX_RANKS) hist[tid] = 0;
            const uint32_t *src = in + t0 * (uint64_t)C;
            for (uint32_t w = tid; w < words; w += CTA_THREADS) rows[w] = ld_table(src + w);
            __syncthreads();
            if (!dup) {
                uint32_t d[RPT], local[RPT];
#pragma unroll
                for (int j = 0; j < RPT; j++) {
                    const uint32_t r = tid + (uint32_t)j * CTA_THREADS;
                    d[j] = 0; local[j] = 0;
                    if (r < nrows) {
                        d[j] = rows[r * C + col] % n;
                        local[j] = atomicAdd(&hist[d[j]], 1u);
                    }
                }
                __syncthreads();
                if (tid == 0) {
                    uint32_t run = 0;
                    for (uint32_t dd = 0; dd < n; dd++) { off[dd] = run * (uint32_t)C; run += hist[dd]; }   // in words
                    off[n] = run * (uint32_t)C;
                }
                if (tid < n && hist[tid]) base[tid] = atomicAdd((unsigned long long *)&loc->cursor[tid], (unsigned long long)hist[tid]);
                __syncthreads();
#pragma unroll
                for (int j = 0; j < RPT; j++) {
                    const uint32_t r = tid + (uint32_t)j * CTA_THREADS;
                    if (r < nrows) {
                        uint32_t *q = stage + off[d[j]] + local[j] * (uint32_t)C;
                        for (int c = 0; c < C; c++) q[c] = rows[r * C + c];
                    }
                }
                __syncthreads();
                for (uint32_t w = tid; w < words; w += CTA_THREADS) {
                    uint32_t dd = 0;
                    while (w >= off[dd + 1]) dd++;
                    t.buf[dst_buf][dd][base[dd] * (uint64_t)C + (w - off[dd])] = stage[w];
                }
            } else {
                if (tid < n) base[tid] = atomicAdd((unsigned long long *)&loc->cursor[tid], (unsigned long long)nrows);
                __syncthreads();
                for (uint32_t dd = 0; dd < n; dd++) {
                    uint32_t *dst = t.buf[dst_buf][dd] + base[dd] * (uint64_t)C;
                    for (uint32_t w = tid; w < words; w += CTA_THREADS) dst[w] = rows[w];
                }
            }
            __syncthreads();
        }
    }
    // completion: fence my stores system-wide, count CTAs, the last one raises the flags
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(&loc->done_ctas, 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    if (last && threadIdx.x < n) {
        __threadfence_system();
        st_sys_u64(&t.ctl[threadIdx.x]->flagB[t.rank], epoch);
    }
}

// a peer of an in-place light query: wait for the owner's verdict; bit 2 of the status word asks for the collective redo
__global__ void p2p_light_wait_kernel(XchCtl *my, int owner, uint64_t epoch, uint32_t *status) {
    if (!wait_flag(&my->flagL[owner], 2 * epoch)) { atomicOr(status, 2u); return; }
    if (ld_sys_u64(&my->flagL[owner]) == 2 * epoch + 1) atomicOr(status, 4u);
}

// barrier 2: every peer has finished pushing into my buffer
__global__ void p2p_wait_kernel(P2PTable t, XchCtl *my, uint64_t epoch, uint32_t *status) {
    const int p = threadIdx.x;
    if (p < t.nranks && !wait_flag(&my->flagB[p], epoch)) atomicOr(status, 2u);
    __threadfence_system();
}

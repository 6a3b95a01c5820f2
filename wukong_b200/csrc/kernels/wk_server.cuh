// Resident light-query server: one CTA that stays on the GPU and answers const-start ("light") plans posted by the
// host through a doorbell in mapped pinned memory, so that a light query pays neither a kernel launch nor a stream
// round trip.  The reference's engine threads are resident in the same way: they poll their message queues and run a
// light query in microseconds (core/engine/engine.hpp:120-221, core/proxy.hpp:298-385).
//
//   request   448 bytes = 28 chunks of 16 bytes {w0, w1, w2, tag}; the host writes the payload words, then the tags
//             (x86 stores are ordered), lane i of warp 0 polls chunk i with one 16-byte volatile load until every tag
//             equals the low 32 bits of the sequence number it expects: the plan arrives with the doorbell, one PCIe
//             round trip in total.  chunk 0 = header, 1-3 = projection columns, 4.. = one step each.
//   segments  steps name their (pid, dir) segment by slot; bucket_start and the modulo magic come from a table the
//             store keeps on the device (staged in shared memory at start-up).
//   reply     the 32-byte checksummed record of the launch-per-query path (wk_light.cuh), the projected table in the
//             mapped staging area.
//   lifetime  the kernel leaves when told to (QUIT) or when no request arrived for idle_ns: a device-wide
//             synchronisation (cudaFree, cudaDeviceSynchronize, a profiler) therefore never waits longer than that, and
//             a dead host process cannot leave a kernel spinning.  The host notices the exit word and relaunches on demand.
#pragma once
#include "wk_light.cuh"

namespace wk {

enum { SRV_HDR_CHUNKS = 4, SRV_CHUNKS = SRV_HDR_CHUNKS + MAX_LIGHT_STEPS, SRV_SMEM_SEGS = 160 };
static_assert(SRV_CHUNKS <= 32, "one lane of the polling warp per request chunk");
enum { SRV_F_PROJECT = 1, SRV_F_STATS = 2, SRV_F_QUIT = 4, SRV_F_TRACE = 8, SRV_F_WAIT = 16 };

struct SrvChunk { uint32_t w0, w1, w2, tag; };

// host-side layout of the mapped page shared with the server
struct SrvMailbox {
    SrvChunk req[SRV_CHUNKS];                 // host -> device
    uint8_t _pad0[512 - sizeof(SrvChunk) * SRV_CHUNKS];
    LightRecord rec;                          // device -> host (32 bytes)
    uint8_t _pad1[64 - sizeof(LightRecord)];
    uint64_t times[2];                        // device -> host: %globaltimer at acquisition / completion of the last request
    uint8_t _pad2[64 - 16];
    uint64_t exit_word;                       // device -> host: launch id of the last server instance that left
    uint8_t _pad3[64 - 8];
};
static_assert(sizeof(SrvMailbox) == 512 + 64 * 3, "mailbox layout");

struct SrvParams {
    const uint4 *vertices;
    const uint32_t *edges;
    const SegLite *segtab;       // device array, one entry per segment slot of the store
    int32_t nsegs, ctl_nwords;
    const SrvChunk *req;         // device pointers into the mapped mailbox
    LightRecord *rec;
    uint64_t *times;
    uint64_t *exit_word;
    uint32_t *host_table;        // mapped staging area of projected results
    uint64_t host_table_words;
    uint32_t *buf[2];            // engine result buffers (spill target)
    uint64_t cap_words;
    uint64_t *counts, *stats, *ctl_words;
    uint32_t *status;
    uint64_t first_seq;          // sequence number of the first request this instance serves
    uint64_t launch_id;
    uint64_t idle_ns;
    long long *trace;            // device buffer of LIGHT_TRACE_WORDS phase clocks (requests with SRV_F_TRACE)
};

// A sharded store (wk_sharded.cuh): the server of the rank that owns a light query's constant walks the other shards through
// peer memory exactly like light_sharded_kernel and tells the peers its verdict; the peers' servers answer a WAIT request
// (header chunk: w0 = SRV_F_WAIT << 8 | owner << 16, w1 / w2 = epoch) by spinning on that verdict -- so an in-place light
// query costs no launch on any rank.  Owner requests carry the epoch in w1 / w2 of the header chunk as well.
struct SrvPeers {
    const uint4 *pv[LIGHT_PEERS];
    const uint32_t *pe[LIGHT_PEERS];
    uint64_t *peer_flag[LIGHT_PEERS];   // &XchCtl::flagL[me] inside every peer's control block (nullptr for myself)
    const uint64_t *my_flag;            // my XchCtl::flagL, indexed by owner
    const SegLite *segr_tab;            // [my segment slot][LIGHT_PEERS]: where every shard keeps that (pid, dir) segment
    uint32_t nranks, rank;
};

__device__ __forceinline__ uint64_t globaltimer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ uint4 ld_volatile_v4(const void *p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}

// step chunk: w0 = kind | col_start << 8 | col_end << 16 | dir << 24 | index_mode << 25 | C << 26,
//             w1 = end_const (known_to_const) or the constant (const_to_unknown), w2 = pid | segment slot << 17
__host__ __device__ __forceinline__ uint32_t srv_pack_w0(int kind, int col_start, int col_end, int dir, int index_mode, int C) {
    return (uint32_t)kind | ((uint32_t)col_start << 8) | ((uint32_t)col_end << 16) | ((uint32_t)(dir & 1) << 24) |
           ((uint32_t)(index_mode & 1) << 25) | ((uint32_t)(C & 31) << 26);
}

struct SrvSmem {
    LightSmem light;
    LightPlan plan;                 // header fields + decoded steps
    SrvChunk chunk[SRV_CHUNKS];
    SegLite seg[SRV_SMEM_SEGS];
    uint32_t ctrl;                  // 0 run, 1 leave
};

struct SrvSmemSharded {
    SrvSmem base;
    SegLite segr[LIGHT_SHARDED_STEPS][LIGHT_PEERS];
};

template <int NT, bool WARPM, bool SHARDED>
__device__ __forceinline__ void light_server_body(const SrvParams &P, const SrvPeers *Q, SrvSmem &S, SegLite (*segr)[LIGHT_PEERS]) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // constant part of the plan header and the segment table
    for (int i = tid; i < P.nsegs && i < SRV_SMEM_SEGS; i += NT) S.seg[i] = P.segtab[i];
    if (tid == 0) {
        LightPlan &lp = S.plan;
        lp.vertices = P.vertices;
        lp.edges = P.edges;
        lp.buf[0] = P.buf[0];
        lp.buf[1] = P.buf[1];
        lp.counts = P.counts;
        lp.stats = P.stats;
        lp.status = P.status;
        lp.ctl_words = P.ctl_words;
        lp.ctl_nwords = P.ctl_nwords;
        lp.rec = P.rec;
        lp.host_table = P.host_table;
        lp.host_table_words = P.host_table_words;
        lp.cap_words = P.cap_words;
        lp.trace = nullptr;
    }
    __syncthreads();
    LocalView sv;
    sv.v = P.vertices;
    sv.e = P.edges;
    uint64_t seq = P.first_seq;
    uint64_t t_idle0 = globaltimer_ns();
    while (true) {
        // ---- acquisition: warp 0 polls the request chunks, everyone else sleeps on the barrier --------------------
        uint64_t t_acq = 0;
        if (warp == 0) {
            const uint32_t want = (uint32_t)seq;
            uint4 v = make_uint4(0, 0, 0, 0);
            bool leave = false;
            uint32_t polls = 0;
            while (true) {
                if (lane < SRV_CHUNKS) v = ld_volatile_v4(P.req + lane);
                if (__all_sync(0xFFFFFFFFu, lane >= SRV_CHUNKS || v.w == want)) break;
                if ((++polls & 7u) == 0) {   // lane 0 decides for the warp: the lanes must not diverge around the vote above
                    int l = (lane == 0 && globaltimer_ns() - t_idle0 > P.idle_ns) ? 1 : 0;
                    l = __shfl_sync(0xFFFFFFFFu, l, 0);
                    if (l) { leave = true; break; }
                }
            }
            if (!leave) {
                if (lane < SRV_CHUNKS) { S.chunk[lane].w0 = v.x; S.chunk[lane].w1 = v.y; S.chunk[lane].w2 = v.z; S.chunk[lane].tag = v.w; }
                const uint32_t hdr = __shfl_sync(0xFFFFFFFFu, v.x, 0);
                if (((hdr >> 8) & SRV_F_QUIT) != 0) leave = true;
            }
            if (lane == 0) S.ctrl = leave ? 1u : 0u;
            t_acq = globaltimer_ns();
            if (lane == 0 && !leave && ((__shfl_sync(0x1u, v.x, 0) >> 8) & SRV_F_TRACE) && P.trace) P.trace[28] = clock64();
        }
        __syncthreads();
        if (S.ctrl != 0) {
            if (tid == 0) {
                __threadfence_system();
                asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(P.exit_word), "l"(P.launch_id) : "memory");
            }
            return;
        }
        // lines of the store (read through the non-coherent path) must not survive from one request to the next: the
        // launch-per-query kernel starts with a cold L1 too, and timing runs flush the L2 between queries.  A gpu-scope
        // fence invalidates the SM's L1 (CCTL.IVALL); one thread does it, everybody passes the barrier below afterwards.
        if (tid == 0) __threadfence();
        // ---- decode ----------------------------------------------------------------------------------------------
        const uint32_t hdr = S.chunk[0].w0;
        const int nsteps = (int)(hdr & 0xFF);
        const uint32_t flags = (hdr >> 8) & 0xFF;
        if (SHARDED && (flags & SRV_F_WAIT)) {
            // a peer owns this query: wait for its verdict (bounded), report it in the record's status word
            // (2 = the owner never answered, 4 = the table outgrew the owner's shared memory: redo collectively)
            if (tid == 0) {
                const uint64_t epoch = (uint64_t)S.chunk[0].w1 | ((uint64_t)S.chunk[0].w2 << 32);
                const uint32_t owner = (hdr >> 16) & 0xFF;
                uint64_t status = 0;
                if (!wait_flag(Q->my_flag + owner, 2 * epoch)) status = 2;
                else if (ld_sys_u64(Q->my_flag + owner) == 2 * epoch + 1) status = 4;
                uint64_t *rec = (uint64_t *)P.rec;
                st_sys_v2u64(P.times, t_acq, globaltimer_ns());
                st_sys_v2u64(rec + 2, status, record_check(seq, 0, status, 0));
                st_sys_v2u64(rec, seq, 0);
            }
            seq++;
            t_idle0 = globaltimer_ns();
            __syncthreads();
            continue;
        }
        if (tid < nsteps) {
            const SrvChunk c = S.chunk[SRV_HDR_CHUNKS + tid];
            LightStep &ls = S.plan.steps[tid];
            const uint32_t slot = c.w2 >> 17;
            const SegLite sg = slot < SRV_SMEM_SEGS ? S.seg[slot] : P.segtab[slot];
            ls.seg.bucket_start = sg.bucket_start;
            ls.seg.fm = sg.fm;
            ls.seg.pid = c.w2 & 0x1FFFFu;
            ls.seg.dir = (c.w0 >> 24) & 1u;
            ls.seg.index_mode = (c.w0 >> 25) & 1u;
            ls.seg._pad = 0;
            ls.kind = (int32_t)(c.w0 & 0xFF);
            ls.col_start = (int32_t)((c.w0 >> 8) & 0xFF);
            ls.col_end = (int32_t)((c.w0 >> 16) & 0xFF);
            ls.C = (int32_t)((c.w0 >> 26) & 31u);
            ls.end_const = c.w1;
            ls._pad0 = 0;
            ls.mt_tid = 0;
            ls.mt_factor = 1;
            ls.key = (ls.kind == LKIND_C2U) ? make_key(c.w1, ls.seg.pid, ls.seg.dir) : 0;
        }
        if (SHARDED) {   // where every shard keeps the segment of every step
            for (int i = tid - 128; i >= 0 && i < nsteps * LIGHT_PEERS; i += NT) {
                const int st = i / LIGHT_PEERS, r = i - st * LIGHT_PEERS;
                const uint32_t slot = S.chunk[SRV_HDR_CHUNKS + st].w2 >> 17;
                if (st < LIGHT_SHARDED_STEPS && r < (int)Q->nranks) segr[st][r] = Q->segr_tab[slot * LIGHT_PEERS + r];
            }
        }
        if (tid >= 32 && tid < 32 + 3 * 12) {   // projection columns: 12 per chunk
            const int j = tid - 32;
            const SrvChunk &c = S.chunk[1 + j / 12];
            const uint32_t w = (j % 12) < 4 ? c.w0 : ((j % 12) < 8 ? c.w1 : c.w2);
            if (j < MAX_COLS) S.plan.proj_cols[j] = (int8_t)((w >> (8 * (j & 3))) & 0xFF);
        }
        if (tid == 64) {
            S.plan.nsteps = nsteps;
            S.plan.seq = seq;
            S.plan.do_project = (flags & SRV_F_PROJECT) ? 1 : 0;
            S.plan.collect_stats = (flags & SRV_F_STATS) ? 1 : 0;
            S.plan.proj_n = (int)((hdr >> 16) & 0xFF);
            S.plan.trace = (flags & SRV_F_TRACE) ? P.trace : nullptr;
            if (S.plan.trace) S.plan.trace[29] = clock64();
        }
        __syncthreads();
        if (SHARDED) {
            PeerView pvw;
            pvw.v = Q->pv;
            pvw.e = Q->pe;
            pvw.segr = segr;
            pvw.n = Q->nranks;
            const bool spilled = light_run<NT, PeerView, WARPM>(S.plan, S.plan.steps, pvw, S.light, (flags & SRV_F_STATS) != 0, P.times, t_acq);
            if (tid < (int)Q->nranks && Q->peer_flag[tid] != nullptr) {
                const uint64_t epoch = (uint64_t)S.chunk[0].w1 | ((uint64_t)S.chunk[0].w2 << 32);
                st_sys_u64_light(Q->peer_flag[tid], 2 * epoch + (spilled ? 1 : 0));
            }
        } else {
            light_run<NT, LocalView, WARPM>(S.plan, S.plan.steps, sv, S.light, (flags & SRV_F_STATS) != 0, P.times, t_acq);
        }
        seq++;
        t_idle0 = globaltimer_ns();
        __syncthreads();
    }
}

template <int NT, bool WARPM>
__global__ void __launch_bounds__(NT) light_server_kernel(const __grid_constant__ SrvParams P) {
    extern __shared__ __align__(16) unsigned char srv_dyn[];
    SrvSmem &S = *reinterpret_cast<SrvSmem *>(srv_dyn);
    light_server_body<NT, WARPM, false>(P, nullptr, S, nullptr);
}

// the server of one shard of a vid % n group (peer memory mapped: wk_comm_p2p_import_store / wk_comm_local_group)
__global__ void __launch_bounds__(LIGHT_SRV_THREADS) light_server_sharded_kernel(const __grid_constant__ SrvParams P, const __grid_constant__ SrvPeers Q) {
    extern __shared__ __align__(16) unsigned char srv_dyn[];
    SrvSmemSharded &S = *reinterpret_cast<SrvSmemSharded *>(srv_dyn);
    light_server_body<LIGHT_SRV_THREADS, true, true>(P, &Q, S.base, S.segr);
}

}  // namespace wk

// Device-side builder of the cluster-hash graph store.
//
// Produces, directly in HBM, the data structure the reference's loader + StaticGStore build on the
// CPU (core/loader/base_loader.hpp:308-378 sort/dedup/partition, core/store/static_gstore.hpp:64-265
// insert_triples/insert_idx, core/store/gstore.hpp:428-472 segment sizing, :789-856 insert_key):
// 16-byte key/pointer slots in 8-way buckets whose last slot chains to an indirect-header bucket,
// one flat 4-byte edge array with a sorted run per key, and one metadata record per segment.
// Only what a probe can observe is contractual; slot placement inside a chain is free, so the
// insertion is a counting pass (atomic rank per bucket), a scan that hands every overflowing
// bucket its run of ext buckets, and a placement pass -- no locks, no retries.
//
// Segment sizing and the edge-array layout restate wukong_b200/csrc/store/host_builder.cpp (the
// host builder that is checked bit-for-bit against the oracle), so both builders emit the same
// segment table and the same edge array.
#include <cub/cub.cuh>
#include <cuda_runtime.h>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <map>
#include <tuple>
#include <vector>

#include "wk_layout.cuh"
#include "wukong_b200.h"

using namespace wk;

namespace {

constexpr int ASSOC = WK_ASSOCIATIVITY;
constexpr int THREADS = 256;

#define B_TRY(expr)                                                                                  \
    do {                                                                                             \
        cudaError_t e_ = (expr);                                                                     \
        if (e_ != cudaSuccess) {                                                                     \
            fprintf(stderr, "wukong_b200 store build: %s failed: %s\n", #expr, cudaGetErrorString(e_)); \
            return WK_ERR_CUDA;                                                                      \
        }                                                                                            \
    } while (0)

__host__ __device__ inline bool is_tpid(uint64_t id) { return id > 1 && id < (1u << WK_NBITS_IDX); }
__host__ __device__ inline uint64_t make_ptr(uint64_t size, uint64_t off) { return size | (off << WK_PTR_SIZE_BITS); }

// scratch allocations of one build, released together
struct Arena {
    std::vector<void *> ptrs;
    ~Arena() { for (void *p : ptrs) cudaFree(p); }
    template <typename T>
    cudaError_t get(T **p, uint64_t n) {
        cudaError_t e = cudaMalloc((void **)p, std::max<uint64_t>(n, 1) * sizeof(T));
        if (e == cudaSuccess) ptrs.push_back(*p);
        return e;
    }
    void drop(void *p) {
        for (size_t i = 0; i < ptrs.size(); i++)
            if (ptrs[i] == p) { cudaFree(p); ptrs.erase(ptrs.begin() + i); return; }
    }
};

struct BuildCtl {
    unsigned long long count;      // triples owned by this server on the side being packed
    unsigned long long dups;       // adjacent duplicates after the sort
    unsigned long long first_non_type;   // IN side: first sorted position whose object is not a type id
    unsigned int bad_pred;         // a predicate id outside [1, num_normal_preds]
    unsigned int big_run;          // a key with >= 2^28 edges (does not fit iptr_t::size)
};

// ---- 1. partition by owner and pack (base_loader.hpp:343-361): OUT edges live with the subject's
//         owner, IN edges with the object's. Order is irrelevant here (the sort follows).
__global__ void __launch_bounds__(THREADS) pack_side_kernel(const uint32_t *__restrict__ tr, uint64_t n, uint32_t S, uint32_t sid,
                                                            uint32_t npreds, int out_side, uint32_t *__restrict__ minor,
                                                            uint64_t *__restrict__ major, BuildCtl *ctl) {
    __shared__ unsigned long long s_base;
    __shared__ uint32_t s_warp[THREADS / 32];
    for (uint64_t i0 = (uint64_t)blockIdx.x * THREADS; i0 < n; i0 += (uint64_t)gridDim.x * THREADS) {
        const uint64_t i = i0 + threadIdx.x;
        uint32_t s = 0, p = 0, o = 0;
        bool own = false;
        if (i < n) {
            s = tr[3 * i]; p = tr[3 * i + 1]; o = tr[3 * i + 2];
            if (p == 0 || p > npreds) atomicOr(&ctl->bad_pred, 1u);
            own = ((out_side ? s : o) % S) == sid;
        }
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, own);
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        if (lane == 0) s_warp[warp] = __popc(m);
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t tot = 0;
            for (int w = 0; w < THREADS / 32; w++) { const uint32_t c = s_warp[w]; s_warp[w] = tot; tot += c; }
            s_base = tot ? atomicAdd(&ctl->count, (unsigned long long)tot) : 0ull;
        }
        __syncthreads();
        if (own) {
            const uint64_t dst = s_base + s_warp[warp] + __popc(m & ((1u << lane) - 1));
            minor[dst] = out_side ? o : s;
            major[dst] = ((uint64_t)p << 32) | (out_side ? s : o);
        }
        __syncthreads();
    }
}

// ---- 2. duplicates (base_loader.hpp:373 dedup) and the leading type-object run of the POS order
//         (static_gstore.hpp:127-130) ----------------------------------------------------------------
__global__ void __launch_bounds__(THREADS) flag_kernel(const uint64_t *__restrict__ K, const uint32_t *__restrict__ E, uint64_t m,
                                                       uint8_t *__restrict__ keep, int in_side, BuildCtl *ctl) {
    unsigned long long dups = 0, first = ~0ull;
    for (uint64_t i = (uint64_t)blockIdx.x * THREADS + threadIdx.x; i < m; i += (uint64_t)gridDim.x * THREADS) {
        const bool k = (i == 0) || K[i] != K[i - 1] || E[i] != E[i - 1];
        keep[i] = k;
        dups += !k;
        if (in_side && !is_tpid((uint32_t)K[i]) && i < first) first = i;
    }
    for (int o = 16; o; o >>= 1) {
        dups += __shfl_down_sync(0xFFFFFFFFu, dups, o);
        const unsigned long long f = __shfl_down_sync(0xFFFFFFFFu, first, o);
        first = f < first ? f : first;
    }
    if ((threadIdx.x & 31) == 0) {
        if (dups) atomicAdd(&ctl->dups, dups);
        if (in_side && first != ~0ull) atomicMin(&ctl->first_non_type, first);
    }
}

__global__ void __launch_bounds__(THREADS) head_kernel(const uint64_t *__restrict__ K, uint64_t m, uint8_t *__restrict__ head) {
    for (uint64_t i = (uint64_t)blockIdx.x * THREADS + threadIdx.x; i < m; i += (uint64_t)gridDim.x * THREADS)
        head[i] = (i == 0) || K[i] != K[i - 1];
}
__global__ void __launch_bounds__(THREADS) runs_kernel(const uint64_t *__restrict__ K, const uint64_t *__restrict__ POS, uint64_t nk,
                                                       uint64_t m, uint64_t *__restrict__ UK, uint32_t *__restrict__ CNT) {
    for (uint64_t i = (uint64_t)blockIdx.x * THREADS + threadIdx.x; i < nk; i += (uint64_t)gridDim.x * THREADS) {
        const uint64_t b = POS[i], e = (i + 1 < nk) ? POS[i + 1] : m;
        UK[i] = K[b];
        const uint64_t c = e - b;
        CNT[i] = c > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)c;   // >= 2^28 is rejected later (iptr_t::size)
    }
}
// (type, instance) pairs of the rdf:type triples, in (instance, type) order
__global__ void __launch_bounds__(THREADS) expand_types_kernel(const uint64_t *__restrict__ UK, const uint32_t *__restrict__ CNT,
                                                               const uint64_t *__restrict__ OFF, const uint32_t *__restrict__ E,
                                                               uint64_t kb, uint64_t ke, uint64_t tb, uint32_t *__restrict__ tk,
                                                               uint32_t *__restrict__ tv) {
    for (uint64_t i = kb + (uint64_t)blockIdx.x * THREADS + threadIdx.x; i < ke; i += (uint64_t)gridDim.x * THREADS) {
        const uint32_t s = (uint32_t)UK[i], c = CNT[i];
        const uint64_t off = OFF[i];
        for (uint32_t j = 0; j < c; j++) {
            tk[off + j - tb] = E[off + j];
            tv[off + j - tb] = s;
        }
    }
}

// first position of every predicate in a (p << 32 | x)-sorted array: out[p] = lower_bound(p << 32), p in [0, npreds + 1]
__global__ void bounds_kernel(const uint64_t *__restrict__ A, uint64_t m, uint32_t nq, int shift, uint64_t *__restrict__ out) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nq) return;
    const uint64_t target = (uint64_t)p << shift;
    uint64_t lo = 0, hi = m;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (A[mid] < target) lo = mid + 1; else hi = mid;
    }
    out[p] = lo;
}
__global__ void bounds32_kernel(const uint32_t *__restrict__ A, uint64_t m, uint32_t nq, uint64_t *__restrict__ out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nq) return;
    uint64_t lo = 0, hi = m;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (A[mid] < t) lo = mid + 1; else hi = mid;
    }
    out[t] = lo;
}

__global__ void __launch_bounds__(THREADS) low32_kernel(const uint64_t *__restrict__ K, uint64_t m, uint32_t *__restrict__ dst) {
    for (uint64_t i = (uint64_t)blockIdx.x * THREADS + threadIdx.x; i < m; i += (uint64_t)gridDim.x * THREADS) dst[i] = (uint32_t)K[i];
}

// ---- 3. insertion ----------------------------------------------------------------------------------
struct SegSlot {   // per (index, pid, dir): what the insertion needs
    uint64_t bucket_start, num_buckets;
    uint64_t edge_base;      // edge_start - first sorted triple of the predicate (normal segments)
    int64_t ext_base;        // ext_start - (scanned ext need at bucket_start)
};
// table index: index * 2 * (npreds + 1) + pid * 2 + dir
__device__ __forceinline__ uint32_t seg_index_of(uint64_t key, uint32_t npreds) {
    const uint32_t dir = (uint32_t)(key & 1), pid = (uint32_t)(key >> 1) & ((1u << WK_NBITS_IDX) - 1);
    const uint32_t index = (key >> WK_KEY_VID_SHIFT) == 0 ? 1u : 0u;
    return index * 2 * (npreds + 1) + pid * 2 + dir;
}

// one entry per (vertex, predicate, direction) key: the key and its pointer into the edge array
__global__ void __launch_bounds__(THREADS) entries_kernel(const uint64_t *__restrict__ UK, const uint32_t *__restrict__ CNT,
                                                          const uint64_t *__restrict__ OFF, uint64_t nk, uint32_t dir, uint32_t npreds,
                                                          const SegSlot *__restrict__ tab, uint64_t *__restrict__ ekey,
                                                          uint64_t *__restrict__ eptr, BuildCtl *ctl) {
    for (uint64_t i = (uint64_t)blockIdx.x * THREADS + threadIdx.x; i < nk; i += (uint64_t)gridDim.x * THREADS) {
        const uint64_t k = UK[i];
        const uint32_t p = (uint32_t)(k >> 32), v = (uint32_t)k, c = CNT[i];
        if (c >= (1u << WK_PTR_SIZE_BITS)) atomicOr(&ctl->big_run, 1u);
        ekey[i] = make_key(v, p, dir);
        eptr[i] = make_ptr(c, tab[(uint64_t)p * 2 + dir].edge_base + OFF[i]);
    }
}

__global__ void __launch_bounds__(THREADS) rank_kernel(const uint64_t *__restrict__ ekey, uint64_t nk, uint32_t npreds,
                                                       const SegSlot *__restrict__ tab, uint32_t *__restrict__ cnt,
                                                       uint32_t *__restrict__ rank) {
    for (uint64_t i = (uint64_t)blockIdx.x * THREADS + threadIdx.x; i < nk; i += (uint64_t)gridDim.x * THREADS) {
        const uint64_t key = ekey[i];
        const SegSlot sg = tab[seg_index_of(key, npreds)];
        const uint64_t bucket = sg.bucket_start + hash_u64(key) % sg.num_buckets;
        rank[i] = atomicAdd(&cnt[bucket], 1u);
    }
}

// ext buckets a main bucket needs: 7 keys fit in the bucket, 7 more in every chained one (gstore.hpp:789-856)
__global__ void __launch_bounds__(THREADS) need_kernel(const uint32_t *__restrict__ cnt, uint64_t nb, uint32_t *__restrict__ need) {
    for (uint64_t b = (uint64_t)blockIdx.x * THREADS + threadIdx.x; b < nb; b += (uint64_t)gridDim.x * THREADS) {
        const uint32_t c = cnt[b];
        need[b] = c > (uint32_t)(ASSOC - 1) ? (c - (ASSOC - 1) + (ASSOC - 2)) / (ASSOC - 1) : 0u;
    }
}

__global__ void gather_kernel(const uint64_t *__restrict__ scan, const uint64_t *__restrict__ idx, uint32_t n, uint64_t *__restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = scan[idx[i]];
}

__global__ void __launch_bounds__(THREADS) place_kernel(const uint64_t *__restrict__ ekey, const uint64_t *__restrict__ eptr,
                                                        const uint32_t *__restrict__ rank, uint64_t nk, uint32_t npreds,
                                                        const SegSlot *__restrict__ tab, const uint64_t *__restrict__ extscan,
                                                        ulonglong2 *__restrict__ V) {
    for (uint64_t i = (uint64_t)blockIdx.x * THREADS + threadIdx.x; i < nk; i += (uint64_t)gridDim.x * THREADS) {
        const uint64_t key = ekey[i];
        const SegSlot sg = tab[seg_index_of(key, npreds)];
        const uint64_t bucket = sg.bucket_start + hash_u64(key) % sg.num_buckets;
        const uint32_t r = rank[i];
        uint64_t slot;
        if (r < (uint32_t)(ASSOC - 1)) {
            slot = bucket * ASSOC + r;
        } else {
            const uint32_t j = r - (ASSOC - 1);
            const uint64_t ext = (uint64_t)(sg.ext_base + (int64_t)extscan[bucket]) + j / (ASSOC - 1);
            slot = ext * ASSOC + j % (ASSOC - 1);
        }
        V[slot] = make_ulonglong2(key, eptr[i]);
    }
}

// chain pointers: last slot of a full bucket names the next bucket in its key.vid field
__global__ void __launch_bounds__(THREADS) link_kernel(const uint32_t *__restrict__ need, const uint64_t *__restrict__ extscan,
                                                       uint64_t nb, const uint64_t *__restrict__ seg_start,
                                                       const int64_t *__restrict__ seg_ext_base, int nseg, ulonglong2 *__restrict__ V) {
    for (uint64_t b = (uint64_t)blockIdx.x * THREADS + threadIdx.x; b < nb; b += (uint64_t)gridDim.x * THREADS) {
        const uint32_t nd = need[b];
        if (nd == 0) continue;
        int lo = 0, hi = nseg - 1;   // last segment whose bucket_start <= b
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (seg_start[mid] <= b) lo = mid; else hi = mid - 1;
        }
        uint64_t ext = (uint64_t)(seg_ext_base[lo] + (int64_t)extscan[b]);
        uint64_t cur = b;
        for (uint32_t e = 0; e < nd; e++, ext++) {
            V[cur * ASSOC + (ASSOC - 1)] = make_ulonglong2(make_key(ext, 0, 0), 0ull);
            cur = ext;
        }
    }
}

struct CastU64 {
    __host__ __device__ uint64_t operator()(uint32_t x) const { return (uint64_t)x; }
};

uint64_t prime_at_most(uint64_t upper) {   // math::hash_prime_u64 (utils/math.hpp:105-131)
    static const uint64_t primes[] = {98317ull, 196613ull, 393241ull, 786433ull, 1572869ull, 3145739ull, 6291469ull,
                                      12582917ull, 25165843ull, 50331653ull, 100663319ull, 201326611ull,
                                      402653189ull, 805306457ull, 1610612741ull};
    if (upper >= (1ull << 31)) return upper;
    uint64_t best = upper;
    for (uint64_t p : primes)
        if (p <= upper) best = p;
    return best;
}

// sorted, de-duplicated triples of one direction, grouped by key
struct Side {
    uint64_t m = 0, nk = 0;        // triples, keys
    uint32_t *E = nullptr;         // m: the far end of every triple, ordered by (p, near, far)
    uint64_t *UK = nullptr;        // nk: (p << 32 | near) of every key
    uint32_t *CNT = nullptr;       // nk: edges of the key
    uint64_t *OFF = nullptr;       // nk: position of the key's run in E
    std::vector<uint64_t> tri_lb;  // npreds + 2: first triple of predicate p
    std::vector<uint64_t> key_lb;  // npreds + 2: first key of predicate p
};

int grid_for(uint64_t n, int sms) {
    const uint64_t want = (n + THREADS - 1) / THREADS;
    return (int)std::max<uint64_t>(1, std::min<uint64_t>(want, (uint64_t)sms * 16));
}

int build_side(Arena &ar, cudaStream_t st, int sms, const uint32_t *d_tr, uint64_t n, const wk_build_opts_t &o, int out_side,
               BuildCtl *d_ctl, Side &sd) {
    const uint32_t npreds = (uint32_t)o.num_normal_preds;
    uint32_t *k1 = nullptr, *k1b = nullptr;
    uint64_t *v1 = nullptr, *v1b = nullptr;
    uint8_t *keep = nullptr;
    B_TRY(ar.get(&k1, n)); B_TRY(ar.get(&k1b, n)); B_TRY(ar.get(&v1, n)); B_TRY(ar.get(&v1b, n));
    B_TRY(cudaMemsetAsync(d_ctl, 0, sizeof(BuildCtl), st));
    {
        const unsigned long long inf = ~0ull;
        B_TRY(cudaMemcpyAsync(&d_ctl->first_non_type, &inf, sizeof(inf), cudaMemcpyHostToDevice, st));
    }
    pack_side_kernel<<<grid_for(n, sms), THREADS, 0, st>>>(d_tr, n, (uint32_t)std::max(1, o.num_servers), (uint32_t)o.sid, npreds,
                                                           out_side, k1, v1, d_ctl);
    BuildCtl h;
    B_TRY(cudaMemcpyAsync(&h, d_ctl, sizeof(h), cudaMemcpyDeviceToHost, st));
    B_TRY(cudaStreamSynchronize(st));
    if (h.bad_pred) return WK_ERR_BAD_ARG;
    uint64_t m = h.count;

    // order by (p, near, far): a stable sort on the major key after a sort on the minor one (base_loader.hpp:363-372)
    cub::DoubleBuffer<uint32_t> dk(k1, k1b);
    cub::DoubleBuffer<uint64_t> dv(v1, v1b);
    size_t tb1 = 0, tb2 = 0;
    B_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tb1, dk, dv, m, 0, 32, st));
    B_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tb2, dv, dk, m, 0, 32 + WK_NBITS_IDX, st));
    void *tmp = nullptr;
    B_TRY(ar.get((uint8_t **)&tmp, std::max(tb1, tb2)));
    B_TRY(cub::DeviceRadixSort::SortPairs(tmp, tb1, dk, dv, m, 0, 32, st));
    B_TRY(cub::DeviceRadixSort::SortPairs(tmp, tb2, dv, dk, m, 0, 32 + WK_NBITS_IDX, st));
    uint64_t *K = dv.Current(), *Kalt = dv.Alternate();
    uint32_t *E = dk.Current(), *Ealt = dk.Alternate();

    B_TRY(ar.get(&keep, m));
    flag_kernel<<<grid_for(m, sms), THREADS, 0, st>>>(K, E, m, keep, !out_side, d_ctl);
    B_TRY(cudaMemcpyAsync(&h, d_ctl, sizeof(h), cudaMemcpyDeviceToHost, st));
    B_TRY(cudaStreamSynchronize(st));
    uint64_t skip = 0;
    if (!out_side) skip = (h.first_non_type == ~0ull) ? m : h.first_non_type;
    if (h.dups) {
        // compact both columns; the leading type-object run is counted on the compacted order below
        unsigned long long *d_num = &d_ctl->count;
        size_t tb = 0;
        B_TRY(cub::DeviceSelect::Flagged(nullptr, tb, K, keep, Kalt, d_num, m, st));
        void *t2 = nullptr;
        B_TRY(ar.get((uint8_t **)&t2, tb));
        B_TRY(cub::DeviceSelect::Flagged(t2, tb, K, keep, Kalt, d_num, m, st));
        B_TRY(cub::DeviceSelect::Flagged(t2, tb, E, keep, Ealt, d_num, m, st));
        B_TRY(cudaMemcpyAsync(&h, d_ctl, sizeof(h), cudaMemcpyDeviceToHost, st));
        B_TRY(cudaStreamSynchronize(st));
        m = h.count;
        std::swap(K, Kalt);
        std::swap(E, Ealt);
        ar.drop(t2);
        if (!out_side) {   // recount the leading run on the compacted arrays
            const unsigned long long inf = ~0ull;
            B_TRY(cudaMemsetAsync(d_ctl, 0, sizeof(BuildCtl), st));
            B_TRY(cudaMemcpyAsync(&d_ctl->first_non_type, &inf, sizeof(inf), cudaMemcpyHostToDevice, st));
            flag_kernel<<<grid_for(m, sms), THREADS, 0, st>>>(K, E, m, keep, 1, d_ctl);
            B_TRY(cudaMemcpyAsync(&h, d_ctl, sizeof(h), cudaMemcpyDeviceToHost, st));
            B_TRY(cudaStreamSynchronize(st));
            skip = (h.first_non_type == ~0ull) ? m : h.first_non_type;
        }
    }
    ar.drop(keep);
    K += skip; E += skip; m -= skip;

    // keys = runs of equal (p, near): positions of the run heads, then (key, count) per run
    {
        uint8_t *head = nullptr;
        uint64_t *POS = nullptr;
        B_TRY(ar.get(&head, m)); B_TRY(ar.get(&POS, m));
        head_kernel<<<grid_for(m, sms), THREADS, 0, st>>>(K, m, head);
        unsigned long long *d_runs = &d_ctl->count;
        B_TRY(cudaMemsetAsync(d_runs, 0, sizeof(*d_runs), st));
        size_t tb = 0;
        thrust::counting_iterator<uint64_t> iota(0);
        B_TRY(cub::DeviceSelect::Flagged(nullptr, tb, iota, head, POS, d_runs, m, st));
        void *t2 = nullptr;
        B_TRY(ar.get((uint8_t **)&t2, tb));
        if (m) B_TRY(cub::DeviceSelect::Flagged(t2, tb, iota, head, POS, d_runs, m, st));
        B_TRY(cudaMemcpyAsync(&h, d_ctl, sizeof(h), cudaMemcpyDeviceToHost, st));
        B_TRY(cudaStreamSynchronize(st));
        sd.nk = h.count;
        ar.drop(t2); ar.drop(head);
        B_TRY(ar.get(&sd.UK, sd.nk)); B_TRY(ar.get(&sd.CNT, sd.nk));
        runs_kernel<<<grid_for(sd.nk, sms), THREADS, 0, st>>>(K, POS, sd.nk, m, sd.UK, sd.CNT);
        // the run's first position is its offset in E; keep an exact-size copy, release the m-sized scratch
        B_TRY(ar.get(&sd.OFF, sd.nk));
        B_TRY(cudaMemcpyAsync(sd.OFF, POS, sd.nk * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
        B_TRY(cudaStreamSynchronize(st));
        ar.drop(POS);
    }
    // per-predicate extents of triples and keys
    uint64_t *d_b = nullptr;
    B_TRY(ar.get(&d_b, 2 * (uint64_t)(npreds + 2)));
    bounds_kernel<<<(npreds + 2 + 127) / 128, 128, 0, st>>>(K, m, npreds + 2, 32, d_b);
    bounds_kernel<<<(npreds + 2 + 127) / 128, 128, 0, st>>>(sd.UK, sd.nk, npreds + 2, 32, d_b + npreds + 2);
    sd.tri_lb.resize(npreds + 2);
    sd.key_lb.resize(npreds + 2);
    B_TRY(cudaMemcpyAsync(sd.tri_lb.data(), d_b, (npreds + 2) * 8, cudaMemcpyDeviceToHost, st));
    B_TRY(cudaMemcpyAsync(sd.key_lb.data(), d_b + npreds + 2, (npreds + 2) * 8, cudaMemcpyDeviceToHost, st));
    // keep only E (compact copy); release the sort buffers
    B_TRY(ar.get(&sd.E, m));
    B_TRY(cudaMemcpyAsync(sd.E, E, m * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    B_TRY(cudaStreamSynchronize(st));
    sd.m = m;
    ar.drop(d_b); ar.drop(tmp); ar.drop(k1); ar.drop(k1b); ar.drop(v1); ar.drop(v1b);
    return WK_SUCCESS;
}

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

extern "C" int wk_store_build(int device, const wk_sid_t *triples, uint64_t n, const wk_build_opts_t *opts, wk_store_t **out,
                              wk_build_stats_t *stats) {
    if (!opts || !out || (!triples && n)) return WK_ERR_BAD_ARG;
    const wk_build_opts_t o = *opts;
    const int npreds = o.num_normal_preds;
    if (npreds <= 0 || npreds >= (1 << WK_NBITS_IDX) || o.num_servers < 1 || o.sid < 0 || o.sid >= o.num_servers) return WK_ERR_BAD_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return WK_ERR_NO_DEVICE;
    if (device < 0 || device >= ndev) return WK_ERR_BAD_ARG;
    B_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    B_TRY(cudaGetDeviceProperties(&prop, device));
    const int sms = prop.multiProcessorCount;
    const double t_begin = now_ms();

    Arena ar;
    cudaStream_t st = nullptr;   // default stream: the build is a bring-up step, not a query path
    const uint32_t *d_tr = triples;
    if (!o.triples_on_device) {
        uint32_t *buf = nullptr;
        B_TRY(ar.get(&buf, 3 * n));
        B_TRY(cudaMemcpy(buf, triples, 3 * n * sizeof(uint32_t), cudaMemcpyHostToDevice));
        d_tr = buf;
    }
    const double t_up = now_ms();
    BuildCtl *d_ctl = nullptr;
    B_TRY(ar.get(&d_ctl, 1));

    // ---- sorted sides ---------------------------------------------------------------------------
    Side so, si;
    int rc = build_side(ar, st, sms, d_tr, n, o, 1, d_ctl, so);
    if (rc) return rc;
    rc = build_side(ar, st, sms, d_tr, n, o, 0, d_ctl, si);
    if (rc) return rc;
    if (!o.triples_on_device) ar.drop((void *)d_tr);

    // ---- type index: instances per type id, ordered by (type, instance) (static_gstore.hpp:217-265) ----
    const uint64_t tb_ = so.tri_lb[WK_TYPE_ID], te_ = so.tri_lb[WK_TYPE_ID + 1], ntype = te_ - tb_;
    std::vector<uint64_t> type_lb(npreds + 2, 0);
    uint32_t *T_inst = nullptr;   // instances grouped by type
    if (ntype) {
        uint32_t *tk = nullptr, *tkb = nullptr, *tv = nullptr, *tvb = nullptr;
        B_TRY(ar.get(&tk, ntype)); B_TRY(ar.get(&tkb, ntype)); B_TRY(ar.get(&tv, ntype)); B_TRY(ar.get(&tvb, ntype));
        const uint64_t kb = so.key_lb[WK_TYPE_ID], ke = so.key_lb[WK_TYPE_ID + 1];
        expand_types_kernel<<<grid_for(ke - kb, sms), THREADS, 0, st>>>(so.UK, so.CNT, so.OFF, so.E, kb, ke, tb_, tk, tv);
        cub::DoubleBuffer<uint32_t> dk(tk, tkb), dv(tv, tvb);
        size_t tb = 0;
        B_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tb, dk, dv, ntype, 0, 32, st));
        void *t2 = nullptr;
        B_TRY(ar.get((uint8_t **)&t2, tb));
        B_TRY(cub::DeviceRadixSort::SortPairs(t2, tb, dk, dv, ntype, 0, 32, st));
        uint64_t *d_b = nullptr;
        B_TRY(ar.get(&d_b, (uint64_t)npreds + 2));
        bounds32_kernel<<<(npreds + 2 + 127) / 128, 128, 0, st>>>(dk.Current(), ntype, npreds + 2, d_b);
        B_TRY(cudaMemcpyAsync(type_lb.data(), d_b, (npreds + 2) * 8, cudaMemcpyDeviceToHost, st));
        B_TRY(cudaStreamSynchronize(st));
        T_inst = dv.Current();
        ar.drop(t2); ar.drop(d_b); ar.drop(dk.Current()); ar.drop(dk.Alternate()); ar.drop(dv.Alternate());
    }
    std::vector<uint64_t> type_cnt(npreds + 1, 0);
    for (int t = 2; t <= npreds; t++) type_cnt[t] = type_lb[t + 1] - type_lb[t];   // is_tpid(t) && t <= npreds
    const double t_sort = now_ms();

    // ---- segment sizing (GStore ctor gstore.hpp:979-1025, init_seg_metas :530-786, alloc :428-472) ----
    auto keys_of = [&](const Side &s, int p) { return s.key_lb[p + 1] - s.key_lb[p]; };
    auto edges_of = [&](const Side &s, int p) { return s.tri_lb[p + 1] - s.tri_lb[p]; };
    std::vector<uint32_t> local_preds;
    uint64_t total_keys = 0, num_typeid = 0;
    for (int p = 1; p <= npreds; p++) {
        if (edges_of(so, p) + edges_of(si, p) > 0) {
            local_preds.push_back((uint32_t)p);
            total_keys += keys_of(so, p) + keys_of(si, p);
        } else if (type_cnt[p] > 0) {
            num_typeid++;
        }
    }
    total_keys += local_preds.size() * 2 + num_typeid;
    const uint64_t nsegs = (uint64_t)npreds * 2 + 2;
    uint64_t total_edges = so.m + si.m;
    for (int p = 1; p <= npreds; p++) total_edges += keys_of(so, p) + keys_of(si, p) + type_cnt[p];
    uint64_t num_slots, num_buckets, num_buckets_ext, num_entries;
    if (o.kvstore_bytes) {
        const uint64_t header = o.kvstore_bytes * (128 * 100 / (128 + 3 * 32)) / 100;
        num_slots = header / sizeof(wk_vertex_t);
        num_buckets = prime_at_most((num_slots / ASSOC) * 80 / 100);
        num_buckets_ext = num_slots / ASSOC - num_buckets;
        num_entries = (o.kvstore_bytes - header) / sizeof(wk_sid_t);
    } else {
        const uint64_t lf = (uint64_t)std::max(1, std::min(100, o.est_load_factor ? o.est_load_factor : 55));
        num_buckets = total_keys * 100 / (ASSOC * lf) + nsegs + 8;
        num_buckets_ext = num_buckets * 15 / 100 + nsegs + 8;
        num_slots = (num_buckets + num_buckets_ext) * ASSOC;
        num_entries = total_edges + 1;
    }
    if (num_buckets <= nsegs || total_edges >= num_entries) return WK_ERR_STORE_FULL;
    if (num_slots / ASSOC >= (1ull << 32)) return WK_ERR_STORE_FULL;

    std::map<std::tuple<uint32_t, int, int>, wk_segmeta_t> segs;   // ordered like segid_t::operator< (pid, index, dir)
    uint64_t last_entry = 0, main_off = 0, last_ext = 0;
    bool ext_fail = false;
    const uint64_t num_free = num_buckets - nsegs;
    auto alloc_edges = [&](uint64_t k) { uint64_t off = k ? last_entry : 0; last_entry += k; return off; };
    auto alloc_buckets = [&](wk_segmeta_t &m) {
        uint64_t nb = 0;
        if (m.num_keys != 0) nb = (uint64_t)((static_cast<double>(m.num_keys) / total_keys) * num_free);
        m.num_buckets = std::max<uint64_t>(nb, 1);
        m.bucket_start = main_off;
        main_off += m.num_buckets;
        const uint64_t el = m.num_buckets * 15 / 100 + 1;   // one extent per segment (USE_GPU, meta.hpp:38-40)
        if (last_ext + el >= num_buckets_ext) { ext_fail = true; return; }
        m.ext_start = num_buckets + last_ext;
        m.ext_num = el;
        last_ext += el;
    };
    auto new_seg = [&](int index, uint32_t pid, int dir) -> wk_segmeta_t & {
        wk_segmeta_t &m = segs[std::make_tuple(pid, index, dir)];
        memset(&m, 0, sizeof(m));
        m.index = index; m.pid = pid; m.dir = dir;
        return m;
    };
    for (int d = 0; d <= 1; d++) new_seg(0, 0, d);
    wk_segmeta_t &idx_in = new_seg(1, WK_PREDICATE_ID, WK_DIR_IN), &idx_out = new_seg(1, WK_PREDICATE_ID, WK_DIR_OUT);
    for (int p = 1; p <= npreds; p++) {
        wk_segmeta_t &mo = new_seg(0, (uint32_t)p, WK_DIR_OUT), &mi = new_seg(0, (uint32_t)p, WK_DIR_IN);
        mo.num_edges = edges_of(so, p);
        mi.num_edges = edges_of(si, p);
        idx_out.num_edges += keys_of(si, p);
        idx_in.num_edges += keys_of(so, p) + type_cnt[p];
        mo.num_keys = mo.num_edges ? keys_of(so, p) : 0;
        mi.num_keys = mi.num_edges ? keys_of(si, p) : 0;
        mo.edge_start = alloc_edges(mo.num_edges);
        mi.edge_start = alloc_edges(mi.num_edges);
        alloc_buckets(mo);
        alloc_buckets(mi);
    }
    idx_out.edge_start = alloc_edges(idx_out.num_edges);
    idx_out.num_keys = local_preds.size();
    alloc_buckets(idx_out);
    idx_in.edge_start = alloc_edges(idx_in.num_edges);
    idx_in.num_keys = local_preds.size() + num_typeid;
    alloc_buckets(idx_in);
    if (ext_fail || main_off > num_buckets) return WK_ERR_STORE_FULL;
    const uint64_t num_edges = last_entry ? last_entry : 1;

    // ---- the two arrays -------------------------------------------------------------------------
    ulonglong2 *V = nullptr;
    uint32_t *EDG = nullptr;
    if (cudaMalloc((void **)&V, num_slots * sizeof(ulonglong2)) != cudaSuccess) return WK_ERR_CUDA;
    if (cudaMalloc((void **)&EDG, num_edges * sizeof(uint32_t)) != cudaSuccess) { cudaFree(V); return WK_ERR_CUDA; }
    struct Guard {   // released unless the store takes them over
        ulonglong2 *&v; uint32_t *&e; bool armed = true;
        ~Guard() { if (armed) { cudaFree(v); cudaFree(e); } }
    } guard{V, EDG};
    B_TRY(cudaMemsetAsync(V, 0, num_slots * sizeof(ulonglong2), st));
    B_TRY(cudaMemsetAsync(EDG, 0, num_edges * sizeof(uint32_t), st));

    // edge runs: the sorted far-end column of every predicate, as is
    for (uint32_t p : local_preds) {
        const wk_segmeta_t &mo = segs[std::make_tuple(p, 0, WK_DIR_OUT)], &mi = segs[std::make_tuple(p, 0, WK_DIR_IN)];
        if (mo.num_edges) B_TRY(cudaMemcpyAsync(EDG + mo.edge_start, so.E + so.tri_lb[p], mo.num_edges * 4, cudaMemcpyDeviceToDevice, st));
        if (mi.num_edges) B_TRY(cudaMemcpyAsync(EDG + mi.edge_start, si.E + si.tri_lb[p], mi.num_edges * 4, cudaMemcpyDeviceToDevice, st));
    }
    // index lists and their keys (insert_idx): [0|p|IN] subjects of p, [0|p|OUT] objects of p, [0|t|IN] instances of t
    std::vector<uint64_t> ikey, iptr;
    for (int d : {WK_DIR_IN, WK_DIR_OUT}) {
        const wk_segmeta_t &sg = (d == WK_DIR_IN) ? idx_in : idx_out;
        const Side &src = (d == WK_DIR_IN) ? so : si;
        uint64_t off = sg.edge_start;
        for (uint32_t p : local_preds) {
            if (p == WK_TYPE_ID) continue;
            const uint64_t nk = keys_of(src, (int)p);
            if (nk == 0) continue;
            if (nk >= (1ull << WK_PTR_SIZE_BITS)) return WK_ERR_STORE_FULL;
            ikey.push_back(make_key(0, p, (uint32_t)d));
            iptr.push_back(make_ptr(nk, off));
            low32_kernel<<<grid_for(nk, sms), THREADS, 0, st>>>(src.UK + src.key_lb[p], nk, EDG + off);
            off += nk;
        }
        if (d == WK_DIR_IN) {
            const uint64_t nt = type_lb[npreds + 1] - type_lb[2];
            if (nt) B_TRY(cudaMemcpyAsync(EDG + off, T_inst + type_lb[2], nt * 4, cudaMemcpyDeviceToDevice, st));
            for (int t = 2; t <= npreds; t++) {
                if (type_cnt[t] == 0) continue;
                if (type_cnt[t] >= (1ull << WK_PTR_SIZE_BITS)) return WK_ERR_STORE_FULL;
                ikey.push_back(make_key(0, (uint32_t)t, WK_DIR_IN));
                iptr.push_back(make_ptr(type_cnt[t], off + (type_lb[t] - type_lb[2])));
            }
            off += nt;
        }
        if (off > sg.edge_start + sg.num_edges) return WK_ERR_STORE_FULL;
    }

    // ---- insertion ------------------------------------------------------------------------------
    const uint64_t tabn = 2ull * 2 * (npreds + 1);
    std::vector<SegSlot> tab(tabn);
    memset(tab.data(), 0, tabn * sizeof(SegSlot));
    for (auto &kv : segs) {
        const wk_segmeta_t &m = kv.second;
        SegSlot &t = tab[(uint64_t)m.index * 2 * (npreds + 1) + (uint64_t)m.pid * 2 + m.dir];
        t.bucket_start = m.bucket_start;
        t.num_buckets = std::max<uint64_t>(m.num_buckets, 1);
        if (m.index == 0 && m.pid >= 1) t.edge_base = m.edge_start - (m.dir == WK_DIR_OUT ? so : si).tri_lb[m.pid];
    }
    // every index key hashes into the one index segment of its direction
    for (int p = 0; p <= npreds; p++)
        for (int d = 0; d <= 1; d++) {
            const wk_segmeta_t &m = d == WK_DIR_IN ? idx_in : idx_out;
            SegSlot &t = tab[1ull * 2 * (npreds + 1) + (uint64_t)p * 2 + d];
            t.bucket_start = m.bucket_start;
            t.num_buckets = m.num_buckets;
        }
    SegSlot *d_tab = nullptr;
    B_TRY(ar.get(&d_tab, tabn));
    B_TRY(cudaMemcpyAsync(d_tab, tab.data(), tabn * sizeof(SegSlot), cudaMemcpyHostToDevice, st));

    const uint64_t nkeys = so.nk + si.nk + ikey.size();
    uint64_t *ekey = nullptr, *eptr = nullptr, *extscan = nullptr;
    uint32_t *rank = nullptr, *cnt = nullptr, *need = nullptr;
    B_TRY(ar.get(&ekey, nkeys)); B_TRY(ar.get(&eptr, nkeys)); B_TRY(ar.get(&rank, nkeys));
    B_TRY(ar.get(&cnt, num_buckets)); B_TRY(ar.get(&need, num_buckets)); B_TRY(ar.get(&extscan, num_buckets + 1));
    B_TRY(cudaMemsetAsync(cnt, 0, num_buckets * sizeof(uint32_t), st));
    B_TRY(cudaMemsetAsync(d_ctl, 0, sizeof(BuildCtl), st));
    entries_kernel<<<grid_for(so.nk, sms), THREADS, 0, st>>>(so.UK, so.CNT, so.OFF, so.nk, WK_DIR_OUT, npreds, d_tab, ekey, eptr, d_ctl);
    entries_kernel<<<grid_for(si.nk, sms), THREADS, 0, st>>>(si.UK, si.CNT, si.OFF, si.nk, WK_DIR_IN, npreds, d_tab, ekey + so.nk,
                                                             eptr + so.nk, d_ctl);
    if (!ikey.empty()) {
        B_TRY(cudaMemcpyAsync(ekey + so.nk + si.nk, ikey.data(), ikey.size() * 8, cudaMemcpyHostToDevice, st));
        B_TRY(cudaMemcpyAsync(eptr + so.nk + si.nk, iptr.data(), iptr.size() * 8, cudaMemcpyHostToDevice, st));
    }
    rank_kernel<<<grid_for(nkeys, sms), THREADS, 0, st>>>(ekey, nkeys, npreds, d_tab, cnt, rank);
    need_kernel<<<grid_for(num_buckets, sms), THREADS, 0, st>>>(cnt, num_buckets, need);
    {
        size_t tb = 0;
        auto it = thrust::make_transform_iterator((const uint32_t *)need, CastU64());
        B_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tb, it, extscan, num_buckets, st));
        void *t2 = nullptr;
        B_TRY(ar.get((uint8_t **)&t2, tb));
        B_TRY(cub::DeviceScan::ExclusiveSum(t2, tb, it, extscan, num_buckets, st));
    }
    // ext demand per segment against its extent
    std::vector<wk_segmeta_t *> order;   // by bucket_start
    for (auto &kv : segs)
        if (kv.second.num_buckets) order.push_back(&kv.second);
    std::sort(order.begin(), order.end(), [](const wk_segmeta_t *a, const wk_segmeta_t *b) { return a->bucket_start < b->bucket_start; });
    const int nso = (int)order.size();
    std::vector<uint64_t> h_idx(2 * nso), h_scan(2 * nso), h_need_last(nso);
    for (int i = 0; i < nso; i++) {
        h_idx[2 * i] = order[i]->bucket_start;
        h_idx[2 * i + 1] = order[i]->bucket_start + order[i]->num_buckets - 1;   // inclusive end = scan[last] + need[last]
    }
    uint64_t *d_idx = nullptr, *d_g = nullptr;
    B_TRY(ar.get(&d_idx, 2 * (uint64_t)nso)); B_TRY(ar.get(&d_g, 2 * (uint64_t)nso));
    B_TRY(cudaMemcpyAsync(d_idx, h_idx.data(), 2 * nso * 8, cudaMemcpyHostToDevice, st));
    gather_kernel<<<(2 * nso + 127) / 128, 128, 0, st>>>(extscan, d_idx, 2 * nso, d_g);
    B_TRY(cudaMemcpyAsync(h_scan.data(), d_g, 2 * nso * 8, cudaMemcpyDeviceToHost, st));
    std::vector<uint32_t> h_last(nso);
    for (int i = 0; i < nso; i++) B_TRY(cudaMemcpyAsync(&h_last[i], need + h_idx[2 * i + 1], 4, cudaMemcpyDeviceToHost, st));
    BuildCtl h;
    B_TRY(cudaMemcpyAsync(&h, d_ctl, sizeof(h), cudaMemcpyDeviceToHost, st));
    B_TRY(cudaStreamSynchronize(st));
    if (h.big_run) return WK_ERR_STORE_FULL;
    uint64_t used_ext = 0;
    std::vector<uint64_t> seg_start(nso);
    std::vector<int64_t> seg_ext_base(nso);
    for (int i = 0; i < nso; i++) {
        const uint64_t used = h_scan[2 * i + 1] + h_last[i] - h_scan[2 * i];
        if (used > order[i]->ext_num) return WK_ERR_STORE_FULL;   // "segment exceeded its single ext extent"
        used_ext += used;
        seg_start[i] = order[i]->bucket_start;
        seg_ext_base[i] = (int64_t)order[i]->ext_start - (int64_t)h_scan[2 * i];
    }
    for (auto &kv : segs) {   // hand every table row its ext base
        const wk_segmeta_t &m = kv.second;
        if (!m.num_buckets) continue;
        const int i = (int)(std::lower_bound(seg_start.begin(), seg_start.end(), m.bucket_start) - seg_start.begin());
        if (m.index == 0) tab[(uint64_t)m.pid * 2 + m.dir].ext_base = seg_ext_base[i];
        else
            for (int p = 0; p <= npreds; p++) tab[1ull * 2 * (npreds + 1) + (uint64_t)p * 2 + m.dir].ext_base = seg_ext_base[i];
    }
    B_TRY(cudaMemcpyAsync(d_tab, tab.data(), tabn * sizeof(SegSlot), cudaMemcpyHostToDevice, st));
    uint64_t *d_ss = nullptr;
    int64_t *d_se = nullptr;
    B_TRY(ar.get(&d_ss, (uint64_t)nso)); B_TRY(ar.get(&d_se, (uint64_t)nso));
    B_TRY(cudaMemcpyAsync(d_ss, seg_start.data(), nso * 8, cudaMemcpyHostToDevice, st));
    B_TRY(cudaMemcpyAsync(d_se, seg_ext_base.data(), nso * 8, cudaMemcpyHostToDevice, st));
    place_kernel<<<grid_for(nkeys, sms), THREADS, 0, st>>>(ekey, eptr, rank, nkeys, npreds, d_tab, extscan, V);
    link_kernel<<<grid_for(num_buckets, sms), THREADS, 0, st>>>(need, extscan, num_buckets, d_ss, d_se, nso, V);
    B_TRY(cudaGetLastError());
    B_TRY(cudaStreamSynchronize(st));
    const double t_ins = now_ms();

    std::vector<wk_segmeta_t> seglist;
    for (auto &kv : segs) seglist.push_back(kv.second);
    rc = wk_store_adopt(device, (wk_vertex_t *)V, num_slots, EDG, num_edges, seglist.data(), (int)seglist.size(), 1, out);
    if (rc) return rc;
    guard.armed = false;
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->num_keys = total_keys;
        stats->num_triples_out = so.m;
        stats->num_triples_in = si.m;
        stats->num_buckets = num_buckets;
        stats->num_buckets_ext = num_buckets_ext;
        stats->used_ext = used_ext;
        stats->num_slots = num_slots;
        stats->num_edges = num_edges;
        stats->ms_upload = (float)(t_up - t_begin);
        stats->ms_sort = (float)(t_sort - t_up);
        stats->ms_insert = (float)(t_ins - t_sort);
        stats->ms_total = (float)(now_ms() - t_begin);
    }
    return WK_SUCCESS;
}

// Result-table post-processing of final_process (reference core/engine/sparql.hpp:1424-1505) on the
// device: DISTINCT (rows ordered like ReduceCmp -- all columns, compared as signed ints -- then
// adjacent rows that agree on the required columns collapse to the first) and OFFSET / LIMIT.
// ORDER BY compares strings from the string server and stays on the host.
//
// DISTINCT = C stable radix-sort passes over a row permutation (last column first), one flag pass
// through the permutation, a stream compaction of the permutation and ONE gather of the rows.
#include <cub/cub.cuh>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

#include "wk_internal.h"
#include "wukong_b200.h"

namespace {

constexpr int THREADS = 256;

#define T_TRY(expr)                                                                                \
    do {                                                                                           \
        cudaError_t e_ = (expr);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            fprintf(stderr, "wukong_b200 table op: %s failed: %s\n", #expr, cudaGetErrorString(e_)); \
            rc = WK_ERR_CUDA;                                                                      \
            goto done;                                                                             \
        }                                                                                          \
    } while (0)

__global__ void __launch_bounds__(THREADS) iota_kernel(uint32_t *idx, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * THREADS + threadIdx.x; i < n; i += (uint64_t)gridDim.x * THREADS) idx[i] = (uint32_t)i;
}
__global__ void __launch_bounds__(THREADS) gather_col_kernel(const uint32_t *__restrict__ in, const uint32_t *__restrict__ idx, uint64_t n,
                                                             int C, int c, int32_t *__restrict__ keys) {
    for (uint64_t i = (uint64_t)blockIdx.x * THREADS + threadIdx.x; i < n; i += (uint64_t)gridDim.x * THREADS)
        keys[i] = (int32_t)in[(uint64_t)idx[i] * C + c];
}
struct ReqCols { int32_t n; int8_t col[32]; };   // 32 = MAX_COLS of the engine (wk_device.cuh)
__global__ void __launch_bounds__(THREADS) distinct_flag_kernel(const uint32_t *__restrict__ in, const uint32_t *__restrict__ idx, uint64_t n,
                                                                int C, ReqCols rq, uint8_t *__restrict__ keep) {
    for (uint64_t i = (uint64_t)blockIdx.x * THREADS + threadIdx.x; i < n; i += (uint64_t)gridDim.x * THREADS) {
        bool k = (i == 0);
        if (!k) {
            const uint32_t *a = in + (uint64_t)idx[i - 1] * C, *b = in + (uint64_t)idx[i] * C;
            for (int j = 0; j < rq.n; j++) k |= a[rq.col[j]] != b[rq.col[j]];
        }
        keep[i] = k;
    }
}
__global__ void __launch_bounds__(THREADS) gather_rows_kernel(const uint32_t *__restrict__ in, const uint32_t *__restrict__ idx,
                                                              const unsigned long long *__restrict__ n_dev, int C, uint32_t *__restrict__ out) {
    const uint64_t words = (uint64_t)(*n_dev) * C;
    for (uint64_t w = (uint64_t)blockIdx.x * THREADS + threadIdx.x; w < words; w += (uint64_t)gridDim.x * THREADS) {
        const uint64_t r = w / C;
        out[w] = in[(uint64_t)idx[r] * C + (w - r * C)];
    }
}
// OFFSET / LIMIT: out = in[offset .. offset + limit), row counts stay on the device
__global__ void __launch_bounds__(THREADS) slice_kernel(const uint32_t *__restrict__ in, const unsigned long long *__restrict__ in_count,
                                                        int C, uint64_t offset, int64_t limit, uint32_t *__restrict__ out,
                                                        unsigned long long *out_count) {
    const uint64_t n = *in_count;
    uint64_t m = n > offset ? n - offset : 0;
    if (limit >= 0 && (uint64_t)limit < m) m = (uint64_t)limit;
    const uint64_t words = m * C;
    const uint32_t *src = in + offset * C;
    for (uint64_t w = (uint64_t)blockIdx.x * THREADS + threadIdx.x; w < words; w += (uint64_t)gridDim.x * THREADS) out[w] = src[w];
    if (blockIdx.x == 0 && threadIdx.x == 0) *out_count = m;
}

int grid_for(uint64_t n, int sms) {
    const uint64_t want = (n + THREADS - 1) / THREADS;
    return (int)(want < 1 ? 1 : (want > (uint64_t)sms * 8 ? (uint64_t)sms * 8 : want));
}

}  // namespace

int wk_internal_distinct(cudaStream_t st, int sms, const uint32_t *in, uint32_t *out, uint64_t rows, int C, const int32_t *req_cols,
                         int nreq, uint64_t *d_out_rows) {
    if (rows >= (1ull << 32) || C <= 0 || nreq <= 0 || nreq > 32) return WK_ERR_BAD_ARG;
    int rc = WK_SUCCESS;
    uint32_t *idx = nullptr, *idx_alt = nullptr;
    int32_t *keys = nullptr, *keys_alt = nullptr;
    uint8_t *keep = nullptr;
    void *tmp = nullptr;
    size_t tb_sort = 0, tb_sel = 0;
    ReqCols rq;
    rq.n = nreq;
    for (int j = 0; j < nreq; j++) rq.col[j] = (int8_t)req_cols[j];
    const int grid = grid_for(rows, sms);
    T_TRY(cudaMallocAsync((void **)&idx, rows * 4, st));
    T_TRY(cudaMallocAsync((void **)&idx_alt, rows * 4, st));
    T_TRY(cudaMallocAsync((void **)&keys, rows * 4, st));
    T_TRY(cudaMallocAsync((void **)&keys_alt, rows * 4, st));
    T_TRY(cudaMallocAsync((void **)&keep, rows, st));
    {
        cub::DoubleBuffer<int32_t> dk(keys, keys_alt);
        cub::DoubleBuffer<uint32_t> dv(idx, idx_alt);
        T_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tb_sort, dk, dv, rows, 0, 32, st));
        T_TRY(cub::DeviceSelect::Flagged(nullptr, tb_sel, idx, keep, idx_alt, (unsigned long long *)d_out_rows, rows, st));
        T_TRY(cudaMallocAsync(&tmp, tb_sort > tb_sel ? tb_sort : tb_sel, st));
        iota_kernel<<<grid, THREADS, 0, st>>>(dv.Current(), rows);
        // least significant column first; every pass is stable, so the final order is ReduceCmp's
        for (int c = C - 1; c >= 0; c--) {
            gather_col_kernel<<<grid, THREADS, 0, st>>>(in, dv.Current(), rows, C, c, dk.Current());
            T_TRY(cub::DeviceRadixSort::SortPairs(tmp, tb_sort, dk, dv, rows, 0, 32, st));
        }
        uint32_t *sorted = dv.Current(), *other = dv.Alternate();
        distinct_flag_kernel<<<grid, THREADS, 0, st>>>(in, sorted, rows, C, rq, keep);
        T_TRY(cub::DeviceSelect::Flagged(tmp, tb_sel, sorted, keep, other, (unsigned long long *)d_out_rows, rows, st));
        gather_rows_kernel<<<grid_for(rows * C, sms), THREADS, 0, st>>>(in, other, (const unsigned long long *)d_out_rows, C, out);
        T_TRY(cudaGetLastError());
    }
done:
    if (idx) cudaFreeAsync(idx, st);
    if (idx_alt) cudaFreeAsync(idx_alt, st);
    if (keys) cudaFreeAsync(keys, st);
    if (keys_alt) cudaFreeAsync(keys_alt, st);
    if (keep) cudaFreeAsync(keep, st);
    if (tmp) cudaFreeAsync(tmp, st);
    return rc;
}

int wk_internal_slice(cudaStream_t st, int sms, const uint32_t *in, const uint64_t *d_in_rows, int C, uint64_t offset, int64_t limit,
                      uint32_t *out, uint64_t *d_out_rows, uint64_t rows_upper_bound) {
    if (C <= 0) return WK_ERR_BAD_ARG;
    slice_kernel<<<grid_for(rows_upper_bound * (uint64_t)C, sms), THREADS, 0, st>>>(in, (const unsigned long long *)d_in_rows, C, offset, limit, out,
                                                                                   (unsigned long long *)d_out_rows);
    return cudaGetLastError() == cudaSuccess ? WK_SUCCESS : WK_ERR_CUDA;
}

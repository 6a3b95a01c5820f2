// Entry points between the translation units of libwukong_b200.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// DISTINCT of final_process (sparql.hpp:1428-1472): `rows` x C table `in` -> `out`, new row count stored to *d_out_rows (device)
int wk_internal_distinct(cudaStream_t st, int sms, const uint32_t *in, uint32_t *out, uint64_t rows, int C, const int32_t *req_cols,
                         int nreq, uint64_t *d_out_rows);
// OFFSET / LIMIT of final_process (sparql.hpp:1487-1499); both row counts live on the device
int wk_internal_slice(cudaStream_t st, int sms, const uint32_t *in, const uint64_t *d_in_rows, int C, uint64_t offset, int64_t limit,
                      uint32_t *out, uint64_t *d_out_rows, uint64_t rows_upper_bound);

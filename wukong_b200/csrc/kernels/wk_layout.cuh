// Bit layout of the cluster-hash store and its hash function, shared by the query kernels and the
// device-side store builder (reference: store/vertex.hpp:47-66, 116-119; utils/math.hpp:58-67).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace wk {

// ---- key / pointer bit layout ---------------------------------------------------------------
// ikey_t: dir:1 | pid:17 | vid:46 (LSB first)  => raw = vid<<18 | pid<<1 | dir
// iptr_t: size:28 | off:34 | type:2
#define WK_KEY_VID_SHIFT 18
#define WK_PTR_SIZE_BITS 28
#define WK_PTR_OFF_BITS 34

__host__ __device__ __forceinline__ uint64_t make_key(uint64_t vid, uint32_t pid, uint32_t dir) {
    return (vid << WK_KEY_VID_SHIFT) | ((uint64_t)pid << 1) | (uint64_t)dir;
}
__host__ __device__ __forceinline__ uint32_t ptr_size(uint64_t p) { return (uint32_t)(p & ((1ull << WK_PTR_SIZE_BITS) - 1)); }
__host__ __device__ __forceinline__ uint64_t ptr_off(uint64_t p) { return (p >> WK_PTR_SIZE_BITS) & ((1ull << WK_PTR_OFF_BITS) - 1); }

// Thomas Wang 64-bit mix (the reference's math::hash_u64, utils/math.hpp:58-67) of the raw key.
__host__ __device__ __forceinline__ uint64_t hash_u64(uint64_t key) {
    key = (~key) + (key << 21);
    key = key ^ (key >> 24);
    key = (key + (key << 3)) + (key << 8);
    key = key ^ (key >> 14);
    key = (key + (key << 2)) + (key << 4);
    key = key ^ (key >> 28);
    key = key + (key << 31);
    return key;
}

}  // namespace wk

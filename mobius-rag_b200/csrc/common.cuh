// common.cuh -- shared device helpers for the mrag kernels (sm_100a only).
//
// Candidate keys.  Every (score, row) candidate travels as one u64
//     key = orderable(score) << 32 | (0xFFFFFFFF - local_row)
// so that "similarity DESC, row ASC" (the ORDER BY of vector_store.py:284 /
// corpus_search.py:1534 with the tie-break this library adds) is plain u64 descending.
// key 0 is "empty".  NaN scores never become keys (they are appended by the NaN tail pass).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#define MRAG_DEVINL __device__ __forceinline__

namespace mrag {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

MRAG_DEVINL uint32_t f2ord(float f) {
    uint32_t u = __float_as_uint(f + 0.0f);                  // -0 -> +0
    return u ^ (uint32_t(int32_t(u) >> 31) | 0x80000000u);
}
MRAG_DEVINL float ord2f(uint32_t o) {
    uint32_t u = (o & 0x80000000u) ? (o ^ 0x80000000u) : ~o;
    return __uint_as_float(u);
}
MRAG_DEVINL uint64_t make_key(float score, uint32_t row) {
    return (uint64_t(f2ord(score)) << 32) | uint64_t(0xFFFFFFFFu - row);
}
MRAG_DEVINL float key_score(uint64_t k) { return ord2f(uint32_t(k >> 32)); }
MRAG_DEVINL uint32_t key_row(uint64_t k) { return 0xFFFFFFFFu - uint32_t(k); }

MRAG_DEVINL int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

// Bitonic sort, descending, of n (power of two) u64 keys in shared memory by ONE warp.
MRAG_DEVINL void warp_sort_desc(uint64_t* a, int n, int lane) {
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = lane; i < n; i += kWarp) {
                int p = i ^ j;
                if (p > i) {
                    uint64_t x = a[i], y = a[p];
                    bool desc = (i & k) == 0;
                    if (desc ? (x < y) : (x > y)) { a[i] = y; a[p] = x; }
                }
            }
            __syncwarp();
        }
    }
}

// Same by a whole thread block (all threads must call; n power of two).
MRAG_DEVINL void block_sort_desc(uint64_t* a, int n) {
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                int p = i ^ j;
                if (p > i) {
                    uint64_t x = a[i], y = a[p];
                    bool desc = (i & k) == 0;
                    if (desc ? (x < y) : (x > y)) { a[i] = y; a[p] = x; }
                }
            }
            __syncthreads();
        }
    }
}

MRAG_DEVINL float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// 128-bit streaming load: read-only path, do not allocate in L1 (each corpus byte is used once).
MRAG_DEVINL uint4 ldg_stream(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

MRAG_DEVINL float bf16lo(uint32_t u) { return __uint_as_float(u << 16); }
MRAG_DEVINL float bf16hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

}  // namespace mrag

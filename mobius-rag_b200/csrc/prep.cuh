// prep.cuh -- write-side and per-call preparation kernels.
//   store_rows_kernel   fp32 rows -> storage dtype (+ zero padding to ld) + 1/|x| of the STORED row
//                       (embedding_worker.py:65-94 / publish.py:327-362 write float4 rows, un-normalised)
//   scatter_meta_kernel mrag_rowmeta AoS -> SoA columns + valid bitmap
//   filter_mask_kernel  K2: WHERE clauses -> row bitmap
//                       (corpus_search.py:516-560, 1471-1523; vector_store.py:245-267)
//   pool_bitmap_kernel  document_id = ANY(:inc_ids) -> document bitmap
//   tombstone_kernel    DELETE .. WHERE document_id = :id
//   query_prep_kernel   pad queries to ld, 1/|q|, bf16 copy for the tensor-core path
#pragma once
#include "common.cuh"
#include "../../include/mrag.h"
#include <math_constants.h>

namespace mrag {

// one warp per row
// `shadow` (fp32 indexes only, may be null): bf16 copy of the row for the candidate-generating
// tensor-core scan; 1/|x| is always that of the PRIMARY row.
template <int DT>
__global__ void __launch_bounds__(256) store_rows_kernel(const float* __restrict__ src, int64_t n, int dim,
                                                        void* __restrict__ dst, int ld, int64_t first_row,
                                                        float* __restrict__ inv_norm,
                                                        __nv_bfloat16* __restrict__ shadow) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (r >= n) return;
    const float* s = src + r * dim;
    float ss = 0.0f;
    for (int e = lane; e < ld; e += 32) {
        float v = (e < dim) ? s[e] : 0.0f;
        if (DT == 1) {
            __nv_bfloat16 b = __float2bfloat16_rn(v);
            reinterpret_cast<__nv_bfloat16*>(dst)[(first_row + r) * ld + e] = b;
            v = __bfloat162float(b);
        } else {
            reinterpret_cast<float*>(dst)[(first_row + r) * ld + e] = v;
            if (shadow) shadow[(first_row + r) * ld + e] = __float2bfloat16_rn(v);
        }
        ss = fmaf(v, v, ss);
    }
    ss = warp_sum(ss);
    if (lane == 0) inv_norm[first_row + r] = (ss > 0.0f) ? 1.0f / sqrtf(ss) : CUDART_INF_F;
}

struct MetaCols {
    uint32_t* doc_idx; uint16_t* payer; uint8_t* state; uint8_t* program; uint8_t* authority;
    uint8_t* source_type; uint32_t* valid;
    uint32_t* live;      // row exists (inserted and not deleted), whether or not it has a vector
};

__global__ void __launch_bounds__(256) scatter_meta_kernel(const mrag_rowmeta* __restrict__ m, int64_t n,
                                                          int64_t first_row, MetaCols c) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const mrag_rowmeta x = m[i];
    const int64_t r = first_row + i;
    c.doc_idx[r] = x.doc_idx; c.payer[r] = x.payer; c.state[r] = x.state; c.program[r] = x.program;
    c.authority[r] = x.authority; c.source_type[r] = x.source_type;
    // the row's bits are WRITTEN, not only set: a slot past `size` may carry bits of an append that failed half way
    const uint32_t bit = 1u << (r & 31);
    if (x.valid) atomicOr(&c.valid[r >> 5], bit); else atomicAnd(&c.valid[r >> 5], ~bit);
    atomicOr(&c.live[r >> 5], bit);
}

// device copy of mrag_filter without the host pointer
struct DevFilter {
    uint32_t flags;
    uint64_t payer_any[MRAG_PAYER_WORDS];
    uint64_t payer_alt_any[MRAG_PAYER_WORDS];
    uint16_t alt_state, state_eq, program_eq, authority_eq, source_type_eq;
    uint32_t doc_eq;
    uint64_t tag_state_any[MRAG_SMALL_WORDS];
    uint64_t tag_program_any[MRAG_SMALL_WORDS];
    uint64_t tag_payer_any[MRAG_PAYER_WORDS];
    uint64_t tag_any[MRAG_TAG_WORDS];
};

MRAG_DEVINL bool bit_in(const uint64_t* set, uint32_t code, uint32_t words) {
    return (code >> 6) < words && ((set[code >> 6] >> (code & 63)) & 1ull);
}

// K2.  One thread per row, one warp per bitmap word (ballot).  Reads 10 B of metadata per row.
__global__ void __launch_bounds__(256) filter_mask_kernel(const __grid_constant__ DevFilter f, MetaCols c, int64_t n,
                                                         const uint32_t* __restrict__ pool_bits,
                                                         const uint64_t* __restrict__ doc_tags, int64_t n_tag_docs,
                                                         uint32_t* __restrict__ mask_out,
                                                         unsigned long long* __restrict__ n_pass, int include_null_vec) {
    const int64_t r = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    bool ok = false;
    if (r < n) {
        // embedding_vec IS NOT NULL -- or, for statements without that clause (the d-tag arm), any live row
        ok = ((include_null_vec ? c.live : c.valid)[r >> 5] >> (r & 31)) & 1u;
        if (ok && f.flags) {
            const uint32_t payer = c.payer[r], state = c.state[r];
            if (f.flags & MRAG_F_PAYER)
                ok = bit_in(f.payer_any, payer, MRAG_PAYER_WORDS) ||
                     (bit_in(f.payer_alt_any, payer, MRAG_PAYER_WORDS) && state == f.alt_state);
            if (ok && (f.flags & MRAG_F_STATE)) ok = state == f.state_eq;
            if (ok && (f.flags & MRAG_F_PROGRAM)) ok = c.program[r] == f.program_eq;
            if (ok && (f.flags & MRAG_F_AUTHORITY)) ok = c.authority[r] == f.authority_eq;
            if (ok && (f.flags & MRAG_F_SOURCE_TYPE)) ok = c.source_type[r] == f.source_type_eq;
            if (ok && (f.flags & (MRAG_F_DOC_EQ | MRAG_F_DOC_POOL | MRAG_F_TAG_RELAXED))) {
                const uint32_t d = c.doc_idx[r];
                if (f.flags & MRAG_F_DOC_EQ) ok = d == f.doc_eq;
                if (ok && (f.flags & MRAG_F_DOC_POOL)) ok = (pool_bits[d >> 5] >> (d & 31)) & 1u;
                if (ok && (f.flags & MRAG_F_TAG_RELAXED)) {
                    bool any = false;
                    if (int64_t(d) < n_tag_docs) {
#pragma unroll
                        for (int w = 0; w < MRAG_TAG_WORDS; ++w) any |= (doc_tags[size_t(d) * MRAG_TAG_WORDS + w] & f.tag_any[w]) != 0;
                    }
                    ok = any;
                }
            }
            if (ok && (f.flags & MRAG_F_TAG_STRICT))
                ok = bit_in(f.tag_state_any, state, MRAG_SMALL_WORDS) ||
                     bit_in(f.tag_program_any, c.program[r], MRAG_SMALL_WORDS) ||
                     bit_in(f.tag_payer_any, payer, MRAG_PAYER_WORDS);
        }
    }
    const uint32_t word = __ballot_sync(kFull, ok);
    if ((threadIdx.x & 31) == 0) {
        if (r < ((n + 31) & ~int64_t(31))) mask_out[r >> 5] = word;
        if (n_pass && word) atomicAdd(n_pass, (unsigned long long)__popc(word));
    }
}

__global__ void pool_bitmap_kernel(const uint32_t* __restrict__ pool, int64_t n_pool, uint32_t* bits, int64_t n_docs) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n_pool) return;
    const uint32_t d = pool[i];
    if (int64_t(d) < n_docs) atomicOr(&bits[d >> 5], 1u << (d & 31));
}

// ---- candidate-pool cascade (build_candidate_pool, corpus_search_agent.py:1762-1888): per document, which levels of
//      J&D&P -> J&D -> AHCA&D -> AHCA it belongs to.  Pure bitset tests over the per-document tag sets that already sit in HBM.
struct DevPoolQuery {
    uint64_t d_all[MRAG_TAG_WORDS], p_all[MRAG_TAG_WORDS];   // bits (document d / p tag sets) that must ALL be present
    uint64_t j_all[MRAG_JTAG_WORDS], ahca[MRAG_JTAG_WORDS];   // same over the j tag sets; the AHCA authority bit
    int has_j, has_d, has_p, has_ahca;                        // kind named in the query and every code known to the vocabulary
};

// out: 4 document bitmaps of `words` u32 each (L1 J&D&P, L2 J&D, L3 AHCA&D, L4 AHCA); counts[0..3] their sizes,
// counts[4] = documents carrying every d: code (the cascade tries L3 only when that set is not empty, :1851)
__global__ void __launch_bounds__(256) pool_cascade_kernel(const __grid_constant__ DevPoolQuery q,
                                                          const uint64_t* __restrict__ doc_tags, int64_t n_tag_docs,
                                                          const uint64_t* __restrict__ doc_jtags, int64_t n_jtag_docs,
                                                          int64_t n_docs, int64_t words,
                                                          uint32_t* __restrict__ out, unsigned long long* __restrict__ counts) {
    const int64_t d = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    bool J = false, D = false, P = false, A = false;
    if (d < n_docs) {
        if (d < n_tag_docs) {
            bool dd = q.has_d != 0, pp = q.has_p != 0;
#pragma unroll
            for (int w = 0; w < MRAG_TAG_WORDS; ++w) {
                const uint64_t t = doc_tags[size_t(d) * MRAG_TAG_WORDS + w];
                dd &= (t & q.d_all[w]) == q.d_all[w];
                pp &= (t & q.p_all[w]) == q.p_all[w];
            }
            D = dd; P = pp;
        }
        if (d < n_jtag_docs) {
            bool jj = q.has_j != 0, aa = q.has_ahca != 0;
#pragma unroll
            for (int w = 0; w < MRAG_JTAG_WORDS; ++w) {
                const uint64_t t = doc_jtags[size_t(d) * MRAG_JTAG_WORDS + w];
                jj &= (t & q.j_all[w]) == q.j_all[w];
                aa &= (t & q.ahca[w]) == q.ahca[w];
            }
            J = jj; A = aa;
        }
    }
    const bool lv[4] = {J && D && P, J && D, A && D, A};
#pragma unroll
    for (int l = 0; l < 4; ++l) {
        const uint32_t word = __ballot_sync(kFull, lv[l]);
        if ((threadIdx.x & 31) == 0 && (d >> 5) < words) {
            out[size_t(l) * words + (d >> 5)] = word;
            if (word) atomicAdd(counts + l, (unsigned long long)__popc(word));
        }
    }
    const uint32_t dword = __ballot_sync(kFull, D);
    if ((threadIdx.x & 31) == 0 && dword) atomicAdd(counts + 4, (unsigned long long)__popc(dword));
}

// keep only the first `cap` set bits (ascending document index) of a bitmap: one block; writes the kept count
__global__ void __launch_bounds__(1024) pool_cap_kernel(uint32_t* __restrict__ bits, int64_t words, int64_t cap, unsigned long long* kept) {
    __shared__ long long part[1024];
    const int t = threadIdx.x;
    const int64_t per = (words + 1023) / 1024;
    const int64_t lo = int64_t(t) * per < words ? int64_t(t) * per : words;
    const int64_t hi = lo + per < words ? lo + per : words;
    long long c = 0;
    for (int64_t w = lo; w < hi; ++w) c += __popc(bits[w]);
    part[t] = c;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {            // inclusive scan
        const long long v = (t >= off) ? part[t - off] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    long long before = part[t] - c;
    for (int64_t w = lo; w < hi; ++w) {
        uint32_t x = bits[w];
        const int n = __popc(x);
        if (before >= cap) x = 0u;
        else if (before + n > cap) {
            int keep = int(cap - before);
            uint32_t y = 0u;
            while (keep-- > 0) { const uint32_t low = x & (0u - x); y |= low; x ^= low; }
            x = y;
        }
        bits[w] = x;
        before += n;
    }
    if (t == 1023) *kept = (unsigned long long)(part[1023] < (long long)cap ? part[1023] : (long long)cap);
}

__global__ void tombstone_kernel(const uint32_t* __restrict__ doc_idx, int64_t n, uint32_t doc, uint32_t* valid,
                                 uint32_t* live, unsigned long long* n_hit) {
    const int64_t r = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (r >= n) return;
    if (doc_idx[r] == doc) {
        uint32_t old = atomicAnd(&valid[r >> 5], ~(1u << (r & 31)));
        atomicAnd(&live[r >> 5], ~(1u << (r & 31)));
        if ((old >> (r & 31)) & 1u) atomicAdd(n_hit, 1ull);
    }
}

// out[0] += rows of [0, n) whose `live` bit is set, out[1] += rows whose `valid` bit is set (mrag_live_rows)
__global__ void __launch_bounds__(256) count_bits_kernel(const uint32_t* __restrict__ live, const uint32_t* __restrict__ valid, int64_t n,
                                                        unsigned long long* __restrict__ out) {
    const int64_t nwords = (n + 31) >> 5;
    unsigned a = 0, b = 0;
    for (int64_t w = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; w < nwords; w += int64_t(gridDim.x) * blockDim.x) {
        const uint32_t tail = (w == nwords - 1 && (n & 31)) ? ((1u << (n & 31)) - 1u) : 0xFFFFFFFFu;
        a += __popc(live[w] & tail);
        b += __popc(valid[w] & tail);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(kFull, a, o); b += __shfl_xor_sync(kFull, b, o); }
    if ((threadIdx.x & 31) == 0) {
        if (a) atomicAdd(out, (unsigned long long)a);
        if (b) atomicAdd(out + 1, (unsigned long long)b);
    }
}

// one warp per query: zero-pad to ld, 1/|q| (float4 query, vector_store.py:272), bf16 copy; qhl (optional): the tensor-memory
// image of the exact scan's A operand -- per query a plane of bf16 `hi` pairs and a plane of bf16 `lo` pairs (q = hi + lo to
// 16 mantissa bits, element 2c in the low half of word c), so that every scan CTA loads packed words instead of splitting
// 64 x ld floats itself
__global__ void __launch_bounds__(128) query_prep_kernel(const float* __restrict__ q, int nq, int dim, int ld,
                                                        float* __restrict__ qpad, float* __restrict__ qinv,
                                                        __nv_bfloat16* __restrict__ qbf, int nq_pad, uint32_t* __restrict__ qhl,
                                                        uint32_t* __restrict__ gthr = nullptr, uint32_t* __restrict__ gmax = nullptr,
                                                        int gk = 0, int gslots = 16, int* __restrict__ flags4 = nullptr) {
    const int lane = threadIdx.x & 31;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= nq_pad) return;
    // per-search scratch the scan expects cleared (saves the memset launches): the cross-CTA bound of query i, its group
    // maxima (slots [gk, 16) can never be the minimum), the four step flags
    if (i < nq) {
        if (gthr && lane == 0) gthr[i] = 0u;
        if (gmax)
            for (int j = lane; j < gslots; j += 32) gmax[size_t(i) * gslots + j] = j < gk ? 0u : 0xFFFFFFFFu;
    }
    if (flags4 && i == 0 && lane < 4) flags4[lane] = 0;
    float ss = 0.0f;
    if ((dim & 3) == 0 && (reinterpret_cast<uintptr_t>(q) & 15u) == 0 && i < nq && !qbf) {
        // vector path: all of a lane's 128-bit loads are independent (one round trip instead of ld / 32 dependent ones)
        const float4* src = reinterpret_cast<const float4*>(q + size_t(i) * dim);
        float4* dst = reinterpret_cast<float4*>(qpad + size_t(i) * ld);
        uint2* hi2 = qhl ? reinterpret_cast<uint2*>(qhl + size_t(i) * ld) : nullptr;          // plane of ld / 2 words
        uint2* lo2 = qhl ? hi2 + ld / 4 : nullptr;
        const int nv = ld >> 2, dv = dim >> 2;
        for (int v0 = 0; v0 < nv; v0 += 32 * 8) {
            float4 f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int v = v0 + 32 * j + lane;
                f[j] = v < dv ? __ldg(src + v) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int v = v0 + 32 * j + lane;
                if (v < nv) {
                    dst[v] = f[j];
                    const float e4[4] = {f[j].x, f[j].y, f[j].z, f[j].w};
                    if (qhl) {
                        __nv_bfloat16 h[4], l[4];
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            h[c] = __float2bfloat16_rn(e4[c]);
                            l[c] = __float2bfloat16_rn(e4[c] - __bfloat162float(h[c]));      // q - hi is exact in fp32
                        }
                        __nv_bfloat162 h01(h[0], h[1]), h23(h[2], h[3]), l01(l[0], l[1]), l23(l[2], l[3]);
                        hi2[v] = make_uint2(*reinterpret_cast<uint32_t*>(&h01), *reinterpret_cast<uint32_t*>(&h23));
                        lo2[v] = make_uint2(*reinterpret_cast<uint32_t*>(&l01), *reinterpret_cast<uint32_t*>(&l23));
                    }
#pragma unroll
                    for (int c = 0; c < 4; ++c) ss = fmaf(e4[c], e4[c], ss);
                }
            }
        }
    } else
    for (int e = lane; e < ld; e += 32) {
        float v = (i < nq && e < dim) ? q[size_t(i) * dim + e] : 0.0f;
        if (i < nq) qpad[size_t(i) * ld + e] = v;
        if (qbf) qbf[size_t(i) * ld + e] = __float2bfloat16_rn(v);
        if (qhl && i < nq) {
            const __nv_bfloat16 hi = __float2bfloat16_rn(v);
            const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));      // q - hi is exact in fp32
            __nv_bfloat16* planes = reinterpret_cast<__nv_bfloat16*>(qhl) + size_t(i) * 2 * ld;
            planes[e] = hi;
            planes[ld + e] = lo;
        }
        ss = fmaf(v, v, ss);
    }
    ss = warp_sum(ss);
    if (lane == 0 && i < nq) qinv[i] = (ss > 0.0f) ? 1.0f / sqrtf(ss) : CUDART_INF_F;
}

}  // namespace mrag

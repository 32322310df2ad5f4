// rescore.cuh -- exact rescoring of the candidates nominated by scan_mma128.cuh, the certificate
// that makes the result exact, and the hand-off of uncertified queries to the exact scan.
//
//   rescore_kernel   one warp per (query, candidate): fp32 dot of the fp32 query with the PRIMARY row
//                    (fp32, or bf16 upcast) -- the arithmetic of scan_gemv.cuh, i.e. of pgvector's float4
//                    loop (corpus_search.py:1525-1536) -- times 1/|x| * 1/|q|.
//   finalize_kernel  one block per query: sorts the K' exact keys, writes the top k, and checks
//                        a_min + eps < tau
//                    (a_min = approximate score of the worst candidate, tau = exact k-th best).  Every row
//                    that was NOT nominated has approx <= a_min, hence exact <= a_min + eps < tau: it cannot
//                    be in the top k and cannot tie with it.  A query that fails the check (dense score
//                    clusters) is appended to the fallback list and rescanned exactly by scan_gemv.
#pragma once
#include "common.cuh"
#include <math_constants.h>

namespace mrag {

struct RescoreArgs {
    const void* rows;        // primary storage [n][ld]
    int ld;
    const float* inv_norm;   // [n] 1/|x| of the primary row
    const float* q;          // [nq][ld] fp32
    const float* qinv;       // [nq]
    const int64_t* cand_rows;    // [nq][kc] local rows nominated (sorted by approximate score)
    const float* cand_scores;    // [nq][kc] approximate scores
    const int32_t* cand_counts;  // [nq]
    int nq, kc;              // kc = K'
    uint64_t* keys;          // [nq][kc] exact keys out (0 = empty)
};

template <int DT>
__global__ void __launch_bounds__(256) rescore_kernel(const RescoreArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (wid >= int64_t(a.nq) * a.kc) return;
    const int q = int(wid / a.kc), j = int(wid - int64_t(q) * a.kc);
    if (j >= a.cand_counts[q]) {
        if (lane == 0) a.keys[wid] = 0ull;
        return;
    }
    const int64_t row = a.cand_rows[wid];
    const float* qrow = a.q + size_t(q) * a.ld;
    float acc = 0.0f;
    if (DT == 1) {
        const uint4* x = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(a.rows) + size_t(row) * a.ld);
        for (int v = lane; v < a.ld / 8; v += 32) {
            const uint4 d = __ldg(x + v);
            const float4 q0 = __ldg(reinterpret_cast<const float4*>(qrow) + 2 * v);
            const float4 q1 = __ldg(reinterpret_cast<const float4*>(qrow) + 2 * v + 1);
            acc = fmaf(bf16lo(d.x), q0.x, acc); acc = fmaf(bf16hi(d.x), q0.y, acc);
            acc = fmaf(bf16lo(d.y), q0.z, acc); acc = fmaf(bf16hi(d.y), q0.w, acc);
            acc = fmaf(bf16lo(d.z), q1.x, acc); acc = fmaf(bf16hi(d.z), q1.y, acc);
            acc = fmaf(bf16lo(d.w), q1.z, acc); acc = fmaf(bf16hi(d.w), q1.w, acc);
        }
    } else {
        const float4* x = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(a.rows) + size_t(row) * a.ld);
        for (int v = lane; v < a.ld / 4; v += 32) {
            const float4 d = __ldg(x + v);
            const float4 q0 = __ldg(reinterpret_cast<const float4*>(qrow) + v);
            acc = fmaf(d.x, q0.x, acc); acc = fmaf(d.y, q0.y, acc);
            acc = fmaf(d.z, q0.z, acc); acc = fmaf(d.w, q0.w, acc);
        }
    }
    acc = warp_sum(acc);
    if (lane == 0) {
        const float s = acc * a.inv_norm[row] * a.qinv[q];
        a.keys[wid] = (s == s) ? make_key(s, uint32_t(row)) : 0ull;
    }
}

// clamp01(cosine) of already selected rows (hybrid search: the rerank score is returned, the similarity beside it)
struct CosRowsArgs {
    const void* rows; int ld; const float* inv_norm; const float* q; const float* qinv;
    const int64_t* sel_rows;     // [nq][k] global row ids (-1 = padding)
    const int32_t* sel_counts;   // [nq]
    int64_t row_base;
    int nq, k;
    float* cos_out;              // [nq][k]
};

template <int DT>
__global__ void __launch_bounds__(256) cos_rows_kernel(const CosRowsArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (wid >= int64_t(a.nq) * a.k) return;
    const int q = int(wid / a.k), j = int(wid - int64_t(q) * a.k);
    if (j >= a.sel_counts[q]) {
        if (lane == 0) a.cos_out[wid] = CUDART_NAN_F;
        return;
    }
    const int64_t row = a.sel_rows[wid] - a.row_base;
    const float* qrow = a.q + size_t(q) * a.ld;
    float acc = 0.0f;
    if (DT == 1) {
        const uint4* x = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(a.rows) + size_t(row) * a.ld);
        for (int v = lane; v < a.ld / 8; v += 32) {
            const uint4 d = __ldg(x + v);
            const float4 q0 = __ldg(reinterpret_cast<const float4*>(qrow) + 2 * v);
            const float4 q1 = __ldg(reinterpret_cast<const float4*>(qrow) + 2 * v + 1);
            acc = fmaf(bf16lo(d.x), q0.x, acc); acc = fmaf(bf16hi(d.x), q0.y, acc);
            acc = fmaf(bf16lo(d.y), q0.z, acc); acc = fmaf(bf16hi(d.y), q0.w, acc);
            acc = fmaf(bf16lo(d.z), q1.x, acc); acc = fmaf(bf16hi(d.z), q1.y, acc);
            acc = fmaf(bf16lo(d.w), q1.z, acc); acc = fmaf(bf16hi(d.w), q1.w, acc);
        }
    } else {
        const float4* x = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(a.rows) + size_t(row) * a.ld);
        for (int v = lane; v < a.ld / 4; v += 32) {
            const float4 d = __ldg(x + v);
            const float4 q0 = __ldg(reinterpret_cast<const float4*>(qrow) + v);
            acc = fmaf(d.x, q0.x, acc); acc = fmaf(d.y, q0.y, acc);
            acc = fmaf(d.z, q0.z, acc); acc = fmaf(d.w, q0.w, acc);
        }
    }
    acc = warp_sum(acc);
    if (lane == 0) {
        const float s = acc * a.inv_norm[row] * a.qinv[q];
        a.cos_out[wid] = (s == s) ? fminf(1.0f, fmaxf(0.0f, s)) : 1.0f;      // corpus_search.py:1569 (NaN -> 1.0)
    }
}

struct FinalizeArgs {
    const uint64_t* keys;        // [nq][kc] exact keys
    const float* cand_scores;    // [nq][kc] approximate scores, descending
    const int32_t* cand_counts;  // [nq]
    int nq, kc, k;
    float eps;                   // bound on |approx - exact|
    float* scores; int64_t* rows; int32_t* counts; int64_t row_base;   // [nq][k] results
    int* need_tail;
    int* fb_list; int* fb_count; // queries that need the exact scan
    uint32_t* gthr;              // [nq] admission bound handed to the exact rescan of a flagged query
};

constexpr int kFinalizeThreads = 128;

__global__ void __launch_bounds__(kFinalizeThreads) finalize_kernel(const FinalizeArgs a) {
    __shared__ uint64_t s[256];
    const int q = blockIdx.x, tid = threadIdx.x;
    const int n2 = next_pow2(a.kc < 2 ? 2 : a.kc);          // <= 256
    const int count = a.cand_counts[q];
    for (int i = tid; i < n2; i += kFinalizeThreads) s[i] = (i < count) ? a.keys[size_t(q) * a.kc + i] : 0ull;
    __syncthreads();
    block_sort_desc(s, n2);
    // keys can be 0 only past `count` (a nominated row always has a finite score)
    const int got = count < a.k ? count : a.k;
    for (int i = tid; i < a.k; i += kFinalizeThreads) {
        const size_t o = size_t(q) * a.k + i;
        if (i < got) {
            const uint64_t key = s[i];
            a.scores[o] = fminf(1.0f, fmaxf(-1.0f, key_score(key)));      // pgvector "keep in range"
            a.rows[o] = int64_t(key_row(key)) + a.row_base;
        } else {
            a.scores[o] = CUDART_NAN_F;
            a.rows[o] = -1;
        }
    }
    if (tid == 0) {
        a.counts[q] = got;
        if (got < a.k && a.need_tail) *a.need_tail = 1;
        // count < kc: no admission bound was ever applied, every passing row was nominated -> exact.
        if (count == a.kc) {
            const float a_min = a.cand_scores[size_t(q) * a.kc + count - 1];
            const float tau = key_score(s[a.k - 1]);
            if (!(a_min + a.eps < tau)) {
                a.fb_list[atomicAdd(a.fb_count, 1)] = q;
                // kc rows have approx >= a_min, hence exact >= a_min - eps: the exact rescan may skip anything below
                a.gthr[q] = f2ord(a_min - a.eps);
            }
        }
    }
}

}  // namespace mrag

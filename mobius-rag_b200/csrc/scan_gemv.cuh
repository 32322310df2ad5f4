// scan_gemv.cuh -- K1 (small query batches) + K3 (fused select), CUDA cores, HBM-bound.
//
// Replaces the sequential-scan plan of
//     ORDER BY embedding_vec <=> :q LIMIT :k      (corpus_search.py:1525-1536, vector_store.py:274-287)
// for 1..4 queries per pass.  One warp owns one row at a time (4 rows in flight): every lane
// streams 16-byte pieces of the row with ld.global.nc.L1::no_allocate, accumulates
// <q_b, x> for each query and |x|^2 in fp32 in the SAME pass (the corpus is stored
// un-normalised, like pgvector stores it), and a shuffle tree finishes the sums.
// Rows whose mask bit is clear are never loaded, so a filtered scan reads only passing rows.
//
// Select: each warp keeps, per query, a private candidate buffer of 2*kp keys in shared
// memory and a register threshold (the k-th best key it has seen).  A row is appended only
// if its key beats the threshold; a full buffer is bitonic-sorted by the warp and cut back
// to k.  At the end the block sorts all its warps' buffers and writes ONE sorted list of kp
// keys per (query, block) -- no N-sized score array ever exists.
#pragma once
#include "common.cuh"
#include "hybrid.cuh"

namespace mrag {

constexpr int kGemvThreads = 512;
constexpr int kGemvWarps = kGemvThreads / 32;
constexpr int kGemvRows = 4;      // rows in flight per warp
constexpr int kGemvVecs = 3;      // 16-byte vectors per lane per row per column block

struct ScanArgs {
    const void* rows;       // [n][ld] storage dtype, row-major
    int64_t n;              // rows in the shard
    int ld;                 // padded row length in elements (multiple of 64, zero padded)
    const uint32_t* mask;   // row bitmap: valid AND filter; bits >= n are zero
    const float* q;         // [*][ld] fp32 queries, zero padded
    const float* qinv;      // [*]  1/|q|  (+inf for a zero query)
    const uint64_t* ub;     // [*]  exclusive upper-bound key per query, or nullptr
    uint64_t* part;         // [*][P][kp] per-block sorted candidate lists
    int q0;                 // first query handled by this launch
    int nq;                 // queries handled by this launch (<= NQ)
    int k, kp, P;
    // exact-fallback mode (rescore.cuh): scan the queries qlist[0 .. *qcount) instead, NQ at a time,
    // one full pass per group; *qcount == 0 makes the launch a no-op.  q0 / nq are ignored.
    const int* qlist;
    const int* qcount;
    int qskip;              // ... starting at list entry qskip (the first entries were served by another kernel)
    int unit_log2;          // work unit = 32 >> unit_log2 consecutive rows (0: a whole bitmap word; 2: 8 rows, for selective
                            // document filters -- see the row loop)
    // hybrid mode (HYB = 1, hybrid.cuh): per-query row bitmaps hmask[query][hwords] replace `mask`, and the
    // candidate key is built from the rerank score of (row, query) instead of the cosine
    const uint32_t* hmask;
    int64_t hwords;
    const mrag_chunkfeat* feat;
    DtagOver dtag_over;
    const DevHyb* hyb;
    const uint32_t* doc_idx;
    const uint8_t* authority;
    const uint64_t* doc_jtags;
    int64_t n_jtag_docs;
};

inline size_t gemv_smem_bytes(int nq_tpl, int ld, int kp, bool hyb = false) {
    return size_t(nq_tpl) * ld * 4 + size_t(nq_tpl) * kGemvWarps * (2 * kp) * 8 + (hyb ? size_t(nq_tpl) * sizeof(DevHyb) : 0);
}

// element j (0..E-1) of 16-byte vector v lives at  qs[plane(j)][v][j%4]
template <int DT> struct VecTraits;
template <> struct VecTraits<0> { static constexpr int E = 4; };   // fp32: 4 elements / 16 B
template <> struct VecTraits<1> { static constexpr int E = 8; };   // bf16: 8 elements / 16 B

template <int DT, int NQ, int HYB = 0>
__global__ void __launch_bounds__(kGemvThreads, 1) scan_gemv_kernel(const ScanArgs a) {
    constexpr int E = VecTraits<DT>::E;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* qs = reinterpret_cast<float*>(smem_raw);
    uint64_t* bufs = reinterpret_cast<uint64_t*>(smem_raw + size_t(NQ) * a.ld * 4);
    DevHyb* hq = reinterpret_cast<DevHyb*>(smem_raw + size_t(NQ) * a.ld * 4 + size_t(NQ) * kGemvWarps * (2 * a.kp) * 8);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ld = a.ld;
    const int nvec = ld / E;               // 16-byte vectors per row
    const int cap = 2 * a.kp;

    const int total = a.qlist ? max(0, *a.qcount - a.qskip) : a.nq;
    for (int g0 = 0; g0 < total; g0 += NQ) {
        const int nq_here = min(NQ, total - g0);
        int qid[NQ];
#pragma unroll
        for (int qi = 0; qi < NQ; ++qi) qid[qi] = qi < nq_here ? (a.qlist ? a.qlist[a.qskip + g0 + qi] : a.q0 + g0 + qi) : 0;
        if (g0) __syncthreads();               // the previous group's buffers are still being read
        // ---- stage the queries: fp32, split in 4-float planes so every LDS.128 is conflict free
        for (int i = tid; i < NQ * ld; i += kGemvThreads) {
            int qi = i / ld, e = i - qi * ld;
            int src_q = 0;
#pragma unroll
            for (int j = 0; j < NQ; ++j) src_q = (qi == j) ? qid[j] : src_q;   // keeps qid[] in registers
            float v = (qi < nq_here) ? a.q[size_t(src_q) * ld + e] : 0.0f;
            int dst;
            if (DT == 1) {
                int vv = e >> 3, w = e & 7;
                dst = qi * ld + (w >> 2) * (ld >> 1) + vv * 4 + (w & 3);
            } else {
                dst = i;
            }
            qs[dst] = v;
        }
        if (HYB) {
            const uint32_t* src = reinterpret_cast<const uint32_t*>(a.hyb);
            uint32_t* dst = reinterpret_cast<uint32_t*>(hq);
            constexpr int W32 = sizeof(DevHyb) / 4;
            for (int i = tid; i < nq_here * W32; i += kGemvThreads) {
                const int qi = i / W32;
                int src_q = 0;
#pragma unroll
                for (int j = 0; j < NQ; ++j) src_q = (qi == j) ? qid[j] : src_q;
                dst[i] = src[size_t(src_q) * W32 + (i - qi * W32)];
            }
        }
        uint64_t thr[NQ], ubk[NQ];
        int cnt[NQ];
        float qinv[NQ];
#pragma unroll
        for (int qi = 0; qi < NQ; ++qi) {
            thr[qi] = 0; cnt[qi] = 0;
            ubk[qi] = (a.ub && qi < nq_here) ? a.ub[qid[qi]] : ~0ull;
            qinv[qi] = (qi < nq_here) ? a.qinv[qid[qi]] : 0.0f;
        }
        __syncthreads();

        const int64_t nwords = (a.n + 31) >> 5;
        const int64_t W = int64_t(gridDim.x) * kGemvWarps;
        const char* base = reinterpret_cast<const char*>(a.rows);
        const size_t row_bytes = size_t(ld) * (DT == 1 ? 2 : 4);

        // row bitmap word w: the WHERE mask, or (hybrid) the union of the group's per-query bitmaps
        auto load_mask = [&](int64_t ww) -> uint32_t {
            if (ww >= nwords) return 0u;
            if (!HYB) return __ldg(a.mask + ww);
            uint32_t u = 0u;
#pragma unroll
            for (int qi = 0; qi < NQ; ++qi)
                if (qi < nq_here) u |= __ldg(a.hmask + size_t(qid[qi]) * a.hwords + ww);
            return u;
        };
        // work unit = a bitmap word (32 rows) or a piece of it, dealt round-robin to the warps of the grid.  A document's
        // chunks are contiguous rows, so with whole words a selective document pool lands on a few warps and the rest of
        // the GPU idles (r1k launch list: 65 us for 3200 rows); 8-row units spread them 4x wider.  Dense and doc-aligned
        // masks keep whole words (measured 10 % faster there).
        const int ul = a.unit_log2;
        const int unit_rows = 32 >> ul;
        const uint32_t unit_bits = unit_rows == 32 ? 0xFFFFFFFFu : ((1u << unit_rows) - 1u);
        const int64_t nunits = nwords << ul;
        auto load_unit = [&](int64_t uu) -> uint32_t {
            return uu < nunits ? (load_mask(uu >> ul) & (unit_bits << (unit_rows * int(uu & ((1 << ul) - 1))))) : 0u;
        };
        int64_t u = int64_t(blockIdx.x) * kGemvWarps + warp;
        uint32_t m_next = load_unit(u);
        for (; u < nunits; u += W) {
            const int64_t w = u >> ul;
            uint32_t m = m_next;
            m_next = load_unit(u + W);
            uint32_t mq[NQ];                       // hybrid: this word of every query's own bitmap
#pragma unroll
            for (int qi = 0; qi < NQ; ++qi) mq[qi] = (HYB && qi < nq_here && m) ? __ldg(a.hmask + size_t(qid[qi]) * a.hwords + w) : 0u;
            while (m) {
                uint32_t r[kGemvRows];
                int nr = 0;
#pragma unroll
                for (int j = 0; j < kGemvRows; ++j) {
                    if (m) { int b = __ffs(m) - 1; m &= m - 1; r[j] = uint32_t(w * 32 + b); ++nr; }
                    else r[j] = r[0];
                }
                float acc[kGemvRows][NQ + 1];
#pragma unroll
                for (int j = 0; j < kGemvRows; ++j)
#pragma unroll
                    for (int x = 0; x <= NQ; ++x) acc[j][x] = 0.0f;

                for (int v0 = 0; v0 < nvec; v0 += kWarp * kGemvVecs) {
                    uint4 d[kGemvRows][kGemvVecs];
#pragma unroll
                    for (int c = 0; c < kGemvVecs; ++c) {
                        int v = v0 + c * kWarp + lane;
                        bool ok = v < nvec;
#pragma unroll
                        for (int j = 0; j < kGemvRows; ++j) {
                            if (ok && j < nr) d[j][c] = ldg_stream(base + size_t(r[j]) * row_bytes + size_t(v) * 16);
                            else d[j][c] = make_uint4(0, 0, 0, 0);
                        }
                    }
#pragma unroll
                    for (int c = 0; c < kGemvVecs; ++c) {
                        int v = v0 + c * kWarp + lane;
                        if (v < nvec) {
                            float x[kGemvRows][E];
#pragma unroll
                            for (int j = 0; j < kGemvRows; ++j) {
                                if (DT == 1) {
                                    x[j][0] = bf16lo(d[j][c].x); x[j][1] = bf16hi(d[j][c].x);
                                    x[j][2] = bf16lo(d[j][c].y); x[j][3] = bf16hi(d[j][c].y);
                                    x[j][4 % E] = bf16lo(d[j][c].z); x[j][5 % E] = bf16hi(d[j][c].z);
                                    x[j][6 % E] = bf16lo(d[j][c].w); x[j][7 % E] = bf16hi(d[j][c].w);
                                } else {
                                    x[j][0] = __uint_as_float(d[j][c].x); x[j][1] = __uint_as_float(d[j][c].y);
                                    x[j][2] = __uint_as_float(d[j][c].z); x[j][3] = __uint_as_float(d[j][c].w);
                                }
#pragma unroll
                                for (int e = 0; e < E; ++e) acc[j][NQ] = fmaf(x[j][e], x[j][e], acc[j][NQ]);
                            }
#pragma unroll
                            for (int qi = 0; qi < NQ; ++qi) {
                                float qv[E];
                                const float4 q0v = *reinterpret_cast<const float4*>(qs + qi * ld + v * 4);
                                qv[0] = q0v.x; qv[1] = q0v.y; qv[2] = q0v.z; qv[3] = q0v.w;
                                if (DT == 1) {
                                    const float4 q1v = *reinterpret_cast<const float4*>(qs + qi * ld + (ld >> 1) + v * 4);
                                    qv[4 % E] = q1v.x; qv[5 % E] = q1v.y; qv[6 % E] = q1v.z; qv[7 % E] = q1v.w;
                                }
#pragma unroll
                                for (int j = 0; j < kGemvRows; ++j)
#pragma unroll
                                    for (int e = 0; e < E; ++e) acc[j][qi] = fmaf(x[j][e], qv[e], acc[j][qi]);
                            }
                        }
                    }
                }
                // ---- finish the sums: every lane ends up with every total
#pragma unroll
                for (int j = 0; j < kGemvRows; ++j)
#pragma unroll
                    for (int x = 0; x <= NQ; ++x) acc[j][x] = warp_sum(acc[j][x]);

                // ---- normalise and select (warp uniform)
#pragma unroll
                for (int j = 0; j < kGemvRows; ++j) {
                    if (j >= nr) break;
                    const float nx = acc[j][NQ];
                    if (!HYB && !(nx > 0.0f)) continue;     // zero-norm row: similarity is NaN (NaN tail pass)
                    const float inv = rsqrtf(nx);
                    // hybrid: lane L scores query L of the group (the queries' rerank scores in parallel, not 4x in
                    // lock step), then the scores travel by shuffle into the warp-uniform select below
                    float hyb_s = CUDART_NAN_F;
                    if (HYB) {
                        float dotL = 0.0f, qinvL = 0.0f;
                        uint32_t mqL = 0u;
#pragma unroll
                        for (int qi = 0; qi < NQ; ++qi) {
                            dotL = (lane == qi) ? acc[j][qi] : dotL;
                            qinvL = (lane == qi) ? qinv[qi] : qinvL;
                            mqL = (lane == qi) ? mq[qi] : mqL;
                        }
                        if (lane < nq_here && ((mqL >> (r[j] & 31u)) & 1u)) {      // passes the filter and this query's coverage floor
                            const mrag_chunkfeat f = a.feat[r[j]];
                            const uint32_t auth_code = a.authority[r[j]];
                            const uint32_t d = a.doc_idx[r[j]];
                            const uint64_t* jt = (a.doc_jtags && int64_t(d) < a.n_jtag_docs) ? a.doc_jtags + size_t(d) * MRAG_JTAG_WORDS : nullptr;
                            const float cs = (nx > 0.0f) ? dotL * inv * qinvL : CUDART_NAN_F;
                            // a NaN similarity reports 1.0 (max(0.0, min(1.0, nan)) in corpus_search.py:1569)
                            const float c01 = (cs == cs) ? cs : 1.0f;
                            hyb_s = hybrid_score(hq[lane], f, hybrid_eval(hq[lane], f, jt, r[j], a.dtag_over), c01, auth_code);
                        }
                    }
#pragma unroll
                    for (int qi = 0; qi < NQ; ++qi) {
                        if (qi >= nq_here) break;
                        float s = (nx > 0.0f) ? acc[j][qi] * inv * qinv[qi] : CUDART_NAN_F;
                        if (HYB) s = __shfl_sync(kFull, hyb_s, qi);           // NaN: row not eligible for this query
                        if (!(s == s)) continue;            // zero-norm query
                        const uint64_t key = make_key(s, r[j]);
                        if (key > thr[qi] && key < ubk[qi]) {
                            uint64_t* b = bufs + size_t(qi * kGemvWarps + warp) * cap;
                            if (lane == 0) b[cnt[qi]] = key;
                            if (++cnt[qi] == cap) {
                                __syncwarp();
                                warp_sort_desc(b, cap, lane);
                                cnt[qi] = a.k;
                                thr[qi] = b[a.k - 1];
                            }
                        }
                    }
                }
            }
        }

        // ---- block merge: clear the unused tail of every warp buffer, sort the block's buffers together
#pragma unroll
        for (int qi = 0; qi < NQ; ++qi) {
            uint64_t* b = bufs + size_t(qi * kGemvWarps + warp) * cap;
            for (int i = cnt[qi] + lane; i < cap; i += kWarp) b[i] = 0;
        }
        __syncthreads();
        for (int qi = 0; qi < nq_here; ++qi) {
            uint64_t* b = bufs + size_t(qi) * kGemvWarps * cap;
            block_sort_desc(b, kGemvWarps * cap);
            uint64_t* out = a.part + (size_t(qid[qi]) * a.P + blockIdx.x) * a.kp;
            for (int i = tid; i < a.kp; i += kGemvThreads) out[i] = (i < a.k) ? b[i] : 0ull;
        }
    }
}

}  // namespace mrag

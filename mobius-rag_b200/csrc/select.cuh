// select.cuh -- K3 block-level merge / finalize, NaN tail, K4 cross-shard k-way merge.
//
// Together with the per-producer lists written by the scan kernels these implement
//     ORDER BY embedding_vec <=> :q LIMIT :k       (vector_store.py:284-285, corpus_search.py:1534-1535)
// with Postgres float8 ordering (NaN last) and ties broken by ascending row.
#pragma once
#include "common.cuh"
#include <math_constants.h>

namespace mrag {

constexpr int kMergeThreads = 512;
constexpr int kMergeSlots = 4096;   // u64 keys held in shared memory by one merge block

struct MergeArgs {
    const uint64_t* part;   // [nq][P][kp] sorted (desc) candidate lists, 0 = empty
    int P, kp;
    int nq;
    int k;                  // results wanted from this round (<= MRAG_FUSED_K)
    int lk;                 // keys a producer list holds at most; 0 means k.  With lk < k (sampling pass
                            // that keeps only each CTA's top 16) no per-list prefilter bound exists.
    int k_total;            // row pitch of the outputs
    int k_off;              // first output slot of this round
    float* scores;          // [nq][k_total]
    int64_t* rows;          // [nq][k_total]
    int32_t* counts;        // [nq]
    int64_t row_base;
    uint64_t* ub_out;       // [nq] last key of this round (next round's exclusive bound), may be null
    int* need_tail;         // set to 1 if some query is still short of k_total after this round
    uint32_t* gthr_out;     // if set: ONLY write orderable(score of the k-th best) per query (0 if fewer
                            // than k candidates) -- the admission bound the sampling pass hands to the scan
    // first level of a two-level merge (wide producer sets x large k): gridDim.y groups of `Pg` lists
    // each; block (q, g) writes the sorted top-k keys of its group to part_out[q][g][kp] and nothing else
    uint64_t* part_out;
    int Pg;
    // exact-fallback mode: block b serves query qlist[q_lo + b] and exits if q_lo + b >= min(*qcount, q_hi)
    const int* qlist;
    const int* qcount;
    int q_lo, q_hi;         // q_hi == 0 means no upper limit
    int no_clamp;           // hybrid rerank scores are not cosines: do not clamp them to [-1, 1]
    // NaN tail folded into the last merge of a search (tail_mask != nullptr): a query that ends short of k_total walks the
    // mask for NaN rows in this block instead of in a nan_tail_kernel launch; need_tail is then left alone
    const uint32_t* tail_mask; const float* tail_inv_norm; const float* tail_qinv; int64_t tail_n;
    const int64_t* seg_off; // if set: query q's keys are the UNSORTED segment part[seg_off[q] .. seg_off[q + 1]) (P, kp ignored)
    const uint32_t* thr_in; // [nq] orderable(score) that k rows of the query are known to reach (the scan's cross-CTA bound), 0 =
                            // unknown; or nullptr.  Replaces the pass over the lists' k-th entries: one dependent round trip less.
};

// Block-cooperative top-k of n UNIQUE non-zero keys in s[0..n): afterwards s[0..min(n,k)) holds the k largest,
// descending; returns min(n, k).  A radix select (8 bits per pass from the top) finds the k-th largest key, the
// k keys >= it are gathered and sorted -- ~10x cheaper than sorting thousands of keys to keep a few dozen.
// k <= 256.  hist: 256 counters, small: 256 keys, misc: 4 ints (all shared memory).  All threads must call.
MRAG_DEVINL int block_topk(uint64_t* s, int n, int k, unsigned* hist, uint64_t* small, int* misc) {
    const int tid = threadIdx.x, nt = blockDim.x;
    if (n > k) {
        uint64_t prefix = 0;
        int krem = k;
        for (int pass = 0; pass < 8; ++pass) {
            const int shift = 56 - 8 * pass;
            const uint64_t hi_mask = pass ? (~0ull << (shift + 8)) : 0ull;
            if (tid < 256) hist[tid] = 0;
            __syncthreads();
            for (int i = tid; i < n; i += nt) {
                const uint64_t key = s[i];
                if ((key & hi_mask) == prefix) atomicAdd(&hist[unsigned(key >> shift) & 255u], 1u);
            }
            __syncthreads();
            if (tid < 32) {
                // lane l owns bins 255-8l .. 248-8l (descending); find the bin holding the krem-th largest
                unsigned c[8], sum = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) { c[j] = hist[255 - 8 * tid - j]; sum += c[j]; }
                unsigned incl = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned v = __shfl_up_sync(kFull, incl, o);
                    if (tid >= o) incl += v;
                }
                unsigned before = incl - sum;
                if (before < unsigned(krem) && unsigned(krem) <= incl) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (before < unsigned(krem) && unsigned(krem) <= before + c[j]) { misc[0] = 255 - 8 * tid - j; misc[1] = krem - int(before); }
                        before += c[j];
                    }
                }
            }
            __syncthreads();
            prefix |= uint64_t(unsigned(misc[0])) << shift;
            krem = misc[1];
        }
        if (tid == 0) misc[2] = 0;
        __syncthreads();
        for (int i = tid; i < n; i += nt) {
            const uint64_t key = s[i];
            if (key >= prefix) small[atomicAdd(&misc[2], 1)] = key;       // exactly k keys (they are unique)
        }
        __syncthreads();
        for (int i = tid; i < k; i += nt) s[i] = small[i];
        n = k;
        __syncthreads();
    }
    const int n2 = next_pow2(n < 2 ? 2 : n);
    for (int i = n + tid; i < n2; i += nt) s[i] = 0;
    __syncthreads();
    block_sort_desc(s, n2);
    return n;
}

// NaN tail.  Postgres sorts a NaN distance after every number, so rows whose similarity is NaN
// (zero-norm row, or every row when the query itself has zero norm) are returned only when
// fewer than k finite rows pass the filter.  One block per query walks the mask in row order.
// inv_norm[r] == +inf marks a stored zero-norm row.
struct TailArgs {
    const float* inv_norm; const uint32_t* mask; int64_t n;
    const float* qinv; int nq; int k_total;
    float* scores; int64_t* rows; int32_t* counts; int64_t row_base;
    const int* need_tail;
};

// the walk for query q, by a whole block of <= 512 threads (all must call); `have` = results the query holds so far
MRAG_DEVINL void nan_tail_block(const TailArgs& a, int q, int have, int* s_scan /*[blockDim]*/, int* s_count /*[1]*/) {
    const int tid = threadIdx.x, nt = blockDim.x;
    if (tid == 0) *s_count = have;
    __syncthreads();
    if (*s_count >= a.k_total) return;
    const bool all_nan = isinf(a.qinv[q]);
    // a zero query makes EVERY similarity NaN; the scan kernels then selected nothing for it
    const int64_t nwords = (a.n + 31) >> 5;
    for (int64_t w0 = 0; w0 < nwords; w0 += nt) {
        int64_t w = w0 + tid;
        uint32_t m = (w < nwords) ? a.mask[w] : 0u;
        if (m && !all_nan) {
            uint32_t keep = 0;
            for (uint32_t mm = m; mm; mm &= mm - 1) {
                int b = __ffs(mm) - 1;
                if (isinf(a.inv_norm[w * 32 + b])) keep |= 1u << b;
            }
            m = keep;
        }
        int c = __popc(m);
        s_scan[tid] = c;
        __syncthreads();
        for (int o = 1; o < nt; o <<= 1) {          // inclusive scan
            int v = (tid >= o) ? s_scan[tid - o] : 0;
            __syncthreads();
            s_scan[tid] += v;
            __syncthreads();
        }
        int pos = *s_count + s_scan[tid] - c;
        for (; m && pos < a.k_total; m &= m - 1, ++pos) {
            int b = __ffs(m) - 1;
            size_t o = size_t(q) * a.k_total + pos;
            a.scores[o] = CUDART_NAN_F;
            a.rows[o] = w * 32 + b + a.row_base;
        }
        __syncthreads();
        if (tid == nt - 1) *s_count = min(a.k_total, *s_count + s_scan[nt - 1]);
        __syncthreads();
        if (*s_count >= a.k_total) break;
    }
    if (tid == 0) a.counts[q] = *s_count;
}

__global__ void __launch_bounds__(256, 1) nan_tail_kernel(const TailArgs a) {
    if (*a.need_tail == 0) return;
    __shared__ int s_scan[256];
    __shared__ int s_count;
    nan_tail_block(a, blockIdx.x, a.counts[blockIdx.x], s_scan, &s_count);
}

// One block per query.  Keys below T = max_p(list_p[k-1]) cannot be in the global top-k (list p
// alone already holds k keys >= T), so only the survivors are gathered and sorted.
__global__ void __launch_bounds__(kMergeThreads, 1) merge_kernel(const MergeArgs a) {
    __shared__ uint64_t s[kMergeSlots];
    __shared__ uint64_t s_small[256];
    __shared__ unsigned s_hist[256];
    __shared__ int s_misc[4];
    __shared__ unsigned long long s_T;
    __shared__ int s_cnt;
    const int tid = threadIdx.x;
    if (a.qlist) {
        const int lim = a.q_hi > 0 ? min(*a.qcount, a.q_hi) : *a.qcount;
        if (a.q_lo + int(blockIdx.x) >= lim) return;
    }
    const int q = a.qlist ? a.qlist[a.q_lo + blockIdx.x] : int(blockIdx.x);
    const int p_lo = a.part_out ? int(blockIdx.y) * a.Pg : 0;
    const int P = a.part_out ? min(a.Pg, a.P - p_lo) : a.P;
    // segments: a first-level block (part_out set) takes the blockIdx.y-th slice of the query's segment
    int64_t seg_lo = 0, seg_hi = 0;
    if (a.seg_off) {
        seg_lo = a.seg_off[q]; seg_hi = a.seg_off[q + 1];
        if (a.part_out) {
            const int64_t per = (seg_hi - seg_lo + gridDim.y - 1) / gridDim.y;
            seg_lo = min(seg_hi, seg_lo + int64_t(blockIdx.y) * per);
            seg_hi = min(seg_hi, seg_lo + per);
        }
    }
    const uint64_t* base = a.seg_off ? a.part + seg_lo : a.part + (size_t(q) * a.P + p_lo) * a.kp;
    if (tid == 0) { s_T = 0ull; s_cnt = 0; }
    __syncthreads();
    uint64_t t = 0;
    if (a.thr_in) {
        if (tid == 0) t = uint64_t(a.thr_in[q]) << 32;          // every key of a row scoring >= the bound compares >= t
    } else if (!a.seg_off && (a.lk == 0 || a.lk >= a.k)) {
        for (int p = tid; p < P; p += kMergeThreads) {
            uint64_t v = base[size_t(p) * a.kp + (a.k - 1)];
            t = v > t ? v : t;
        }
    }
    if (t) atomicMax(&s_T, (unsigned long long)t);
    __syncthreads();
    const uint64_t T = s_T;
    const int total = a.seg_off ? int(seg_hi - seg_lo) : P * a.kp;
    int done = 0;
    while (done < total) {
        int kept = s_cnt;                       // uniform: read between two barriers
        __syncthreads();
        if (kept > kMergeSlots / 2) {           // make room: keep the best k of what we have
            kept = block_topk(s, kept, a.k, s_hist, s_small, s_misc);
            if (tid == 0) s_cnt = kept;
            __syncthreads();
        }
        int room = kMergeSlots - kept;
        int take = total - done < room ? total - done : room;
        for (int i = tid; i < take; i += kMergeThreads) {
            uint64_t v = base[done + i];
            if (v != 0 && v >= T) { int pos = atomicAdd(&s_cnt, 1); s[pos] = v; }
        }
        done += take;
        __syncthreads();
    }
    int got;
    if (s_cnt <= 32) {
        // the usual case with a tight bound: a handful of survivors -- one warp ranks them with shuffles
        got = s_cnt < a.k ? s_cnt : a.k;
        __syncthreads();
        if (tid < 32) {
            const int n = s_cnt;
            const uint64_t mine = tid < n ? s[tid] : 0ull;
            int rank = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const uint64_t other = __shfl_sync(kFull, mine, j);
                rank += other > mine ? 1 : 0;                   // keys are unique
            }
            __syncwarp();
            if (tid < n) s[rank] = mine;
        }
        __syncthreads();
    } else {
        got = block_topk(s, s_cnt, a.k, s_hist, s_small, s_misc);
    }
    if (a.part_out) {
        uint64_t* out = a.part_out + (size_t(q) * gridDim.y + blockIdx.y) * a.kp;
        for (int i = tid; i < a.kp; i += kMergeThreads) out[i] = (i < got) ? s[i] : 0ull;
        return;
    }
    if (a.gthr_out) {
        if (tid == 0) a.gthr_out[q] = (got == a.k) ? uint32_t(s[a.k - 1] >> 32) : 0u;
        return;
    }
    // round 0 starts at slot 0 and pads the whole row; later rounds continue at counts[q]
    const int prev = (a.k_off == 0) ? 0 : a.counts[q];
    __syncthreads();
    const int hi = (a.k_off == 0) ? a.k_total : got;
    for (int i = tid; i < hi; i += kMergeThreads) {
        size_t o = size_t(q) * a.k_total + prev + i;
        if (i < got) {
            uint64_t key = s[i];
            float sc = key_score(key);
            if (!a.no_clamp) sc = fminf(1.0f, fmaxf(-1.0f, sc));          // pgvector "keep in range"
            a.scores[o] = sc;
            a.rows[o] = int64_t(key_row(key)) + a.row_base;
        } else {
            a.scores[o] = CUDART_NAN_F;
            a.rows[o] = -1;
        }
    }
    if (tid == 0) {
        a.counts[q] = prev + got;
        if (a.ub_out) a.ub_out[q] = (got == a.k) ? s[a.k - 1] : 0ull;
        if (prev + got < a.k_total && a.need_tail && !a.tail_mask) *a.need_tail = 1;
    }
    if (a.tail_mask && prev + got < a.k_total) {          // block-uniform
        __syncthreads();                                    // s[] is free from here on
        TailArgs t;
        t.inv_norm = a.tail_inv_norm; t.mask = a.tail_mask; t.n = a.tail_n; t.qinv = a.tail_qinv; t.nq = a.nq; t.k_total = a.k_total;
        t.scores = a.scores; t.rows = a.rows; t.counts = a.counts; t.row_base = a.row_base; t.need_tail = nullptr;
        nan_tail_block(t, q, prev + got, reinterpret_cast<int*>(s), &s_cnt);
    }
}

// K4: k-way merge of per-shard results after the allgather.  Entries are (score, global row);
// order = score DESC, NaN last, row ASC.  Dynamic shared memory: next_pow2(n_lists*k) * 12 bytes.
struct XMergeArgs {
    int n_lists, nq, k;
    const float* scores_in; const int64_t* rows_in; const int32_t* counts_in;
    int64_t stride_scores, stride_rows, stride_counts;      // elements between consecutive lists
    float* scores_out; int64_t* rows_out; int32_t* counts_out;
};

constexpr int kXMergeMaxSlots = 16384;

MRAG_DEVINL bool xm_before(uint32_t ca, int64_t ra, uint32_t cb, int64_t rb) {
    // class: 0 = empty, 1 = NaN, >= 2 = orderable(score) (a finite score never maps below 2)
    if (ca != cb) return ca > cb;
    return ra < rb;
}

__global__ void __launch_bounds__(kMergeThreads, 1) xmerge_kernel(const XMergeArgs a) {
    extern __shared__ __align__(16) unsigned char xm_smem[];
    const int q = blockIdx.x, tid = threadIdx.x;
    const int total = a.n_lists * a.k;
    const int n2 = next_pow2(total < 2 ? 2 : total);
    int64_t* sr = reinterpret_cast<int64_t*>(xm_smem);
    uint32_t* sc = reinterpret_cast<uint32_t*>(xm_smem + size_t(n2) * 8);
    for (int i = tid; i < n2; i += kMergeThreads) {
        uint32_t c = 0; int64_t r = INT64_MAX;
        if (i < total) {
            int l = i / a.k, j = i - l * a.k;
            if (j < a.counts_in[l * a.stride_counts + q]) {
                size_t o = size_t(q) * a.k + j;
                float s = a.scores_in[l * a.stride_scores + o];
                r = a.rows_in[l * a.stride_rows + o];
                c = (s == s) ? f2ord(s) : 1u;
                if (c < 2u) c = (s == s) ? 2u : 1u;
            }
        }
        sc[i] = c; sr[i] = r;
    }
    __syncthreads();
    for (int kk = 2; kk <= n2; kk <<= 1) {
        for (int j = kk >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < n2; i += kMergeThreads) {
                int p = i ^ j;
                if (p > i) {
                    uint32_t ci = sc[i], cp = sc[p]; int64_t ri = sr[i], rp = sr[p];
                    bool first_half = (i & kk) == 0;
                    bool swap = first_half ? xm_before(cp, rp, ci, ri) : xm_before(ci, ri, cp, rp);
                    if (swap) { sc[i] = cp; sc[p] = ci; sr[i] = rp; sr[p] = ri; }
                }
            }
            __syncthreads();
        }
    }
    int got = 0;
    for (int l = 0; l < a.n_lists; ++l) got += a.counts_in[l * a.stride_counts + q];
    got = got < a.k ? got : a.k;
    for (int i = tid; i < a.k; i += kMergeThreads) {
        size_t o = size_t(q) * a.k + i;
        if (i < got) {
            uint32_t c = sc[i];
            a.scores_out[o] = (c == 1u) ? CUDART_NAN_F : ord2f(c);
            a.rows_out[o] = sr[i];
        } else {
            a.scores_out[o] = CUDART_NAN_F;
            a.rows_out[o] = -1;
        }
    }
    if (tid == 0) a.counts_out[q] = got;
}

// K4 fused with the exchange: instead of an allgather followed by the merge, every rank's block q STORES its k
// results for query q straight into each peer's gather buffer over NVLink (peer-mapped memory), raises a flag per
// (peer, query), waits for the peers' flags for the same query and merges -- one kernel, no collective call.
// Slot layout (bytes): rows i64[nq*k] | scores f32[nq*k] | counts i32[nq], slot r of a gather area = rank r's results.
// flags: [world][nq] u32 per rank, monotonically increasing epochs; two gather areas alternate by epoch parity (a rank
// can be at most one step ahead of a peer, because finishing a step needs that peer's flag for it).
struct XchgArgs {
    int world, rank, nq, k;
    unsigned char* bufs[8];        // peer-mapped base of every rank's buffer (bufs[rank] = own)
    int64_t area_bytes;            // one gather area = world * slot_bytes
    int64_t slot_bytes, scores_off, counts_off;
    int64_t flags_off;             // flags start (after the two gather areas)
    uint32_t epoch;
    float* scores_out; int64_t* rows_out; int32_t* counts_out;
};

__global__ void __launch_bounds__(kMergeThreads, 1) xchg_merge_kernel(const XchgArgs a) {
    extern __shared__ __align__(16) unsigned char xm_smem[];
    const int q = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;      // 128 threads for small merges: the block then fits beside a resident scan CTA
    const int64_t area = int64_t(a.epoch & 1u) * a.area_bytes;
    unsigned char* own_slot = a.bufs[a.rank] + area + int64_t(a.rank) * a.slot_bytes;
    // 1. my results for query q -> every peer's gather area (slot = my rank)
    const int64_t* my_rows = reinterpret_cast<const int64_t*>(own_slot) + size_t(q) * a.k;
    const float* my_scores = reinterpret_cast<const float*>(own_slot + a.scores_off) + size_t(q) * a.k;
    const int32_t my_count = *(reinterpret_cast<const int32_t*>(own_slot + a.counts_off) + q);
    for (int p = 0; p < a.world; ++p) {
        if (p == a.rank) continue;
        unsigned char* dst = a.bufs[p] + area + int64_t(a.rank) * a.slot_bytes;
        for (int i = tid; i < a.k; i += nt) {
            reinterpret_cast<int64_t*>(dst)[size_t(q) * a.k + i] = my_rows[i];
            reinterpret_cast<float*>(dst + a.scores_off)[size_t(q) * a.k + i] = my_scores[i];
        }
        if (tid == 0) reinterpret_cast<int32_t*>(dst + a.counts_off)[q] = my_count;
    }
    __threadfence_system();
    __syncthreads();
    // 2. raise my flag for query q at every peer; 3. wait for every peer's flag for query q
    if (tid < a.world && tid != a.rank) {
        uint32_t* peer_flag = reinterpret_cast<uint32_t*>(a.bufs[tid] + a.flags_off) + size_t(a.rank) * a.nq + q;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peer_flag), "r"(a.epoch) : "memory");
        const uint32_t* mine = reinterpret_cast<const uint32_t*>(a.bufs[a.rank] + a.flags_off) + size_t(tid) * a.nq + q;
        const long long t0 = clock64();
        for (;;) {
            uint32_t v;
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
            if (int32_t(v - a.epoch) >= 0) break;
            if (clock64() - t0 > 20000000000ll) __trap();      // ~10 s: a peer died
            __nanosleep(200);
        }
    }
    __syncthreads();
    // 4. k-way merge of the world lists of query q (same order as xmerge_kernel)
    const unsigned char* g = a.bufs[a.rank] + area;
    const int total = a.world * a.k;
    const int n2 = next_pow2(total < 2 ? 2 : total);
    int64_t* sr = reinterpret_cast<int64_t*>(xm_smem);
    uint32_t* sc = reinterpret_cast<uint32_t*>(xm_smem + size_t(n2) * 8);
    int got = 0;
    for (int l = 0; l < a.world; ++l) got += reinterpret_cast<const int32_t*>(g + int64_t(l) * a.slot_bytes + a.counts_off)[q];
    got = got < a.k ? got : a.k;
    for (int i = tid; i < n2; i += nt) {
        uint32_t c = 0; int64_t r = INT64_MAX;
        if (i < total) {
            const int l = i / a.k, j = i - l * a.k;
            const unsigned char* slot = g + int64_t(l) * a.slot_bytes;
            if (j < reinterpret_cast<const int32_t*>(slot + a.counts_off)[q]) {
                const float s = reinterpret_cast<const float*>(slot + a.scores_off)[size_t(q) * a.k + j];
                r = reinterpret_cast<const int64_t*>(slot)[size_t(q) * a.k + j];
                c = (s == s) ? f2ord(s) : 1u;
                if (c < 2u) c = (s == s) ? 2u : 1u;
            }
        }
        sc[i] = c; sr[i] = r;
    }
    __syncthreads();
    for (int kk = 2; kk <= n2; kk <<= 1) {
        for (int j = kk >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < n2; i += nt) {
                const int p = i ^ j;
                if (p > i) {
                    const uint32_t ci = sc[i], cp = sc[p]; const int64_t ri = sr[i], rp = sr[p];
                    const bool first_half = (i & kk) == 0;
                    const bool swap = first_half ? xm_before(cp, rp, ci, ri) : xm_before(ci, ri, cp, rp);
                    if (swap) { sc[i] = cp; sc[p] = ci; sr[i] = rp; sr[p] = ri; }
                }
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < a.k; i += nt) {
        const size_t o = size_t(q) * a.k + i;
        if (i < got) {
            const uint32_t c = sc[i];
            a.scores_out[o] = (c == 1u) ? CUDART_NAN_F : ord2f(c);
            a.rows_out[o] = sr[i];
        } else {
            a.scores_out[o] = CUDART_NAN_F;
            a.rows_out[o] = -1;
        }
    }
    if (tid == 0) a.counts_out[q] = got;
}

}  // namespace mrag

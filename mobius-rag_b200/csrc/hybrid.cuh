// hybrid.cuh -- K5: the rerank score of `_rerank` (app/services/corpus_search.py:1909-2297) fused into the scan.
//
// The reference scores <= ~3k candidates per query in a Python loop (substring tests per phrase per chunk).
// Here every text-dependent test is a bit computed once per row at insert (mrag_chunkfeat), so the score of a
// (row, query) pair is a few dozen integer / fp32 instructions and can be taken over ALL rows:
//   hybrid_mask_kernel   the coverage floor (:2183-2247) AND the WHERE mask, per query -> one bitmap per query;
//                        with required phrases only the few rows that cover them all survive, and the scan
//                        (scan_gemv.cuh, HYB mode) loads nothing else.
//   hybrid_score()       sim' / authority / length / jpd / coverage mix and the chunk d-tag boost (:2020-2137),
//                        called from the scan's epilogue instead of the plain cosine.
#pragma once
#include "common.cuh"
#include "../../include/mrag.h"
#include <math_constants.h>

namespace mrag {

// number of patterns per category of _JPD_PATTERNS (corpus_search.py:233-309), in dictionary order, empty
// category ("other_important") left out
__constant__ float kJpdPatterns[MRAG_JPD_CATS] = {20.f, 13.f, 12.f, 13.f, 8.f, 16.f, 11.f, 6.f, 9.f, 4.f, 19.f};

struct DevHyb {
    mrag_hybrid_query q;
    float total_weight;          // sum of phrase weights, or 1 (:1986)
    float max_weight;            // w_sim + w_auth + w_len + w_jpd + w_cov (:2011)
    float qcat_sum;              // sum of qcat
    uint64_t need[MRAG_PHRASE_WORDS];   // dictionary bits of the phrases that can only be present through their bit
    uint32_t impossible;         // some phrase has neither a dictionary bit nor a j-code: coverage can never be complete
    uint32_t has_dcodes;         // some phrase carries a d: code (chunk d-tag boost / exemption possible)
    uint32_t src_restrict;       // source_type_any is not empty
};

// Per chunk of 32 queries: what phase 1 of hybrid_mask_kernel needs, transposed so that a thread (= row) settles all the
// chunk's queries with a handful of word operations.  Built on the host (fill_hyb_chunks in mrag.cu).
struct HybChunk {
    uint32_t qneed[MRAG_PHRASE_WORDS * 64];   // bit q: query q0 + q can only see dictionary phrase p through its bit (DevHyb::need)
    uint32_t cnt[5];                          // popcount(need) of every query, bit-sliced (slice b = bit b of the count)
    uint32_t live, phr, imp, src, contact, dcodes, lowfloor;   // queries that: exist / have phrases / are impossible /
                                              // restrict source_type / are contact queries / carry d: codes / have a floor < 1
    uint32_t ncodes;                          // distinct d: codes of the chunk's queries (0xFFFFFFFF: more than 64, no table)
    uint16_t code[64];
    uint32_t codemask[64];                    // bit q: query q0 + q carries code[i]
};

struct HybEval {
    float cov;
    bool dtag_match;
};

// chunk_d_tags keys beyond the four inline slots of mrag_chunkfeat: (row, code) pairs sorted by row.  `chunk_d_tags ? :key`
// (corpus_search.py:1637-1672) matches ANY key of the JSONB map, so a chunk with more than four keys keeps the rest here
// and carries MRAG_CF_DTAG_OVERFLOW; the table is only consulted for such rows.
struct DtagOver {
    const uint32_t* rows;
    const uint16_t* codes;
    int64_t n;
};

MRAG_DEVINL bool has_dtag(const mrag_chunkfeat& f, uint32_t code, int64_t row, const DtagOver& ov) {
    if (f.dtags[0] == code || f.dtags[1] == code || f.dtags[2] == code || f.dtags[3] == code) return true;
    if (!(f.flags & MRAG_CF_DTAG_OVERFLOW) || ov.n == 0) return false;
    int64_t lo = 0, hi = ov.n;                            // first pair of this row
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (int64_t(ov.rows[mid]) < row) lo = mid + 1; else hi = mid;
    }
    for (; lo < ov.n && int64_t(ov.rows[lo]) == row; ++lo)
        if (ov.codes[lo] == code) return true;
    return false;
}

// coverage and chunk d-tag match of one (row, query) pair.  jt: the document's j-tag words or nullptr.
MRAG_DEVINL HybEval hybrid_eval(const DevHyb& h, const mrag_chunkfeat& f, const uint64_t* jt, int64_t row, const DtagOver& ov) {
    HybEval e;
    e.cov = 0.0f;
    e.dtag_match = false;
    float acc = 0.0f;
    for (int i = 0; i < h.q.n_phrases; ++i) {
        bool present = false;
        const int jb = h.q.phrase_jbit[i];
        if (jb >= 0 && jt) present = (jt[jb >> 6] >> (jb & 63)) & 1ull;
        if (!present) {
            const int pb = h.q.phrase_bit[i];
            if (pb >= 0) present = (f.phrase_bits[pb >> 6] >> (pb & 63)) & 1ull;
        }
        if (present) acc += h.q.phrase_weight[i];          // same order as the reference's sum: complete coverage is exactly 1
        const uint32_t dc = h.q.phrase_dcode[i];
        if (dc != 0u) e.dtag_match = e.dtag_match || has_dtag(f, dc, row, ov);
    }
    if (h.q.n_phrases > 0) e.cov = acc / h.total_weight;
    return e;
}

MRAG_DEVINL bool hybrid_keep(const DevHyb& h, const mrag_chunkfeat& f, const HybEval& e) {
    if (h.q.n_phrases == 0) return true;
    if (!(e.cov < h.q.floor)) return true;
    if (f.flags & MRAG_CF_PROMOTED) return true;
    if (h.q.contact_query && (f.flags & MRAG_CF_CONTACT_VALUE)) return true;
    return e.dtag_match;
}

// sim = `_best_arm_sim` of the candidate (:1787-1814)
MRAG_DEVINL float hybrid_score_sim(const DevHyb& h, const mrag_chunkfeat& f, const HybEval& e, float sim, uint32_t auth_code) {
    const float auth = h.q.auth_score[auth_code < 31u ? auth_code : 31u];
    float jpd = 0.0f;
    if (h.q.w_jpd > 0.0f) {
        float num = 0.0f;
#pragma unroll
        for (int c = 0; c < MRAG_JPD_CATS; ++c) {
            const float hits = float(f.jpd_hits[c]);
            const float den = (f.flags & MRAG_CF_SHORT_TEXT) ? sqrtf(kJpdPatterns[c]) : kJpdPatterns[c];
            num += h.q.qcat[c] * fminf(1.0f, hits / den);
        }
        jpd = fminf(1.0f, num / h.qcat_sum);
    }
    const float raw = h.q.w_sim * sim + h.q.w_auth * auth + h.q.w_len * f.length_score + h.q.w_jpd * jpd + h.q.w_cov * e.cov;
    float score = h.max_weight > 0.0f ? raw / h.max_weight : raw;
    if (e.dtag_match) score *= h.q.boost;
    return score;
}

MRAG_DEVINL float hybrid_score(const DevHyb& h, const mrag_chunkfeat& f, const HybEval& e, float cos, uint32_t auth_code) {
    const float c01 = fminf(1.0f, fmaxf(0.0f, cos));                       // _vector_arm clamp (:1569)
    return hybrid_score_sim(h, f, e, fmaxf(0.0f, (c01 - 0.5f) * 2.0f), auth_code);     // vector arm of _best_arm_sim (:1808-1810)
}

// `_rerank` over a candidate LIST (the RRF output of several arms, corpus_search.py:3519-3622): one thread per candidate.
// Everything textual arrives as bits (mrag_candidate.feat, built by the host shim from the candidate's haystacks incl. any
// neighbour text), sim is `_best_arm_sim` over the arms that found it.  Writes the score, the coverage and whether the
// coverage floor keeps it; the per-(arm, source_type) decay and the sort are a host pass over <= ~600 numbers.
__global__ void __launch_bounds__(128) rerank_candidates_kernel(const mrag_candidate* __restrict__ cands, int64_t n, const DevHyb* __restrict__ hq,
                                                               const uint64_t* __restrict__ doc_jtags, int64_t n_jtag_docs,
                                                               float* __restrict__ scores, float* __restrict__ cov, uint8_t* __restrict__ keep) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const DevHyb& h = hq[0];
    const mrag_candidate c = cands[i];
    const uint64_t* jt = (doc_jtags && int64_t(c.doc_idx) < n_jtag_docs) ? doc_jtags + size_t(c.doc_idx) * MRAG_JTAG_WORDS : nullptr;
    HybEval e = hybrid_eval(h, c.feat, jt, -1, DtagOver{nullptr, nullptr, 0});
    e.dtag_match = e.dtag_match || c.dtag_match != 0;
    scores[i] = hybrid_score_sim(h, c.feat, e, c.sim, c.authority);
    cov[i] = e.cov;
    keep[i] = hybrid_keep(h, c.feat, e) ? 1 : 0;
}

// One thread per row.  Bit (q, r) = row r passes the WHERE mask and query q's coverage floor.  hmask: [nq][nwords].
// Phase 1 (thread = row): all 32 queries of a chunk at once, on words whose bit q belongs to query q (HybChunk): which
// queries find every dictionary bit they need in this row (bit-sliced count of the row's present bits against the
// queries' need counts), which keep it through an exemption that does not depend on the coverage (promoted, contact
// value of a contact query, an inline chunk d-tag that matches one of the query's d: codes) -- and which must run the
// phrase loop (weighted coverage with j-tag credit, overflow d-tags, floors below 1): those pairs are only MARKED.
// One ballot per query turns the per-row words into the per-query bitmap words.
// Phase 2 (lane = query): the warp takes its marked rows one at a time; the lanes that marked a row run hybrid_eval.
// History (10M rows x 22 queries): evaluation inline in a per-query loop 1.47 ms (nearly every warp holds a row that may
// be exempt and ran it with 2 lanes active, once per query); per-query bit tests + phase 2: 0.96-1.0 ms, ALU bound at
// 50-65 instructions per (warp, query); the same with lane = query in phase 1: 1.15 ms (shuffle latency).
__global__ void __launch_bounds__(256) hybrid_mask_kernel(const DevHyb* __restrict__ hq, const HybChunk* __restrict__ chunks, int nq,
                                                         const mrag_chunkfeat* __restrict__ feat,
                                                         const uint32_t* __restrict__ base_mask, const uint32_t* __restrict__ doc_idx,
                                                         const uint8_t* __restrict__ source_type,
                                                         const uint64_t* __restrict__ doc_jtags, int64_t n_jtag_docs, int64_t n,
                                                         uint32_t* __restrict__ hmask, int64_t nwords, const DtagOver ov) {
    const int64_t r = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool in_range = r < n;
    bool base = false;
    mrag_chunkfeat f;
    const uint64_t* jt = nullptr;
    uint32_t src = 0xFFu;
    if (in_range) {
        base = (base_mask[r >> 5] >> (r & 31)) & 1u;
        if (base) {
            f = feat[r];
            src = source_type[r];
            const uint32_t d = doc_idx[r];
            if (doc_jtags && int64_t(d) < n_jtag_docs) jt = doc_jtags + size_t(d) * MRAG_JTAG_WORDS;
        }
    }
    const bool anydtag = base && (f.dtags[0] | f.dtags[1] | f.dtags[2] | f.dtags[3]) != 0u;
    union FeatWords { mrag_chunkfeat f; uint32_t w[10]; };
    static_assert(sizeof(mrag_chunkfeat) == 40, "mrag_chunkfeat is broadcast as 10 words");
    for (int q0 = 0; q0 < nq; q0 += 32) {
        const int nqc = min(32, nq - q0);
        const HybChunk& c = chunks[q0 >> 5];
        uint32_t keepm = 0u, slow = 0u;          // bit i: query q0 + i keeps this row / needs the phrase loop for it
        if (base) {
            uint32_t livem = c.live;
            if (c.src) {                          // rare: some query restricts source_type
                for (uint32_t m = c.src; m; m &= m - 1) {
                    const int i = __ffs(m) - 1;
                    if (!((hq[q0 + i].q.source_type_any[src >> 6] >> (src & 63)) & 1ull)) livem &= ~(1u << i);
                }
            }
            // queries whose needed dictionary bits are all present: count the row's bits per query, compare with popcount(need)
            uint32_t s0 = 0u, s1 = 0u, s2 = 0u, s3 = 0u, s4 = 0u;
#pragma unroll
            for (int w = 0; w < MRAG_PHRASE_WORDS; ++w) {
                for (uint64_t m = f.phrase_bits[w]; m; m &= m - 1) {
                    uint32_t a = c.qneed[w * 64 + (__ffsll((long long)m) - 1)], t;
                    t = s0 & a; s0 ^= a; a = t;
                    t = s1 & a; s1 ^= a; a = t;
                    t = s2 & a; s2 ^= a; a = t;
                    t = s3 & a; s3 ^= a; a = t;
                    s4 ^= a;
                }
            }
            const uint32_t okm = ~((s0 ^ c.cnt[0]) | (s1 ^ c.cnt[1]) | (s2 ^ c.cnt[2]) | (s3 ^ c.cnt[3]) | (s4 ^ c.cnt[4])) & ~c.imp;
            // exemptions that do not depend on the coverage
            uint32_t exm = (f.flags & MRAG_CF_PROMOTED) ? c.phr : ((f.flags & MRAG_CF_CONTACT_VALUE) ? c.contact : 0u);
            uint32_t dslow = 0u;                  // queries whose d: codes need the phrase loop (overflow table / no code table)
            if (anydtag && c.dcodes) {
                if (c.ncodes == 0xFFFFFFFFu || (f.flags & MRAG_CF_DTAG_OVERFLOW)) dslow = c.dcodes;
                if (c.ncodes != 0xFFFFFFFFu) {
                    for (uint32_t i = 0; i < c.ncodes; ++i) {
                        const uint32_t code = c.code[i];
                        if (f.dtags[0] == code || f.dtags[1] == code || f.dtags[2] == code || f.dtags[3] == code) exm |= c.codemask[i];
                    }
                }
            }
            exm &= c.phr;
            keepm = livem & (~c.phr | exm);
            // (a floor below 1 -- the reference's is the constant 1.0, corpus_search.py:614 -- can be met with a phrase
            //  missing, so such a query takes the phrase loop for every row)
            slow = livem & c.phr & ~exm & (okm | dslow | c.lowfloor);
        }
        uint32_t* hm = hmask + size_t(q0) * nwords + (r >> 5);
        const bool writer = lane == 0 && (r >> 5) < nwords;
        for (int i = 0; i < nqc; ++i) {
            const uint32_t word = __ballot_sync(kFull, (keepm >> i) & 1u);
            if (writer) hm[size_t(i) * nwords] = word;
        }
        unsigned todo = __ballot_sync(kFull, slow != 0u);
        __syncwarp();                            // the words above are in place before any bit is OR-ed into them
        while (todo) {
            const int L = __ffs(todo) - 1;
            todo &= todo - 1;
            FeatWords fw;
            fw.f = f;
#pragma unroll
            for (int k = 0; k < 10; ++k) fw.w[k] = __shfl_sync(kFull, fw.w[k], L);
            const uint32_t slowL = __shfl_sync(kFull, slow, L);
            const long long rL = __shfl_sync(kFull, (long long)r, L);
            const unsigned long long jtL = __shfl_sync(kFull, (unsigned long long)reinterpret_cast<uintptr_t>(jt), L);
            if (lane < nqc && ((slowL >> lane) & 1u)) {
                const DevHyb& h = hq[q0 + lane];
                const uint64_t* jp = reinterpret_cast<const uint64_t*>(uintptr_t(jtL));
                if (hybrid_keep(h, fw.f, hybrid_eval(h, fw.f, jp, rL, ov)))
                    atomicOr(&hmask[size_t(q0 + lane) * nwords + (rL >> 5)], 1u << (rL & 31));
            }
        }
    }
}

// ---- pair path: when the coverage floors leave each query a small share of the rows (the usual case: required phrases),
// the rerank is taken over the LIST of surviving (query, row) pairs instead of a scan that walks every mask word and
// serves 4 queries per pass:  count -> (host: segment offsets) -> fill -> score (one warp per pair) -> segmented merge.
// counts[q] += rows set in hmask[q][*].   grid = (blocks over words, nq)
__global__ void __launch_bounds__(256) hybrid_count_kernel(const uint32_t* __restrict__ hmask, int64_t nwords, unsigned long long* __restrict__ counts) {
    __shared__ unsigned s_part[8];
    const int q = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t* m = hmask + size_t(q) * nwords;
    unsigned c = 0;
    for (int64_t w = int64_t(blockIdx.x) * 256 + threadIdx.x; w < nwords; w += int64_t(gridDim.x) * 256) c += __popc(__ldg(m + w));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(kFull, c, o);
    if (lane == 0) s_part[warp] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
        for (int i = 0; i < 8; ++i) t += s_part[i];
        if (t) atomicAdd(counts + q, (unsigned long long)t);
    }
}

// rows_out[seg_off[q] + ...] = the rows set in hmask[q][*] (order within a segment is arbitrary: the keys carry the row).
// One warp per 32 words of a query.   grid = (blocks over word groups, nq)
__global__ void __launch_bounds__(256) hybrid_fill_kernel(const uint32_t* __restrict__ hmask, int64_t nwords, const int64_t* __restrict__ seg_off,
                                                         unsigned long long* __restrict__ cursor, uint32_t* __restrict__ rows_out) {
    const int q = blockIdx.y, lane = threadIdx.x & 31;
    const int64_t warps = (int64_t(gridDim.x) * 256) >> 5;
    const uint32_t* m = hmask + size_t(q) * nwords;
    for (int64_t w0 = ((int64_t(blockIdx.x) * 256 + threadIdx.x) >> 5) * 32; w0 < nwords; w0 += warps * 32) {
        const int64_t w = w0 + lane;
        uint32_t word = w < nwords ? __ldg(m + w) : 0u;
        const unsigned c = __popc(word);
        unsigned incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned v = __shfl_up_sync(kFull, incl, o);
            if (lane >= o) incl += v;
        }
        const unsigned total = __shfl_sync(kFull, incl, 31);
        if (total == 0) continue;
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(cursor + q, (unsigned long long)total);
        base = __shfl_sync(kFull, base, 0);
        uint32_t* out = rows_out + seg_off[q] + int64_t(base) + (incl - c);
        for (; word; word &= word - 1) *out++ = uint32_t(w * 32 + (__ffs(word) - 1));
    }
}

struct PairArgs {
    const void* rows; int ld; const float* inv_norm; const float* q; const float* qinv;
    const uint32_t* pair_rows; const int64_t* seg_off; int nq; int64_t total;
    const mrag_chunkfeat* feat; const DevHyb* hyb; const uint32_t* doc_idx; const uint8_t* authority;
    const uint64_t* doc_jtags; int64_t n_jtag_docs; DtagOver ov;
    uint64_t* keys;                 // [total] key of the pair, 0 = not eligible (NaN score)
};

// dot product of one stored row with one padded fp32 query, by a whole warp (every lane returns the sum)
template <int DT>
MRAG_DEVINL float pair_dot(const void* rows, int ld, uint32_t row, const float* qrow, int lane) {
    float acc = 0.0f;
    if (DT == 1) {
        const uint4* x = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(rows) + size_t(row) * ld);
        for (int v = lane; v < ld / 8; v += 32) {
            const uint4 d = ldg_stream(x + v);
            const float4 q0 = __ldg(reinterpret_cast<const float4*>(qrow) + 2 * v);
            const float4 q1 = __ldg(reinterpret_cast<const float4*>(qrow) + 2 * v + 1);
            acc = fmaf(bf16lo(d.x), q0.x, acc); acc = fmaf(bf16hi(d.x), q0.y, acc);
            acc = fmaf(bf16lo(d.y), q0.z, acc); acc = fmaf(bf16hi(d.y), q0.w, acc);
            acc = fmaf(bf16lo(d.z), q1.x, acc); acc = fmaf(bf16hi(d.z), q1.y, acc);
            acc = fmaf(bf16lo(d.w), q1.z, acc); acc = fmaf(bf16hi(d.w), q1.w, acc);
        }
    } else {
        const float4* x = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(rows) + size_t(row) * ld);
        for (int v = lane; v < ld / 4; v += 32) {
            const uint4 du = ldg_stream(x + v);
            const float4 q0 = __ldg(reinterpret_cast<const float4*>(qrow) + v);
            acc = fmaf(__uint_as_float(du.x), q0.x, acc); acc = fmaf(__uint_as_float(du.y), q0.y, acc);
            acc = fmaf(__uint_as_float(du.z), q0.z, acc); acc = fmaf(__uint_as_float(du.w), q0.w, acc);
        }
    }
    return warp_sum(acc);
}

// One warp per 32 consecutive pairs.  Phase A: the warp takes the dot products of its pairs one after the other (four rows
// in flight), lane j keeps pair j's.  Phase B: every lane turns ITS pair's cosine into the rerank score -- the same
// arithmetic as the scan's epilogue, 32 pairs at a time instead of one lane working while 31 wait (r2z: the one-pair-per-
// warp version issued 932 instructions per pair, ~700 of them on a single lane, and was issue bound at 0.92 ms).
template <int DT>
__global__ void __launch_bounds__(256) hybrid_pair_score_kernel(const PairArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t p0 = ((int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5) * 32;
    if (p0 >= a.total) return;
    const int np = int(min(int64_t(32), a.total - p0));
    int q = 0;
    uint32_t row = 0u;
    if (lane < np) {
        const int64_t p = p0 + lane;
        int lo = 0, hi = a.nq;                      // the segment that holds p: seg_off[lo] <= p < seg_off[lo + 1]
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(a.seg_off + mid) <= p) lo = mid; else hi = mid;
        }
        q = lo;
        row = __ldg(a.pair_rows + p);
    }
    float mydot = 0.0f;
    for (int j0 = 0; j0 < np; j0 += 4) {
        float d4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {                // independent loads of up to four rows (np is warp uniform)
            const int j = j0 + u < np ? j0 + u : j0;
            const uint32_t rj = __shfl_sync(kFull, row, j);
            const int qj = __shfl_sync(kFull, q, j);
            d4[u] = pair_dot<DT>(a.rows, a.ld, rj, a.q + size_t(qj) * a.ld, lane);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (lane == j0 + u) mydot = d4[u];
    }
    if (lane < np) {
        const float inv = a.inv_norm[row];
        const float cs = isinf(inv) ? CUDART_NAN_F : mydot * inv * a.qinv[q];
        // a NaN similarity reports 1.0 (max(0.0, min(1.0, nan)) in corpus_search.py:1569)
        const float c01 = (cs == cs) ? cs : 1.0f;
        const mrag_chunkfeat f = a.feat[row];
        const uint32_t d = a.doc_idx[row];
        const uint64_t* jt = (a.doc_jtags && int64_t(d) < a.n_jtag_docs) ? a.doc_jtags + size_t(d) * MRAG_JTAG_WORDS : nullptr;
        const DevHyb& h = a.hyb[q];
        const float s = hybrid_score(h, f, hybrid_eval(h, f, jt, int64_t(row), a.ov), c01, a.authority[row]);
        a.keys[p0 + lane] = (s == s) ? make_key(s, row) : 0ull;
    }
}

// The d-tag arm's WHERE (corpus_search.py:1632-1640, 1667-1672): row passes the base mask (filters over LIVE rows --
// that statement has no "embedding_vec IS NOT NULL") and chunk_d_tags holds any of the codes.
// counts[0] = rows passing the base mask ("n_total" of the IDF count, :1649-1653), counts[1+i] = of those, rows
// holding code i.
__global__ void __launch_bounds__(256) dtag_mask_kernel(const mrag_chunkfeat* __restrict__ feat, const uint32_t* __restrict__ base_mask,
                                                       int64_t n, const uint16_t* __restrict__ codes, int n_codes,
                                                       uint32_t* __restrict__ mask_out, unsigned long long* __restrict__ counts,
                                                       const DtagOver ov) {
    const int64_t r = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    bool base = false, any = false;
    uint32_t has = 0;                       // bit i: the row holds code i
    if (r < n) {
        base = (base_mask[r >> 5] >> (r & 31)) & 1u;
        if (base) {
            const mrag_chunkfeat f = feat[r];
            for (int i = 0; i < n_codes; ++i)
                if (has_dtag(f, codes[i], r, ov)) has |= 1u << i;
            any = has != 0u;
        }
    }
    const uint32_t word = __ballot_sync(kFull, any);
    const uint32_t bword = __ballot_sync(kFull, base);
    if (lane == 0 && r < ((n + 31) & ~int64_t(31))) mask_out[r >> 5] = word;
    if (lane == 0 && bword) atomicAdd(counts, (unsigned long long)__popc(bword));
    for (int i = 0; i < n_codes; ++i) {
        const uint32_t w = __ballot_sync(kFull, (has >> i) & 1u);
        if (lane == 0 && w) atomicAdd(counts + 1 + i, (unsigned long long)__popc(w));
    }
}

}  // namespace mrag

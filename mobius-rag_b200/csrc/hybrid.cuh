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

// One thread per row, all queries in a loop (the row's features are read once).  Bit (q, r) = row r passes the
// WHERE mask and query q's coverage floor.  hmask: [nq][nwords].
__global__ void __launch_bounds__(256) hybrid_mask_kernel(const DevHyb* __restrict__ hq, int nq, const mrag_chunkfeat* __restrict__ feat,
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
    const bool maybe_exempt = base && ((f.flags & (MRAG_CF_PROMOTED | MRAG_CF_CONTACT_VALUE)) ||
                                       (f.dtags[0] | f.dtags[1] | f.dtags[2] | f.dtags[3]) != 0u);
    for (int q = 0; q < nq; ++q) {
        const DevHyb& h = hq[q];
        bool keep = base;
        if (keep && h.src_restrict) keep = (h.q.source_type_any[src >> 6] >> (src & 63)) & 1ull;
        if (keep && h.q.n_phrases > 0) {
            // quick reject: a phrase that only its dictionary bit can satisfy is missing, and no exemption can apply
            const bool bits_ok = !h.impossible && (f.phrase_bits[0] & h.need[0]) == h.need[0] && (f.phrase_bits[1] & h.need[1]) == h.need[1];
            if (!bits_ok && !maybe_exempt) keep = false;
            else keep = hybrid_keep(h, f, hybrid_eval(h, f, jt, r, ov));
        }
        const uint32_t word = __ballot_sync(kFull, keep);
        if (lane == 0 && (r >> 5) < nwords) hmask[size_t(q) * nwords + (r >> 5)] = word;
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

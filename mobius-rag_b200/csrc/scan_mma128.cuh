// scan_mma128.cuh -- K1'' : candidate-generating tensor-core scan, 128 queries per pass.
//
// Same statement as scan_mma.cuh (corpus_search.py:1525-1536, vector_store.py:274-287) for large
// query batches.  One pass over the bf16 rows (the corpus itself, or the bf16 shadow copy of an
// fp32 corpus) serves 128 queries: every TMEM lane holds ONE query, normalised and rounded to bf16
// (tcgen05.mma kind::f16 needs A and B in the same format: an f16 x bf16 descriptor raises an illegal
// instruction on B200 -- measured, r1g).  The scores are therefore APPROXIMATE:
//     |approx - exact| <= eps,   eps = 2^-9 (query rounding) [+ 2^-9 (shadow rounding) for fp32 corpora]
// (Cauchy-Schwarz on the rounding error vectors), so this kernel only nominates the K' = k + slack best
// rows per query; rescore.cuh recomputes those exactly from the primary rows, proves that nothing
// outside the candidate set can reach the top k (certificate) or hands the query to the exact scan.
//
// Layout of the machine (192 threads, 1 CTA per SM, persistent over 64-row tiles):
//   warp 0     TMA producer: corpus tile boxes into the ring + 256 B of 1/|x| per tile (bulk copy)
//   warp 1     MMA issuer, TMEM owner:  D[128 queries x 64 rows] += A[128 x K] * B[64 x K]^T
//   warps 2-5  select: thread = query, TMEM lane = thread; scores of a tile in registers, per-query
//              candidate buffer of K' + 32 keys in shared memory, warp-cooperative compaction.
// TMEM: 384 columns of A (K = 768, bf16 pairs) + 2 x 64 columns of D = 512.
#pragma once
#include "scan_mma.cuh"

namespace mrag {

constexpr int kMma128Queries = 128;
constexpr int kMma128InvSlots = 8;
constexpr int kMma128Slack = 32;          // candidate buffer = K' + slack keys per query

// a pipeline stage = kbs k-blocks (kbs * 8 KB): one full / empty barrier pair, one wait and one commit per stage
// (the MMA issue thread paces this kernel; fewer waits and commits per tile shorten its loop)
inline size_t mma128_smem_bytes(int stages, int cap, int kbs = 1) {
    return 1024 /*align slack*/ + size_t(stages) * kbs * kMmaStageBytes + size_t(kMma128InvSlots) * 64 * 4 +
           size_t(kMma128Queries) * cap * 8 + 1024 /*barriers*/;
}

MRAG_DEVINL void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// rank sort like warp_rank_select, with the per-lane work bounded by the lanes actually in use; up to 256 keys
// (K' = 2k nominees + 32 slack keys for k up to 106: with K' = 1.5k a few certificates per 256 queries failed at k = 100
//  and each failure cost one exact pass over the shard)
constexpr int kRank128PerLane = 8;
MRAG_DEVINL void warp_rank_select_n(uint64_t* buf, int n, int keep, int lane) {
    uint64_t key[kRank128PerLane];
    int rank[kRank128PerLane];
    const int ne = (n + 31) >> 5;             // warp uniform
#pragma unroll
    for (int e = 0; e < kRank128PerLane; ++e) {
        const int i = lane + 32 * e;
        key[e] = (i < n) ? buf[i] : 0ull;
        rank[e] = 0;
    }
    for (int j = 0; j < n; ++j) {
        const uint64_t kj = buf[j];           // broadcast read
#pragma unroll
        for (int e = 0; e < kRank128PerLane; ++e)
            if (e < ne) rank[e] += (kj > key[e]) ? 1 : 0;
    }
    __syncwarp();
#pragma unroll
    for (int e = 0; e < kRank128PerLane; ++e) {
        const int i = lane + 32 * e;
        if (i < n && rank[e] < keep) buf[rank[e]] = key[e];
    }
    __syncwarp();
}

__device__ __noinline__ SelState select_compact128(SelState st, unsigned full, uint64_t* cand_warp, int cap, int k, int lane) {
    while (full) {
        const int L = __ffs(full) - 1;
        full &= full - 1;
        __syncwarp();
        uint64_t* b = cand_warp + size_t(L) * cap;
        warp_rank_select_n(b, cap, k, lane);
        const uint64_t kth = b[k - 1];
        if (lane == L) { st.cnt = k; st.thr_s = key_score(kth); }
    }
    return st;
}

// rows 8G .. 8G+7 of the tile: append what beats the threshold; a full buffer is compacted and the walk resumes.
// (a function template so that every index into sc[] is a compile-time constant: the array must stay in registers)
template <int G>
MRAG_DEVINL void walk_group(const float (&sc)[64], float& thr, SelState& st, uint64_t* mybuf, uint64_t* cand_warp, int cap,
                            int k, int lane, int64_t r0, uint32_t* gslot, unsigned& n_keys, unsigned& n_compact) {
    const int cnt0 = st.cnt;
    int c_start = 0;
    for (;;) {
        int ovf = 8;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            if (c >= c_start && sc[8 * G + c] > thr) {
                if (st.cnt < cap) mybuf[st.cnt++] = make_key(sc[8 * G + c], uint32_t(r0 + 8 * G + c));
                else ovf = min(ovf, c);
            }
        }
        const unsigned full = __ballot_sync(kFull, st.cnt == cap);
        if (!full) { n_keys += st.cnt - cnt0; break; }
        n_compact += __popc(full);
        st = select_compact128(st, full, cand_warp, cap, k, lane);
        if ((full >> lane) & 1u) atomicMax(gslot, f2ord(st.thr_s));
        thr = fmaxf(thr, st.thr_s);
        c_start = ovf;
        if (!__any_sync(kFull, ovf < 8)) break;
    }
}

// a.q: fp32 queries; a.qinv: 1/|q|; a.k = K' (candidates wanted), a.kp = pow2 >= K', a.cap = K' + slack.
// a.ub is ignored.
// KREG = 0: per-query shared-memory buffers of `cap` keys (the full pass).
// KREG > 0: every thread keeps the sorted top-KREG of its query in registers (a.k <= KREG, cap = 0): the
//           sampling pass, where a buffer would be compacted over and over before any bound exists.
template <int KREG>
__global__ void __launch_bounds__(kMmaThreads, 1) scan_mma128_kernel(const __grid_constant__ CUtensorMap tmap, const MmaArgs a) {
    extern __shared__ __align__(1024) unsigned char mma_smem[];
    unsigned char* smem = mma_smem + ((1024u - (smem_u32(mma_smem) & 1023u)) & 1023u);
    unsigned char* stage_base = smem;                                                    // stages * 8 KB
    const int kbs = a.kbs;
    const int stage_bytes = kbs * kMmaStageBytes;
    float* xinv = reinterpret_cast<float*>(smem + size_t(a.stages) * stage_bytes);      // [8][64] 1/|x| of a tile's rows
    uint64_t* smem_cand = reinterpret_cast<uint64_t*>(xinv + kMma128InvSlots * 64);      // [128 queries][cap] unless a.gcand
    uint64_t* cand = a.gcand ? a.gcand + size_t(blockIdx.x) * kMma128Queries * a.cap : smem_cand;
    uint64_t* bars = smem_cand + (a.gcand ? size_t(0) : size_t(kMma128Queries) * a.cap);
    uint64_t* full_bar = bars;                       // [stages]   TMA -> MMA
    uint64_t* empty_bar = full_bar + a.stages;       // [stages]   MMA -> TMA
    uint64_t* tfull_bar = empty_bar + a.stages;      // [2]        MMA -> select
    uint64_t* tempty_bar = tfull_bar + 2;            // [2]        select -> MMA
    uint64_t* ifull_bar = tempty_bar + 2;            // [4]        TMA (1/|x|) -> select
    uint64_t* iempty_bar = ifull_bar + kMma128InvSlots;   // [4]   select -> TMA
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(iempty_bar + kMma128InvSlots);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t slp = a.sleep_ns;
    const int kblocks = a.ld / kMmaKBlock;
    const int64_t num_tiles = (a.n + kMmaTileRows - 1) / kMmaTileRows;
    const int64_t nwords = (a.n + 31) >> 5;
    const int64_t tmul = a.tile_mul;
    const int64_t G = int64_t(gridDim.x) * tmul;
    const int64_t t_first = int64_t(blockIdx.x) * tmul;

    auto tile_mask = [&](int64_t t) -> uint2 {
        uint2 m = make_uint2(0u, 0u);
        if (t < num_tiles) {
            m.x = __ldg(a.mask + 2 * t);
            if (2 * t + 1 < nwords) m.y = __ldg(a.mask + 2 * t + 1);
        }
        return m;
    };

    if (tid == 0) {
        for (int s = 0; s < a.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 128); }
        for (int i = 0; i < kMma128InvSlots; ++i) { mbar_init(&ifull_bar[i], 1); mbar_init(&iempty_bar[i], 4); }
        fence_barrier_init();
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap)) : "memory");
    }
    if (warp == 1) tmem_alloc(tmem_holder, kMmaTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    const int quarter = warp & 3;                       // TMEM lanes [32*quarter, +32)
    const int qi = quarter * 32 + lane;                 // query owned by this thread (select warps)

    // ---- queries -> tensor memory: normalised (|q| = 1), rounded to bf16 once
    if (warp >= 2) {
        const bool live = qi < a.nq;
        const float* qrow = a.q + size_t(a.q0 + (live ? qi : 0)) * a.ld;
        float qs = live ? a.qinv[a.q0 + qi] : 0.0f;
        if (isinf(qs)) qs = 0.0f;                       // zero query: admits nothing anyway
        for (int c0 = 0; c0 < a.ld / 2; c0 += 32) {
            uint32_t r[32];
#pragma unroll
            for (int v = 0; v < 16; ++v) {
                float4 f = live ? __ldg(reinterpret_cast<const float4*>(qrow + c0 * 2) + v) : make_float4(0.f, 0.f, 0.f, 0.f);
                r[2 * v] = pack_bf16x2(f.x * qs, f.y * qs);
                r[2 * v + 1] = pack_bf16x2(f.z * qs, f.w * qs);
            }
            const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(c0);
            MRAG_TMEM_ST32(taddr, r);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp == 0) {
        // ================= TMA producer =================
        int s = 0, is = 0;
        uint32_t ph = 0, iph = 0;
        uint2 m = tile_mask(t_first);
        for (int64_t t = t_first; t < num_tiles; t += G) {
            const uint2 mn = tile_mask(t + G);
            if ((m.x | m.y) != 0u) {
                mbar_wait(&iempty_bar[is], iph ^ 1u, slp);
                if (elect_one()) {
                    mbar_expect_tx(&ifull_bar[is], 256);
                    bulk_load_1d(xinv + is * 64, a.inv_norm + t * kMmaTileRows, 256, &ifull_bar[is]);
                }
                __syncwarp();
                if (++is == kMma128InvSlots) { is = 0; iph ^= 1u; }
                for (int kb = 0; kb < kblocks; kb += kbs) {
                    mbar_wait(&empty_bar[s], ph ^ 1u, slp);
                    if (elect_one()) {
                        mbar_expect_tx(&full_bar[s], uint32_t(stage_bytes));
                        for (int j = 0; j < kbs; ++j)
                            tma_load_2d(stage_base + size_t(s) * stage_bytes + size_t(j) * kMmaStageBytes, &tmap,
                                        (kb + j) * kMmaKBlock, int(t * kMmaTileRows), &full_bar[s]);
                    }
                    __syncwarp();
                    if (++s == a.stages) { s = 0; ph ^= 1u; }
                }
            }
            m = mn;
        }
    } else if (warp == 1) {
        // ================= MMA issuer: ONE elected thread runs the whole loop =================
        // (the issue loop is the critical path of the kernel -- ncu r1g: ~75 instructions per k-block with a
        //  per-iteration election; a single-thread loop drops the election, the reconvergence and the warp syncs)
        if (elect_one()) {
            int s = 0, as = 0;
            uint32_t ph = 0, aph = 0;
            const uint64_t bdesc0 = make_sw128_desc(smem_u32(stage_base));
            const uint32_t idesc = kMmaIdesc;
            uint2 m = tile_mask(t_first);
            for (int64_t t = t_first; t < num_tiles; t += G) {
                const uint2 mn = tile_mask(t + G);
                if ((m.x | m.y) != 0u) {
                    mbar_wait(&tempty_bar[as], aph ^ 1u, slp);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + uint32_t(kMmaDCol0 + as * kMmaTileRows);
                    for (int kb = 0; kb < kblocks; kb += kbs) {
                        mbar_wait(&full_bar[s], ph, slp);
                        tc_fence_after();
                        for (int j = 0; j < kbs; ++j) {
                            const uint64_t bdesc = bdesc0 + uint64_t((size_t(s) * stage_bytes + size_t(j) * kMmaStageBytes) >> 4);
                            const uint32_t a_tmem = tmem_base + uint32_t((kb + j) * (kMmaKBlock / 2));
                            umma_ts_bf16(d_tmem, a_tmem, bdesc, idesc, (kb + j) != 0 ? 1u : 0u);
                            umma_ts_bf16(d_tmem, a_tmem + 8, bdesc + 2, idesc, 1u);
                            umma_ts_bf16(d_tmem, a_tmem + 16, bdesc + 4, idesc, 1u);
                            umma_ts_bf16(d_tmem, a_tmem + 24, bdesc + 6, idesc, 1u);
                        }
                        umma_commit(&empty_bar[s]);
                        if (kb + kbs >= kblocks) umma_commit(&tfull_bar[as]);
                        if (++s == a.stages) { s = 0; ph ^= 1u; }
                    }
                    if (++as == 2) { as = 0; aph ^= 1u; }
                }
                m = mn;
            }
        }
        __syncwarp();
    } else {
        // ================= select (warps 2..5): thread = query =================
        const uint32_t lane_addr = uint32_t(quarter * 32) << 16;
        int as = 0, is = 0;
        uint32_t aph = 0, iph = 0;
        const bool live = qi < a.nq;
        const bool warp_live = quarter * 32 < a.nq;
        const float qinv = live ? a.qinv[a.q0 + qi] : 0.0f;
        uint64_t* cand_warp = cand + size_t(quarter * 32) * a.cap;
        uint64_t* mybuf = cand_warp + size_t(lane) * a.cap;
        const int cap = a.cap;
        SelState st;
        st.cnt = 0;
        st.thr_s = (live && !isinf(qinv)) ? -CUDART_INF_F : CUDART_INF_F;
        uint32_t* gslot = a.gthr + a.q0 + (live ? qi : 0);
        unsigned n_tiles = 0, n_groups = 0, n_keys = 0, n_compact = 0;
        // KREG mode: only SCORES are kept (the pass exists to produce a bound, rows do not matter): sorted
        // orderable scores, the k live slots are the LAST k of top[] (sentinels before them), so the k-th best
        // is always top[KREG-1]; every index is a compile-time constant (the array must stay in registers)
        uint32_t top[KREG > 0 ? KREG : 1];
        const int top_off = (KREG > 0 ? KREG : 1) - a.k;
#pragma unroll
        for (int i = 0; i < (KREG > 0 ? KREG : 1); ++i) top[i] = (i < top_off) ? ~0u : 0u;

        uint2 m = tile_mask(t_first);
        for (int64_t t = t_first; t < num_tiles; t += G) {
            const uint2 mn = tile_mask(t + G);
            if ((m.x | m.y) != 0u) {
                const int64_t r0 = t * kMmaTileRows;
                float sc[64];
                float bestg[8];                                      // max of each group of 8 rows
                uint32_t gord;
                asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(gord) : "l"(gslot) : "memory");
                mbar_wait(&ifull_bar[is], iph, slp);
                mbar_wait(&tfull_bar[as], aph, slp);
                tc_fence_after();
                if (warp_live) {
                    const float4* inv4 = reinterpret_cast<const float4*>(xinv + is * 64);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        uint32_t d[32];
                        MRAG_TMEM_LD32(d, tmem_base + lane_addr + uint32_t(kMmaDCol0 + as * kMmaTileRows + h * 32));
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int c4 = 0; c4 < 8; ++c4) {
                            const float4 iv = inv4[h * 8 + c4];                  // broadcast
                            const float ivv[4] = {iv.x, iv.y, iv.z, iv.w};
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const int c = h * 32 + c4 * 4 + i;
                                sc[c] = __uint_as_float(d[c4 * 4 + i]) * ivv[i];  // query already normalised
                            }
                        }
                    }
                    // rows whose mask bit is clear (filtered, NULL vector, past the end) never score;
                    // a zero row has 1/|x| = +inf and scores NaN (0 * inf), which no comparison admits
                    if ((m.x & m.y) != 0xffffffffu) {
#pragma unroll
                        for (int c = 0; c < 64; ++c) {
                            const uint32_t w = c < 32 ? m.x : m.y;
                            if (!((w >> (c & 31)) & 1u)) sc[c] = -CUDART_INF_F;
                        }
                    }
#pragma unroll
                    for (int g = 0; g < 8; ++g) {
                        const float m01 = fmaxf(sc[8 * g + 0], sc[8 * g + 1]), m23 = fmaxf(sc[8 * g + 2], sc[8 * g + 3]);
                        const float m45 = fmaxf(sc[8 * g + 4], sc[8 * g + 5]), m67 = fmaxf(sc[8 * g + 6], sc[8 * g + 7]);
                        bestg[g] = fmaxf(fmaxf(m01, m23), fmaxf(m45, m67));      // fmaxf drops NaN
                    }
                }
                tc_fence_before();
                mbar_arrive(&tempty_bar[as]);
                if (++as == 2) { as = 0; aph ^= 1u; }
                __syncwarp();
                if (lane == 0) mbar_arrive(&iempty_bar[is]);
                if (++is == kMma128InvSlots) { is = 0; iph ^= 1u; }
                if (warp_live) {
                    float thr = fmaxf(st.thr_s, gord ? ord2f(gord - 1u) : -CUDART_INF_F);
                    // the warp walks only the groups of 8 rows in which some query admits a row
                    // (warp-uniform votes: in steady state a tile has no such group, or one)
                    ++n_tiles;
                    if constexpr (KREG > 0) {
                        const float thr_in = st.thr_s;
#pragma unroll
                        for (int c = 0; c < 64; ++c) {
                            const bool ins = sc[c] > thr;
                            if (__any_sync(kFull, ins)) {                        // warp-uniform
                                uint32_t key = ins ? f2ord(sc[c]) : 0u;
#pragma unroll
                                for (int i = 0; i < KREG; ++i) {                 // sorted insertion, a zero key falls through
                                    const uint32_t hi = max(key, top[i]);
                                    key = min(key, top[i]);
                                    top[i] = hi;
                                }
                                const uint32_t kth = top[KREG - 1];
                                if (kth) { st.thr_s = ord2f(kth); thr = fmaxf(thr, st.thr_s); }
                            }
                        }
                        if (st.thr_s > thr_in) atomicMax(gslot, f2ord(st.thr_s));
                    } else {
#define MRAG_WALK_GROUP(G)                                                                                        \
                    if (__any_sync(kFull, bestg[G] > thr)) {                                                      \
                        ++n_groups;                                                                               \
                        walk_group<G>(sc, thr, st, mybuf, cand_warp, cap, a.k, lane, r0, gslot, n_keys, n_compact); \
                    }
                    MRAG_WALK_GROUP(0) MRAG_WALK_GROUP(1) MRAG_WALK_GROUP(2) MRAG_WALK_GROUP(3)
                    MRAG_WALK_GROUP(4) MRAG_WALK_GROUP(5) MRAG_WALK_GROUP(6) MRAG_WALK_GROUP(7)
#undef MRAG_WALK_GROUP
                    }
                }
            }
            m = mn;
        }

        if (a.stats) {
            // [0] tiles (per select warp), [1] groups of 8 rows walked, [2] keys appended (uncompacted tiles only),
            // [3] buffer compactions
            if (lane == 0) { atomicAdd(a.stats + 0, (unsigned long long)n_tiles); atomicAdd(a.stats + 1, (unsigned long long)n_groups);
                             atomicAdd(a.stats + 3, (unsigned long long)n_compact); }
            if (n_keys) atomicAdd(a.stats + 2, (unsigned long long)n_keys);
        }
        // ---- this CTA's sorted candidate list per query
        __syncwarp();
        if constexpr (KREG > 0) {
            if (live) {
                uint64_t* out = a.part + (size_t(a.q0 + qi) * a.P + blockIdx.x) * a.kp;
                // keys stay unique across CTAs and slots (the merge's radix select relies on it); 0 = empty
#pragma unroll
                for (int i = 0; i < KREG; ++i)
                    if (i >= top_off) out[i - top_off] = top[i] ? (uint64_t(top[i]) << 32) | uint64_t((blockIdx.x << 8) | unsigned(i)) : 0ull;
                for (int i = a.k; i < a.kp; ++i) out[i] = 0ull;
            }
        } else {
            for (int L = 0; L < 32; ++L) {
                const int qL = quarter * 32 + L;
                if (qL >= a.nq) break;
                const int n = __shfl_sync(kFull, st.cnt, L);
                uint64_t* b = cand_warp + size_t(L) * a.cap;
                warp_rank_select_n(b, n, a.kp, lane);
                uint64_t* out = a.part + (size_t(a.q0 + qL) * a.P + blockIdx.x) * a.kp;
                const int have = n < a.k ? n : a.k;
                for (int i = lane; i < a.kp; i += 32) out[i] = (i < have) ? b[i] : 0ull;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kMmaTmemCols);
    }
}

}  // namespace mrag

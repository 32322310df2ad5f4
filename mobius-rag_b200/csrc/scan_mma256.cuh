// scan_mma256.cuh -- K1''' : the candidate scan on CTA PAIRS (tcgen05 cta_group::2): 256 queries per pass.
//
// Same statement and same role as scan_mma128.cuh (nominate the K' best rows per query; rescore.cuh makes the
// result exact), but two CTAs of a thread-block cluster share every corpus tile:
//     D[256 queries x 64 rows] += A[256 x K] * B[64 x K]^T        tcgen05.mma.cta_group::2.kind::f16, M = 256, N = 64
//   A: each CTA keeps ITS 128 queries in its own tensor memory (lanes), as in scan_mma128;
//   B: each CTA streams HALF of the tile (32 rows, 4 KB per k-block) into its own shared memory, and the pair's
//      tensor cores read both halves -- a corpus byte is fetched from HBM once per 256 queries;
//   D: each CTA's tensor memory receives the 64 scores of its own 128 queries.
// Only the leader CTA (cluster rank 0) issues MMAs.  Barriers:
//   full[s]    (leader)   both CTAs' TMA loads complete_tx on the LEADER's barrier (address with the peer bit cleared)
//   empty[s]   (each CTA) tcgen05.commit ... multicast::cluster, mask 0b11: the stage is free in both CTAs
//   tfull[b]   (each CTA) same multicast commit: the accumulator buffer is ready in both CTAs
//   tempty[b]  (leader)   one arrive per select warp of BOTH CTAs (8), remote for the peer
// The select half of the kernel is scan_mma128's (walk_group / select_compact128).
#pragma once
#include "scan_mma128.cuh"

namespace mrag {

constexpr int kMma256HalfRows = 32;                          // rows of a tile each CTA loads
constexpr int kMma256StageBytes = kMma256HalfRows * 128;     // 4 KB per k-block per CTA
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;               // shared::cluster address of the even CTA of a pair
// kind::f16 instruction descriptor: D fp32, A/B bf16, K-major, M = 256, N = 64
constexpr uint32_t kMma256Idesc = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(kMmaTileRows >> 3) << 17) | (uint32_t(256 >> 4) << 24);

// a pipeline stage = kbs k-blocks (kbs * 4 KB per CTA): one full / empty barrier pair and ONE multicast commit per
// stage -- a commit per k-block made the issue thread the bottleneck (measured r1n: 8 us per tile instead of 2.3)
inline size_t mma256_smem_bytes(int stages, int cap, int kbs = 1) {
    return 1024 /*align slack*/ + size_t(stages) * kbs * kMma256StageBytes + size_t(kMma128InvSlots) * 64 * 4 +
           size_t(kMma128Queries) * cap * 8 + 2048 /*barriers*/;
}

MRAG_DEVINL uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
MRAG_DEVINL void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
MRAG_DEVINL void tmem_alloc_2cta(uint32_t* dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
MRAG_DEVINL void tmem_dealloc_2cta(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// this CTA's half tile -> its own shared memory; the bytes are counted on the LEADER's barrier
MRAG_DEVINL void tma_load_2d_2cta(void* dst, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar) {
    const uint64_t evict_first = 0x12F0000000000000ull;
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "l"(evict_first)
        : "memory");
}
// arrive (count 1) on the LEADER's copy of `bar` from either CTA of the pair
MRAG_DEVINL void mbar_arrive_leader(uint64_t* bar) {
    // default semantics (release at CTA scope), as CUTLASS's ClusterBarrier::arrive(cta_id): a .release.cluster arrive
    // compiles to MEMBAR.ALL.GPU + ERRBAR + CGAERRBAR in front of the remote arrive and took 16 % of all stall samples
    // (ncu r1n); what has to be ordered here are tensor-memory reads, which tcgen05.fence::before_thread_sync covers
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}
MRAG_DEVINL void umma_ts_bf16_2cta(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
MRAG_DEVINL void umma_commit_2cta(uint64_t* bar) {     // arrives on `bar` in BOTH CTAs when the pair's MMAs so far retire
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"(uint16_t(3))
                 : "memory");
}

// MmaArgs as scan_mma128 (KREG = 0 only): a.q0 = first query of the PAIR (CTA rank r serves [q0 + 128 r, +128)),
// a.nq <= 256, a.P = number of pairs; tmap32 = the corpus tensor map with 64 x 32 boxes.
// KBS = k-blocks per pipeline stage (compile time, so that the 4 * KBS MMAs of a stage are issued from ONE descriptor
// base with immediate offsets: the issue thread is the critical path -- ncu r1o: ~80 cycles per MMA with run-time
// descriptor arithmetic, against 32 cycles of tensor work)
template <int KBS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kMmaThreads, 1)
scan_mma256_kernel(const __grid_constant__ CUtensorMap tmap32, const MmaArgs a) {
    extern __shared__ __align__(1024) unsigned char mma_smem[];
    unsigned char* smem = mma_smem + ((1024u - (smem_u32(mma_smem) & 1023u)) & 1023u);
    unsigned char* stage_base = smem;                                                     // stages * 4 KB
    constexpr int kbs = KBS;
    const int stage_bytes = kbs * kMma256StageBytes;
    float* xinv = reinterpret_cast<float*>(smem + size_t(a.stages) * stage_bytes);       // [8][64] 1/|x| of a tile's rows
    uint64_t* smem_cand = reinterpret_cast<uint64_t*>(xinv + kMma128InvSlots * 64);
    uint64_t* cand = a.gcand ? a.gcand + size_t(blockIdx.x) * kMma128Queries * a.cap : smem_cand;
    uint64_t* bars = smem_cand + (a.gcand ? size_t(0) : size_t(kMma128Queries) * a.cap);
    uint64_t* full_bar = bars;                       // [stages]   (leader's copy is the one in use)
    uint64_t* empty_bar = full_bar + a.stages;       // [stages]
    uint64_t* tfull_bar = empty_bar + a.stages;      // [2]
    uint64_t* tempty_bar = tfull_bar + 2;            // [2]        (leader's copy)
    uint64_t* ifull_bar = tempty_bar + 2;            // [8]        local
    uint64_t* iempty_bar = ifull_bar + kMma128InvSlots;   // [8]   local
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(iempty_bar + kMma128InvSlots);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t slp = a.sleep_ns;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair = blockIdx.x >> 1;
    const int npairs = gridDim.x >> 1;
    const int kblocks = a.ld / kMmaKBlock;
    const int64_t num_tiles = (a.n + kMmaTileRows - 1) / kMmaTileRows;
    const int64_t nwords = (a.n + 31) >> 5;
    const int64_t G = npairs;
    const int64_t t_first = pair;

    auto tile_mask = [&](int64_t t) -> uint2 {
        uint2 m = make_uint2(0u, 0u);
        if (t < num_tiles) {
            m.x = __ldg(a.mask + 2 * t);
            if (2 * t + 1 < nwords) m.y = __ldg(a.mask + 2 * t + 1);
        }
        return m;
    };

    if (tid == 0) {
        for (int s = 0; s < a.stages; ++s) { mbar_init(&full_bar[s], 2); mbar_init(&empty_bar[s], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 8); }
        for (int i = 0; i < kMma128InvSlots; ++i) { mbar_init(&ifull_bar[i], 1); mbar_init(&iempty_bar[i], 4); }
        fence_barrier_init();
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap32)) : "memory");
    }
    cluster_sync_all();                                  // both CTAs' barriers exist before anything remote touches them
    if (warp == 1) tmem_alloc_2cta(tmem_holder, kMmaTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    const int quarter = warp & 3;
    const int qi = quarter * 32 + lane;                  // query slot inside this CTA
    const int qbase = a.q0 + int(rank) * kMma128Queries; // first query of this CTA
    const int nq_cta = max(0, min(kMma128Queries, a.q0 + a.nq - qbase));

    // ---- this CTA's queries -> its tensor memory
    if (warp >= 2) {
        const bool live = qi < nq_cta;
        const float* qrow = a.q + size_t(live ? qbase + qi : a.q0) * a.ld;
        float qs = live ? a.qinv[qbase + qi] : 0.0f;
        if (isinf(qs)) qs = 0.0f;
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
    cluster_sync_all();                                  // A is in place in BOTH CTAs before the leader issues MMAs
    tc_fence_after();

    if (warp == 0) {
        // ================= TMA producer (each CTA: its half of every tile + the tile's 1/|x|) =================
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
                        // the leader announces the bytes of BOTH halves; the peer only arrives
                        if (leader) mbar_expect_tx(&full_bar[s], 2 * stage_bytes);
                        else mbar_arrive_leader(&full_bar[s]);
                        for (int j = 0; j < kbs; ++j)
                            tma_load_2d_2cta(stage_base + size_t(s) * stage_bytes + size_t(j) * kMma256StageBytes, &tmap32,
                                             (kb + j) * kMmaKBlock, int(t * kMmaTileRows + rank * kMma256HalfRows), &full_bar[s]);
                    }
                    __syncwarp();
                    if (++s == a.stages) { s = 0; ph ^= 1u; }
                }
            }
            m = mn;
        }
    } else if (warp == 1) {
        // ================= MMA issuer: one thread of the LEADER CTA =================
        if (leader && elect_one()) {
            int s = 0, as = 0;
            uint32_t ph = 0, aph = 0;
            const uint64_t bdesc0 = make_sw128_desc(smem_u32(stage_base));
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
                        const uint64_t sdesc = bdesc0 + uint64_t((size_t(s) * stage_bytes) >> 4);
                        const uint32_t abase = tmem_base + uint32_t(kb * (kMmaKBlock / 2));
                        umma_ts_bf16_2cta(d_tmem, abase, sdesc, kMma256Idesc, kb != 0 ? 1u : 0u);
#pragma unroll
                        for (int ji = 1; ji < 4 * KBS; ++ji) {
                            constexpr int kStageDesc = kMma256StageBytes >> 4;
                            const int j = ji >> 2, i = ji & 3;
                            umma_ts_bf16_2cta(d_tmem, abase + uint32_t(j * (kMmaKBlock / 2) + i * 8),
                                              sdesc + uint64_t(j * kStageDesc + i * 2), kMma256Idesc, 1u);
                        }
                        umma_commit_2cta(&empty_bar[s]);
                        if (kb + kbs >= kblocks) umma_commit_2cta(&tfull_bar[as]);
                        if (++s == a.stages) { s = 0; ph ^= 1u; }
                    }
                    if (++as == 2) { as = 0; aph ^= 1u; }
                }
                m = mn;
            }
        }
        __syncwarp();
    } else {
        // ================= select (warps 2..5 of each CTA): thread = query =================
        const uint32_t lane_addr = uint32_t(quarter * 32) << 16;
        int as = 0, is = 0;
        uint32_t aph = 0, iph = 0;
        const bool live = qi < nq_cta;
        const bool warp_live = quarter * 32 < nq_cta;
        const float qinv = live ? a.qinv[qbase + qi] : 0.0f;
        uint64_t* cand_warp = cand + size_t(quarter * 32) * a.cap;
        uint64_t* mybuf = cand_warp + size_t(lane) * a.cap;
        const int cap = a.cap;
        SelState st;
        st.cnt = 0;
        st.thr_s = (live && !isinf(qinv)) ? -CUDART_INF_F : CUDART_INF_F;
        uint32_t* gslot = a.gthr + (live ? qbase + qi : a.q0);
        unsigned n_groups = 0, n_keys = 0, n_compact = 0;

        uint2 m = tile_mask(t_first);
        for (int64_t t = t_first; t < num_tiles; t += G) {
            const uint2 mn = tile_mask(t + G);
            if ((m.x | m.y) != 0u) {
                const int64_t r0 = t * kMmaTileRows;
                float sc[64];
                float bestg[8];
                uint32_t gord;
                asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(gord) : "l"(gslot) : "memory");
                mbar_wait(&ifull_bar[is], iph, slp);
                mbar_wait(&tfull_bar[as], aph, slp);
                tc_fence_after();
                if (warp_live) {
                    const float4* inv4 = reinterpret_cast<const float4*>(xinv + is * 64);
                    // both halves of the accumulator row are requested before the single wait
                    uint32_t d0[32], d1[32];
                    MRAG_TMEM_LD32(d0, tmem_base + lane_addr + uint32_t(kMmaDCol0 + as * kMmaTileRows));
                    MRAG_TMEM_LD32(d1, tmem_base + lane_addr + uint32_t(kMmaDCol0 + as * kMmaTileRows + 32));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int c4 = 0; c4 < 8; ++c4) {
                        const float4 iv0 = inv4[c4], iv1 = inv4[8 + c4];
                        sc[c4 * 4 + 0] = __uint_as_float(d0[c4 * 4 + 0]) * iv0.x; sc[c4 * 4 + 1] = __uint_as_float(d0[c4 * 4 + 1]) * iv0.y;
                        sc[c4 * 4 + 2] = __uint_as_float(d0[c4 * 4 + 2]) * iv0.z; sc[c4 * 4 + 3] = __uint_as_float(d0[c4 * 4 + 3]) * iv0.w;
                        sc[32 + c4 * 4 + 0] = __uint_as_float(d1[c4 * 4 + 0]) * iv1.x; sc[32 + c4 * 4 + 1] = __uint_as_float(d1[c4 * 4 + 1]) * iv1.y;
                        sc[32 + c4 * 4 + 2] = __uint_as_float(d1[c4 * 4 + 2]) * iv1.z; sc[32 + c4 * 4 + 3] = __uint_as_float(d1[c4 * 4 + 3]) * iv1.w;
                    }
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
                        bestg[g] = fmaxf(fmaxf(m01, m23), fmaxf(m45, m67));
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive_leader(&tempty_bar[as]);          // the leader's MMA thread may overwrite this buffer
                    mbar_arrive(&iempty_bar[is]);
                }
                if (++as == 2) { as = 0; aph ^= 1u; }
                if (++is == kMma128InvSlots) { is = 0; iph ^= 1u; }
                if (warp_live) {
                    float thr = fmaxf(st.thr_s, gord ? ord2f(gord - 1u) : -CUDART_INF_F);
#define MRAG_WALK_GROUP(G_)                                                                                         \
                    if (__any_sync(kFull, bestg[G_] > thr)) {                                                       \
                        ++n_groups;                                                                                 \
                        walk_group<G_>(sc, thr, st, mybuf, cand_warp, cap, a.k, lane, r0, gslot, n_keys, n_compact); \
                    }
                    MRAG_WALK_GROUP(0) MRAG_WALK_GROUP(1) MRAG_WALK_GROUP(2) MRAG_WALK_GROUP(3)
                    MRAG_WALK_GROUP(4) MRAG_WALK_GROUP(5) MRAG_WALK_GROUP(6) MRAG_WALK_GROUP(7)
#undef MRAG_WALK_GROUP
                }
            }
            m = mn;
        }
        (void)n_groups;

        // ---- this pair's sorted candidate list per query (P = number of pairs)
        __syncwarp();
        for (int L = 0; L < 32; ++L) {
            const int qL = quarter * 32 + L;
            if (qL >= nq_cta) break;
            const int n = __shfl_sync(kFull, st.cnt, L);
            uint64_t* b = cand_warp + size_t(L) * a.cap;
            warp_rank_select_n(b, n, a.kp, lane);
            uint64_t* out = a.part + (size_t(qbase + qL) * a.P + pair) * a.kp;
            const int have = n < a.k ? n : a.k;
            for (int i = lane; i < a.kp; i += 32) out[i] = (i < have) ? b[i] : 0ull;
        }
    }

    tc_fence_before();
    cluster_sync_all();                                  // nobody leaves (or frees tensor memory) while the peer still works
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2cta(tmem_base, kMmaTmemCols);
    }
}

}  // namespace mrag

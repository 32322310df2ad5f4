// scan_mma256w.cuh -- the CTA-pair candidate scan with 128-row tiles (tcgen05 cta_group::2, M = 256, N = 128).
//
// Same statement, same barriers and same select code as scan_mma256.cuh; what changes is the shape of one MMA.
// Measured on the 64-row kernel (r1q, idesc N overridden, results discarded, 10M x 768, 256 queries): a pass takes
// 2.8 / 2.9 / 3.96 ms at N = 16 / 32 / 64 and 5.1 ms at N = 128 for TWICE the rows, i.e. most of a 64-row MMA is a
// fixed cost and the pair cannot keep up with HBM (2.35 ms per pass).  Twice the rows per MMA halves that fixed cost
// per corpus byte:
//     D[256 queries x 128 rows] += A[256 x 16] * B[128 x 16]^T        each CTA streams 64 rows of every tile
// Two 128-column accumulators (MMAs of tile t+1 overlap the read-out of tile t: with a single accumulator the
// tensor cores idled ~2000 cycles per tile, measured) leave 256 tensor-memory columns = 8 k-blocks for the queries.
// The FIRST ks = kblocks - 8 k-blocks of the queries therefore live in SHARED memory (canonical K-major
// SWIZZLE_128B tiles, 16 KB per k-block, written once by the select warps) and their MMAs take both operands from
// shared memory; the rest stay in tensor memory as in the other scans.  (768 dims: 4 k-blocks = 64 KB.)
// EIGHT select warps per CTA -- two per lane quarter, one for accumulator columns 0..63 and one for 64..127 -- so a
// query has two threads (two candidate lists, two partial results: P = 2 * pairs).  Candidate buffers live in global
// memory (2 x 128 x cap keys per CTA).
// Result (profiles/r1q_*): 3113 SM cycles per 128-row tile for 48 MMAs of 64 cycles (128 x 128 x 16 MACs per SM at
// 4096 MAC/clk), tensor pipe 96 % active, 2.88 ms per pass
// at the ~1.17 GHz the SMs hold under this load = 1365 TFLOP/s, 0.98 of the sustained cuBLAS bf16 rate on this pool.
#pragma once
#include "scan_mma256.cuh"

namespace mrag {

constexpr int kMmaWThreads = 320;                            // TMA, MMA, 8 select warps
constexpr int kMmaWTileRows = 128;                           // UMMA N
constexpr int kMmaWHalfRows = 64;                            // rows of a tile each CTA loads (= the 64 x 64 box of tmap)
constexpr int kMmaWStageBytes = kMmaWHalfRows * 128;         // 8 KB per k-block per CTA
constexpr uint32_t kMmaWIdesc = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(kMmaWTileRows >> 3) << 17) | (uint32_t(256 >> 4) << 24);

constexpr int kMmaWTmemKBlocks = 8;                          // query k-blocks kept in tensor memory (256 columns)
constexpr int kMmaWDCol0 = 256;                              // two accumulators: columns 256..383, 384..511
constexpr int kMmaWATileBytes = kMma128Queries * 128;        // 16 KB: one k-block of this CTA's queries in shared memory

inline int mma256w_smem_kblocks(int ld) { return std::max(0, ld / kMmaKBlock - kMmaWTmemKBlocks); }
inline size_t mma256w_smem_bytes(int stages, int kbs, int ks) {
    return 1024 /*align slack*/ + size_t(ks) * kMmaWATileBytes + size_t(stages) * kbs * kMmaWStageBytes +
           size_t(kMma128InvSlots) * kMmaWTileRows * 4 + 2048 /*barriers*/;
}

// both operands from shared memory (the query k-blocks that do not fit into tensor memory)
MRAG_DEVINL void umma_ss_bf16_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}

// tcgen05.ld of 32 columns into r[o .. o+31] (o a compile-time constant: the array must stay in registers)
#define MRAG_TMEM_LD32O(r, o, taddr)                                                                                  \
    asm volatile(                                                                                                    \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                    \
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28," \
        "%29,%30,%31}, [%32];"                                                                                       \
        : "=r"(r[o + 0]), "=r"(r[o + 1]), "=r"(r[o + 2]), "=r"(r[o + 3]), "=r"(r[o + 4]), "=r"(r[o + 5]), "=r"(r[o + 6]),     \
          "=r"(r[o + 7]), "=r"(r[o + 8]), "=r"(r[o + 9]), "=r"(r[o + 10]), "=r"(r[o + 11]), "=r"(r[o + 12]), "=r"(r[o + 13]), \
          "=r"(r[o + 14]), "=r"(r[o + 15]), "=r"(r[o + 16]), "=r"(r[o + 17]), "=r"(r[o + 18]), "=r"(r[o + 19]),               \
          "=r"(r[o + 20]), "=r"(r[o + 21]), "=r"(r[o + 22]), "=r"(r[o + 23]), "=r"(r[o + 24]), "=r"(r[o + 25]),               \
          "=r"(r[o + 26]), "=r"(r[o + 27]), "=r"(r[o + 28]), "=r"(r[o + 29]), "=r"(r[o + 30]), "=r"(r[o + 31])                \
        : "r"(taddr)                                                                                                 \
        : "memory")

// -DMRAG_WSTATS=1: cycle accounting of pair 0's leader (debug builds only; a.stats[8..16))
#if defined(MRAG_WSTATS) && MRAG_WSTATS
#define MRAG_WSTAT_BEGIN() const bool ws_on = a.stats && blockIdx.x == 0 && lane == 0 && (warp == 1 || warp == 2); long long ws_t0 = 0, ws_begin = clock64(); unsigned long long ws_acc[16] = {}
#define MRAG_WSTAT_T0() ws_t0 = clock64()
#define MRAG_WSTAT_ADD(i) ws_acc[i] += (unsigned long long)(clock64() - ws_t0)
#define MRAG_WSTAT_END(i) do { if (ws_on) { a.stats[i] = (unsigned long long)(clock64() - ws_begin); for (int j_ = 0; j_ < 16; ++j_) if (ws_acc[j_]) a.stats[j_] = ws_acc[j_]; } } while (0)
#else
#define MRAG_WSTAT_BEGIN()
#define MRAG_WSTAT_T0()
#define MRAG_WSTAT_ADD(i)
#define MRAG_WSTAT_END(i)
#endif

MRAG_DEVINL float4 lds128_opaque(const float* p) {
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(smem_u32(p)) : "memory");
    return r;
}

// rows 8G .. 8G+7 of a thread's 64 accumulator columns, kept RAW (d = q.x, not yet scaled by 1/|x|): the scores are
// formed here, on the rare path, so that the hot loop holds one 64-register array instead of two
template <int G>
MRAG_DEVINL void walk_group_raw(const uint32_t (&d)[64], const float* inv64, uint32_t mbits, float& thr, SelState& st,
                                uint64_t* mybuf, uint64_t* cand_warp, int cap, int k, int lane, int64_t r0, uint32_t* gslot) {
    // (an opaque load: otherwise the compiler shares these products with the hot loop's and keeps all 64 on the stack)
    const float4 iva = lds128_opaque(inv64 + 8 * G), ivb = lds128_opaque(inv64 + 8 * G + 4);
    const float iv[8] = {iva.x, iva.y, iva.z, iva.w, ivb.x, ivb.y, ivb.z, ivb.w};
    float v[8];
    asm volatile("" : "+r"(mbits));                  // likewise for the mask bits
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] = ((mbits >> c) & 1u) ? __uint_as_float(d[8 * G + c]) * iv[c] : -CUDART_INF_F;
    int c_start = 0;
    for (;;) {
        int ovf = 8;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            if (c >= c_start && v[c] > thr) {
                if (st.cnt < cap) mybuf[st.cnt++] = make_key(v[c], uint32_t(r0 + 8 * G + c));
                else ovf = min(ovf, c);
            }
        }
        const unsigned full = __ballot_sync(kFull, st.cnt == cap);
        if (!full) break;
        st = select_compact128(st, full, cand_warp, cap, k, lane);
        if ((full >> lane) & 1u) atomicMax(gslot, f2ord(st.thr_s));
        thr = fmaxf(thr, st.thr_s);
        c_start = ovf;
        if (!__any_sync(kFull, ovf < 8)) break;
    }
}

// MmaArgs as scan_mma256: a.q0 = first query of the PAIR, a.nq <= 256, a.P = 2 * pairs (partial lists per query),
// a.gcand = [gridDim][2][128][cap] keys (required), tmap = the corpus tensor map with 64 x 64 boxes.
template <int KBS>
// (10 warps = 3 on two of the SM's four sub-partitions: 16384 / 3 registers per warp = 168 per thread)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kMmaWThreads, 1)
scan_mma256w_kernel(const __grid_constant__ CUtensorMap tmap, const MmaArgs a) {
    extern __shared__ __align__(1024) unsigned char mma_smem[];
    unsigned char* smem = mma_smem + ((1024u - (smem_u32(mma_smem) & 1023u)) & 1023u);
    const int kblocks = a.ld / kMmaKBlock;
    const int ks = max(0, kblocks - kMmaWTmemKBlocks);                                    // query k-blocks in shared memory
    unsigned char* a_smem = smem;                                                         // [ks][128 queries x 128 B], swizzled
    unsigned char* stage_base = smem + size_t(ks) * kMmaWATileBytes;
    constexpr int kbs = KBS;
    constexpr int stage_bytes = kbs * kMmaWStageBytes;
    float* xinv = reinterpret_cast<float*>(stage_base + size_t(a.stages) * stage_bytes); // [8][128] 1/|x| of a tile's rows
    uint64_t* bars = reinterpret_cast<uint64_t*>(xinv + kMma128InvSlots * kMmaWTileRows);
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
    const int64_t num_tiles = (a.n + kMmaWTileRows - 1) / kMmaWTileRows;
    const int64_t nwords = (a.n + 31) >> 5;
    const int64_t G = npairs;
    const int64_t t_first = pair;

    auto tile_mask = [&](int64_t t) -> uint4 {
        uint4 m = make_uint4(0u, 0u, 0u, 0u);
        if (t < num_tiles) {
            const int64_t w0 = 4 * t;
            m.x = __ldg(a.mask + w0);
            if (w0 + 1 < nwords) m.y = __ldg(a.mask + w0 + 1);
            if (w0 + 2 < nwords) m.z = __ldg(a.mask + w0 + 2);
            if (w0 + 3 < nwords) m.w = __ldg(a.mask + w0 + 3);
        }
        return m;
    };
    auto any_row = [](const uint4& m) -> bool { return (m.x | m.y | m.z | m.w) != 0u; };

    if (tid == 0) {
        for (int s = 0; s < a.stages; ++s) { mbar_init(&full_bar[s], 2); mbar_init(&empty_bar[s], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 16); }
        for (int i = 0; i < kMma128InvSlots; ++i) { mbar_init(&ifull_bar[i], 1); mbar_init(&iempty_bar[i], 8); }
        fence_barrier_init();
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap)) : "memory");
    }
    cluster_sync_all();                                  // both CTAs' barriers exist before anything remote touches them
    if (warp == 1) tmem_alloc_2cta(tmem_holder, kMmaTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    const int quarter = warp & 3;                        // TMEM lanes [32 * quarter, +32) (hardware: warp id mod 4)
    const int half = (warp - 2) >> 2;                    // select warps 2..5: accumulator columns 0..63, 6..9: 64..127
    const int qi = quarter * 32 + lane;                  // query slot inside this CTA
    const int qbase = a.q0 + int(rank) * kMma128Queries; // first query of this CTA
    const int nq_cta = max(0, min(kMma128Queries, a.q0 + a.nq - qbase));

    // ---- this CTA's queries -> shared memory (k-blocks < ks) and tensor memory (the two warps of a quarter take alternate k-blocks)
    if (warp >= 2) {
        const bool live = qi < nq_cta;
        const float* qrow = a.q + size_t(live ? qbase + qi : a.q0) * a.ld;
        float qs = live ? a.qinv[qbase + qi] : 0.0f;
        if (isinf(qs)) qs = 0.0f;
        for (int c0 = half * 32; c0 < a.ld / 2; c0 += 64) {
            uint32_t r[32];
#pragma unroll
            for (int v = 0; v < 16; ++v) {
                float4 f = live ? __ldg(reinterpret_cast<const float4*>(qrow + c0 * 2) + v) : make_float4(0.f, 0.f, 0.f, 0.f);
                r[2 * v] = pack_bf16x2(f.x * qs, f.y * qs);
                r[2 * v + 1] = pack_bf16x2(f.z * qs, f.w * qs);
            }
            const int kb = c0 >> 5;
            if (kb < ks) {
                // shared memory, the layout TMA's SWIZZLE_128B produces: row qi = 128 B, 16-byte chunk c at c ^ (qi & 7)
                const uint32_t row = smem_u32(a_smem) + uint32_t(kb) * kMmaWATileBytes + uint32_t(qi >> 3) * 1024u + uint32_t(qi & 7) * 128u;
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + (uint32_t(c ^ (qi & 7)) << 4)), "r"(r[4 * c]),
                                 "r"(r[4 * c + 1]), "r"(r[4 * c + 2]), "r"(r[4 * c + 3])
                                 : "memory");
            } else {
                const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t((kb - ks) * 32);
                MRAG_TMEM_ST32(taddr, r);
            }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the tensor cores read a_smem through the async proxy
    }
    tc_fence_before();
    cluster_sync_all();                                  // A is in place in BOTH CTAs before the leader issues MMAs
    tc_fence_after();

    if (warp == 0) {
        // ================= TMA producer (each CTA: its 64 rows of every tile + the tile's 1/|x|) =================
        int s = 0, is = 0;
        uint32_t ph = 0, iph = 0;
        uint4 m = tile_mask(t_first);
        for (int64_t t = t_first; t < num_tiles; t += G) {
            const uint4 mn = tile_mask(t + G);
            if (any_row(m)) {
                mbar_wait(&iempty_bar[is], iph ^ 1u, slp);
                if (elect_one()) {
                    mbar_expect_tx(&ifull_bar[is], kMmaWTileRows * 4);
                    bulk_load_1d(xinv + is * kMmaWTileRows, a.inv_norm + t * kMmaWTileRows, kMmaWTileRows * 4, &ifull_bar[is]);
                }
                __syncwarp();
                if (++is == kMma128InvSlots) { is = 0; iph ^= 1u; }
                for (int kb = 0; kb < kblocks; kb += kbs) {
                    mbar_wait(&empty_bar[s], ph ^ 1u, slp);
                    if (elect_one()) {
                        // the leader announces the bytes of BOTH halves; the peer only arrives
                        if (leader) mbar_expect_tx(&full_bar[s], 2 * stage_bytes);
                        else mbar_arrive_leader(&full_bar[s]);
#pragma unroll
                        for (int j = 0; j < kbs; ++j)
                            tma_load_2d_2cta(stage_base + size_t(s) * stage_bytes + size_t(j) * kMmaWStageBytes, &tmap,
                                             (kb + j) * kMmaKBlock, int(t * kMmaWTileRows + rank * kMmaWHalfRows), &full_bar[s]);
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
            const uint64_t adesc0 = make_sw128_desc(smem_u32(a_smem));
            MRAG_WSTAT_BEGIN();
            uint4 m = tile_mask(t_first);
            for (int64_t t = t_first; t < num_tiles; t += G) {
                const uint4 mn = tile_mask(t + G);
                if (any_row(m)) {
                    MRAG_WSTAT_T0();
                    mbar_wait(&tempty_bar[as], aph ^ 1u, slp);     // all 16 select warps of the pair have read this buffer
                    MRAG_WSTAT_ADD(9);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + uint32_t(kMmaWDCol0 + as * kMmaWTileRows);
                    for (int kb = 0; kb < kblocks; kb += kbs) {
                        MRAG_WSTAT_T0();
                        mbar_wait(&full_bar[s], ph, slp);
                        MRAG_WSTAT_ADD(10);
                        tc_fence_after();
                        const uint64_t sdesc = bdesc0 + uint64_t((size_t(s) * stage_bytes) >> 4);
#pragma unroll
                        for (int j = 0; j < KBS; ++j) {
                            constexpr int kStageDesc = kMmaWStageBytes >> 4;
                            const uint64_t bd = sdesc + uint64_t(j * kStageDesc);
                            const uint32_t acc0 = (kb + j) != 0 ? 1u : 0u;
                            if (kb + j < ks) {
                                const uint64_t ad = adesc0 + uint64_t((kb + j) * (kMmaWATileBytes >> 4));
                                umma_ss_bf16_2cta(d_tmem, ad, bd, kMmaWIdesc, acc0);
                                umma_ss_bf16_2cta(d_tmem, ad + 2, bd + 2, kMmaWIdesc, 1u);
                                umma_ss_bf16_2cta(d_tmem, ad + 4, bd + 4, kMmaWIdesc, 1u);
                                umma_ss_bf16_2cta(d_tmem, ad + 6, bd + 6, kMmaWIdesc, 1u);
                            } else {
                                const uint32_t at = tmem_base + uint32_t((kb + j - ks) * (kMmaKBlock / 2));
                                umma_ts_bf16_2cta(d_tmem, at, bd, kMmaWIdesc, acc0);
                                umma_ts_bf16_2cta(d_tmem, at + 8, bd + 2, kMmaWIdesc, 1u);
                                umma_ts_bf16_2cta(d_tmem, at + 16, bd + 4, kMmaWIdesc, 1u);
                                umma_ts_bf16_2cta(d_tmem, at + 24, bd + 6, kMmaWIdesc, 1u);
                            }
                        }
                        umma_commit_2cta(&empty_bar[s]);
                        if (kb + kbs >= kblocks) umma_commit_2cta(&tfull_bar[as]);
                        if (++s == a.stages) { s = 0; ph ^= 1u; }
                    }
                    if (++as == 2) { as = 0; aph ^= 1u; }
                }
                m = mn;
            }
            MRAG_WSTAT_END(8);
        }
        __syncwarp();
    } else {
        // ================= select (warps 2..9 of each CTA): thread = (query, half of the tile's rows) =================
        const uint32_t lane_addr = uint32_t(quarter * 32) << 16;
        const uint32_t dcol0 = uint32_t(kMmaWDCol0 + half * 64);
        int is = 0, as = 0;
        uint32_t aph = 0, iph = 0;
        const bool live = qi < nq_cta;
        const bool warp_live = quarter * 32 < nq_cta;
        const float qinv = live ? a.qinv[qbase + qi] : 0.0f;
        uint64_t* cand_warp = a.gcand + ((size_t(blockIdx.x) * 2 + half) * kMma128Queries + quarter * 32) * a.cap;
        uint64_t* mybuf = cand_warp + size_t(lane) * a.cap;
        const int cap = a.cap;
        SelState st;
        st.cnt = 0;
        st.thr_s = (live && !isinf(qinv)) ? -CUDART_INF_F : CUDART_INF_F;
        uint32_t* gslot = a.gthr + (live ? qbase + qi : a.q0);
        MRAG_WSTAT_BEGIN();
        uint4 m4 = tile_mask(t_first);
        for (int64_t t = t_first; t < num_tiles; t += G) {
            const uint4 mn = tile_mask(t + G);
            if (any_row(m4)) {
                const uint2 m = half ? make_uint2(m4.z, m4.w) : make_uint2(m4.x, m4.y);
                const int64_t r0 = t * kMmaWTileRows + half * 64;
                uint32_t d[64];                                     // this thread's 64 accumulator columns, raw
                float bestg[8];
                uint32_t gord;
                asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(gord) : "l"(gslot) : "memory");
                MRAG_WSTAT_T0();
                mbar_wait(&ifull_bar[is], iph, slp);
                MRAG_WSTAT_ADD(14);
                MRAG_WSTAT_T0();
                mbar_wait(&tfull_bar[as], aph, slp);
                MRAG_WSTAT_ADD(13);
                MRAG_WSTAT_T0();
                tc_fence_after();
                if (warp_live) {
                    const uint32_t dcol = dcol0 + uint32_t(as * kMmaWTileRows);
                    MRAG_TMEM_LD32O(d, 0, tmem_base + lane_addr + dcol);
                    MRAG_TMEM_LD32O(d, 32, tmem_base + lane_addr + dcol + 32u);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                }
                // the accumulator is in registers: hand the buffer back before any arithmetic
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_leader(&tempty_bar[as]);
                MRAG_WSTAT_ADD(15);
                if (++as == 2) { as = 0; aph ^= 1u; }
                const float* inv64 = xinv + is * kMmaWTileRows + half * 64;
                if (warp_live) {
                    const float4* inv4 = reinterpret_cast<const float4*>(inv64);
                    const bool all_rows = (m.x & m.y) == 0xffffffffu;
#pragma unroll
                    for (int g = 0; g < 8; ++g) {                 // rows 8g .. 8g+7: scale, (mask,) group maximum
                        const float4 iva = inv4[2 * g], ivb = inv4[2 * g + 1];
                        const float iv[8] = {iva.x, iva.y, iva.z, iva.w, ivb.x, ivb.y, ivb.z, ivb.w};
                        const uint32_t mw = (g < 4 ? m.x : m.y) >> ((8 * g) & 31);
                        float v[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            v[i] = __uint_as_float(d[8 * g + i]) * iv[i];
                            if (!all_rows && !((mw >> i) & 1u)) v[i] = -CUDART_INF_F;
                        }
                        bestg[g] = fmaxf(fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3])), fmaxf(fmaxf(v[4], v[5]), fmaxf(v[6], v[7])));
                    }
                    float thr = fmaxf(st.thr_s, gord ? ord2f(gord - 1u) : -CUDART_INF_F);
#define MRAG_WALK_GROUP(G_)                                                                                         \
                    if (__any_sync(kFull, bestg[G_] > thr))                                                         \
                        walk_group_raw<G_>(d, inv64, ((G_) < 4 ? m.x : m.y) >> ((8 * (G_)) & 31), thr, st, mybuf, cand_warp, cap, a.k, lane, r0, gslot);
                    MRAG_WALK_GROUP(0) MRAG_WALK_GROUP(1) MRAG_WALK_GROUP(2) MRAG_WALK_GROUP(3)
                    MRAG_WALK_GROUP(4) MRAG_WALK_GROUP(5) MRAG_WALK_GROUP(6) MRAG_WALK_GROUP(7)
#undef MRAG_WALK_GROUP
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&iempty_bar[is]);          // the tile's 1/|x| slot is free
                if (++is == kMma128InvSlots) { is = 0; iph ^= 1u; }
            }
            m4 = mn;
        }
        if (warp == 2) { MRAG_WSTAT_END(12); }

        // ---- this (pair, half)'s sorted candidate list per query
        __syncwarp();
        for (int L = 0; L < 32; ++L) {
            const int qL = quarter * 32 + L;
            if (qL >= nq_cta) break;
            const int n = __shfl_sync(kFull, st.cnt, L);
            uint64_t* b = cand_warp + size_t(L) * a.cap;
            warp_rank_select_n(b, n, a.kp, lane);
            uint64_t* out = a.part + (size_t(qbase + qL) * a.P + size_t(pair) * 2 + half) * a.kp;
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

// scan_mma.cuh -- K1' + K3: the tensor-core scan for bf16 corpora (sm_100a: TMA + tcgen05 + TMEM).
//
// Replaces the same sequential-scan plan as scan_gemv.cuh (corpus_search.py:1525-1536,
// vector_store.py:274-287) for up to 64 queries per pass while staying HBM-bound: every corpus
// byte is read once, by TMA, into a shared-memory ring; the dot products run on the tensor cores.
//
// Orientation (transposed w.r.t. a textbook GEMM so that a thread owns a QUERY, not a row):
//     D[128 lanes x 64 rows] += A[128 lanes x K] * B[64 rows x K]^T          (tcgen05.mma kind::f16)
//   A = the queries, resident in TENSOR MEMORY for the whole kernel (K/2 columns of packed bf16x2).
//       An fp32 query q is split EXACTLY-ish into bf16 hi + bf16 lo (q = hi + lo to 16 mantissa
//       bits); lane 64 + j holds hi of query j, lane j holds lo of query j, so the final score
//       carries fp32-level accuracy although the MMA operands are bf16.
//   B = a 64-row tile of the corpus, K-major, 64 bf16 (128 B) per shared-memory row, SWIZZLE_128B,
//       streamed by TMA: one 8 KB box per (tile, k-block), ring of `stages` boxes.
//   D = fp32 accumulators in tensor memory, two buffers of 64 columns (MMA of tile t+1 overlaps the
//       select of tile t).  TMEM budget: 384 (A, K = 768) + 2 x 64 (D) = 512 columns.
//
// Warp roles (192 threads, 1 CTA per SM, persistent over row tiles):
//   warp 0      TMA producer            warp 1      MMA issuer (one elected lane) + TMEM owner
//   warps 4,5   "lo" epilogue: TMEM lanes 0..63 (lo halves) -> shared-memory exchange buffer
//   warps 2,3   "hi" epilogue: TMEM lanes 64..127 (hi halves), add lo, scale by 1/|x| * 1/|q|,
//               per-THREAD top-k (a thread sees all scores of its query in row order: no
//               cross-thread traffic except the warp-cooperative compaction of a full buffer).
//               The two heavy warps sit alone on scheduler partitions 2 and 3 (warp % 4); the
//               waiting roles share partitions 0 and 1.
// Tiles whose 64 mask bits are all clear are skipped by every role (never loaded).
//
// KS = true ("k-split", rows of 769 .. 1536 elements, kind "mma_ks"): the queries of such rows need ld/2 > 384 tensor-
// memory columns, so TWO CTAs of a thread-block cluster share every row tile along K: CTA c keeps the k-blocks
// [c * ceil(kb/2), ...) of the 64 hi/lo queries in ITS tensor memory, streams only those columns of the tile (a TMA box
// is a column block, so each corpus byte is still read once, by one CTA) and accumulates a PARTIAL dot product.  The
// two CTAs alternate as "leader" tile by tile: the helper's hi warps add their hi + lo partials into a staging buffer
// and ONE bulk copy (cp.async.bulk shared::cta -> shared::cluster, complete_tx on the leader's mbarrier: 16 KB per
// tile against 192 KB of corpus) moves the 64 x 64 sums into the leader's shared memory; the leader adds them to its
// own partials and runs the select.  Each CTA thus selects every other tile of the pair and writes its own candidate
// lists, exactly like a CTA of the one-CTA kernel.  Sending is the job of two EXTRA warps (6, 7: tensor-memory quarters
// 2, 3 like the hi warps), so a CTA's select of tile i and its send of tile i + 1 overlap; with the hi warps doing both
// the two CTAs ping-ponged (A sends -> B selects -> B sends -> A selects ...) and one tile cost send + copy + select.  (A first version stored the sums with st.shared::cluster and
// signalled with release / acquire at cluster scope: MEMBAR.ALL.GPU + CCTL.IVALL per tile, 0.47 of the HBM rate.)
#pragma once
#include "common.cuh"
#include <cuda.h>
#include <math_constants.h>

namespace mrag {

#ifndef MRAG_SLEEP_REG
#define MRAG_SLEEP_REG 1
#endif
#ifndef MRAG_QREADY_BAR
// 1: the TMA producer starts streaming while the epilogue warps still load / split the queries into tensor memory; only
// the MMA issuer waits for them (an mbarrier instead of a block-wide sync): hides the ring's fill behind the query load
#define MRAG_QREADY_BAR 1
#endif
#ifndef MRAG_STAMPS
// %globaltimer stamps of CTA 0's phases (tools/stats_probe.py).  OFF in the shipped build: merely
// compiling the (never taken) stamp branches in costs the hot loop 2.3x (measured, r1f A/B run).
#define MRAG_STAMPS 0
#endif
constexpr int kMmaThreads = 192;
constexpr int kMmaKsThreads = 256;        // k-split pair: two more warps (6, 7) that SEND this CTA's partials on the tiles its peer leads
constexpr int kMmaTileRows = 64;          // UMMA N
constexpr int kMmaKBlock = 64;            // bf16 elements per smem row = 128 B = one swizzle span
constexpr int kMmaStageBytes = kMmaTileRows * 128;
constexpr int kMmaQueries = 64;           // queries per pass (hi / lo halves of the 128 TMEM lanes)
constexpr int kMmaMaxLd = 768;            // A needs ld/2 TMEM columns; 384 + 128 (D) = 512
constexpr int kMmaTmemCols = 512;
constexpr int kMmaDCol0 = 384;
constexpr int kMmaSlack = 64;             // candidate buffer = k + slack keys per query
constexpr long long kSpinCycles = 4000000000ll;   // ~2 s: a protocol bug traps instead of hanging the GPU

struct MmaArgs {
    int64_t n;               // rows in the shard
    int ld;                  // padded row length (multiple of 64, <= kMmaMaxLd)
    const uint32_t* mask;    // row bitmap (valid AND filter), ceil(n/32) words
    const float* inv_norm;   // [n] 1/|x| of the stored row (+inf for a zero row)
    const float* q;          // [*][ld] fp32 queries, zero padded
    const uint32_t* qhl;     // [*][2][ld/2] the same queries as packed bf16 hi / lo planes (query_prep_kernel); nullptr: split here
    const float* qinv;       // [*] 1/|q| (+inf for a zero query)
    const uint64_t* ub;      // [*] exclusive upper-bound key per query, or nullptr
    uint64_t* part;          // [*][P][kp] per-CTA sorted candidate lists
    uint32_t* gthr;          // [*] orderable(score) of the best k-th score any CTA has proven so far (0 = none)
    uint32_t* gmax;          // [*][gslots] group maxima (GroupBound below), or nullptr
    int ngroups;             // groups in use (= the k of the launch <= gslots); slots [ngroups, gslots) hold 0xFFFFFFFF
    int gslots;              // 16 or 64
    int q0, nq;              // queries [q0, q0+nq) handled by this launch (nq <= 64)
    int k, kp, P;
    int cap;                 // candidate buffer capacity per query (k + kMmaSlack)
    int stages;              // TMA ring depth
    unsigned long long* stats;   // optional counters: [0] tiles, [1] tiles with a candidate, [2] keys appended,
                                 // [3] compactions, [4] overflow retries  (summed over hi warps / lanes)
    int tile_mul;            // 1: every 64-row tile; m > 1: only tiles 0, m, 2m, ... (threshold sampling pass)
    // exact-rescan mode (rescore.cuh): serve the queries qlist[0 .. min(64, *qcount)) instead of [q0, q0 + nq);
    // *qcount == 0 makes the launch a no-op
    const int* qlist;
    const int* qcount;
    int kbs;                 // scan_mma256: k-blocks per pipeline stage (divides ld / 64)
    uint64_t* gcand;         // scan_mma128, large k: candidate buffers in GLOBAL memory [gridDim][128][cap] (appends are
                             // rare once a bound exists; the buffers of one CTA stay in its L1 / L2); nullptr = shared memory
    uint32_t sleep_ns;       // suspend-time hint of the barrier waits
    unsigned long long* tstamps;   // optional [16] %globaltimer stamps of CTA 0 (debugging aid)
};

constexpr int kMmaKsMaxLd = 1536;         // k-split pair: two CTAs x 384 tensor-memory columns of queries
inline size_t mma_smem_bytes(int stages, int cap, bool ksplit = false) {
    return 1024 /*align slack*/ + size_t(stages) * kMmaStageBytes + 2 * 64 * 64 * 4 /*exchange*/ +
           (ksplit ? 2 * 64 * 64 * 4 + 2 * 64 * 4 : 0) /*partials received from / staged for the peer CTA; 1/|x| ring of 4*/ +
           2 * 64 * 4 /*1/|x| of the tile*/ + size_t(kMmaQueries) * cap * 8 + 1024 /*barriers*/;
}

// ---- PTX wrappers ----------------------------------------------------------------------------
MRAG_DEVINL uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }

MRAG_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
MRAG_DEVINL void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
MRAG_DEVINL void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
MRAG_DEVINL bool mbar_try_wait(uint64_t* bar, uint32_t parity, uint32_t sleep_ns) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
#if MRAG_SLEEP_REG
        : "r"(smem_u32(bar)), "r"(parity), "r"(sleep_ns)      // suspend-time hint: sleep, do not spin
#else
        : "r"(smem_u32(bar)), "r"(parity), "n"(0x989680)
#endif
        : "memory");
    return ok != 0;
}
MRAG_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity, uint32_t sleep_ns = 0x989680u) {
    if (mbar_try_wait(bar, parity, sleep_ns)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity, sleep_ns)) {
        if (clock64() - t0 > kSpinCycles) __trap();
    }
}
MRAG_DEVINL unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// one lane of a CONVERGED warp (the compiler then knows the guarded block runs on a single lane and
// emits the uniform-datapath tcgen05 / TMA instructions without per-instruction election loops)
MRAG_DEVINL bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
        "elect.sync rx|px, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, px;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
MRAG_DEVINL void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
MRAG_DEVINL void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
MRAG_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
MRAG_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- thread-block cluster helpers of the k-split variant
MRAG_DEVINL uint32_t ks_cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
MRAG_DEVINL void ks_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
MRAG_DEVINL uint32_t ks_map_to_cta(uint32_t local_smem_addr, uint32_t cta_rank) {      // shared::cluster address in a peer CTA
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(cta_rank));
    return r;
}
// plain remote arrive (the consumer-release of a multicast pipeline): tells the peer its send buffer may be reused
MRAG_DEVINL void ks_arrive_remote(uint32_t cluster_bar_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar_addr) : "memory");
}
// bulk copy of `bytes` from this CTA's shared memory into a peer CTA's, completing on the PEER's mbarrier
MRAG_DEVINL void ks_bulk_copy_to_peer(uint32_t dst_cluster_addr, const void* src, uint32_t bytes, uint32_t bar_cluster_addr) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster_addr),
                 "r"(smem_u32(src)), "r"(bytes), "r"(bar_cluster_addr)
                 : "memory");
}
MRAG_DEVINL void ks_bar_send() { asm volatile("bar.sync 1, 64;" ::: "memory"); }    // the 64 threads of the two send warps

MRAG_DEVINL void tma_load_2d(void* dst, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar) {
    const uint64_t evict_first = 0x12F0000000000000ull;      // each corpus byte is used once
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(evict_first)
        : "memory");
}

MRAG_DEVINL void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
MRAG_DEVINL void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem desc]
MRAG_DEVINL void umma_ts_bf16(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
MRAG_DEVINL void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// K-major, SWIZZLE_128B operand tile: rows of 128 B, 8-row groups 1024 B apart (SBO), version 1
MRAG_DEVINL uint64_t make_sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= uint64_t((smem_addr & 0x3FFFFu) >> 4);            // start address
    d |= uint64_t(0) << 16;                                // leading byte offset (unused for swizzled K-major)
    d |= uint64_t(1024 >> 4) << 32;                        // stride byte offset
    d |= uint64_t(1) << 46;                                // descriptor version (sm_100)
    d |= uint64_t(2) << 61;                                // SWIZZLE_128B
    return d;
}
// kind::f16 instruction descriptor: D fp32, A/B bf16, both K-major, M = 128, N = 64
constexpr uint32_t kMmaIdesc = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(kMmaTileRows >> 3) << 17) | (uint32_t(128 >> 4) << 24);

#define MRAG_TMEM_LD64(r, taddr)                                                                                      \
    asm volatile(                                                                                                    \
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "                                                                    \
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28," \
        "%29,%30,%31,%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55," \
        "%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"                                                                   \
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),  \
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),     \
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),    \
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]),    \
          "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]),    \
          "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]),    \
          "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]),    \
          "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])                  \
        : "r"(taddr)                                                                                                 \
        : "memory")

#define MRAG_TMEM_LD32(r, taddr)                                                                                      \
    asm volatile(                                                                                                    \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                    \
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28," \
        "%29,%30,%31}, [%32];"                                                                                       \
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),  \
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),     \
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),    \
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                  \
        : "r"(taddr)                                                                                                 \
        : "memory")

#define MRAG_TMEM_ST32(taddr, r)                                                                                      \
    asm volatile(                                                                                                    \
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "                                                              \
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"   \
        "%29,%30,%31,%32};" ::"r"(taddr),                                                                            \
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), \
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),  \
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),  \
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])                                                               \
        : "memory")

MRAG_DEVINL uint32_t pack_bf16x2(float lo_elem, float hi_elem) {
    // element with the LOWER k index in the low half
    __nv_bfloat162 v = __floats2bfloat162_rn(lo_elem, hi_elem);
    return *reinterpret_cast<uint32_t*>(&v);
}

// ---- warp-cooperative selection of the `keep` largest of n unique keys in buf (rank sort) -----
// afterwards buf[0..min(n,keep)) holds them in descending order.  n <= 32 * kRankPerLane.
constexpr int kRankPerLane = 6;           // cap <= 192
MRAG_DEVINL void warp_rank_select(uint64_t* buf, int n, int keep, int lane) {
    uint64_t key[kRankPerLane];
    int rank[kRankPerLane];
#pragma unroll
    for (int e = 0; e < kRankPerLane; ++e) {
        const int i = lane + 32 * e;
        key[e] = (i < n) ? buf[i] : 0ull;
        rank[e] = 0;
    }
    for (int j = 0; j < n; ++j) {
        const uint64_t kj = buf[j];           // broadcast read
#pragma unroll
        for (int e = 0; e < kRankPerLane; ++e) rank[e] += (kj > key[e]) ? 1 : 0;
    }
    __syncwarp();
#pragma unroll
    for (int e = 0; e < kRankPerLane; ++e) {
        const int i = lane + 32 * e;
        if (i < n && rank[e] < keep) buf[rank[e]] = key[e];
    }
    __syncwarp();
}

// ---- cross-CTA admission bound from group maxima (replaces a threshold-sampling launch + its merge) ----------------
// The producers of a launch (CTAs, CTA pairs or half-tiles of a pair: whatever owns DISJOINT rows) are dealt into k
// groups (producer % k); gmax[q][g] is the best score any producer of group g has seen for query q.  The k group
// maxima are k DIFFERENT rows, so their minimum is a score that k rows reach: nothing below it can be in the top-k.
// Since the k best rows of what has been scanned so far mostly sit in different producers, this tracks the k-th best of
// the UNION of all producers' rows (within a small factor in rank) and keeps tightening during the scan.  A thread
// posts only when its running maximum rises (~ln(tiles) times); the poster then folds the new minimum into gthr[q], the
// one word every thread reads per tile.  (A first version had every thread read the group words every tile: +0.37 us
// per tile, r2l.  Measured r2m, 1.25M x 768, 64 queries, k = 10: 0.349 ms per step against 0.378 ms with the sampling
// launch; 6.7k of 39k tiles take the slow path against 34.5k.)
struct GroupBound {
    const uint32_t* gq;      // the query's slots
    uint32_t* gpost;         // this producer's slot
    int nvec;                // slots / 4
    float posted;            // best score this thread has posted
    bool on;
};
MRAG_DEVINL GroupBound group_bound_init(const MmaArgs& a, int qsrc, int producer, bool can_post) {
    GroupBound g;
    g.on = a.gmax != nullptr && can_post;
    g.gq = a.gmax + size_t(qsrc) * a.gslots;
    g.gpost = a.gmax + size_t(qsrc) * a.gslots + (a.ngroups > 0 ? producer % a.ngroups : 0);
    g.nvec = a.gslots >> 2;
    g.posted = -CUDART_INF_F;
    return g;
}
MRAG_DEVINL void group_bound_post(GroupBound& g, float best, uint32_t* gslot) {
    if (g.on && best > g.posted) {
        g.posted = best;
        const uint32_t mine = f2ord(best);
        if (atomicMax(g.gpost, mine) < mine) {
            // my group's maximum rose, so the minimum over the groups may have (a torn snapshot only under-estimates; of
            // two concurrent posters at least one sees both updates, because each reads after its own atomic returned)
            uint32_t gmin = 0xFFFFFFFFu;
            for (int i = 0; i < g.nvec; ++i) {
                uint4 v;
                asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];"
                             : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(g.gq + 4 * i) : "memory");
                gmin = min(gmin, min(min(v.x, v.y), min(v.z, v.w)));
            }
            if (gmin) atomicMax(gslot, gmin);            // 0: some group has not posted yet
        }
    }
}

struct SelState {
    float thr_s;     // admit scores > thr_s (score of the k-th best key of the last compaction)
    int cnt;         // keys in this thread's buffer
};

// Rare path of the select: the buffers of the lanes in `full` are at capacity; the whole warp
// compacts each of them to its k best keys and raises that lane's threshold.
__device__ __noinline__ SelState select_compact(SelState st, unsigned full, uint64_t* cand_warp, int cap, int k, int lane) {
    while (full) {
        const int L = __ffs(full) - 1;
        full &= full - 1;
        __syncwarp();
        uint64_t* b = cand_warp + size_t(L) * cap;
        warp_rank_select(b, cap, k, lane);
        const uint64_t kth = b[k - 1];
        if (lane == L) { st.cnt = k; st.thr_s = key_score(kth); }
    }
    return st;
}

// rows 8G .. 8G+7 of the tile: append what beats the threshold; a full buffer is compacted and the walk resumes.
// (a function template so that every index into sc[] is a compile-time constant: the array must stay in registers)
template <int G>
MRAG_DEVINL void walk_group64(const float (&sc)[64], float& thr, SelState& st, uint64_t* mybuf, uint64_t* cand_warp,
                              int cap, int k, int lane, int64_t r0, uint32_t* gslot, uint64_t ubk,
                              unsigned long long& n_keys, unsigned long long& n_compact, unsigned long long& n_retry) {
    const int cnt_before = st.cnt;
    int c_start = 0;
    for (;;) {
        int ovf = 8;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            if (c >= c_start && sc[8 * G + c] > thr) {
                const uint64_t key = make_key(sc[8 * G + c], uint32_t(r0 + 8 * G + c));
                if (key < ubk) {
                    if (st.cnt < cap) mybuf[st.cnt++] = key;
                    else ovf = min(ovf, c);
                }
            }
        }
        const unsigned full = __ballot_sync(kFull, st.cnt == cap);
        if (!full) { n_keys += st.cnt - cnt_before; break; }
        n_compact += (full >> lane) & 1u;
        ++n_retry;
        st = select_compact(st, full, cand_warp, cap, k, lane);
        if ((full >> lane) & 1u) atomicMax(gslot, f2ord(st.thr_s));
        thr = fmaxf(thr, st.thr_s);
        c_start = ovf;
        if (!__any_sync(kFull, ovf < 8)) break;
    }
}

// ----------------------------------------------------------------------------------------------
// KREG = 0: candidates go to per-query shared-memory buffers of k + 64 keys, compacted when full
//           (k up to 128; on large shards the admission bound comes from the cross-CTA group maxima -- GroupBound above --
//            or, with MRAG_GMAX=0 and in the rounds of k > 128, from a sampling pass first).
// KREG > 0: k <= KREG; every thread keeps the sorted top-KREG keys of its query in REGISTERS
//           (branch-free insertion), so its threshold is always the exact k-th best so far:
//           ~k ln(n/k) insertions per query and CTA, no buffers, no compaction, no sampling pass.
constexpr int kMmaRegK = 16;
#ifndef MRAG_SAMPLE_K
#define MRAG_SAMPLE_K 4
#endif
// The threshold-sampling passes keep only the best kMmaSampleK scores per (query, CTA): the bound is the k'-th best of
// the UNION over ~148 CTAs, of which a CTA holds k'/148 on average, so 4 per CTA lose (almost) nothing -- and a list
// that does truncate only makes the bound slightly looser, never wrong.  The sorted insertion is the sampling kernels'
// whole epilogue (fully unrolled: 64 rows x K compare-exchange steps, instruction-fetch bound at K = 16, r1q ncu).
constexpr int kMmaSampleK = MRAG_SAMPLE_K;

// SO (score only, KREG > 0): the sampling pass needs a bound, not rows: 32-bit orderable scores in the
//           registers (half the insertion work); the lists it writes carry synthetic unique low words.
template <int KREG, bool SO = false, bool KS = false>
__global__ void __launch_bounds__(KS ? kMmaKsThreads : kMmaThreads, 1) scan_mma_kernel(const __grid_constant__ CUtensorMap tmap, const MmaArgs a) {
    extern __shared__ __align__(1024) unsigned char mma_smem[];
    // SWIZZLE_128B tiles need 1024-byte alignment; stay in the shared address space (no integer casts)
    unsigned char* smem = mma_smem + ((1024u - (smem_u32(mma_smem) & 1023u)) & 1023u);
    unsigned char* stage_base = smem;                                            // stages * 8 KB
    float* xbuf = reinterpret_cast<float*>(smem + size_t(a.stages) * kMmaStageBytes);   // [2][64 rows][64 queries] lo parts
    float* ybuf = xbuf + 2 * 64 * 64;                                            // KS: [64 rows][64 queries] partial dots received from the peer CTA,
    float* sbuf = ybuf + 64 * 64;                                                //     followed by the staging buffer of the partials this CTA sends
    float* xinv = ybuf + (KS ? 2 * 64 * 64 : 0);                                 // [2][64 rows] 1/|x|, NaN = masked row
    uint64_t* cand = reinterpret_cast<uint64_t*>(xinv + (KS ? 4 : 2) * 64);      // [64 queries][cap]   (KS: 1/|x| ring of 4 tiles)
    uint64_t* bars = cand + size_t(kMmaQueries) * a.cap;
    uint64_t* full_bar = bars;                       // [stages]   TMA -> MMA
    uint64_t* empty_bar = full_bar + a.stages;       // [stages]   MMA -> TMA
    uint64_t* tfull_bar = empty_bar + a.stages;      // [2]        MMA -> epilogue
    uint64_t* tempty_bar = tfull_bar + 2;            // [2]        epilogue -> MMA
    uint64_t* xfull_bar = tempty_bar + 2;            // [2]        lo warps -> hi warps
    uint64_t* xempty_bar = xfull_bar + 2;            // [2]        hi warps -> lo warps
    uint64_t* yfull_bar = xempty_bar + 2;            // [1]  KS:   the peer's bulk copy has landed in ybuf (expect_tx by me + complete_tx)
    uint64_t* yempty_bar = yfull_bar + 2;            // [1]  KS:   the peer has consumed what I sent (its hi warps arrive here)
    uint64_t* qready_bar = yempty_bar + 2;           // [1]        the queries are in tensor memory (epilogue warps -> MMA issuer)
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(qready_bar + 1);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t slp = a.sleep_ns;
    const int nq_eff = a.qlist ? min(kMmaQueries, *a.qcount) : a.nq;
    if (nq_eff == 0) return;                            // nothing to rescan (uniform over the grid: no barrier is pending)
#if MRAG_STAMPS
    auto stamp = [&](int i) { if (a.tstamps && blockIdx.x == 0 && lane == 0) a.tstamps[i] = globaltimer_ns(); };
#else
    auto stamp = [&](int) {};
#endif
    if (tid == 0) stamp(0);
    // KS: this CTA's share of the k-blocks (the pair splits every row tile along K)
    const uint32_t ks_rank = KS ? ks_cluster_rank() : 0u;
    const int all_kblocks = a.ld / kMmaKBlock;
    const int kb0 = KS ? (ks_rank ? (all_kblocks + 1) / 2 : 0) : 0;
    const int kblocks = KS ? (ks_rank ? all_kblocks / 2 : (all_kblocks + 1) / 2) : all_kblocks;
    const int64_t all_tiles = (a.n + kMmaTileRows - 1) / kMmaTileRows;
    const int64_t nwords = (a.n + 31) >> 5;
    // tile index space of this launch: t = 0, tile_mul, 2*tile_mul, ...; CTA b (KS: pair b) takes every gridDim-th of them
    const int64_t tmul = a.tile_mul;
    const int64_t num_tiles = all_tiles;
    const int64_t G = int64_t(KS ? gridDim.x / 2 : gridDim.x) * tmul;
    const int64_t t_first = int64_t(KS ? blockIdx.x / 2 : blockIdx.x) * tmul;

    // 64 mask bits of tile t (zero past the end); every role skips a tile whose bits are all clear
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
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], 128);
            mbar_init(&xfull_bar[i], 64);
            mbar_init(&xempty_bar[i], 64);
            if constexpr (KS) { mbar_init(&yfull_bar[i], 1); mbar_init(&yempty_bar[i], 2); }   // [0] used; yempty: one arrival per peer hi warp
        }
        mbar_init(qready_bar, 128);
        fence_barrier_init();
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap)) : "memory");
    }
    if (warp == 1) tmem_alloc(tmem_holder, kMmaTmemCols);
    tc_fence_before();
    __syncthreads();
    if constexpr (KS) ks_cluster_sync();               // the peer's barriers exist before anything remote touches them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    if (tid == 0) stamp(1);

    // epilogue geometry: TMEM lanes [32*quarter, +32); quarters 2,3 hold the hi halves
    const int quarter = warp & 3;
    const int qi = (quarter & 1) * 32 + lane;           // query owned by this thread
    const bool hi_part = quarter >= 2;

    // ---- queries -> tensor memory (epilogue warps; lane of TMEM = thread)
    if (warp >= 2 && warp < 6) {
        const bool live = qi < nq_eff;
        const int qsrc = live ? (a.qlist ? a.qlist[qi] : a.q0 + qi) : a.q0;
        const int ncols = kblocks * (kMmaKBlock / 2);
        if (a.qhl) {
            // packed planes: four 32-column chunks (32 x 128-bit loads) in flight per round, no conversion -- the loads hit
            // L2 (every CTA reads the same planes) and a round costs about one L2 round trip whatever its size
            const uint4* src = reinterpret_cast<const uint4*>(a.qhl + (size_t(qsrc) * 2 + (hi_part ? 0 : 1)) * (a.ld / 2) + size_t(kb0) * (kMmaKBlock / 2));
            for (int c0 = 0; c0 < ncols; c0 += 128) {
                uint32_t r[4][32];
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) {
                    const bool on = live && c0 + 32 * ch < ncols;
#pragma unroll
                    for (int v = 0; v < 8; ++v) {
                        const uint4 u = on ? __ldg(src + c0 / 4 + 8 * ch + v) : make_uint4(0u, 0u, 0u, 0u);
                        r[ch][4 * v] = u.x; r[ch][4 * v + 1] = u.y; r[ch][4 * v + 2] = u.z; r[ch][4 * v + 3] = u.w;
                    }
                }
                const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(c0);
#pragma unroll
                for (int ch = 0; ch < 4; ++ch)
                    if (c0 + 32 * ch < ncols) MRAG_TMEM_ST32(taddr + 32u * ch, r[ch]);
            }
        } else {
        const float* qrow = a.q + size_t(qsrc) * a.ld + size_t(kb0) * kMmaKBlock;
        for (int c0 = 0; c0 < ncols; c0 += 32) {     // 32 columns = 64 elements per store
            uint32_t r[32];
#pragma unroll
            for (int v = 0; v < 16; ++v) {
                float4 f = live ? __ldg(reinterpret_cast<const float4*>(qrow + c0 * 2) + v) : make_float4(0.f, 0.f, 0.f, 0.f);
                float e[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float hi = __bfloat162float(__float2bfloat16_rn(e[i]));
                    e[i] = hi_part ? hi : (e[i] - hi);      // lo = q - hi is exact in fp32
                }
                r[2 * v] = pack_bf16x2(e[0], e[1]);
                r[2 * v + 1] = pack_bf16x2(e[2], e[3]);
            }
            const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(c0);
            MRAG_TMEM_ST32(taddr, r);
        }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
#if MRAG_QREADY_BAR
        tc_fence_before();
        mbar_arrive(qready_bar);
#endif
    }
#if !MRAG_QREADY_BAR
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
#endif
    if (tid == 0) stamp(2);

    if (warp == 0) {
        // ================= TMA producer (whole warp walks the loop, one elected lane issues) =================
        int s = 0;
        uint32_t ph = 0;
        uint2 m = tile_mask(t_first);
        for (int64_t t = t_first; t < num_tiles; t += G) {
            const uint2 mn = tile_mask(t + G);                  // next tile's bits, off the critical path
            if ((m.x | m.y) != 0u) {
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&empty_bar[s], ph ^ 1u, slp);
                    if (elect_one()) {
                        mbar_expect_tx(&full_bar[s], kMmaStageBytes);
                        tma_load_2d(stage_base + size_t(s) * kMmaStageBytes, &tmap, (kb0 + kb) * kMmaKBlock, int(t * kMmaTileRows),
                                    &full_bar[s]);
                    }
                    __syncwarp();
                    if (++s == a.stages) { s = 0; ph ^= 1u; }
                }
            }
            m = mn;
        }
        stamp(3);
    } else if (warp == 1) {
        // ================= MMA issuer: ONE elected thread runs the whole loop =================
        // (the issue loop is the critical path of the kernel -- ncu r1g: ~75 instructions per k-block with a
        //  per-iteration election; a single-thread loop drops the election, the reconvergence and the warp syncs)
        if (elect_one()) {
            int s = 0, as = 0;
            uint32_t ph = 0, aph = 0;
#if MRAG_QREADY_BAR
            mbar_wait(qready_bar, 0u, slp);
            tc_fence_after();
#endif
            const uint64_t bdesc0 = make_sw128_desc(smem_u32(stage_base));
            uint2 m = tile_mask(t_first);
            for (int64_t t = t_first; t < num_tiles; t += G) {
                const uint2 mn = tile_mask(t + G);
                if ((m.x | m.y) != 0u) {
                    mbar_wait(&tempty_bar[as], aph ^ 1u, slp);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + uint32_t(kMmaDCol0 + as * kMmaTileRows);
                    for (int kb = 0; kb < kblocks; ++kb) {
                        mbar_wait(&full_bar[s], ph, slp);
                        tc_fence_after();
                        const uint64_t bdesc = bdesc0 + uint64_t(s) * (kMmaStageBytes >> 4);
                        const uint32_t a_tmem = tmem_base + uint32_t(kb * (kMmaKBlock / 2));
                        // K = 16 per instruction: 8 TMEM columns of A, 32 bytes of the swizzled B rows
                        umma_ts_bf16(d_tmem, a_tmem, bdesc, kMmaIdesc, kb != 0 ? 1u : 0u);
                        umma_ts_bf16(d_tmem, a_tmem + 8, bdesc + 2, kMmaIdesc, 1u);
                        umma_ts_bf16(d_tmem, a_tmem + 16, bdesc + 4, kMmaIdesc, 1u);
                        umma_ts_bf16(d_tmem, a_tmem + 24, bdesc + 6, kMmaIdesc, 1u);
                        umma_commit(&empty_bar[s]);             // frees the smem slot when these MMAs retire
                        if (kb == kblocks - 1) umma_commit(&tfull_bar[as]);   // accumulator ready
                        if (++s == a.stages) { s = 0; ph ^= 1u; }
                    }
                    if (++as == 2) { as = 0; aph ^= 1u; }
                }
                m = mn;
            }
        }
        __syncwarp();
        stamp(4);
    } else if (KS && warp >= 6) {
        // ================= KS send warps (6,7): on the tiles the PEER leads, hi + lo partials of my k-blocks ->
        //                   staging buffer -> one bulk copy into the leader's shared memory =================
        const uint32_t lane_addr = uint32_t(quarter * 32) << 16;
        uint32_t it = 0;
        uint2 m = tile_mask(t_first);
        for (int64_t t = t_first; t < num_tiles; t += G) {
            const uint2 mn = tile_mask(t + G);
            if ((m.x | m.y) != 0u) {
                if ((it & 1u) != ks_rank) {
                    const int as = int(it & 1u), xs = as;            // slot and phase are functions of the tile count
                    const uint32_t aph = (it >> 1) & 1u, yj = it >> 1;
                    mbar_wait(&tfull_bar[as], aph, slp);
                    tc_fence_after();
                    mbar_wait(&xfull_bar[xs], aph, slp);
                    const float* xb = xbuf + size_t(xs) * 64 * 64;
                    float part[64];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        uint32_t d[32];
                        MRAG_TMEM_LD32(d, tmem_base + lane_addr + uint32_t(kMmaDCol0 + as * kMmaTileRows + h * 32));
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int c = 0; c < 32; ++c) part[h * 32 + c] = __uint_as_float(d[c]) + xb[(h * 32 + c) * 64 + qi];
                    }
                    // the accumulator and the exchange slot go back BEFORE anything that depends on the peer
                    tc_fence_before();
                    mbar_arrive(&tempty_bar[as]);
                    mbar_arrive(&xempty_bar[xs]);
                    if (yj > 0) mbar_wait(&yempty_bar[0], (yj - 1u) & 1u, slp);      // the leader has consumed my previous send: sbuf and its ybuf are free
#pragma unroll
                    for (int c = 0; c < 64; ++c) sbuf[c * 64 + qi] = part[c];
                    fence_proxy_async();                                     // my stores -> visible to the bulk copy engine
                    ks_bar_send();
                    if (warp == 6 && lane == 0)
                        ks_bulk_copy_to_peer(ks_map_to_cta(smem_u32(ybuf), ks_rank ^ 1u), sbuf, 64 * 64 * 4,
                                             ks_map_to_cta(smem_u32(&yfull_bar[0]), ks_rank ^ 1u));
                }
                ++it;
            }
            m = mn;
        }
    } else if (!hi_part) {
        // ================= lo epilogue (warps 4,5): TMEM lanes 0..63 -> exchange buffer =================
        const uint32_t lane_addr = uint32_t(quarter * 32) << 16;
        int as = 0, xs = 0;
        uint32_t aph = 0, xph = 0;
        [[maybe_unused]] uint32_t vs = 0;                   // KS: slot of the 1/|x| ring (4 tiles: the leader keeps reading it after it released xbuf)
        uint2 m = tile_mask(t_first);
        for (int64_t t = t_first; t < num_tiles; t += G) {
            const uint2 mn = tile_mask(t + G);
            if ((m.x | m.y) != 0u) {
                // warp 4 also publishes 1/|x| of the tile's rows (NaN for a masked row: the mask is
                // folded into the score); fetched before the accumulator is waited for
                const int64_t r0 = t * kMmaTileRows;
                float in0 = CUDART_NAN_F, in1 = CUDART_NAN_F;
                if (quarter == 0) {
                    if ((m.x >> lane) & 1u) in0 = __ldg(a.inv_norm + r0 + lane);
                    if ((m.y >> lane) & 1u) in1 = __ldg(a.inv_norm + r0 + 32 + lane);
                }
                uint32_t d[64];
                mbar_wait(&tfull_bar[as], aph, slp);
                tc_fence_after();
                MRAG_TMEM_LD64(d, tmem_base + lane_addr + uint32_t(kMmaDCol0 + as * kMmaTileRows));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                tc_fence_before();
                mbar_arrive(&tempty_bar[as]);
                if (++as == 2) { as = 0; aph ^= 1u; }
                mbar_wait(&xempty_bar[xs], xph ^ 1u, slp);
                float* xb = xbuf + size_t(xs) * 64 * 64;
#pragma unroll
                for (int c = 0; c < 64; ++c) xb[c * 64 + qi] = __uint_as_float(d[c]);
                if (quarter == 0) {
                    const int vslot = KS ? int(vs) : xs;
                    xinv[vslot * 64 + lane] = in0;
                    xinv[vslot * 64 + 32 + lane] = in1;
                }
                if constexpr (KS) vs = (vs + 1u) & 3u;
                mbar_arrive(&xfull_bar[xs]);
                if (++xs == 2) { xs = 0; xph ^= 1u; }
            }
            m = mn;
        }
        if (warp == 4) stamp(5);
    } else {
        // ================= hi epilogue (warps 2,3): score + per-thread top-k =================
        const uint32_t lane_addr = uint32_t(quarter * 32) << 16;
        int as = 0, xs = 0;
        uint32_t aph = 0, xph = 0;
        const bool live = qi < nq_eff;
        const bool warp_live = (quarter & 1) * 32 < nq_eff;     // any query in this warp at all?
        const int qsrc = live ? (a.qlist ? a.qlist[qi] : a.q0 + qi) : a.q0;
        const float qinv = live ? a.qinv[qsrc] : 0.0f;
        const uint64_t ubk = (a.ub && live) ? a.ub[qsrc] : ~0ull;
        uint64_t* cand_warp = cand + size_t((quarter & 1) * 32) * a.cap;
        uint64_t* mybuf = cand_warp + size_t(lane) * a.cap;
        const int cap = a.cap;
        SelState st;
        st.cnt = 0;
        // a thread without a query, or with a zero-norm query (every similarity NaN), admits nothing
        st.thr_s = (live && !isinf(qinv)) ? -CUDART_INF_F : CUDART_INF_F;
        // register-resident sorted top-k (KREG mode): the k live slots are the LAST k of top[], the
        // slots before them hold an unbeatable sentinel, so the k-th best is always top[KREG-1]
        // (every index stays a compile-time constant: the array must not fall into local memory)
        uint64_t top[(KREG > 0 && !SO) ? KREG : 1];
        uint32_t tops[(KREG > 0 && SO) ? KREG : 1];
        const int top_off = (KREG > 0 ? KREG : 1) - a.k;
#pragma unroll
        for (int i = 0; i < ((KREG > 0 && !SO) ? KREG : 1); ++i) top[i] = (i < top_off) ? ~0ull : 0ull;
#pragma unroll
        for (int i = 0; i < ((KREG > 0 && SO) ? KREG : 1); ++i) tops[i] = (i < top_off) ? ~0u : 0u;
        const uint32_t ub_ord = uint32_t(ubk >> 32);            // SO: rows scoring >= the bound's score are left out
        // Cross-CTA bound: once ANY CTA holds k keys with score >= g, a row scoring below g cannot be
        // in the global top-k.  Rows scoring exactly g may still win a tie by row index, so the bound
        // admits s >= g, i.e. s > prev(g).  Read relaxed once per tile, raised after each compaction.
        uint32_t* gslot = a.gthr + qsrc;
        GroupBound gb = group_bound_init(a, qsrc, int(blockIdx.x), live && !isinf(qinv));
        const bool use_g = a.gmax != nullptr;
        unsigned long long n_tiles = 0, n_slow = 0, n_keys = 0, n_compact = 0, n_retry = 0;
        uint32_t it = 0;                                         // KS: tiles this pair has processed (leader = it & 1)

        uint2 m = tile_mask(t_first);
        for (int64_t t = t_first; t < num_tiles; t += G) {
            const uint2 mn = tile_mask(t + G);
            if ((m.x | m.y) != 0u) {
                const int64_t r0 = t * kMmaTileRows;
                // KS: this CTA sends (helper) or receives (leader) its (it / 2)-th partial
                const uint32_t yj = it >> 1;
                if constexpr (KS) {
                    if ((it & 1u) != ks_rank) {
                        // the peer leads this tile; my share of it is sent by warps 6, 7 (below)
                        if (++as == 2) { as = 0; aph ^= 1u; }
                        if (++xs == 2) { xs = 0; xph ^= 1u; }
                        ++it;
                        m = mn;
                        continue;
                    }
                    if (warp == 2 && lane == 0) mbar_expect_tx(&yfull_bar[0], 64 * 64 * 4);       // leader: the peer's partials of this tile
                }
                float sc[64];
                float bestg[8];                                      // max of each group of 8 rows
#pragma unroll
                for (int g = 0; g < 8; ++g) bestg[g] = -CUDART_INF_F;
                uint32_t gord;
                asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(gord) : "l"(gslot) : "memory");
                mbar_wait(&tfull_bar[as], aph, slp);
                tc_fence_after();
                mbar_wait(&xfull_bar[xs], xph, slp);
                if constexpr (KS) {
                    // ---- leader of this tile.  Phase 1: my own hi + lo partials into registers, then the accumulator and
                    //      the exchange slot go back at once -- the MMA pipeline must not wait for the peer
                    const float* xb = xbuf + size_t(xs) * 64 * 64;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        uint32_t d[32];
                        MRAG_TMEM_LD32(d, tmem_base + lane_addr + uint32_t(kMmaDCol0 + as * kMmaTileRows + h * 32));
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int c = 0; c < 32; ++c) sc[h * 32 + c] = __uint_as_float(d[c]) + xb[(h * 32 + c) * 64 + qi];
                    }
                    tc_fence_before();
                    mbar_arrive(&tempty_bar[as]);
                    if (++as == 2) { as = 0; aph ^= 1u; }
                    mbar_arrive(&xempty_bar[xs]);
                    // ---- phase 2: the peer's partials of this tile (its k-blocks), 1/|x| from the 4-deep ring
                    mbar_wait(&yfull_bar[0], yj & 1u, slp);
                    if (warp_live) {
                        const float4* inv4 = reinterpret_cast<const float4*>(xinv + (it & 3u) * 64);
#pragma unroll
                        for (int c4 = 0; c4 < 16; ++c4) {
                            const float4 iv = inv4[c4];                          // broadcast
                            const float ivv[4] = {iv.x, iv.y, iv.z, iv.w};
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const int c = c4 * 4 + i;
                                sc[c] = (sc[c] + ybuf[c * 64 + qi]) * ivv[i] * qinv;
                                bestg[c >> 3] = fmaxf(bestg[c >> 3], sc[c]);     // fmaxf drops NaN
                            }
                        }
                    }
                    __syncwarp();                                   // every lane has consumed its column of ybuf
                    if (lane == 0) ks_arrive_remote(ks_map_to_cta(smem_u32(&yempty_bar[0]), ks_rank ^ 1u));
                    ++it;
                } else {
                if (warp_live) {
                    const float* xb = xbuf + size_t(xs) * 64 * 64;
                    const float4* inv4 = reinterpret_cast<const float4*>(xinv + xs * 64);
                    // ---- all 64 scores of this thread's query, branch free (masked / zero rows give
                    //      NaN), in two halves of 32 accumulator columns to bound register pressure
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
                                const float dot = __uint_as_float(d[c4 * 4 + i]) + xb[c * 64 + qi];
                                sc[c] = dot * ivv[i] * qinv;
                                bestg[c >> 3] = fmaxf(bestg[c >> 3], sc[c]);     // fmaxf drops NaN
                            }
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(&tempty_bar[as]);
                if (++as == 2) { as = 0; aph ^= 1u; }
                }
                if (warp_live) {
                    if constexpr (!KS) mbar_arrive(&xempty_bar[xs]);
                    // ---- rows arrive in increasing order, so a later row never beats an equal score:
                    //      only scores strictly above the threshold can enter
                    const float best = fmaxf(fmaxf(fmaxf(bestg[0], bestg[1]), fmaxf(bestg[2], bestg[3])),
                                             fmaxf(fmaxf(bestg[4], bestg[5]), fmaxf(bestg[6], bestg[7])));
                    group_bound_post(gb, best, gslot);
                    // the first tiles of a launch: the other CTAs post their first maxima at about the same moment, so the
                    // word read at the top of the tile is still empty -- read it again, after the posts
                    if (use_g && n_tiles < 3)
                        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(gord) : "l"(gslot) : "memory");
                    float thr = fmaxf(st.thr_s, gord ? ord2f(gord - 1u) : -CUDART_INF_F);
                    ++n_tiles;
                    if (__any_sync(kFull, best > thr)) {
                        ++n_slow;
                        if constexpr (KREG > 0 && SO) {
                            const float thr_in = st.thr_s;
#pragma unroll
                            for (int c = 0; c < 64; ++c) {
                                const bool ins = sc[c] > thr;
                                if (__any_sync(kFull, ins)) {                    // warp-uniform
                                    uint32_t key = ins ? f2ord(sc[c]) : 0u;
                                    if (key >= ub_ord) key = 0u;
#pragma unroll
                                    for (int i = 0; i < KREG; ++i) {             // sorted insertion, a zero key falls through
                                        const uint32_t hi = max(key, tops[i]);
                                        key = min(key, tops[i]);
                                        tops[i] = hi;
                                    }
                                    const uint32_t kth = tops[KREG - 1];
                                    if (kth) { st.thr_s = ord2f(kth); thr = fmaxf(thr, st.thr_s); }
                                }
                            }
                            if (st.thr_s > thr_in) atomicMax(gslot, f2ord(st.thr_s));
                        } else if constexpr (KREG > 0) {
                            const float thr_in = st.thr_s;
#pragma unroll
                            for (int c = 0; c < 64; ++c) {
                                const bool ins = sc[c] > thr;
                                if (__any_sync(kFull, ins)) {                    // warp-uniform
                                    uint64_t key = 0ull;
                                    if (ins) {
                                        key = make_key(sc[c], uint32_t(r0 + c));
                                        if (key >= ubk) key = 0ull;
                                    }
                                    n_keys += key != 0ull;
                                    // sorted insertion into top[] (descending); a zero key falls through
#pragma unroll
                                    for (int i = 0; i < KREG; ++i) {
                                        const bool gt = key > top[i];
                                        const uint64_t lower = gt ? top[i] : key;
                                        top[i] = gt ? key : top[i];
                                        key = lower;
                                    }
                                    const uint64_t kth = top[KREG - 1];
                                    if (kth) {                                   // k keys held: exact k-th best so far
                                        st.thr_s = key_score(kth);
                                        thr = fmaxf(thr, st.thr_s);
                                    }
                                }
                            }
                            if (st.thr_s > thr_in) atomicMax(gslot, f2ord(st.thr_s));
                        } else {
                        // the warp walks only the groups of 8 rows in which some query admits a row
#define MRAG_WALK_GROUP(G)                                                                                          \
                        if (__any_sync(kFull, bestg[G] > thr))                                                      \
                            walk_group64<G>(sc, thr, st, mybuf, cand_warp, cap, a.k, lane, r0, gslot, ubk, n_keys, n_compact, n_retry);
                        MRAG_WALK_GROUP(0) MRAG_WALK_GROUP(1) MRAG_WALK_GROUP(2) MRAG_WALK_GROUP(3)
                        MRAG_WALK_GROUP(4) MRAG_WALK_GROUP(5) MRAG_WALK_GROUP(6) MRAG_WALK_GROUP(7)
#undef MRAG_WALK_GROUP
                        }
                    }
                } else {
                    if constexpr (!KS) mbar_arrive(&xempty_bar[xs]);
                }
                if (++xs == 2) { xs = 0; xph ^= 1u; }
            }
            m = mn;
        }

        if (warp == 2) stamp(6);
        if (a.stats) {
            if (lane == 0) { atomicAdd(a.stats + 0, n_tiles); atomicAdd(a.stats + 1, n_slow); atomicAdd(a.stats + 4, n_retry); }
            if (n_keys) atomicAdd(a.stats + 2, n_keys);
            if (n_compact) atomicAdd(a.stats + 3, n_compact);
        }
        // ---- this CTA's sorted list per query
        __syncwarp();
        if constexpr (KREG > 0) {
            if (live) {
                uint64_t* out = a.part + (size_t(qsrc) * a.P + blockIdx.x) * a.kp;
                if constexpr (SO) {
                    // keys stay unique across CTAs and slots (the merge's radix select relies on it); 0 = empty
#pragma unroll
                    for (int i = 0; i < KREG; ++i)
                        if (i >= top_off) out[i - top_off] = tops[i] ? (uint64_t(tops[i]) << 32) | uint64_t((blockIdx.x << 8) | unsigned(i)) : 0ull;
                } else {
#pragma unroll
                    for (int i = 0; i < KREG; ++i)
                        if (i >= top_off) out[i - top_off] = top[i];
                }
                for (int i = a.k; i < a.kp; ++i) out[i] = 0ull;
            }
        } else if (a.k <= 16) {
            // small k: every thread extracts the k best of ITS buffer by repeated maximum (k * n compares, all 32
            // queries of the warp at once) instead of 32 warp-wide rank sorts one after the other
            if (live) {
                uint64_t* out = a.part + (size_t(qsrc) * a.P + blockIdx.x) * a.kp;
                const int n = st.cnt;
                const int have = n < a.k ? n : a.k;
                for (int i = 0; i < a.kp; ++i) {
                    uint64_t best = 0ull;
                    if (i < have) {
                        int bi = 0;
                        for (int j = 0; j < n; ++j) {
                            const uint64_t v = mybuf[j];
                            if (v > best) { best = v; bi = j; }
                        }
                        mybuf[bi] = 0ull;
                    }
                    out[i] = best;
                }
            }
        } else {
            for (int L = 0; L < 32; ++L) {
                const int qL = (quarter & 1) * 32 + L;
                if (qL >= nq_eff) break;
                const int n = __shfl_sync(kFull, st.cnt, L);
                uint64_t* b = cand_warp + size_t(L) * a.cap;
                warp_rank_select(b, n, a.kp, lane);
                uint64_t* out = a.part + (size_t(__shfl_sync(kFull, qsrc, L)) * a.P + blockIdx.x) * a.kp;
                const int have = n < a.k ? n : a.k;
                for (int i = lane; i < a.kp; i += 32) out[i] = (i < have) ? b[i] : 0ull;
            }
        }
    }

    if (warp == 2) stamp(7);
    tc_fence_before();
    __syncthreads();
    if constexpr (KS) ks_cluster_sync();               // nobody leaves while the peer may still store into this CTA
    if (tid == 0) stamp(8);
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kMmaTmemCols);
    }
}

}  // namespace mrag

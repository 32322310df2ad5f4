"""compute-sanitizer driver (GPU box): one small search through every kernel family.

    compute-sanitizer --tool memcheck python tools/sanitize_small.py      (compute-sanitizer is closed on the
    round-1 GPU pool; the script doubles as a plain smoke run of every kernel family)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import mrag_b200  # noqa: F401
from mrag_b200 import _native as N
from mrag_b200 import synth
from mrag_b200.index import FEAT_DTYPE, Filter, Index

n, dim = 3000, 128
X, valid = synth.make_corpus(n, dim, seed=1)
meta, doc_tags, info = synth.make_metadata(n, seed=2, rows_per_doc=16, valid=valid)
for dtype in ("f32", "bf16"):
    idx = Index(dim, dtype, 0, n + 7)
    idx.append(X, meta)
    idx.set_doc_tags(0, doc_tags)
    for nq, k, opt, flt in ((1, 10, N.OPT_FORCE_GEMV, None), (3, 100, N.OPT_FORCE_GEMV, Filter().state_eq(0)),
                            (5, 10, 0, None), (70, 10, N.OPT_FORCE_MMA128, None), (70, 10, N.OPT_FORCE_MMA128, Filter().tag_relaxed([0, 3])),
                            (2, 200, 0, None)):
        Q = synth.make_queries(X, nq, seed=nq)
        s, r, c = idx.search(Q, k, flt, options=opt)
        print(dtype, nq, k, idx.last_scan_kind(), int(c[0]))
    if dtype == "bf16":
        s, r, c = idx.search(synth.make_queries(X, 9, seed=3), 10, options=N.OPT_FORCE_MMA)
        print("bf16 mma", idx.last_scan_kind(), int(c[0]))
    feat = np.zeros(n, dtype=FEAT_DTYPE)
    feat["phrase_bits"][::7, 0] = 3
    feat["dtags"][::11, 0] = 2
    idx.set_chunk_features(0, feat)
    hq = (N.HybridQuery * 2)()
    for h in hq:
        h.n_phrases = 1; h.phrase_weight[0] = 1.0; h.phrase_bit[0] = 1; h.phrase_jbit[0] = -1
        h.w_sim, h.w_auth, h.w_len, h.w_cov, h.boost, h.floor = 0.25, 0.1, 0.05, 0.55, 1.5, 1.0
    out = idx.search_hybrid(synth.make_queries(X, 2, seed=5), 20, hq)
    print("hybrid", int(out[3][0]))
    idx.tombstone_doc(3)
    idx.close()
# round 2: the k-split pair kernel (rows of 1536 elements), the pool cascade, the candidate rerank, explicit row ids
Xw, vw = synth.make_corpus(700, 1536, seed=9)
idx = Index(1536, "bf16", 0, 800)
ids = np.arange(1000, 1700, dtype=np.int64)
N.check(idx._lib.mrag_set_row_ids(idx._h, 0, ids.ctypes.data, 700))
idx.append(Xw, None)
for nq, k in ((1, 10), (40, 10), (70, 100)):
    s, r, c = idx.search(synth.make_queries(Xw, nq, seed=nq), k)
    print("ks", nq, k, idx.last_scan_kind(), int(c[0]), int(r[0, 0]))
idx.set_doc_tags(0, np.ones((700, N.MRAG_TAG_WORDS), dtype=np.uint64))
idx.set_doc_jtags(0, np.ones((700, N.MRAG_JTAG_WORDS), dtype=np.uint64))
import ctypes as C
q = N.PoolQuery()
q.d_all[0] = 1; q.j_all[0] = 1; q.ahca[0] = 1; q.has_d = q.has_j = q.has_ahca = 1
h = C.c_void_p(); counts = (C.c_int64 * 5)()
N.check(idx._lib.mrag_pool_build(idx._h, C.byref(q), C.byref(h), counts))
kept = C.c_int64()
N.check(idx._lib.mrag_pool_select(h, 1, 100, C.byref(kept)))
print("pool", list(counts), kept.value)
s, r, c = idx.search(synth.make_queries(Xw, 3, seed=1), 5, Filter().pool_handle(h))
print("pool search", int(c[0]))
N.check(idx._lib.mrag_pool_destroy(h))
cands = (N.Candidate * 50)()
hq1 = N.HybridQuery()
hq1.w_sim, hq1.w_auth, hq1.w_len = 0.25, 0.1, 0.05
for i in range(50):
    cands[i].sim = i / 50.0
print("rerank", float(idx.rerank_candidates(cands, 50, hq1)[0][49]))
idx.close()
print("sanitize driver done")

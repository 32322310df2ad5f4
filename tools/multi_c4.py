"""Config 4 through the PLUGIN boundary: ONE process owns all GPUs of the box (MultiIndex behind B200VectorStore /
vector_arm), 6.25M x 1536 bf16 rows per GPU, per-payor bitset filter, top-10.

    python tools/multi_c4.py [rows_per_gpu] [reps]

Prints one JSON line: single-query latency of store.search() / vector_arm() (host lists in, dicts out) and of the bare
MultiIndex.search (host arrays), against the HBM floor of one pass over a shard (all GPUs scan concurrently).
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import mrag_b200
from mrag_b200 import index as mi
from mrag_b200 import synth
from mrag_b200.corpus_search import CorpusFilters

rows_per_gpu = int(sys.argv[1]) if len(sys.argv) > 1 else 6_250_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
devs = [int(x) for x in os.environ.get("MRAG_DEVICES", ",".join(str(g) for g in range(torch.cuda.device_count()))).split(",")]
dim, G = 1536, len(devs)                              # MRAG_DEVICES=0,0 : two shards on one GPU (a 1-GPU box can validate the path)
n = rows_per_gpu * G
t0 = time.perf_counter()
pt = mrag_b200.PublishedTable(dim, "bf16", 0, n, devices=devs)
m_idx = pt.index
plant = None
first_doc = 0
for g in range(G):
    dev = torch.device(f"cuda:{devs[g]}")
    with torch.cuda.device(dev):
        for first, X in synth.cuda_corpus_chunks(rows_per_gpu, dim, dev, seed=1234 + g, chunk=1 << 18):
            if plant is None:
                plant = X[:64].cpu().numpy()
            m = X.shape[0]
            docs = (first_doc + (first + np.arange(m)) // 64).astype(np.uint32)
            meta = mi.make_meta(m, doc_idx=docs, payer=(docs % 13).astype(np.uint16))
            m_idx.append_device_shard(g, X, meta)
        torch.cuda.synchronize(dev)
    first_doc += (rows_per_gpu + 63) // 64
# host half of the table, columnar: ids / document ids only (what the store returns)
pt.vocab.payer.values.extend(f"payer-{i}" for i in range(13))
pt.vocab.payer._code.update({f"payer-{i}": i for i in range(13)})
pt.id.extend_raw(np.full(n, 9, dtype=np.int64), np.frombuffer(b"".join(b"r%08d" % i for i in range(n)), dtype=np.uint8)) if n <= 4_000_000 else \
    pt.id.extend_raw(np.full(n, 9, dtype=np.int64), np.tile(np.frombuffer(b"r00000000", dtype=np.uint8), n))
pt.source_id.extend_raw(np.zeros(n, dtype=np.int64), np.zeros(0, dtype=np.uint8), np.ones(n, dtype=bool))
docs_per_gpu = (rows_per_gpu + 63) // 64
ar = np.arange(n, dtype=np.int64)
pt.row_doc = ((ar // rows_per_gpu) * docs_per_gpu + (ar % rows_per_gpu) // 64).astype(np.uint32)     # as the loader above numbered them
n_docs = int(pt.row_doc.max()) + 1
pt.doc_ids = [f"doc-{d}" for d in range(n_docs)]
pt.doc_idx = {v: i for i, v in enumerate(pt.doc_ids)}
pt.source_type.extend(np.full(n, 0xFF, dtype=np.uint8))
pt.document_payer.extend((pt.row_doc % 13).astype(np.uint16))
for col in (pt.document_state, pt.document_program, pt.document_authority_level):
    col.extend(np.full(n, 0xFF, dtype=np.uint8))
for name, col in pt.extra.items():
    if hasattr(col, "s"):
        col.s.extend_raw(np.zeros(n, dtype=np.int64), np.zeros(0, dtype=np.uint8), np.ones(n, dtype=bool))
    elif hasattr(col, "val"):
        col.extend(np.zeros(n, dtype=np.int64))
    else:
        col.extend_raw(np.zeros(n, dtype=np.int64), np.zeros(0, dtype=np.uint8), np.ones(n, dtype=bool))
pt._n = pt._host_n = n
build_s = time.perf_counter() - t0

store = mrag_b200.B200VectorStore(table=pt)
q = (plant[7] + 0.05 * np.random.default_rng(1).standard_normal(dim)).astype(np.float32)
ql = q.tolist()
flt_store = {"payer": "payer-3"}


def timed(fn, reps):
    for _ in range(3):
        out = fn()
    t = time.perf_counter()
    for _ in range(reps):
        out = fn()
    return (time.perf_counter() - t) / reps * 1e3, out


ms_store, hits = timed(lambda: store.search(ql, 10, filters=flt_store), reps)
ms_store_nf, hits_nf = timed(lambda: store.search(ql, 10), reps)
ms_arm, arm = timed(lambda: mrag_b200.vector_arm(pt, ql, 10, CorpusFilters(payer="payer-3"), None), reps)
ms_raw, raw = timed(lambda: m_idx.search(q[None, :], 10, pt.filter_pg_store(None, flt_store)), reps)
ms_raw64, raw64 = timed(lambda: m_idx.search(np.repeat(q[None, :], 64, 0) + 0.01 * np.random.default_rng(2).standard_normal((64, dim)).astype(np.float32), 10), max(3, reps // 4))
assert hits and all(int(h["document_id"].split("-")[1]) % 13 == 3 for h in hits)
assert hits_nf[0]["distance"] > 0.9
floor_ms = rows_per_gpu * dim * 2 / 6543.7e9 * 1e3
print(json.dumps({
    "workload": f"{n}x{dim} bf16 over {G} GPUs in ONE process (MultiIndex), top-10, single query", "gpus": G, "rows_per_gpu": rows_per_gpu,
    "hbm_floor_ms_per_pass": floor_ms, "hbm_floor_ms_payer_filter_1_of_13": floor_ms / 13,
    "store_search_ms_payer_filter": ms_store, "store_search_ms_unfiltered": ms_store_nf, "vector_arm_ms_payer_filter": ms_arm,
    "multiindex_search_ms_payer_filter": ms_raw, "multiindex_search_ms_b64_unfiltered": ms_raw64,
    "unfiltered_frac_of_hbm_floor": floor_ms / ms_store_nf, "kernel": m_idx.last_scan_kind(), "build_s": build_s,
    "shard_sizes": m_idx.shard_sizes(), "top_hit": hits_nf[0]}))

"""Profiling driver (GPU box): single-query searches with a document-pool filter on a production-shaped corpus."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import mrag_b200  # noqa: F401
from mrag_b200 import index as mi
from mrag_b200 import synth

n, dim = 1_900_000, 1536
dev = torch.device("cuda:0")
idx = mi.Index(dim, "f32", 0, n)
plant = None
for first, X in synth.cuda_corpus_chunks(n, dim, dev, chunk=1 << 16):
    if first == 0:
        plant = X[:4096].clone()
    idx.append_device(X, mi.make_meta(X.shape[0], doc_idx=(np.arange(first, first + X.shape[0]) // 64).astype(np.uint32)))
pool = np.random.default_rng(5).choice(n // 64, size=int(sys.argv[1]) if len(sys.argv) > 1 else 50, replace=False).astype(np.uint32)
flt = mi.Filter().doc_pool(pool)
Q = synth.cuda_queries(plant, 1, dim, dev).cpu().numpy()
for _ in range(4):
    idx.search(Q, 10, flt)
print("kind", idx.last_scan_kind(), [round(idx.last_kernel_ms(i), 4) for i in range(4)])

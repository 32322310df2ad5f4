"""Profiling driver (GPU box): builds the 10M x 768 bf16 bench corpus and runs a few searches.

    ncu --set full --clock-control none --import-source on -k regex:scan_mma128 -s 5 -c 1 -o out \\
        python tools/one_search.py <batch> <k> <reps> [rows] [dtype] [dim]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import mrag_b200  # noqa: F401
from mrag_b200 import index as mi
from mrag_b200 import synth

B, k, reps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
n = int(sys.argv[4]) if len(sys.argv) > 4 else 10_000_000
dtype = sys.argv[5] if len(sys.argv) > 5 else "bf16"
dim = int(sys.argv[6]) if len(sys.argv) > 6 else 768
dev = torch.device("cuda:0")
idx = mi.Index(dim, dtype, 0, n)
plant = None
for first, X in synth.cuda_corpus_chunks(n, dim, dev):
    if first == 0:
        plant = X[:4096].clone()
    idx.append_device(X, mi.make_meta(X.shape[0], doc_idx=(np.arange(first, first + X.shape[0]) // 64).astype(np.uint32)))
Q = synth.cuda_queries(plant, B, dim, dev)
for _ in range(reps):
    idx.search_device(Q, k)
torch.cuda.synchronize()
print("kind", idx.last_scan_kind(), "scan_ms", idx.last_kernel_ms(1), "total_ms", idx.last_kernel_ms(3))
from mrag_b200 import _native as N
print("fallbacks", N.load().mrag_debug_fallback_count())

"""Debugging aid (GPU box): select counters of the 128-query candidate scan on a 10M x 768 bf16 corpus."""
import ctypes as C
import os
import sys

os.environ["MRAG_SCAN_STATS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import mrag_b200  # noqa: F401
from mrag_b200 import _native as N
from mrag_b200 import index as mi
from mrag_b200 import synth

dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
dim = 768
idx = mi.Index(dim, "bf16", 0, n)
plant = None
for first, X in synth.cuda_corpus_chunks(n, dim, dev):
    if first == 0:
        plant = X[:4096].clone()
    idx.append_device(X, mi.make_meta(X.shape[0], doc_idx=(np.arange(first, first + X.shape[0]) // 64).astype(np.uint32)))
lib = N.load()
for B in (128, 256):
    Q = synth.cuda_queries(plant, B, dim, dev)
    for it in range(3):
        idx.search_device(Q, 10)
    torch.cuda.synchronize()
    st = (C.c_ulonglong * 40)()
    rc = lib.mrag_debug_scan_stats(st)
    print(f"B={B} kind={idx.last_scan_kind()} rc={rc} tiles(warp)={st[0]} groups={st[1]} keys={st[2]} compactions={st[3]} "
          f"scan_ms={idx.last_kernel_ms(1):.3f} total_ms={idx.last_kernel_ms(3):.3f} fallbacks={lib.mrag_debug_fallback_count()}")

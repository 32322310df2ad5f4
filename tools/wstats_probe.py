"""Debugging aid (GPU box): cycle accounting of the wide CTA-pair scan (library built with -DMRAG_WSTATS=1).

    MRAG_LIB=mobius-rag_b200/libmrag_wstats.so python tools/wstats_probe.py [rows] [batch]
"""
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
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dim = 768
idx = mi.Index(dim, "bf16", 0, n)
plant = None
for first, X in synth.cuda_corpus_chunks(n, dim, dev):
    if first == 0:
        plant = X[:4096].clone()
    idx.append_device(X, mi.make_meta(X.shape[0], doc_idx=(np.arange(first, first + X.shape[0]) // 64).astype(np.uint32)))
lib = N.load()
Q = synth.cuda_queries(plant, B, dim, dev)
for it in range(3):
    idx.search_device(Q, 10)
torch.cuda.synchronize()
st = (C.c_ulonglong * 40)()
lib.mrag_debug_scan_stats(st)
tiles = (n + 127) // 128 / 74
print(f"scan_ms={idx.last_kernel_ms(1):.3f} tiles/pair={tiles:.0f}")
print(f"MMA thread : total={st[8]} cyc ({st[8] / tiles:.0f}/tile)  wait tempty={st[9]} ({st[9] / tiles:.0f}/tile)  wait full={st[10]} ({st[10] / tiles:.0f}/tile)")
print(f"select w2  : total={st[12]} cyc ({st[12] / tiles:.0f}/tile)  wait tfull={st[13]} ({st[13] / tiles:.0f}/tile)  wait ifull={st[14]} ({st[14] / tiles:.0f}/tile)  "
      f"ld+arrive={st[15]} ({st[15] / tiles:.0f}/tile)")

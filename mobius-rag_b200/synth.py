"""Deterministic synthetic corpora, metadata and queries.

numpy versions for the parity tests (the oracle and the GPU see the very same arrays) and a
chunked torch-CUDA generator for the full-size benchmark corpus (10M x 768 would take minutes
and 30 GB of host memory to build and ship from the CPU).
"""
from __future__ import annotations

import numpy as np

from . import _native as N
from .index import make_meta

# vocabulary shapes taken from the reference's canonical tables
PAYERS = ["Sunshine Health", "Humana", "Humana Healthy Horizons", "United Healthcare", "Molina Healthcare",
          "Molina Healthcare of Florida", "Aetna", "Centene", "WellCare", "Simply Healthcare", "AHCA",
          "Ahca.myflorida", "Florida Medicaid"]                       # metadata_canonical.py:46-71
STATES = ["FL", "AL", "AK", "AZ", "AR", "CA", "CO", "CT", "DE", "GA", "HI", "ID", "IL", "IN", "IA", "KS", "KY",
          "LA", "ME", "MD", "MA", "MI", "MN", "MS", "MO", "MT", "NE", "NV", "NH", "NJ", "NM", "NY", "NC", "ND",
          "OH", "OK", "OR", "PA", "RI", "SC", "SD", "TN", "TX", "UT", "VT", "VA", "WA", "WV", "WI", "WY", "DC", "PR"]
PROGRAMS = ["Medicaid", "Medicare", "Marketplace", "CHIP", "MMA", "LTC", "Dental", "Dual Eligible", "Commercial",
            "Medicare Advantage", "Behavioral Health", "Pharmacy", "Vision", "Child Welfare", "HIV/AIDS"]
AUTHORITIES = ["contract_source_of_truth", "payer_website", "operational_suggested", "payer_policy",
               "fyi_not_citable"]                                     # corpus_search.py:215-221
SOURCE_TYPES = ["hierarchical", "fact", "policy_paragraph"]


def make_corpus(n: int, dim: int, seed: int = 1234, dup_frac: float = 0.005, null_frac: float = 1e-4,
                zero_norm_rows: int = 1) -> tuple[np.ndarray, np.ndarray]:
    """X float32 [n, dim] (un-normalised, like the reference's writes) and valid uint8 [n].

    0.5 % of the rows are exact copies of other rows in clusters of 2..200 (the boilerplate
    clusters of corpus_search.py:3546-3553), 0.01 % have no vector (embedding_vec IS NULL) and
    one row is all zeros (cosine distance NaN)."""
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, dim), dtype=np.float32)
    X *= np.exp(0.25 * rng.standard_normal(n, dtype=np.float32))[:, None]
    n_dup = int(n * dup_frac)
    placed = 0
    while placed < n_dup and n > 4:
        size = int(min(rng.integers(2, 201), n_dup - placed + 1))
        src = int(rng.integers(0, n))
        dst = rng.integers(0, n, size=size - 1)
        X[dst] = X[src]
        placed += size - 1
    valid = np.ones(n, dtype=np.uint8)
    n_null = int(round(n * null_frac))
    if n_null:
        valid[rng.choice(n, size=n_null, replace=False)] = 0
    for _ in range(min(zero_norm_rows, n)):
        X[int(rng.integers(0, n))] = 0.0
    return X, valid


def make_queries(X: np.ndarray, nq: int, seed: int = 4321) -> np.ndarray:
    """Half random directions, half planted neighbours X[j] + 0.1 * noise."""
    rng = np.random.default_rng(seed)
    n, dim = X.shape
    Q = rng.standard_normal((nq, dim), dtype=np.float32)
    for i in range(nq // 2, nq):
        j = int(rng.integers(0, n))
        Q[i] = X[j] + np.float32(0.1) * rng.standard_normal(dim, dtype=np.float32) * max(
            float(np.linalg.norm(X[j])) / np.sqrt(dim), 1e-3)
    return Q


def make_metadata(n: int, seed: int = 99, rows_per_doc: int = 64, valid: np.ndarray | None = None):
    """Doc-contiguous metadata as codes.  Returns (meta META_DTYPE[n], doc_tags uint64[n_docs, TAG_WORDS],
    info dict with the string tables so a test can build the oracle's string columns)."""
    rng = np.random.default_rng(seed)
    n_docs = max(1, (n + rows_per_doc - 1) // rows_per_doc)
    bounds = np.sort(rng.choice(np.arange(1, n), size=min(n_docs - 1, max(n - 1, 0)), replace=False)) if n > 1 else np.array([], dtype=np.int64)
    doc_of_row = np.searchsorted(bounds, np.arange(n), side="right").astype(np.uint32)
    n_docs = int(doc_of_row.max()) + 1 if n else 1
    zipf = 1.0 / np.arange(1, len(PAYERS) + 1) ** 1.2
    d_payer = rng.choice(len(PAYERS), size=n_docs, p=zipf / zipf.sum()).astype(np.uint16)
    p_state = np.full(len(STATES), 0.4 / (len(STATES) - 1)); p_state[0] = 0.6
    d_state = rng.choice(len(STATES), size=n_docs, p=p_state).astype(np.uint8)
    d_program = rng.integers(0, len(PROGRAMS), size=n_docs).astype(np.uint8)
    d_auth = rng.choice(len(AUTHORITIES), size=n_docs, p=[0.3, 0.2, 0.15, 0.25, 0.1]).astype(np.uint8)
    # 5 % of the documents have no payer / authority (NULL columns)
    none_p = rng.random(n_docs) < 0.05
    none_a = rng.random(n_docs) < 0.05
    payer = d_payer.copy(); payer[none_p] = N.MRAG_CODE_NONE
    auth = d_auth.copy(); auth[none_a] = 0xFF
    # tag bits: bit b set with probability chosen so single-tag filters pass ~10 %, 1 %, 0.1 % of docs
    nbits = N.MRAG_TAG_WORDS * 64
    dens = np.where(np.arange(nbits) % 3 == 0, 0.10, np.where(np.arange(nbits) % 3 == 1, 0.01, 0.001))
    tagmat = rng.random((n_docs, nbits)) < dens[None, :]
    no_tag_row = rng.random(n_docs) < 0.1                 # documents without a document_tags row
    tagmat[no_tag_row] = False
    doc_tags = np.zeros((n_docs, N.MRAG_TAG_WORDS), dtype=np.uint64)
    for w in range(N.MRAG_TAG_WORDS):
        chunk = tagmat[:, w * 64:(w + 1) * 64].astype(np.uint64)
        doc_tags[:, w] = (chunk << np.arange(64, dtype=np.uint64)[None, :]).sum(axis=1, dtype=np.uint64)
    src = rng.integers(0, len(SOURCE_TYPES), size=n).astype(np.uint8)
    meta = make_meta(n, doc_idx=doc_of_row, payer=payer[doc_of_row], state=d_state[doc_of_row],
                     program=d_program[doc_of_row], authority=auth[doc_of_row], source_type=src,
                     valid=valid if valid is not None else 1)
    info = {"n_docs": n_docs, "doc_of_row": doc_of_row, "tagmat": tagmat}
    return meta, doc_tags, info


def cuda_corpus_chunks(n: int, dim: int, device, seed: int = 1234, chunk: int = 1 << 18, dup_frac: float = 0.005):
    """Yields (first_row, X_chunk float32 CUDA [m, dim]) for the benchmark corpus: same
    distribution as make_corpus, duplicates drawn inside each chunk."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    for first in range(0, n, chunk):
        m = min(chunk, n - first)
        X = torch.randn((m, dim), generator=g, device=device, dtype=torch.float32)
        X.mul_(torch.exp(0.25 * torch.randn((m, 1), generator=g, device=device, dtype=torch.float32)))
        n_dup = int(m * dup_frac)
        if n_dup > 1 and m > 256:
            n_src = max(1, n_dup // 20)
            src = torch.randint(0, m, (n_src,), generator=g, device=device)
            dst = torch.randint(0, m, (n_dup,), generator=g, device=device)
            X[dst] = X[src[torch.randint(0, n_src, (n_dup,), generator=g, device=device)]]
        yield first, X


def cuda_queries(index_rows_sample, nq: int, dim: int, device, seed: int = 4321):
    """Half random, half planted near rows of `index_rows_sample` (float32 CUDA [m, dim])."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    Q = torch.randn((nq, dim), generator=g, device=device, dtype=torch.float32)
    m = index_rows_sample.shape[0]
    if m > 0 and nq > 1:
        j = torch.randint(0, m, (nq - nq // 2,), generator=g, device=device)
        base = index_rows_sample[j]
        noise = torch.randn(base.shape, generator=g, device=device, dtype=torch.float32)
        Q[nq // 2:] = base + 0.1 * noise * (base.norm(dim=1, keepdim=True) / dim ** 0.5)
    return Q.contiguous()

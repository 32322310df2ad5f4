"""Candidate-pool builder on the GPU, mirroring ``build_candidate_pool`` (app/services/corpus_search_agent.py:1762-1888).

The reference fetches one document set per lexicon tag with a SQL statement each (`_doc_ids_with_tag`, :1461-1482),
intersects them in Python, walks the cascade  L1 J&D&P -> L2 J&D -> L3 AHCA&D -> L4 AHCA -> L5 empty  and hands
up to 5000 document UUIDs to every later search as ``include_document_ids``.  Here the per-document tag sets already sit
in HBM as bitsets, so all four levels are ONE kernel over the documents (``mrag_pool_build``); the level that wins stays
on the device as a document bitmap and the returned ``CandidatePool`` can be passed as ``include_document_ids`` to
``vector_arm`` / ``dtag_arm`` / ``hybrid_rerank`` as it is -- no UUID list is marshalled per search.  ``document_ids``
(the reference's field) is still there, materialised on first use.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Any, Sequence

import numpy as np

from . import _native as N
from .table import PublishedTable

AHCA_TAG = "j:regulatory_authority.ahca"                  # corpus_search_agent.py:1458
POOL_CAP = 5000                                           # list(L)[:5000], :1818
LEVELS = ("L1_JDP", "L2_JD", "L3_AHCA_D", "L4_AHCA")


@dataclass
class TermAssignment:
    """corpus_search_agent.py:1317-1324 (the fields the pool builder reads)."""
    term: str = ""
    kind: str = "tag"                 # "tag" | "literal"
    full_code: str | None = None      # "j:payor.sunshine_health", set when kind == "tag"
    selectivity: float = 0.0


@dataclass
class TermPartition:
    """corpus_search_agent.py:1327-1331."""
    required: list = field(default_factory=list)
    boosted: list = field(default_factory=list)
    dropped: list = field(default_factory=list)


class CandidatePool:
    """corpus_search_agent.py:1412-1454: same attributes (document_ids, cascade_level, cascade_steps, intersect_codes,
    inherited_document_ids, required_codes_used, relaxed, relaxed_dropped_codes) + the device handle behind them."""

    def __init__(self, table: PublishedTable | None, handle, cascade_level: str, cascade_steps: list, intersect_codes: list,
                 size: int, inherited_document_ids: list | None = None, document_ids: list | None = None):
        self._table, self._handle = table, handle
        self.cascade_level, self.cascade_steps, self.intersect_codes = cascade_level, cascade_steps, intersect_codes
        self.size = int(size)
        self.inherited_document_ids = list(inherited_document_ids or [])
        self._document_ids = document_ids
        self._doc_idx: np.ndarray | None = None
        self._unknown_ids: list[str] = []        # inherited ids that name no document of this table (kept, like the reference)

    # -- the reference's fields -----------------------------------------------------------------
    @property
    def document_ids(self) -> list[str]:
        if self._document_ids is None:
            t = self._table
            self._document_ids = [t.doc_ids[int(d)] for d in self.doc_indices()] + list(self._unknown_ids)
        return self._document_ids

    @property
    def required_codes_used(self) -> list[str]:
        return self.intersect_codes

    @property
    def relaxed(self) -> bool:
        return self.cascade_level not in ("L1_JDP", "L5_empty")

    @property
    def relaxed_dropped_codes(self) -> list[str]:
        return []

    # -- device side ------------------------------------------------------------------------------
    def doc_indices(self) -> np.ndarray:
        """Dense document indices of the pool (ascending), downloaded once."""
        if self._doc_idx is None:
            if self._handle is None:
                self._doc_idx = np.zeros(0, dtype=np.uint32)
            else:
                out = np.zeros(max(self.size, 1) + 8, dtype=np.uint32)
                n = C.c_int64(0)
                N.check(N.load().mrag_pool_docs(self._handle, out.ctypes.data, out.shape[0], C.byref(n)))
                self._doc_idx = out[:int(n.value)]
        return self._doc_idx

    def close(self) -> None:
        if self._handle is not None:
            N.load().mrag_pool_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self) -> int:
        return self.size

    def __bool__(self) -> bool:          # `if include_document_ids:` keeps meaning "a pool restriction is present"
        return True


def _kind_bits(table: PublishedTable, codes: list[str], kind: str, words: int):
    """(bitset words, known): the bits of `codes` (all of one kind) in the document tag sets; known = every code has
    a bit, i.e. some document carries it (a code nobody carries makes the intersection empty, as in the reference)."""
    bits = [0] * words
    known = True
    for c in codes:
        key = c.split(":", 1)[1]
        b = table.vocab.jtag_bit(key, False) if kind == "j" else table.vocab.tag_bit(kind, key, False)
        if b is None:
            known = False
            continue
        bits[b >> 6] |= 1 << (b & 63)
    return bits, known


def build_candidate_pool(table: PublishedTable, partition: TermPartition, *, min_pool_size: int = 5) -> CandidatePool:
    """Cascading pool builder (same name, argument meaning, levels and trace strings as the reference's)."""
    all_tag_codes = [t.full_code for t in (list(partition.required) + list(partition.boosted))
                     if getattr(t, "kind", None) == "tag" and getattr(t, "full_code", None)]
    j_codes = [c for c in all_tag_codes if c.startswith("j:")]
    d_codes = [c for c in all_tag_codes if c.startswith("d:")]
    p_codes = [c for c in all_tag_codes if c.startswith("p:")]

    q = N.PoolQuery()
    for name, codes, kind, words in (("j", j_codes, "j", N.MRAG_JTAG_WORDS), ("d", d_codes, "d", N.MRAG_TAG_WORDS),
                                     ("p", p_codes, "p", N.MRAG_TAG_WORDS)):
        bits, known = _kind_bits(table, codes, kind, words)
        arr = getattr(q, name + "_all")
        for i, w in enumerate(bits):
            arr[i] = w
        setattr(q, "has_" + name, 1 if (codes and known) else 0)
    ab, aknown = _kind_bits(table, [AHCA_TAG], "j", N.MRAG_JTAG_WORDS)
    for i, w in enumerate(ab):
        q.ahca[i] = w
    q.has_ahca = 1 if aknown else 0

    index = getattr(table.index, "shards", [table.index])[0]       # every shard holds all per-document tag sets
    handle = C.c_void_p()
    counts = (C.c_int64 * (N.MRAG_POOL_LEVELS + 1))()
    lib = N.load()
    N.check(lib.mrag_pool_build(index._h, C.byref(q), C.byref(handle), counts))
    n1, n2, n3, n4, n_d = (int(counts[i]) for i in range(5))
    has_j, has_d, has_p = bool(j_codes), bool(d_codes), bool(p_codes)

    def done(level: int, codes: list[str], steps: list) -> CandidatePool:
        kept = C.c_int64(0)
        N.check(lib.mrag_pool_select(handle, level, POOL_CAP, C.byref(kept)))
        return CandidatePool(table, handle, LEVELS[level], steps, codes, int(kept.value))

    steps: list[tuple[str, Any]] = []
    # L1: J & D & P
    if has_j and has_d and has_p:
        steps.append(("L1_JDP", n1))
        if n1:
            return done(0, j_codes + d_codes + p_codes, steps)
    else:
        steps.append(("L1_JDP", f"skip: missing kind ({'' if has_j else 'J'}{'' if has_d else 'D'}{'' if has_p else 'P'})"))
    # L2: J & D
    if has_j and has_d:
        steps.append(("L2_JD", n2))
        if n2:
            return done(1, j_codes + d_codes, steps)
    else:
        steps.append(("L2_JD", f"skip: missing {'J' if not has_j else 'D'}"))
    # L3: AHCA & D -- only when the D intersection itself is non-empty
    if has_d and n_d:
        steps.append(("L3_AHCA_D", n3))
        if n3:
            return done(2, [AHCA_TAG] + d_codes, steps)
    else:
        steps.append(("L3_AHCA_D", "skip: no D-tag"))
    # L4: AHCA only
    steps.append(("L4_AHCA", n4))
    if n4:
        return done(3, [AHCA_TAG], steps)
    # L5: empty -- the agent bootstraps through the broad vector search
    steps.append(("L5_empty", 0))
    lib.mrag_pool_destroy(handle)
    return CandidatePool(table, None, "L5_empty", steps, [], 0, document_ids=[])


def augment_pool_with_inheritance(pool: CandidatePool, inherited_ids: Sequence[str]) -> CandidatePool:
    """`_augment_pool_with_inheritance` (corpus_search_agent.py:1966-2002): union the inherited-authority documents into a
    plan-scoped (L1 / L2) pool; they are tracked separately so a caller can rerank them without the payer floor."""
    if not inherited_ids or pool.cascade_level not in ("L1_JDP", "L2_JD") or pool._handle is None:
        return pool
    t = pool._table
    have = set(int(d) for d in pool.doc_indices())
    seen_ids = set()
    added, add_idx = [], []
    for did in inherited_ids:
        did = str(did)
        d = t.doc_idx.get(did)
        if did in seen_ids or (d is not None and d in have):
            continue
        seen_ids.add(did)
        added.append(did)                                 # the reference does not check that the id is a known document
        if d is not None:
            have.add(d)
            add_idx.append(d)
    added = added[:max(0, POOL_CAP - pool.size)]
    if not added:
        return pool
    keep = set(added)
    arr = np.asarray([d for d in add_idx if t.doc_ids[d] in keep], dtype=np.uint32)
    if arr.size:
        N.check(N.load().mrag_pool_add_docs(pool._handle, arr.ctypes.data, arr.shape[0]))
    out = CandidatePool(t, pool._handle, pool.cascade_level,
                        list(pool.cascade_steps or []) + [("inherited_authority_union", len(added))],
                        pool.intersect_codes, pool.size + len(added), inherited_document_ids=added)
    out._unknown_ids = [x for x in added if x not in t.doc_idx]
    pool._handle = None                                   # ownership of the device bitmap moves to the new object
    return out

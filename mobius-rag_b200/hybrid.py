"""Hybrid rerank fused with the vector scan (BASELINE.json config 5).

Reference: ``_rerank`` + ``_best_arm_sim`` (app/services/corpus_search.py:1787-1814, 1909-2297), which
score <= ~3k candidates per query in a Python loop of substring tests.  Here everything that needs
TEXT is evaluated once per row when it is indexed (``chunk_features``) and shipped to the GPU as
bits; the score of a (row, query) pair is then a few integer / fp32 instructions inside the scan
(csrc/hybrid.cuh), taken over ALL rows that pass the filter, and the top k come back.

Host side of that split:
  * the text rules of the reference, restated: haystacks (1844-1906), JPD patterns and classifier
    (233-349), length / authority scores (215-225, 1773-1784), contact regexes (666-683);
  * ``HybridTable``: per-row features + document j-tags on top of a PublishedTable;
  * ``hybrid_rerank(...)``: the call with `_rerank`'s argument names; returns the reranked dicts.
"""
from __future__ import annotations

import ctypes as C
import math
import re
from typing import Any, Sequence

import numpy as np

from . import _native as N
from .corpus_search import _row_to_base_dict
from .index import FEAT_DTYPE, Filter
from .table import PublishedTable, to_float4

# ---------------------------------------------------------------------------------------------
# data tables of the reference (corpus_search.py:215-309) -- values, not code
# ---------------------------------------------------------------------------------------------
AUTHORITY_WEIGHTS = {
    "contract_source_of_truth": 1.0, "payer_website": 0.75, "operational_suggested": 0.65,
    "payer_policy": 0.50, "fyi_not_citable": 0.20,
}
AUTHORITY_DEFAULT = 0.10
CHUNK_TAG_BOOST = 1.5            # app/config.py:129
TAG_COVERAGE_FLOOR = 1.0         # corpus_search.py:604
CONFIDENCE = (("high", 0.55), ("medium", 0.35), ("low", 0.18))      # corpus_search.py:381-383

JPD_PATTERNS: dict[str, list[str]] = {
    "prior_authorization_required": [
        "prior authorization", "prior auth", "pre-authorization", "pre auth", "pa required",
        "requires authorization", "authorization required", "authorization criteria", "medical necessity",
        "utilization management", "um criteria", "medically necessary", "clinical criteria", "level of care",
        "admission criteria", "inpatient criteria", "coverage criteria", "criteria for", "clinical guidelines",
        "covered criteria"],
    "claims_authorization_submissions": [
        "claims submission", "claim form", "billing code", "cpt code", "hcpcs", "procedure code", "revenue code",
        "submit claim", "claim adjudication", "authorization number", "pa number", "claims processing", "remittance"],
    "member_eligibility_molina": [
        "member eligibility", "eligibility verification", "enrollment", "eligible member", "covered services",
        "plan benefit", "benefit coverage", "covered under", "who is eligible", "who qualifies", "eligible for",
        "beneficiary"],
    "benefit_access_limitations": [
        "limitation", "exclusion", "not covered", "non-covered", "benefit limit", "annual limit", "visit limit",
        "frequency limit", "coverage limit", "service limit", "maximum benefit", "out-of-network", "out of network"],
    "coordination_of_benefits": [
        "coordination of benefits", "cob", "dual coverage", "other insurance", "third party liability", "tpl",
        "primary payer", "secondary payer"],
    "compliant_claim_requirements": [
        "documentation required", "required documentation", "supporting documentation", "clinical documentation",
        "medical records", "clinical notes", "progress notes", "treatment plan", "discharge summary",
        "clinical record", "what documentation", "documentation needed", "records required", "supporting evidence",
        "clinical evidence", "chart notes"],
    "credentialing": [
        "credentialing", "credential", "provider enrollment", "network enrollment", "network participation",
        "in-network", "provider qualification", "licensure", "certification", "provider manual",
        "participating provider"],
    "claim_submission_important": [
        "timely filing", "filing deadline", "claim deadline", "corrected claim", "claim adjustment", "resubmission"],
    "claim_disputes": [
        "appeal", "grievance", "dispute", "reconsideration", "denial", "denied claim", "adverse determination",
        "fair hearing", "redetermination"],
    "contacting_marketing_members": [
        "contact member", "member outreach", "member communication", "member notification"],
    "contact_info": [
        "phone", "fax", "telephone", "call", "hotline", "toll-free", "toll free", "contact number",
        "provider services phone", "member services phone", "provider phone", "member phone", "edi", "payer id",
        "payer number", "1-800", "1-866", "1-877", "1-888"],
}
JPD_CATS = list(JPD_PATTERNS)            # dictionary order == order of the GPU's category table
assert len(JPD_CATS) == N.MRAG_JPD_CATS
assert [len(JPD_PATTERNS[c]) for c in JPD_CATS] == [20, 13, 12, 13, 8, 16, 11, 6, 9, 4, 19]   # csrc/hybrid.cuh kJpdPatterns

CONTACT_QUERY_RE = re.compile(
    r"\b(phone|fax|telephone|hotline|toll[\s.\-]?free|"
    r"edi\s+payer[\s\-]?id|payer[\s\-]?id|edi[\s\-]?id|"
    r"contact\s+number|provider\s+(services|phone|contact)|"
    r"member\s+(services|phone|contact))\b", re.IGNORECASE)
CONTACT_VALUE_RE = re.compile(
    r"(?<!\d)(?:1-\d{3}-\d{3}-\d{4}|\(\d{3}\)\s*\d{3}[-.\s]\d{4}|\d{3}-\d{3}-\d{4}|\d{3}\.\d{3}\.\d{4})(?!\d)")


# ---------------------------------------------------------------------------------------------
# text rules: the haystacks `_rerank` substring-tests (corpus_search.py:1844-1906), produced from cached pieces
# ---------------------------------------------------------------------------------------------
_SEP = " | "
_PATHLIKE = str.maketrans({"_": " ", "-": " ", "/": " ", ".": " "})      # separators that pack several words into one token
_TAGLIKE = str.maketrans({".": " ", "_": " "})


def _squash(value) -> str:
    """lower case, runs of white space -> one blank; '' for None / empty (what every haystack piece goes through)"""
    return " ".join(str(value).lower().split()) if value else ""


def _tag_words(tags) -> list[str]:
    """An inherited document tag contributes its leaf (``a.b.prior_authorization`` -> ``prior authorization``) and, when
    different, its whole dotted path with the separators blanked (:1893-1905)."""
    words: list[str] = []
    for tag in tags or ():
        text = str(tag)
        leaf = text.rsplit(".", 1)[-1].replace("_", " ").strip().lower()
        whole = text.translate(_TAGLIKE).strip().lower()
        if leaf:
            words.append(leaf)
        if whole and whole != leaf:
            words.append(whole)
    return words


def _field_pieces(value, pathlike: bool) -> list[str]:
    """one metadata field -> its squashed text, plus a separator-blanked copy for filenames / paths (:1879-1890)"""
    if not value:
        return []
    pieces = [_squash(value)]
    if pathlike:
        pieces.append(_squash(str(value).translate(_PATHLIKE)))
    return [p for p in pieces if p]


def body_haystack(c: dict) -> str:
    """the chunk body and, when the enrichment step attached them, its neighbour paragraphs"""
    return _SEP.join(p for p in (_squash(c.get("text")), _squash(c.get("_neighbor_text"))) if p)


def meta_haystack(c: dict) -> str:
    """document-level text that belongs to the chunk: names, payer / state, paths, summary, inherited tag words --
    in the reference's field order, so the joined string is the same string"""
    pieces: list[str] = []
    for key, pathlike in (("document_name", False), ("document_filename", True), ("document_display_name", False),
                          ("payer", False), ("state", False), ("section_path", True), ("chapter_path", True), ("summary", False)):
        pieces += _field_pieces(c.get(key), pathlike)
    for key in ("_doc_d_tags", "_doc_j_tags", "_doc_p_tags"):
        pieces += _tag_words(c.get(key))
    return _SEP.join(pieces)


def normalise_for_haystack(s) -> str:
    return _squash(s)


def jpd_hits(text: str) -> tuple[list[int], bool]:
    """Per category: how many of its patterns occur in ``text``; and whether the text is 'short'
    (<= 20 words, scored hits / sqrt(n) instead of hits / n) -- _classify_jpd, :324-349."""
    lower = text.lower()
    short = len(lower.split()) <= 20
    return [sum(1 for p in JPD_PATTERNS[cat] if p in lower) for cat in JPD_CATS], short


def classify_jpd(text: str) -> dict[str, float]:             # :324-349
    hits, short = jpd_hits(text)
    out = {}
    for cat, h in zip(JPD_CATS, hits):
        if h:
            n = len(JPD_PATTERNS[cat])
            out[cat] = min(1.0, h / math.sqrt(n)) if short else min(1.0, h / n)
    return out


def length_score(t: str | None) -> float:                    # :1779-1784
    n = len(t or "")
    if n < 50:
        return 0.0
    return min(1.0, (n - 50) / 450)


def authority_score(level: str | None) -> float:             # :1773-1776
    if not level:
        return AUTHORITY_DEFAULT
    return AUTHORITY_WEIGHTS.get((level or "").strip().lower(), AUTHORITY_DEFAULT)


def confidence_label(score: float) -> str:                   # :2307-2314
    for name, lo in CONFIDENCE:
        if score >= lo:
            return name
    return "abstain"


# ---------------------------------------------------------------------------------------------
# the table with text features
# ---------------------------------------------------------------------------------------------
class HybridTable:
    """PublishedTable + the per-row features the fused rerank reads.

    ``phrases`` is the phrase dictionary (<= 128 entries): the required phrases the features carry a presence bit for.
    A query phrase outside it is ADDED on first use (one pass over the texts), never silently treated as absent.
    Features are derived data: a change of a document's tags (d / j / p: they feed the meta haystack) or of the
    promoted set marks the rows concerned dirty and ``build_features`` recomputes them."""

    def __init__(self, table: PublishedTable, phrases: Sequence[str]):
        self.table = table
        self.phrases: list[str] = []
        for p in phrases:
            p = (p or "").lower()
            if p and p not in self.phrases:
                self.phrases.append(p)
        if len(self.phrases) > N.MRAG_PHRASE_WORDS * 64:
            raise ValueError(f"phrase dictionary holds at most {N.MRAG_PHRASE_WORDS * 64} phrases")
        self.phrase_index = {p: i for i, p in enumerate(self.phrases)}
        self.dcodes: dict[str, int] = {}       # chunk d-tag key -> code >= 1
        self.promoted: set[int] = set()
        self._built = 0
        self._dirty_rows: set[int] = set()
        self._dirty_docs: set[int] = set()
        self._overflow: dict[int, list[int]] = {}      # row -> d-tag codes beyond the four inline slots
        self._id_rank: np.ndarray | None = None        # row -> rank of its id in ascending id order (the d-tag arm's ORDER BY)
        table.doc_listeners.append(self._dirty_docs.add)

    # -- document j-tags live in the table (document_tags.j_tags, app/models.py:535-537) ----------
    @property
    def jbits(self) -> dict[str, int]:
        return self.table.vocab._jtag_bit

    @property
    def doc_j_tags(self) -> dict[str, list]:
        return self.table.doc_j_tags

    def set_document_j_tags(self, document_id: str, j_tags: Sequence[str]) -> None:
        self.table.set_document_j_tags(document_id, j_tags)

    def set_promoted(self, rows: Sequence[int]) -> None:
        """rows promoted from a seed / inheriting a BM25 score: exempt from the coverage floor (:2205-2208)"""
        new = set(int(r) for r in rows)
        self._dirty_rows |= (new ^ self.promoted)
        self.promoted = new

    def candidate_dict(self, r: int) -> dict:
        """The dict `_rerank` would see for row r before scoring (base dict + inherited doc tags)."""
        t = self.table
        c = _row_to_base_dict(t, r)
        did = t.document_id[r]
        if did in t.doc_d_tags or did in t.doc_j_tags:
            c["_doc_d_tags"] = sorted(t.doc_d_tags.get(did, ()))
            c["_doc_j_tags"] = list(t.doc_j_tags.get(did, ()))
            c["_doc_p_tags"] = sorted(t.doc_p_tags.get(did, ()))
        if r in self.promoted:
            c["_promoted_from_seed"] = "seed"
        return c

    def chunk_features(self, r: int) -> tuple:
        c = self.candidate_dict(r)
        body = body_haystack(c)
        meta = meta_haystack(c)
        bits = [0] * N.MRAG_PHRASE_WORDS
        for p, i in self.phrase_index.items():
            if p in body or (meta and p in meta):
                bits[i >> 6] |= 1 << (i & 63)
        hits, short = jpd_hits(body)
        codes = [self.dcodes.setdefault(key, len(self.dcodes) + 1) for key in (c.get("chunk_d_tags") or {})]
        if len(self.dcodes) >= 0xFFFF:
            raise ValueError("more than 65534 distinct chunk d-tag keys")
        flags = (N.CF_SHORT_TEXT if short else 0) | (N.CF_CONTACT_VALUE if CONTACT_VALUE_RE.search(c.get("text") or "") else 0) \
            | (N.CF_PROMOTED if r in self.promoted else 0) | (N.CF_DTAG_OVERFLOW if len(codes) > 4 else 0)
        if len(codes) > 4:
            self._overflow[r] = codes[4:]                 # `chunk_d_tags ? key` matches ANY key: nothing is dropped
        else:
            self._overflow.pop(r, None)
        return bits, [min(255, h) for h in hits], flags, length_score(c.get("text") or ""), (codes + [0, 0, 0, 0])[:4]

    def _features_of(self, rows: Sequence[int]) -> np.ndarray:
        feat = np.zeros(len(rows), dtype=FEAT_DTYPE)
        for i, r in enumerate(rows):
            bits, hits, flags, ls, dt = self.chunk_features(int(r))
            feat["phrase_bits"][i] = bits
            feat["jpd_hits"][i] = hits
            feat["flags"][i] = flags
            feat["length_score"][i] = ls
            feat["dtags"][i] = dt
        return feat

    def build_features(self) -> None:
        """Compute the features of every row not yet covered and of every row whose inputs changed; upload them."""
        t = self.table
        with t.lock:
            n = len(t)
            over_before = dict(self._overflow)
            stale = set(r for r in self._dirty_rows if r < self._built)
            if self._dirty_docs:
                docs = np.fromiter(self._dirty_docs, dtype=np.uint32)
                stale |= set(np.flatnonzero(np.isin(t.row_doc[:self._built], docs)).tolist())
            self._dirty_rows.clear()
            self._dirty_docs.clear()
            if stale:
                rows = np.asarray(sorted(stale), dtype=np.int64)
                cuts = np.flatnonzero(np.diff(rows) != 1) + 1          # upload run by run (a document's rows are contiguous)
                for run in np.split(rows, cuts):
                    t.index.set_chunk_features(int(run[0]), self._features_of(run))
            if n > self._built:
                t.index.set_chunk_features(self._built, self._features_of(range(self._built, n)))
                self._built = n
            if self._overflow != over_before:
                self._upload_overflow()

    def _upload_overflow(self) -> None:
        pairs = sorted((r, c) for r, codes in self._overflow.items() for c in codes)
        self.table.index.set_dtag_overflow(np.asarray([p[0] for p in pairs], dtype=np.int64), np.asarray([p[1] for p in pairs], dtype=np.uint16))

    def ensure_phrases(self, phrases: Sequence[str]) -> None:
        """Every required phrase of a query must be in the dictionary.  New ones are added (their presence bit is computed
        for all rows: one substring pass over the haystacks) -- `_rerank` would substring-test them, so they must not be
        treated as absent.  A full dictionary is an error, not a silently wrong coverage."""
        new = []
        for p in phrases or ():
            p = (p or "").lower()
            if p and p not in self.phrase_index and p not in new:
                new.append(p)
        if not new:
            return
        if len(self.phrases) + len(new) > N.MRAG_PHRASE_WORDS * 64:
            raise ValueError(f"phrase dictionary is full ({N.MRAG_PHRASE_WORDS * 64}); rebuild the HybridTable with the query bank's phrases")
        with self.table.lock:
            for p in new:
                self.phrase_index[p] = len(self.phrases)
                self.phrases.append(p)
            self._dirty_rows |= set(range(self._built))            # recomputed by the next build_features()

    def id_rank(self) -> np.ndarray:
        """row -> position of its id in ascending id order (uuid order == order of the canonical lowercase text), cached
        until the table grows: the d-tag arm orders by (authority tier, id) and must not sort strings per call."""
        t = self.table
        n = len(t)
        if self._id_rank is None or self._id_rank.shape[0] != n:
            col = t.id
            lens = np.diff(col.off[:n + 1])
            if n and (lens == lens[0]).all() and lens[0] > 0:
                keys = col.buf[:int(col.off[n])].reshape(n, int(lens[0])).view(f"S{int(lens[0])}").ravel()
            else:
                keys = np.asarray([col[i] or "" for i in range(n)], dtype=object)
            order = np.argsort(keys, kind="stable")
            rank = np.empty(n, dtype=np.int64)
            rank[order] = np.arange(n)
            self._id_rank = rank
        return self._id_rank

    # -- query side -----------------------------------------------------------------------------
    def hybrid_query(self, query: str, required_phrases, required_phrase_weights, required_phrase_tag_codes) -> N.HybridQuery:
        """The per-query constants of `_rerank` (:1950-2011) as an mrag_hybrid_query."""
        hq = N.HybridQuery()
        query_cats = classify_jpd(query) if query else {}
        required_lower = [p.lower() for p in (required_phrases or []) if p]
        if len(required_lower) > N.MRAG_HYB_MAX_PHRASES:
            raise ValueError(f"at most {N.MRAG_HYB_MAX_PHRASES} required phrases")
        w = required_phrase_weights
        if w and len(w) == len(required_lower) and any(x > 0 for x in w):
            weights = [max(0.0, float(x)) for x in w]
        else:
            weights = [1.0] * len(required_lower)
        codes = list(required_phrase_tag_codes) if (required_phrase_tag_codes and len(required_phrase_tag_codes) == len(required_lower)) \
            else [None] * len(required_lower)
        hq.n_phrases = len(required_lower)
        for i, p in enumerate(required_lower):
            hq.phrase_weight[i] = weights[i]
            hq.phrase_bit[i] = self.phrase_index.get(p, -1)
            hq.phrase_jbit[i] = -1
            hq.phrase_dcode[i] = 0
            code = codes[i]
            if code and code.startswith("j:"):
                hq.phrase_jbit[i] = self.jbits.get(code.split(":", 1)[1], -1)
            if code and code.startswith("d:"):
                hq.phrase_dcode[i] = self.dcodes.get(code[2:], 0)
        for c, cat in enumerate(JPD_CATS):
            hq.qcat[c] = float(query_cats.get(cat, 0.0))
        v = self.table.vocab.authority
        for code in range(32):
            hq.auth_score[code] = AUTHORITY_DEFAULT
        for code, s in enumerate(v.values):
            if code < 31:
                hq.auth_score[code] = authority_score(s)        # codes >= 31 fall to the default, like unknown levels
        hq.w_sim, hq.w_auth, hq.w_len = 0.25, 0.10, 0.05
        hq.w_jpd = 0.20 if query_cats else 0.0
        hq.w_cov = 0.55 if required_lower else 0.0
        hq.boost = CHUNK_TAG_BOOST
        hq.floor = TAG_COVERAGE_FLOOR
        hq.contact_query = 1 if (query and CONTACT_QUERY_RE.search(query)) else 0
        return hq


def hybrid_rerank(ht: HybridTable, query_embedding: Sequence[float], k: int, query: str = "",
                  required_phrases: Sequence[str] | None = None, required_phrase_weights: Sequence[float] | None = None,
                  required_phrase_tag_codes: Sequence[str | None] | None = None, filters: Any = None,
                  include_document_ids: Sequence[str] | None = None) -> list[dict]:
    """`_rerank` over EVERY row that passes the filters (vector arm only), top k by rerank score.

    Returns dicts shaped like `_rerank`'s output: base dict + similarity / arm_scores / _arm /
    rerank_score / confidence_label, sorted by rerank_score descending.  The per-(arm, source_type)
    0.6 x best decay (:2258-2285) is exact: every category gets its own query slot in the fused scan, so
    each category's best and its top k are known; the host decays each list and merges them."""
    t = ht.table
    ht.ensure_phrases(required_phrases)
    ht.build_features()
    q = to_float4(query_embedding)
    base = ht.hybrid_query(query, required_phrases, required_phrase_weights, required_phrase_tag_codes)
    # one query slot per decay category: "vector_<source_type>", where _row_to_base_dict maps NULL / ''
    # to 'hierarchical' (corpus_search.py:573).  The scan serves up to 4 slots per pass.
    cats: dict[str, list[int]] = {}
    v = t.vocab.source_type
    for code, s in enumerate(v.values):
        cats.setdefault(s or "hierarchical", []).append(code)
    cats.setdefault("hierarchical", []).append(v.none_code)
    names = sorted(cats)
    hq = (N.HybridQuery * len(names))()
    for i, name in enumerate(names):
        C.memmove(C.addressof(hq[i]), C.addressof(base), C.sizeof(N.HybridQuery))
        for code in cats[name]:
            hq[i].source_type_any[code >> 6] |= 1 << (code & 63)
    flt: Filter = t.filter_corpus(filters, include_document_ids)
    k = int(k)
    Q = np.repeat(q[None, :], len(names), axis=0)
    scores, cos, rows, counts = t.index.search_hybrid(Q, min(k, N.MRAG_FUSED_K), hq, flt if flt.active else None)
    out = []
    for i in range(len(names)):
        n_i = int(counts[i])
        if n_i == 0:
            continue
        best = float(scores[i, 0])
        for j in range(n_i):
            sc = float(scores[i, j])
            if best > 0 and sc < 0.6 * best:               # per-category decay (:2258-2285); the list is sorted, so stop
                break
            out.append((sc, int(rows[i, j]), float(cos[i, j])))
    out.sort(key=lambda x: (-x[0], x[1]))
    res = []
    for sc, r, cs in out[:k]:
        c = ht.candidate_dict(r)
        c["similarity"] = cs
        c["arm_scores"] = {"vector": cs}
        c["_arm"] = "vector"
        c["retrieval_arms"] = ["vector"]
        c["rerank_score"] = sc
        c["confidence_label"] = confidence_label(sc)
        res.append(c)
    return res


# ---------------------------------------------------------------------------------------------
# the d-tag arm and reciprocal rank fusion (corpus_search.py:1605-1766)
# ---------------------------------------------------------------------------------------------
RRF_K = 60                        # corpus_search.py:394


def _authority_tier(level: str | None) -> int:
    """ORDER BY CASE document_authority_level ... of the d-tag arm (corpus_search.py:1674-1678)."""
    return 0 if level == "contract_source_of_truth" else 1 if level == "operational" else 2


def dtag_arm(ht: "HybridTable", dtag_keys: Sequence[str], k: int, filters: Any = None,
             include_document_ids: Sequence[str] | None = None, search_id: str = "", idf_mode: bool = False) -> list[dict]:
    """`_dtag_arm`: chunks whose chunk_d_tags hold any of the keys, ordered by (authority tier, id), LIMIT k;
    similarity is the constant 0.5; with ``idf_mode`` each chunk carries ``_dtag_idf`` = ln(n_total / pool_size)
    of its highest-IDF key.  The WHERE and the counts run on the GPU (dtag_mask_kernel); ordering k rows by
    (tier, id) is done here over the matches."""
    if not dtag_keys:
        return []
    t = ht.table
    try:
        ht.build_features()
        keys = list(dtag_keys)
        flt: Filter = t.filter_corpus(filters, include_document_ids)
        n = len(t)
        rows, n_total, counts = t.index.dtag_rows(flt, [ht.dcodes.get(key, 0xFFFF) for key in keys])   # 0xFFFF: a key no chunk has
        per_key = dict(zip(keys, counts))
        idf_weights: dict[str, float] = {}
        if idf_mode:
            for key in keys:
                idf_weights[key] = math.log(max(1, n_total or 1) / max(1, per_key.get(key, 0) or 1))
        # ORDER BY CASE authority tier, id LIMIT k -- on integer keys: tier from the authority code, id through its cached rank
        tier_of_code = np.full(256, 2, dtype=np.int64)
        for code, level in enumerate(t.vocab.authority.values):
            tier_of_code[code] = _authority_tier(level)
        key64 = tier_of_code[t.document_authority_level.view()[rows]] * max(n, 1) + ht.id_rank()[rows]
        kk = max(0, int(k))
        if rows.size > kk > 0:
            sel = np.argpartition(key64, kk - 1)[:kk]
            rows, key64 = rows[sel], key64[sel]
        order = rows[np.argsort(key64, kind="stable")][:kk].tolist()
    except Exception as exc:                       # fail-soft like the reference (:1684-1686)
        import logging
        logging.getLogger(__name__).warning("corpus_search dtag arm failed: %s", exc)
        return []
    out = []
    for r in order:
        c = _row_to_base_dict(t, r)
        c["similarity"] = 0.5
        c["match_score"] = 0.5
        c["_arm"] = "dtag"
        if idf_mode and idf_weights:
            chunk_dtags = c.get("chunk_d_tags") or {}
            c["_dtag_idf"] = max((idf_weights[key] for key in keys if key in chunk_dtags), default=1.0)
        out.append(c)
    return out


def rrf_merge(arms: dict[str, list[dict]], k: int = RRF_K, search_id: str = "") -> list[dict]:
    """Reciprocal rank fusion with the reference's outputs (`_rrf_merge`, corpus_search.py:1708-1766), done as array
    accumulation over chunk slots.

    Pass 1 gives every distinct chunk id a slot in first-seen order and records, per arm, the (slot, rank, weight,
    similarity) of each hit; the per-slot score  sum over arms of  idf / (k + rank)  is accumulated in that same
    order, so the float sums are the reference's.  Pass 2 builds one dict per slot: the first arm's dict wins, blanks
    (None, '', []) are filled from later arms.  Order: score descending, best rank ascending, first-seen order last
    (a stable sort over the slots); `similarity` becomes the fused score."""
    BOOKKEEPING = ("retrieval_arms", "arm_ranks", "arm_scores", "rrf_score")
    slot_of: dict[str, int] = {}
    merged: list[dict] = []
    score: list[float] = []
    for arm, ranked in arms.items():
        for rank, hit in enumerate(ranked, start=1):
            cid = hit.get("id") or ""
            if not cid:
                continue
            share = float(hit.get("_dtag_idf") or 1.0) / (k + rank)
            sim = float(hit.get("similarity", 0.0))
            slot = slot_of.get(cid)
            if slot is None:
                slot_of[cid] = len(merged)
                first = dict(hit)
                first.update(retrieval_arms=[arm], arm_ranks={arm: rank}, arm_scores={arm: sim})
                merged.append(first)
                score.append(share)
                continue
            entry = merged[slot]
            for field_name, value in hit.items():
                if field_name not in BOOKKEEPING and entry.get(field_name) in (None, "", []) and value not in (None, "", []):
                    entry[field_name] = value
            if arm not in entry["retrieval_arms"]:
                entry["retrieval_arms"].append(arm)
            entry["arm_ranks"][arm] = rank
            entry["arm_scores"][arm] = sim
            score[slot] += share
    if not merged:
        return []
    total = np.asarray(score, dtype=np.float64)
    best = np.fromiter((min(e["arm_ranks"].values()) for e in merged), dtype=np.int64, count=len(merged))
    order = np.lexsort((best, -total))                   # stable: ties keep first-seen order, like sorted() over the dict
    out = []
    for slot in order:
        entry = merged[int(slot)]
        entry["rrf_score"] = float(total[slot])
        entry["similarity"] = entry["rrf_score"]
        out.append(entry)
    return out

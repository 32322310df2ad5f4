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
# text rules
# ---------------------------------------------------------------------------------------------
def normalise_for_haystack(s) -> str:                        # :1844-1847
    if not s:
        return ""
    return " ".join(str(s).lower().split())


def body_haystack(c: dict) -> str:                           # :1850-1856
    parts = [normalise_for_haystack(c.get("text"))]
    nbr = normalise_for_haystack(c.get("_neighbor_text"))
    if nbr:
        parts.append(nbr)
    return " | ".join(p for p in parts if p)


def meta_haystack(c: dict) -> str:                           # :1859-1906
    parts: list[str] = []
    for key in ("document_name", "document_filename", "document_display_name", "payer", "state", "section_path",
                "chapter_path", "summary"):
        v = c.get(key)
        if not v:
            continue
        parts.append(normalise_for_haystack(v))
        if key in ("document_filename", "section_path", "chapter_path"):
            split = normalise_for_haystack(str(v).replace("_", " ").replace("-", " ").replace("/", " ").replace(".", " "))
            if split:
                parts.append(split)
    for tag_key in ("_doc_d_tags", "_doc_j_tags", "_doc_p_tags"):
        for tag in (c.get(tag_key) or []):
            leaf = str(tag).split(".")[-1].replace("_", " ").strip().lower()
            if leaf:
                parts.append(leaf)
            full = str(tag).replace(".", " ").replace("_", " ").strip().lower()
            if full and full != leaf:
                parts.append(full)
    return " | ".join(p for p in parts if p)


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
    """PublishedTable + the per-row features / document j-tags the fused rerank reads.

    ``phrases`` is the phrase dictionary: the required phrases of the query bank (<= 128), fixed when
    the features are built.  A query phrase outside the dictionary is treated as present nowhere."""

    def __init__(self, table: PublishedTable, phrases: Sequence[str]):
        self.table = table
        self.phrases = []
        for p in phrases:
            p = (p or "").lower()
            if p and p not in self.phrases:
                self.phrases.append(p)
        if len(self.phrases) > N.MRAG_PHRASE_WORDS * 64:
            raise ValueError(f"phrase dictionary holds at most {N.MRAG_PHRASE_WORDS * 64} phrases")
        self.phrase_index = {p: i for i, p in enumerate(self.phrases)}
        self.dcodes: dict[str, int] = {}       # chunk d-tag key -> code >= 1
        self.jbits: dict[str, int] = {}        # document j-tag key -> bit
        self.doc_j_tags: dict[str, list[str]] = {}
        self.promoted: set[int] = set()
        self._built = 0

    # -- document j-tags (document_tags.j_tags, app/models.py:535-537) --------------------------
    def set_document_j_tags(self, document_id: str, j_tags: Sequence[str]) -> None:
        t = self.table
        with t.lock:
            self.doc_j_tags[str(document_id)] = list(j_tags or ())
            bits = np.zeros((1, N.MRAG_JTAG_WORDS), dtype=np.uint64)
            for key in j_tags or ():
                b = self.jbits.setdefault(key, len(self.jbits))
                if b >= N.MRAG_JTAG_WORDS * 64:
                    raise ValueError("too many distinct j-tag codes")
                bits[0, b >> 6] |= np.uint64(1 << (b & 63))
            t.index.set_doc_jtags(t._doc(str(document_id)), bits)

    def candidate_dict(self, r: int) -> dict:
        """The dict `_rerank` would see for row r before scoring (base dict + inherited doc tags)."""
        t = self.table
        c = _row_to_base_dict(t, r)
        did = t.document_id[r]
        if did in t.doc_d_tags or did in self.doc_j_tags:
            c["_doc_d_tags"] = sorted(t.doc_d_tags.get(did, ()))
            c["_doc_j_tags"] = list(self.doc_j_tags.get(did, ()))
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
        flags = (N.CF_SHORT_TEXT if short else 0) | (N.CF_CONTACT_VALUE if CONTACT_VALUE_RE.search(c.get("text") or "") else 0) \
            | (N.CF_PROMOTED if r in self.promoted else 0)
        dt = [0, 0, 0, 0]
        for j, key in enumerate(list(c.get("chunk_d_tags") or {})[:4]):
            dt[j] = self.dcodes.setdefault(key, len(self.dcodes) + 1)
        return bits, [min(255, h) for h in hits], flags, length_score(c.get("text") or ""), dt

    def build_features(self) -> None:
        """(Re)compute the features of every row not yet covered and upload them."""
        t = self.table
        with t.lock:
            n = len(t)
            if n == self._built:
                return
            feat = np.zeros(n - self._built, dtype=FEAT_DTYPE)
            for i, r in enumerate(range(self._built, n)):
                bits, hits, flags, ls, dt = self.chunk_features(r)
                feat["phrase_bits"][i] = bits
                feat["jpd_hits"][i] = hits
                feat["flags"][i] = flags
                feat["length_score"][i] = ls
                feat["dtags"][i] = dt
            t.index.set_chunk_features(self._built, feat)
            self._built = n

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
        if len(keys) > 32:
            raise ValueError("at most 32 d-tag keys")
        codes = (C.c_uint16 * max(1, len(keys)))(*[ht.dcodes.get(key, 0xFFFF) for key in keys])   # 0xFFFF: a key no chunk has
        flt: Filter = t.filter_corpus(filters, include_document_ids)
        n = len(t)
        mask = np.zeros((n + 31) // 32 + 1, dtype=np.uint32)
        counts = (C.c_int64 * (len(keys) + 1))()
        N.check(t.index._lib.mrag_dtag_mask(t.index._h, flt.ref() if flt.active else None, codes, len(keys),
                                            mask.ctypes.data, counts))
        rows = np.flatnonzero(np.unpackbits(mask.view(np.uint8), bitorder="little")[:n])
        idf_weights: dict[str, float] = {}
        if idf_mode:
            n_total = max(1, int(counts[0]) or 1)
            for i, key in enumerate(keys):
                idf_weights[key] = math.log(n_total / max(1, int(counts[1 + i]) or 1))
        order = sorted(rows.tolist(), key=lambda r: (_authority_tier(t.document_authority_level[r]), t.id[r]))[:max(0, int(k))]
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
    """`_rrf_merge` (corpus_search.py:1708-1766): sum over arms of idf / (k + rank), first arm's dict wins,
    blanks filled from later arms, ordered by (-rrf, best rank); `similarity` becomes the RRF score."""
    fused: dict[str, dict] = {}
    for arm_name, ranked in arms.items():
        for rank0, chunk in enumerate(ranked):
            cid = chunk.get("id") or ""
            if not cid:
                continue
            rank1 = rank0 + 1
            contribution = float(chunk.get("_dtag_idf") or 1.0) / (k + rank1)
            f = fused.get(cid)
            if f is None:
                f = fused[cid] = dict(chunk)
                f["retrieval_arms"] = [arm_name]
                f["arm_ranks"] = {arm_name: rank1}
                f["arm_scores"] = {arm_name: float(chunk.get("similarity", 0.0))}
                f["rrf_score"] = contribution
            else:
                for key, val in chunk.items():
                    if key in ("retrieval_arms", "arm_ranks", "arm_scores", "rrf_score"):
                        continue
                    if f.get(key) in (None, "", []) and val not in (None, "", []):
                        f[key] = val
                if arm_name not in f["retrieval_arms"]:
                    f["retrieval_arms"].append(arm_name)
                f["arm_ranks"][arm_name] = rank1
                f["arm_scores"][arm_name] = float(chunk.get("similarity", 0.0))
                f["rrf_score"] += contribution
    out = sorted(fused.values(), key=lambda c: (-float(c.get("rrf_score", 0.0)), min(c.get("arm_ranks", {}).values() or [999])))
    for c in out:
        c["similarity"] = float(c.get("rrf_score", 0.0))
    return out

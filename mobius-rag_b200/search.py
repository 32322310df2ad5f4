"""The retrieval driver: ``corpus_search`` (app/services/corpus_search.py:3280-3826) over the B200 table.

Same request / response models (:81-199), same stages in the same order, same clamps:

    0  k = clamp(request.k, 1, 100)                                                        (:3297)
    1  arms per mode -- corpus: bm25 + vector (k*2 each) + d-tag arm (k); precision: bm25 (+ d-tag arm on pools of
       <= 200 documents); recall: vector only, k*2 with over_fetch_factor 8 and min_similarity, skipped when the pinned
       pool holds more than 2000 chunks                                                     (:3343-3514)
    2  fuse: reciprocal rank fusion of the arms that ran, or the single arm as it is        (:3517-3543)
    3  content de-duplication on the first 400 characters of the squashed body              (:3561-3579)
    4  with required phrases: neighbour text (+-1 paragraph, +-1 page), inherited document tags, topic-block merge into
       seeds with similarity >= 0.7                                                         (:3599-3619)
    5  rerank (score on the GPU: mrag_rerank_candidates; floor, per-category decay, sort)   (:3622-3629)
    6  assemble k chunks (score | canonical_first | balanced; confidence threshold, page and content de-duplication),
       expand with neighbours                                                               (:3631-3668)
    7  CorpusChunk list + telemetry                                                         (:3670-3826)

What the GPU does: the vector arm (exact cosine scan), the d-tag arm's WHERE + IDF counts, the rerank score of every
candidate.  What stays host Python, as in the reference: RRF over <= 600 dicts, de-duplication, assembly, tracing.  What is
NOT here and is injected by the caller: the BM25 arm (Postgres full-text search) and the query
embedding (an API call) -- `bm25_arm` / `embed` callables with the reference's return shapes.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import time
import uuid
from typing import Any, Callable, Sequence

import numpy as np
from pydantic import BaseModel

from . import _native as N
from . import hybrid as H
from .corpus_search import CorpusFilters as _ArmFilters
from .corpus_search import LexiconExpansion, _log_stage, vector_arm
from .neighbors import NeighborIndex

DTAG_ARM_IDF = os.getenv("DTAG_ARM_IDF", "").lower() in ("1", "true", "yes")       # :405
DTAG_ARM_MAX_POOL_DOCS = 200                                                        # :412
TINY_POOL_CHUNK_MAX = 2000                                                          # :420
MODE_MIN = {"corpus": "low", "precision": "low", "recall": "abstain"}               # :385-389
LABEL_ORDER = ["abstain", "low", "medium", "high"]
AUTHORITY_TIER = {"contract_source_of_truth": 0, "payer_policy": 1, "operational_suggested": 1, "fyi_not_citable": 2}   # :2334-2339
NEIGHBOR_TOTAL_CAP, NEIGHBOR_PER_DOC_CAP, NEIGHBOR_SCORE_FLOOR = 50, 20, 0.15      # :2547-2553
JPD_FAMILY = {"prior_authorization_required": "P", "claims_authorization_submissions": "P", "member_eligibility_molina": "J",
              "benefit_access_limitations": "J", "coordination_of_benefits": "J", "compliant_claim_requirements": "D",
              "credentialing": "D", "claim_submission_important": "D", "claim_disputes": "D", "contacting_marketing_members": "O",
              "contact_info": "C", "other_important": "O"}                          # :309-322


class CorpusFilters(BaseModel):
    payer: str | None = None
    state: str | None = None
    program: str | None = None
    authority_level: str | None = None


class CorpusSearchRequest(BaseModel):
    query: str
    k: int = 10
    mode: str = "corpus"
    filters: CorpusFilters | None = None
    include_document_ids: list[str] | None = None
    min_similarity: float | None = None
    tag_mode: str = "auto"
    assembly_strategy: str = "score"
    canonical_floor: float = 0.5
    required_phrases: list[str] | None = None
    required_phrase_weights: list[float] | None = None
    required_phrase_tag_codes: list[str | None] | None = None
    neighbor_paragraph_window: int = 2
    neighbor_page_window: int = 1


class CorpusChunk(BaseModel):
    id: str
    text: str
    document_id: str
    document_name: str
    page_number: int | None
    paragraph_index: int | None
    source_type: str
    similarity: float
    rerank_score: float
    confidence_label: str
    retrieval_arms: list[str]
    authority_level: str | None
    payer: str | None
    state: str | None
    jpd_tags: list[str] = []
    section_path: str | None = None
    chapter_path: str | None = None
    summary: str | None = None
    is_neighbor: bool = False


class CorpusSearchResponse(BaseModel):
    chunks: list[CorpusChunk]
    telemetry: dict[str, Any]


def authority_tier(level: str | None) -> int:
    return AUTHORITY_TIER.get((level or "").strip().lower(), 3)


def preview(text: str, n: int = 120) -> str:
    s = (text or "").replace("\n", " ").strip()
    return s[:n] + ("…" if len(s) > n else "")


def best_arm_sim(c: dict) -> float:
    """`_best_arm_sim` (:1787-1814): the vector arm's cosine rescaled from [0.5, 1] to [0, 1], every other arm's score raw;
    a single-arm candidate carries its raw score under `similarity`."""
    arms = c.get("arm_scores") or {}
    if not arms:
        return float(c.get("similarity") or 0.0)
    return max([0.0] + [max(0.0, (float(s) - 0.5) * 2.0) if a == "vector" else float(s) for a, s in arms.items()])


# ---------------------------------------------------------------------------------------------
# stage 5: rerank of a candidate list
# ---------------------------------------------------------------------------------------------
def rerank_candidates(ht: "H.HybridTable", chunks: list[dict], search_id: str = "", query: str = "",
                      required_phrases: Sequence[str] | None = None, required_phrase_weights: Sequence[float] | None = None,
                      required_phrase_tag_codes: Sequence[str | None] | None = None) -> list[dict]:
    """`_rerank` (:1909-2297) with the scoring on the GPU.  The host turns each candidate's haystacks into presence bits
    (the request's own phrases are the dictionary), the kernel evaluates coverage / floor / score, the host applies the
    per-(arm, source_type) 0.6 x best decay and sorts."""
    if not chunks:
        return chunks
    t = ht.table
    phrases = [p.lower() for p in (required_phrases or []) if p]
    if len(phrases) > N.MRAG_HYB_MAX_PHRASES:
        raise ValueError(f"at most {N.MRAG_HYB_MAX_PHRASES} required phrases")
    local = H.HybridTable.__new__(H.HybridTable)          # a per-request phrase dictionary: bit i = phrase i
    local.table, local.phrases, local.phrase_index = t, phrases, {p: i for i, p in enumerate(dict.fromkeys(phrases))}
    local.dcodes = {}
    hq = local.hybrid_query(query, required_phrases, required_phrase_weights, required_phrase_tag_codes)
    codes = list(required_phrase_tag_codes) if (required_phrase_tag_codes and len(required_phrase_tag_codes) == len(phrases)) \
        else [None] * len(phrases)
    for i, code in enumerate(codes):                       # j-codes resolve against the TABLE's j-tag bits
        hq.phrase_jbit[i] = t.vocab._jtag_bit.get(code.split(":", 1)[1], -1) if (code and code.startswith("j:")) else -1
        hq.phrase_dcode[i] = 0                             # chunk d-tag matches are resolved on the host (dtag_match below)
    d_bodies = [code[2:] for code in codes if code and code.startswith("d:")]
    levels: dict[str | None, int] = {}
    cands = (N.Candidate * len(chunks))()
    jpd_tags: list[list[str]] = []
    for i, c in enumerate(chunks):
        body, meta = H.body_haystack(c), (H.meta_haystack(c) if phrases else "")
        cd = cands[i]
        for p, b in local.phrase_index.items():
            if p in body or (meta and p in meta):
                cd.feat.phrase_bits[b >> 6] |= 1 << (b & 63)
        hits, short = H.jpd_hits(body)
        for j, h in enumerate(hits):
            cd.feat.jpd_hits[j] = min(255, h)
        promoted = c.get("_promoted_from_seed") is not None or "bm25_inherited" in (c.get("retrieval_arms") or [])
        cd.feat.flags = (N.CF_SHORT_TEXT if short else 0) | (N.CF_CONTACT_VALUE if H.CONTACT_VALUE_RE.search(c.get("text") or "") else 0) \
            | (N.CF_PROMOTED if promoted else 0)
        cd.feat.length_score = H.length_score(c.get("text") or "")
        cd.sim = best_arm_sim(c)
        # binary j-tag credit needs the inherited tags to have been attached (stage 4), like the reference's c["_doc_j_tags"]
        d = t.doc_idx.get(str(c.get("document_id") or "")) if c.get("_doc_j_tags") else None
        cd.doc_idx = 0xFFFFFFFF if d is None else d
        level = c.get("authority_level")
        code = levels.setdefault(level, len(levels))
        if code >= 31:
            code = 31
        else:
            hq.auth_score[code] = H.authority_score(level)
        cd.authority = code
        chunk_d = c.get("chunk_d_tags") or {}
        cd.dtag_match = 1 if any(b in chunk_d for b in d_bodies) else 0
        cats = H.classify_jpd(body) if hq.w_jpd > 0 else {}
        jpd_tags.append(sorted({JPD_FAMILY.get(k, "O") for k in sorted(cats, key=lambda k: -cats[k])[:2]}))
    hq.auth_score[31] = H.AUTHORITY_DEFAULT
    scores, cov, keep = t.index.rerank_candidates(cands, len(chunks), hq)
    survivors = []
    for i, c in enumerate(chunks):
        c["rerank_score"] = float(scores[i])
        c["_jpd_tags"] = jpd_tags[i]
        c["_combined_coverage"] = float(cov[i]) if phrases else 0.0
        if phrases and not keep[i]:
            continue
        survivors.append(c)
    best: dict[str, float] = {}
    for c in survivors:
        cat = f"{c.get('_arm') or 'vector'}_{c.get('source_type', 'hierarchical')}"
        best[cat] = max(best.get(cat, 0.0), c["rerank_score"])
    out = [c for c in survivors
           if not (best[f"{c.get('_arm') or 'vector'}_{c.get('source_type', 'hierarchical')}"] > 0
                   and c["rerank_score"] < 0.6 * best[f"{c.get('_arm') or 'vector'}_{c.get('source_type', 'hierarchical')}"])]
    out.sort(key=lambda c: -c["rerank_score"])
    _log_stage("rerank_summary", search_id, input=len(chunks), after_floor=len(survivors), after_decay=len(out))
    return out


# ---------------------------------------------------------------------------------------------
# stage 4: enrichment (neighbour text, inherited tags, topic-block merge)
# ---------------------------------------------------------------------------------------------
def enrich_with_neighbor_text(nb: NeighborIndex, candidates: list[dict], *, paragraph_window: int = 1, page_window: int = 1,
                              max_neighbor_chars: int = 1500) -> None:
    """`_enrich_candidates_with_neighbor_text` (:2823-2918): c["_neighbor_text"] = bodies of the +-N paragraph siblings."""
    if not candidates or paragraph_window <= 0:
        return
    by_doc: dict[str, list[dict]] = {}
    for s in nb.fetch_siblings(candidates, paragraph_window=paragraph_window, page_window=page_window):
        by_doc.setdefault(str(s.get("document_id") or ""), []).append(s)
    for c in candidates:
        doc, page, para = str(c.get("document_id") or ""), c.get("page_number"), c.get("paragraph_index")
        if not doc or page is None or para is None:
            continue
        seen, bodies, total = set(), [], 0
        for s in by_doc.get(doc) or []:
            sp, si = s.get("page_number"), s.get("paragraph_index")
            if sp is None or si is None or abs(int(si) - int(para)) > paragraph_window or abs(int(sp) - int(page)) > page_window:
                continue
            text = (s.get("text") or "").strip()
            key = " ".join(text.lower().split())[:200]
            if not text or key in seen:
                continue
            seen.add(key)
            bodies.append(text)
            total += len(text)
            if total >= max_neighbor_chars:
                break
        if bodies:
            c["_neighbor_text"] = " || ".join(bodies)[:max_neighbor_chars]


def attach_inherited_doc_tags(ht: "H.HybridTable", candidates: list[dict]) -> None:
    """`_attach_inherited_doc_tags` (:2732-2807): d / j / p tags of the parent document on every candidate whose document
    has a document_tags row.  (The payor_inherited_authority view is a separate store; its j-tag injection is the caller's.)"""
    t = ht.table
    for c in candidates:
        did = str(c.get("document_id") or "")
        if did in t.doc_d_tags or did in t.doc_p_tags or did in t.doc_j_tags:
            c["_doc_d_tags"] = sorted(t.doc_d_tags.get(did, ()))
            c["_doc_j_tags"] = list(t.doc_j_tags.get(did, ()))
            c["_doc_p_tags"] = sorted(t.doc_p_tags.get(did, ()))


def merge_topic_blocks(nb: NeighborIndex, candidates: list[dict], *, sim_threshold: float = 0.7, paragraph_window: int = 3,
                       page_window: int = 0, max_per_seed: int = 5) -> int:
    """`_promote_high_sim_neighbors_to_candidates` (:2921-3076): same-page siblings of a high-similarity seed are merged
    into the seed's text in reading order (one citation, one k slot)."""
    seeds = []
    for c in candidates:
        sc = (c.get("arm_scores") or {}).get("bm25", 0.0) or float(c.get("similarity") or 0.0)
        if sc >= sim_threshold:
            seeds.append(c)
    if not seeds:
        return 0
    blocks: dict[tuple[str, int], dict[int, dict]] = {}
    for s in nb.fetch_siblings(seeds, paragraph_window=paragraph_window, page_window=page_window):
        doc, page, pi = str(s.get("document_id") or ""), s.get("page_number"), s.get("paragraph_index")
        if not doc or page is None or pi is None:
            continue
        text = (s.get("text") or "").strip()
        if len(text) < 80 and ("Provider Manual" in text or "SH_" in text):      # page headers are noise
            continue
        blocks.setdefault((doc, int(page)), {}).setdefault(int(pi), s)          # first chunk per paragraph (multi-ingest corpora)
    extended = 0
    for s in seeds:
        doc, page, para = str(s.get("document_id") or ""), s.get("page_number"), s.get("paragraph_index") or 0
        if not doc or page is None:
            continue
        sibs = blocks.get((doc, int(page))) or {}
        picked = [sibs[pi] for pi in sorted(sibs) if abs(pi - para) <= paragraph_window and (sibs[pi].get("text") or "").strip()][:max_per_seed]
        if not picked:
            continue
        before = [(p.get("text") or "").strip() for p in picked if int(p.get("paragraph_index") or 0) < para]
        after = [(p.get("text") or "").strip() for p in picked if int(p.get("paragraph_index") or 0) > para]
        s["text"] = "\n\n".join(x for x in before + [(s.get("text") or "").strip()] + after if x)
        s["_topic_block_merged"] = {"n_siblings": len(picked), "before": len(before), "after": len(after)}
        extended += 1
    return extended


# ---------------------------------------------------------------------------------------------
# stage 6: assembly and neighbour expansion
# ---------------------------------------------------------------------------------------------
def _content_key(c: dict) -> str:
    sha = (c.get("content_sha") or "").strip()
    return f"sha:{sha}" if sha else "body:" + " ".join((c.get("text") or "").lower().split())[:200]


def assemble(candidates: list[dict], k: int, strategy: str, canonical_floor: float, seen_pages: set, min_label: str):
    """`_assemble` (:2348-2520)."""
    score = lambda c: c.get("rerank_score", 0)
    if strategy == "canonical_first":
        ordered = sorted(candidates, key=lambda c: (-LABEL_ORDER.index(H.confidence_label(score(c))), authority_tier(c.get("authority_level")), -score(c)))
    elif strategy == "balanced":
        tiered = sorted(candidates, key=lambda c: -score(c))
        n_canon = math.ceil(k * max(0.0, min(1.0, canonical_floor)))
        canon = [c for c in tiered if authority_tier(c.get("authority_level")) <= 1][:n_canon * 3]
        rest = [c for c in tiered if authority_tier(c.get("authority_level")) > 1][:max(0, k - n_canon) * 3]
        ordered = canon + rest
    else:
        ordered = candidates
    floor = LABEL_ORDER.index(min_label)
    selected, extra, seen_content = [], [], set()
    for c in ordered:
        if LABEL_ORDER.index(H.confidence_label(score(c))) < floor:
            continue
        key = _content_key(c)
        if key in seen_content:
            continue
        if c.get("_promoted_from_seed") is not None or "bm25_inherited" in (c.get("retrieval_arms") or []):
            if len(extra) < 10:                               # promoted context: outside the k slots and the page rule
                seen_content.add(key)
                extra.append(c)
            continue
        page_key = (c["document_id"], c.get("page_number"))
        if len(selected) >= k or page_key in seen_pages:
            continue
        seen_pages.add(page_key)
        seen_content.add(key)
        selected.append(c)
    selected += extra
    tiers = {"contract_source_of_truth": 0, "payer_policy": 0, "fyi_not_citable": 0, "untagged": 0}
    for c in selected:
        level = (c.get("authority_level") or "").strip().lower()
        tiers["payer_policy" if level == "operational_suggested" else level if level in tiers and level != "untagged" else "untagged"] += 1
    total = len(selected)
    meta = {"strategy": strategy, "canonical_floor": canonical_floor if strategy == "balanced" else None,
            "canonical_ratio": round((tiers["contract_source_of_truth"] + tiers["payer_policy"]) / total if total else 0.0, 3),
            "strict_canonical_ratio": round(tiers["contract_source_of_truth"] / total if total else 0.0, 3),
            "tier_breakdown": tiers, "total_selected": total}
    return selected, meta


def expand_with_neighbors(nb: NeighborIndex, seeds: list[dict], *, paragraph_window: int = 2, page_window: int = 1):
    """`_expand_with_neighbors` + `_apply_neighbor_caps` (:2690-2729, 3079-3184)."""
    if not seeds or paragraph_window <= 0:
        return seeds, {"requested": False, "fetched": 0, "kept": 0}
    raw = nb.fetch_siblings(seeds, paragraph_window=paragraph_window, page_window=page_window)
    seed_ids = {str(s.get("id")) for s in seeds if s.get("id")}
    seen = {k for k in (_content_key(s) for s in seeds) if k != "body:"}
    fresh = []
    for s in raw:
        key = _content_key(s)
        if str(s.get("id")) in seed_ids or key == "body:" or key in seen:
            continue
        seen.add(key)
        fresh.append(s)
    parent: dict[str, float] = {}
    for s in seeds:
        doc = str(s.get("document_id") or "")
        if doc:
            parent[doc] = max(parent.get(doc, 0.0), float(s.get("rerank_score") or 0.0))
    for s in fresh:
        p = parent.get(str(s.get("document_id") or ""), 0.0)
        s["rerank_score"] = 0.5 * p if p > 0.0 else NEIGHBOR_SCORE_FLOOR
        s["confidence_label"] = "low"
    per_doc: dict[str, int] = {}
    combined = []
    for c in seeds:
        if len(combined) >= NEIGHBOR_TOTAL_CAP:
            break
        key = str(c.get("document_id") or c.get("document_name") or "_unknown")
        per_doc[key] = per_doc.get(key, 0) + 1
        combined.append(c)
    for c in sorted(fresh, key=lambda c: -float(c.get("rerank_score") or 0.0)):
        if len(combined) >= NEIGHBOR_TOTAL_CAP:
            break
        key = str(c.get("document_id") or c.get("document_name") or "_unknown")
        if per_doc.get(key, 0) >= NEIGHBOR_PER_DOC_CAP:
            continue
        per_doc[key] = per_doc.get(key, 0) + 1
        combined.append(c)
    return combined, {"requested": True, "fetched": len(raw), "kept": len(combined) - len(seeds), "para_window": paragraph_window,
                      "page_window": page_window}


# ---------------------------------------------------------------------------------------------
# the driver
# ---------------------------------------------------------------------------------------------
EMPTY_EXPANSION = {"matched_codes": [], "expansion_phrases": [], "expansion_phrases_count": 0, "final_tsquery": "", "log": [],
                   "domain_tags": [], "jurisdiction_tags": [], "process_tags": []}


def corpus_search(ht: "H.HybridTable", request: CorpusSearchRequest, *, embed: Callable[[str], Sequence[float] | None],
                  bm25_arm: Callable[..., tuple[list[dict], str | None, dict]] | None = None,
                  expand: Callable[[str], Any] | None = None, neighbors: NeighborIndex | None = None,
                  caller: str = "api", caller_id: str | None = None) -> CorpusSearchResponse:
    """Run bm25 / vector / hybrid search and return ranked, labelled chunks -- `corpus_search` of the reference, with the
    table (`ht`) in the place of the database session.

    embed(query) -> embedding or None; bm25_arm(query, k, filters, include_document_ids, search_id=, tag_mode=) ->
    (chunks, normalized_query, expansion dict) -- `_bm25_arm`'s return shape (:806); expand(query) -> LexiconExpansion or
    None -- `expand_query_via_lexicon` (recall mode)."""
    search_id = uuid.uuid4().hex[:12]
    if not (request.query or "").strip():
        return CorpusSearchResponse(chunks=[], telemetry={"mode": request.mode, "k": request.k, "error": "empty query", "search_id": search_id})
    t = ht.table
    nb = neighbors or NeighborIndex(t)
    mode = request.mode or "corpus"
    k = max(1, min(100, request.k))
    t0 = time.monotonic()
    filters = _ArmFilters(**request.filters.model_dump()) if request.filters else None
    pool_ids = request.include_document_ids
    bm25_chunks: list[dict] = []
    vec_chunks: list[dict] = []
    dtag_chunks: list[dict] = []
    embed_ms = bm25_ms = vec_ms = 0.0
    normalized: str | None = None
    expansion: dict[str, Any] = dict(EMPTY_EXPANSION)

    def run_bm25():
        nonlocal bm25_chunks, normalized, expansion, bm25_ms
        if bm25_arm is None:
            return
        tb = time.monotonic()
        bm25_chunks, normalized, expansion = bm25_arm(request.query, k * 2, filters, pool_ids, search_id=search_id, tag_mode=request.tag_mode)
        bm25_ms = (time.monotonic() - tb) * 1000

    def timed_embed():
        nonlocal embed_ms
        te = time.monotonic()
        try:
            e = embed(request.query)
        except Exception:                                      # embedding failure: the vector arm is simply skipped (:478-484)
            e = None
        embed_ms = (time.monotonic() - te) * 1000
        return list(e) if e is not None and len(e) else None

    def dtag_keys():
        req = [c[2:] for c in (request.required_phrase_tag_codes or []) if c and c.startswith("d:")]
        return req or [c[2:] for c in (expansion.get("domain_tags") or []) if c and c.startswith("d:")]

    if mode == "corpus":
        run_bm25()
        q = timed_embed()
        if q:
            exp = LexiconExpansion(matched_codes=expansion.get("matched_codes") or [], expansion_phrases=expansion.get("expansion_phrases") or [],
                                   domain_tags=expansion.get("domain_tags") or [], jurisdiction_tags=expansion.get("jurisdiction_tags") or [],
                                   process_tags=expansion.get("process_tags") or [], log=[])
            tv = time.monotonic()
            vec_chunks = vector_arm(t, q, k * 2, filters, pool_ids, search_id=search_id, expansion=exp, tag_mode=request.tag_mode)
            dtag_chunks = H.dtag_arm(ht, dtag_keys(), k, filters, pool_ids, search_id=search_id, idf_mode=DTAG_ARM_IDF)
            vec_ms = (time.monotonic() - tv) * 1000
    elif mode == "precision":
        run_bm25()
        keys = dtag_keys()
        n_pool = len(pool_ids) if pool_ids else 0
        if keys and (n_pool == 0 or n_pool <= DTAG_ARM_MAX_POOL_DOCS):
            dtag_chunks = H.dtag_arm(ht, keys, k, filters, pool_ids, search_id=search_id, idf_mode=DTAG_ARM_IDF)
    else:
        skip = False
        if pool_ids:
            n_chunks = int(sum(int(t.doc_rows[d]) for d in (t.doc_idx.get(str(x)) for x in pool_ids) if d is not None))
            skip = n_chunks > TINY_POOL_CHUNK_MAX
            if skip:
                _log_stage("recall_arm_skipped", search_id, reason="chunk_heavy_pool_hnsw_disabled", chunk_count=n_chunks)
        if not skip:
            q = timed_embed()
            if q:
                try:
                    exp = expand(request.query) if expand else None
                except Exception:
                    exp = None
                tv = time.monotonic()
                vec_chunks = vector_arm(t, q, k * 2, filters, pool_ids, search_id=search_id, expansion=exp, tag_mode=request.tag_mode,
                                        min_similarity=request.min_similarity, over_fetch_factor=8)
                vec_ms = (time.monotonic() - tv) * 1000

    # ---- fuse
    tr = time.monotonic()
    if mode == "corpus":
        arms = {"bm25": bm25_chunks, "vector": vec_chunks}
        if dtag_chunks:
            arms["dtag"] = dtag_chunks
        candidates = H.rrf_merge(arms, search_id=search_id)
    elif mode == "precision" and dtag_chunks:
        candidates = H.rrf_merge({"bm25": bm25_chunks, "dtag": dtag_chunks}, search_id=search_id)
    else:
        candidates = bm25_chunks if mode == "precision" else vec_chunks
        for c in candidates:
            c.setdefault("retrieval_arms", ["bm25" if mode == "precision" else "vector"])

    # ---- content de-duplication (boilerplate pages share text AND embedding, and tie at one similarity)
    seen_bodies, unique = set(), []
    for c in candidates:
        key = " ".join((c.get("text") or "").lower().split())[:400]
        if key and key in seen_bodies:
            continue
        if key:
            seen_bodies.add(key)
        unique.append(c)
    candidates = unique

    # ---- enrichment before the rerank
    if request.required_phrases:
        enrich_with_neighbor_text(nb, candidates, paragraph_window=1, page_window=1)
        attach_inherited_doc_tags(ht, candidates)
        merge_topic_blocks(nb, candidates, sim_threshold=0.7, paragraph_window=3, page_window=0, max_per_seed=5)

    reranked = rerank_candidates(ht, candidates, search_id=search_id, query=request.query, required_phrases=request.required_phrases,
                                 required_phrase_weights=request.required_phrase_weights,
                                 required_phrase_tag_codes=request.required_phrase_tag_codes)
    rerank_ms = (time.monotonic() - tr) * 1000

    # ---- assemble
    min_label = MODE_MIN.get(mode, "low")
    if request.min_similarity is not None:
        min_label = H.confidence_label(request.min_similarity)
    assembled, assembly_meta = assemble(reranked[:k * 3], k, request.assembly_strategy or "score",
                                        max(0.0, min(1.0, request.canonical_floor)), set(), min_label)
    if request.neighbor_paragraph_window and request.neighbor_paragraph_window > 0 and assembled:
        assembled, nmeta = expand_with_neighbors(nb, assembled, paragraph_window=request.neighbor_paragraph_window,
                                                 page_window=max(0, int(request.neighbor_page_window)))
        assembly_meta["neighbor_expansion"] = nmeta
    else:
        assembly_meta["neighbor_expansion"] = {"requested": False}

    # ---- output
    chunks_out, trace = [], []
    for rank, c in enumerate(assembled, 1):
        score = float(c.get("rerank_score") or 0.0)
        label = H.confidence_label(score)
        chunk = CorpusChunk(
            id=c["id"], text=c["text"], document_id=c["document_id"], document_name=c["document_name"], page_number=c.get("page_number"),
            paragraph_index=c.get("paragraph_index"), source_type=c.get("source_type") or "hierarchical",
            similarity=round(float(c.get("similarity") or 0.0), 4), rerank_score=round(score, 4), confidence_label=label,
            retrieval_arms=c.get("retrieval_arms") or ["unknown"], authority_level=c.get("authority_level"), payer=c.get("payer"),
            state=c.get("state"), jpd_tags=c.get("_jpd_tags") or [], section_path=c.get("section_path"), chapter_path=c.get("chapter_path"),
            summary=c.get("summary"), is_neighbor=bool(c.get("is_neighbor")))
        chunks_out.append(chunk)
        entry = {"rank": rank, "chunk_id": chunk.id, "document_name": chunk.document_name, "document_id": chunk.document_id,
                 "page_number": chunk.page_number, "paragraph_index": chunk.paragraph_index, "retrieval_arms": chunk.retrieval_arms,
                 "authority_level": chunk.authority_level, "authority_tier": authority_tier(chunk.authority_level),
                 "confidence_label": label, "text_preview": preview(chunk.text)}
        if "arm_scores" in c:
            entry["arm_scores"] = {a: round(s, 4) for a, s in c["arm_scores"].items()}
            entry["arm_ranks"] = c.get("arm_ranks")
            entry["rrf_score"] = round(float(c.get("rrf_score") or c.get("similarity") or 0.0), 4)
        else:
            entry["arm_scores"] = {(c.get("retrieval_arms") or ["unknown"])[0]: round(float(c.get("similarity") or 0.0), 4)}
        trace.append(entry)
    total_ms = round((time.monotonic() - t0) * 1000, 1)

    def arm_summary(chunks, key):
        return [{"chunk_id": c["id"], "document_name": c["document_name"][:60], "document_id": c["document_id"], "page_number": c["page_number"],
                 "authority_level": c.get("authority_level"), "authority_tier": authority_tier(c.get("authority_level")), "payer": c.get("payer"),
                 key: round(float(c.get("similarity") or 0.0), 4), "text_preview": preview(c.get("text") or "")} for c in chunks]

    telemetry = {
        "search_id": search_id, "query": request.query, "bm25_normalized_query": normalized, "mode": mode, "k": k,
        "embed_ms": round(embed_ms, 1), "bm25_ms": round(bm25_ms, 1), "vec_ms": round(vec_ms, 1), "rerank_ms": round(rerank_ms, 1),
        "total_ms": total_ms, "arm_hits": {"bm25": len(bm25_chunks), "vector": len(vec_chunks)},
        "arm_results": {"bm25": arm_summary(bm25_chunks, "ts_rank"), "vector": arm_summary(vec_chunks, "cosine")},
        "candidates": len(reranked), "returned": len(chunks_out), "min_label_applied": min_label,
        "reranker": "score+authority+length (phase1)", "assembly": assembly_meta, "scoring_trace": trace, "bm25_expansion": expansion,
        "filters": request.filters.model_dump() if request.filters else None, "caller": caller, "caller_id": caller_id,
    }
    return CorpusSearchResponse(chunks=chunks_out, telemetry=telemetry)


async def acorpus_search(ht, request: CorpusSearchRequest, **kw) -> CorpusSearchResponse:
    """awaitable form for the FastAPI routers (`await corpus_search(db, request)`, app/main.py)"""
    import asyncio
    return await asyncio.to_thread(corpus_search, ht, request, **kw)

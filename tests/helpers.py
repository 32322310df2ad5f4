"""Shared builders for the parity tests: the SAME synthetic rows go into the oracle's string
table (oracle.Table) and into the product's PublishedTable / Index."""
from __future__ import annotations

import uuid

import numpy as np

import mrag_b200
from mrag_b200 import synth
from mrag_b200 import _native as N


def string_columns(meta: np.ndarray):
    """Decode synth's codes back to the strings of rag_published_embeddings (None = SQL NULL)."""
    payer = [None if c == N.MRAG_CODE_NONE else synth.PAYERS[c] for c in meta["payer"]]
    state = [None if c == 0xFF else synth.STATES[c] for c in meta["state"]]
    program = [None if c == 0xFF else synth.PROGRAMS[c] for c in meta["program"]]
    auth = [None if c == 0xFF else synth.AUTHORITIES[c] for c in meta["authority"]]
    src = [None if c == 0xFF else synth.SOURCE_TYPES[c] for c in meta["source_type"]]
    return payer, state, program, auth, src


def tag_key(bit: int) -> tuple[str, str]:
    """synthetic document tag key for bit b: even bits are d_tags keys, odd bits p_tags keys."""
    return ("d" if bit % 2 == 0 else "p", f"topic_{bit:03d}.leaf")


def build_tables(oracle, n: int, dim: int, seed: int = 7, dtype: str = "f32", null_frac: float = 2e-3,
                 rows_per_doc: int = 16, device: int = 0, with_product: bool = True, n_tag_bits: int = 24, devices=None):
    """Returns (oracle.Table, PublishedTable or None, X, valid, meta, info)."""
    X, valid = synth.make_corpus(n, dim, seed=seed, null_frac=null_frac)
    meta, doc_tags, info = synth.make_metadata(n, seed=seed + 1, rows_per_doc=rows_per_doc, valid=valid)
    payer, state, program, auth, src = string_columns(meta)
    rng = np.random.default_rng(seed + 2)
    doc_uuid = [str(uuid.UUID(int=int(rng.integers(0, 2**63)) << 64 | d)) for d in range(info["n_docs"])]
    ids = [str(uuid.UUID(int=(int(rng.integers(0, 2**63)) << 64) | i)) for i in range(n)]
    document_id = [doc_uuid[d] for d in info["doc_of_row"]]
    extra = {
        "text": [f"chunk {i} text" for i in range(n)],
        "page_number": [int(i % 40) + 1 for i in range(n)],
        "paragraph_index": [int(i % 7) for i in range(n)],
        "section_path": [("" if i % 5 == 0 else f"sec/{i % 11}") for i in range(n)],
        "chapter_path": [None if i % 3 == 0 else f"ch/{i % 4}" for i in range(n)],
        "summary": [None] * n,
        "content_sha": [f"{i:040x}" for i in range(n)],
        "document_display_name": [("" if d % 4 == 0 else f"Doc {d}") for d in info["doc_of_row"]],
        "document_filename": [f"doc_{d}.pdf" for d in info["doc_of_row"]],
        "chunk_d_tags": [({"benefits.dme": 1} if i % 20 == 0 else None) for i in range(n)],
        "chunk_p_tags": [None] * n,
        "chunk_j_tags": [None] * n,
    }
    Xs = oracle.round_bf16(X) if dtype == "bf16" else X
    # document_tags rows: only the first n_tag_bits bits are used as keys
    d_tags, p_tags = {}, {}
    for d in range(info["n_docs"]):
        bits = np.nonzero(info["tagmat"][d, :n_tag_bits])[0]
        has_row = info["tagmat"][d].any() or (d % 10 != 0)
        if not has_row:
            continue
        d_tags[doc_uuid[d]] = {tag_key(b)[1] for b in bits if tag_key(b)[0] == "d"}
        p_tags[doc_uuid[d]] = {tag_key(b)[1] for b in bits if tag_key(b)[0] == "p"}
    ot = oracle.Table(id=ids, document_id=document_id, source_type=src, source_id=[f"src-{i}" for i in range(n)],
                      document_payer=payer, document_state=state, document_program=program,
                      document_authority_level=auth, has_vec=valid.astype(bool), X=Xs,
                      doc_d_tags=d_tags, doc_p_tags=p_tags, extra=extra)
    pt = None
    if with_product:
        pt = mrag_b200.PublishedTable(dim, dtype=dtype, device=device, capacity=n + 64, devices=devices)
        rows = []
        for i in range(n):
            r = {"id": ids[i], "document_id": document_id[i], "source_type": src[i], "source_id": f"src-{i}",
                 "document_payer": payer[i], "document_state": state[i], "document_program": program[i],
                 "document_authority_level": auth[i]}
            for c, col in extra.items():
                r[c] = col[i]
            rows.append(r)
        embs = [None if not valid[i] else X[i].tolist() for i in range(n)]
        step = 4096
        for lo in range(0, n, step):
            pt.insert(rows[lo:lo + step], embs[lo:lo + step])
        for did in d_tags:
            pt.set_document_tags(did, sorted(d_tags[did]), sorted(p_tags[did]))
    return ot, pt, X, valid, meta, info


GOLDEN_DIR = __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "golden")


def load_golden_json(name: str):
    import gzip
    import json
    import os
    with gzip.open(os.path.join(GOLDEN_DIR, name + ".gz"), "rb") as f:
        return json.loads(f.read().decode())


def load_golden_table(oracle, with_product: bool = False, dtype: str = "f32", device: int = 0, devices=None):
    """The table the golden fixtures were generated on (tests/golden/make_golden.py), rebuilt from the
    committed files -- NOT from synth, so generator changes cannot silently move the fixtures."""
    import json
    import os
    tj = load_golden_json("table.json")
    vz = np.load(os.path.join(GOLDEN_DIR, "table_vectors.npz"))
    X, has_vec = vz["X"], vz["has_vec"].astype(bool)
    c = tj["columns"]
    ot = oracle.Table(id=c["id"], document_id=c["document_id"], source_type=c["source_type"], source_id=c["source_id"],
                      document_payer=c["document_payer"], document_state=c["document_state"],
                      document_program=c["document_program"], document_authority_level=c["document_authority_level"],
                      has_vec=has_vec, X=X, doc_d_tags={k: set(v) for k, v in tj["doc_d_tags"].items()},
                      doc_p_tags={k: set(v) for k, v in tj["doc_p_tags"].items()}, extra=tj["extra"])
    pt = None
    if with_product:
        n, dim = tj["n"], tj["dim"]
        pt = mrag_b200.PublishedTable(dim, dtype=dtype, device=device, capacity=n + 64, devices=devices)
        rows = []
        for i in range(n):
            r = {name: col[i] for name, col in c.items()}
            for name, col in tj["extra"].items():
                r[name] = col[i]
            rows.append(r)
        embs = [X[i].tolist() if has_vec[i] else None for i in range(n)]
        pt.insert(rows, embs)
        for did in tj["doc_d_tags"]:
            pt.set_document_tags(did, tj["doc_d_tags"][did], tj["doc_p_tags"].get(did, []))
    return ot, pt, X, has_vec

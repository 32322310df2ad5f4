"""Index: thin object wrapper over the C ABI (include/mrag.h) -- one row shard resident in HBM.

Host buffers are numpy arrays; device buffers are torch CUDA tensors (torch is plumbing here:
device memory and streams).  All arithmetic happens in libmrag.so's CUDA kernels.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Sequence

import numpy as np

from . import _native as N

FEAT_DTYPE = np.dtype([
    ("phrase_bits", "<u8", (N.MRAG_PHRASE_WORDS,)), ("jpd_hits", "u1", (N.MRAG_JPD_CATS,)), ("flags", "u1"),
    ("length_score", "<f4"), ("dtags", "<u2", (4,)),
])
assert FEAT_DTYPE.itemsize == C.sizeof(N.ChunkFeat) == 40

META_DTYPE = np.dtype([
    ("doc_idx", "<u4"), ("payer", "<u2"), ("state", "u1"), ("program", "u1"),
    ("authority", "u1"), ("source_type", "u1"), ("valid", "u1"), ("reserved", "u1"),
])
assert META_DTYPE.itemsize == C.sizeof(N.RowMeta) == 12

DTYPES = {"f32": N.MRAG_F32, "fp32": N.MRAG_F32, "float32": N.MRAG_F32, "bf16": N.MRAG_BF16, "bfloat16": N.MRAG_BF16}


def _bitset(codes: Iterable[int], words: int) -> list[int]:
    out = [0] * words
    for c in codes:
        c = int(c)
        if 0 <= c < words * 64:
            out[c >> 6] |= 1 << (c & 63)
    return out


class Filter:
    """Builder for mrag_filter: the WHERE clauses of the statement on dictionary codes.

    Each method corresponds to one clause of the reference SQL (see include/mrag.h)."""

    def __init__(self):
        self.s = N.FilterStruct()
        self._pool = None   # keeps the numpy array behind doc_pool alive

    # corpus_search.py:524-535
    def payer_in(self, codes: Iterable[int], alt_codes: Iterable[int] = (), alt_state: int | None = None) -> "Filter":
        self.s.flags |= N.F_PAYER
        for i, w in enumerate(_bitset(codes, N.MRAG_PAYER_WORDS)):
            self.s.payer_any[i] = w
        if alt_state is not None:
            for i, w in enumerate(_bitset(alt_codes, N.MRAG_PAYER_WORDS)):
                self.s.payer_alt_any[i] = w
            self.s.alt_state = int(alt_state)
        return self

    def state_eq(self, code: int) -> "Filter":
        self.s.flags |= N.F_STATE
        self.s.state_eq = int(code)
        return self

    def program_eq(self, code: int) -> "Filter":
        self.s.flags |= N.F_PROGRAM
        self.s.program_eq = int(code)
        return self

    def authority_eq(self, code: int) -> "Filter":
        self.s.flags |= N.F_AUTHORITY
        self.s.authority_eq = int(code)
        return self

    def source_type_eq(self, code: int) -> "Filter":
        self.s.flags |= N.F_SOURCE_TYPE
        self.s.source_type_eq = int(code)
        return self

    # vector_store.py:247-249
    def doc_eq(self, doc_idx: int) -> "Filter":
        self.s.flags |= N.F_DOC_EQ
        self.s.doc_eq = int(doc_idx)
        return self

    # corpus_search.py:546-558
    def doc_pool(self, doc_idxs: Sequence[int]) -> "Filter":
        self.s.flags |= N.F_DOC_POOL
        self._pool = np.ascontiguousarray(np.asarray(list(doc_idxs) if not isinstance(doc_idxs, np.ndarray) else doc_idxs,
                                                     dtype=np.uint32))
        self.s.doc_pool = self._pool.ctypes.data if self._pool.size else None
        self.s.n_doc_pool = int(self._pool.size)
        return self

    # corpus_search_agent.py:1762-1888: the pool is a document bitmap resident on the device (mrag_pool_build)
    def pool_handle(self, handle) -> "Filter":
        self.s.flags |= N.F_DOC_POOL_HANDLE
        self.s.pool = handle
        return self

    # corpus_search.py:1478-1496
    def tag_strict(self, state_codes: Iterable[int] = (), program_codes: Iterable[int] = (),
                   payer_codes: Iterable[int] = ()) -> "Filter":
        self.s.flags |= N.F_TAG_STRICT
        for i, w in enumerate(_bitset(state_codes, N.MRAG_SMALL_WORDS)):
            self.s.tag_state_any[i] = w
        for i, w in enumerate(_bitset(program_codes, N.MRAG_SMALL_WORDS)):
            self.s.tag_program_any[i] = w
        for i, w in enumerate(_bitset(payer_codes, N.MRAG_PAYER_WORDS)):
            self.s.tag_payer_any[i] = w
        return self

    # corpus_search.py:1497-1510
    def tag_relaxed(self, tag_bits: Iterable[int]) -> "Filter":
        self.s.flags |= N.F_TAG_RELAXED
        for i, w in enumerate(_bitset(tag_bits, N.MRAG_TAG_WORDS)):
            self.s.tag_any[i] = w
        return self

    @property
    def active(self) -> bool:
        return self.s.flags != 0

    def ref(self):
        return C.byref(self.s)


def make_meta(n: int, doc_idx=None, payer=None, state=None, program=None, authority=None,
              source_type=None, valid=None) -> np.ndarray:
    """Structured array of mrag_rowmeta; missing columns get the 'empty' codes."""
    m = np.zeros(n, dtype=META_DTYPE)
    m["doc_idx"] = np.arange(n, dtype=np.uint32) if doc_idx is None else doc_idx
    m["payer"] = N.MRAG_CODE_NONE if payer is None else payer
    m["state"] = 0xFF if state is None else state
    m["program"] = 0xFF if program is None else program
    m["authority"] = 0xFF if authority is None else authority
    m["source_type"] = 0xFF if source_type is None else source_type
    m["valid"] = 1 if valid is None else valid
    return m


class Index:
    """One row shard of the corpus on one GPU."""

    def __init__(self, dim: int, dtype: str | int = "f32", device: int = 0, capacity: int = 1 << 20):
        self._lib = N.load()
        self._h = C.c_void_p()
        dt = DTYPES[dtype] if isinstance(dtype, str) else int(dtype)
        N.check(self._lib.mrag_create(C.byref(self._h), int(dim), dt, int(device), int(capacity)))
        self.dim, self.dtype, self.device, self.capacity = int(dim), dt, int(device), int(capacity)

    # -- snapshot --------------------------------------------------------------------------
    def save(self, path: str, version: int = 0) -> None:
        """Write this shard to one file (mrag_save); ``version`` = corpus_state.corpus_version."""
        N.check(self._lib.mrag_save(self._h, str(path).encode(), int(version)))

    @classmethod
    def load(cls, path: str, device: int = 0, capacity: int = 0) -> tuple["Index", int]:
        """(Index, version) from a snapshot file (mrag_load)."""
        self = cls.__new__(cls)
        self._lib = N.load()
        self._h = C.c_void_p()
        ver = C.c_int64(0)
        N.check(self._lib.mrag_load(C.byref(self._h), str(path).encode(), int(device), int(capacity), C.byref(ver)))
        self.dim = int(self._lib.mrag_dim(self._h))
        self.dtype = int(self._lib.mrag_index_dtype(self._h))
        self.device = int(device)
        self.capacity = int(self._lib.mrag_capacity(self._h))
        return self, int(ver.value)

    # -- lifecycle -------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.mrag_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self) -> int:
        return int(self._lib.mrag_size(self._h))

    # -- write side ------------------------------------------------------------------------
    def append(self, X: np.ndarray, meta: np.ndarray | None = None) -> int:
        """Append host rows (float32 [n, dim]).  Returns the first new row index."""
        X = np.ascontiguousarray(X, dtype=np.float32)
        if X.ndim != 2 or X.shape[1] != self.dim:
            raise ValueError(f"expected [n, {self.dim}] rows, got {X.shape}")
        if not np.isfinite(X).all():
            raise ValueError("NaN/Inf not allowed in a vector")       # pgvector rejects them on input
        first = C.c_int64(-1)
        N.check(self._lib.mrag_append(self._h, X.ctypes.data, X.shape[0], self._meta_ptr(meta, X.shape[0]), C.byref(first)))
        return int(first.value)

    def append_device(self, X, meta: np.ndarray | None = None) -> int:
        """Append rows already on this index's GPU (torch float32 CUDA tensor [n, dim])."""
        import torch
        if not (X.is_cuda and X.dtype == torch.float32 and X.dim() == 2 and X.shape[1] == self.dim and X.is_contiguous()):
            raise ValueError("append_device needs a contiguous float32 CUDA tensor [n, dim]")
        if X.device.index != self.device:
            raise ValueError("tensor is on another device")
        first = C.c_int64(-1)
        stream = torch.cuda.current_stream(X.device).cuda_stream
        N.check(self._lib.mrag_append_device(self._h, X.data_ptr(), X.shape[0], self._meta_ptr(meta, X.shape[0]),
                                             C.byref(first), stream))
        return int(first.value)

    def _meta_ptr(self, meta, n):
        if meta is None:
            return None
        meta = np.ascontiguousarray(meta)
        if meta.dtype != META_DTYPE or meta.shape != (n,):
            raise ValueError("meta must be a META_DTYPE array of length n")
        self._keep = meta
        return meta.ctypes.data

    def set_doc_tags(self, first_doc: int, bits: np.ndarray) -> None:
        bits = np.ascontiguousarray(bits, dtype=np.uint64)
        if bits.ndim != 2 or bits.shape[1] != N.MRAG_TAG_WORDS:
            raise ValueError(f"bits must be [n_docs, {N.MRAG_TAG_WORDS}] uint64")
        N.check(self._lib.mrag_set_doc_tags(self._h, int(first_doc), bits.ctypes.data, bits.shape[0]))

    def set_chunk_features(self, first_row: int, feat: np.ndarray) -> None:
        """Per-row text features of the hybrid rerank (FEAT_DTYPE array)."""
        feat = np.ascontiguousarray(feat)
        if feat.dtype != FEAT_DTYPE or feat.ndim != 1:
            raise ValueError("feat must be a 1-D FEAT_DTYPE array")
        N.check(self._lib.mrag_set_chunk_features(self._h, int(first_row), feat.ctypes.data, feat.shape[0]))

    def set_dtag_overflow(self, rows: np.ndarray, codes: np.ndarray) -> None:
        """chunk d-tag keys beyond the four inline slots: (row, code) pairs sorted by row; replaces the whole table"""
        rows = np.ascontiguousarray(rows, dtype=np.uint32)
        codes = np.ascontiguousarray(codes, dtype=np.uint16)
        N.check(self._lib.mrag_set_dtag_overflow(self._h, rows.ctypes.data if rows.size else None,
                                                 codes.ctypes.data if codes.size else None, int(rows.size)))

    def dtag_rows(self, flt: "Filter | None", codes: Sequence[int]):
        """The d-tag arm's WHERE on the GPU (mrag_dtag_mask): (rows whose chunk_d_tags hold any of `codes` among the LIVE rows
        that pass flt, ascending; n_total = live rows passing flt; per-code counts).  Any number of codes (32 per pass)."""
        n = len(self)
        words = (n + 31) // 32
        mask = np.zeros(words + 1, dtype=np.uint32)
        n_total, per_code = 0, []
        codes = [int(c) for c in codes]
        for lo in range(0, max(len(codes), 1), 32):
            part = codes[lo:lo + 32]
            arr = (C.c_uint16 * max(1, len(part)))(*part)
            m = np.zeros(words + 1, dtype=np.uint32)
            counts = (C.c_int64 * (len(part) + 1))()
            N.check(self._lib.mrag_dtag_mask(self._h, flt.ref() if (flt is not None and flt.active) else None, arr, len(part),
                                             m.ctypes.data, counts))
            mask |= m
            n_total = int(counts[0])
            per_code += [int(counts[1 + i]) for i in range(len(part))]
        wz = np.flatnonzero(mask[:words])                      # expand the non-zero words only (no N-sized temporary)
        bits = (mask[wz, None] >> np.arange(32, dtype=np.uint32)[None, :]) & 1
        rows = (wz[:, None] * 32 + np.arange(32)[None, :])[bits.astype(bool)]
        return rows[rows < n], n_total, per_code

    def rerank_candidates(self, cands, n: int, hq):
        """mrag_rerank_candidates: (scores f32 [n], coverage f32 [n], keep u8 [n]) for a ctypes array of N.Candidate"""
        scores, cov, keep = np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros(n, np.uint8)
        N.check(self._lib.mrag_rerank_candidates(self._h, cands, int(n), C.byref(hq), scores.ctypes.data, cov.ctypes.data, keep.ctypes.data))
        return scores, cov, keep

    def set_doc_jtags(self, first_doc: int, bits: np.ndarray) -> None:
        bits = np.ascontiguousarray(bits, dtype=np.uint64)
        if bits.ndim != 2 or bits.shape[1] != N.MRAG_JTAG_WORDS:
            raise ValueError(f"bits must be [n_docs, {N.MRAG_JTAG_WORDS}] uint64")
        N.check(self._lib.mrag_set_doc_jtags(self._h, int(first_doc), bits.ctypes.data, bits.shape[0]))

    def search_hybrid(self, Q: np.ndarray, k: int, hq, flt: "Filter | None" = None):
        """Fused hybrid rerank.  hq: ctypes array of N.HybridQuery (one per query).
        Returns (rerank scores f32 [nq,k], clamp01(cos) f32 [nq,k], rows i64 [nq,k], counts i32 [nq])."""
        Q = np.ascontiguousarray(np.atleast_2d(np.asarray(Q, dtype=np.float32)))
        if Q.shape[1] != self.dim:
            raise ValueError(f"query dim {Q.shape[1]} != index dim {self.dim}")
        nq = Q.shape[0]
        if len(hq) != nq:
            raise ValueError("one HybridQuery per query")
        scores = np.empty((nq, k), dtype=np.float32)
        cos = np.empty((nq, k), dtype=np.float32)
        rows = np.empty((nq, k), dtype=np.int64)
        counts = np.zeros(nq, dtype=np.int32)
        N.check(self._lib.mrag_search_hybrid(self._h, Q.ctypes.data, nq, int(k), flt.ref() if flt is not None else None,
                                             C.addressof(hq), scores.ctypes.data, cos.ctypes.data, rows.ctypes.data,
                                             counts.ctypes.data, None))
        return scores, cos, rows, counts

    def tombstone_doc(self, doc_idx: int) -> int:
        n = C.c_int64(0)
        N.check(self._lib.mrag_tombstone_doc(self._h, int(doc_idx), C.byref(n)))
        return int(n.value)

    def live_rows(self) -> tuple[int, int]:
        """(rows that still exist, rows that still have a vector) of the len(self) slots in use."""
        a, b = C.c_int64(0), C.c_int64(0)
        N.check(self._lib.mrag_live_rows(self._h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def set_row_base(self, base: int) -> None:
        N.check(self._lib.mrag_set_row_base(self._h, int(base)))

    # -- read side -------------------------------------------------------------------------
    def search(self, Q: np.ndarray, k: int, flt: Filter | None = None, options: int = 0):
        """Host-buffer search.  Returns (scores f32 [nq,k], rows i64 [nq,k], counts i32 [nq])."""
        Q = np.ascontiguousarray(np.atleast_2d(np.asarray(Q, dtype=np.float32)))
        if Q.shape[1] != self.dim:
            raise ValueError(f"query dim {Q.shape[1]} != index dim {self.dim}")
        nq = Q.shape[0]
        scores = np.empty((nq, k), dtype=np.float32)
        rows = np.empty((nq, k), dtype=np.int64)
        counts = np.zeros(nq, dtype=np.int32)
        N.check(self._lib.mrag_search(self._h, Q.ctypes.data, nq, int(k), flt.ref() if flt is not None else None,
                                      scores.ctypes.data, rows.ctypes.data, counts.ctypes.data, int(options), None))
        return scores, rows, counts

    def search_pinned(self, q_host, k: int, out_scores, out_rows, out_counts, flt: Filter | None = None,
                      options: int = 0) -> None:
        """Host-buffer search on caller-provided (ideally pinned) torch CPU tensors -- the call the
        end-to-end benchmark times: H2D of the queries and D2H of the results are inside."""
        nq = q_host.shape[0]
        N.check(self._lib.mrag_search(self._h, q_host.data_ptr(), nq, int(k), flt.ref() if flt is not None else None,
                                      out_scores.data_ptr(), out_rows.data_ptr(), out_counts.data_ptr(),
                                      int(options), None))

    def search_device(self, q, k: int, flt: Filter | None = None, out=None, sync: bool = True, options: int = 0):
        """Device-buffer search on torch's current stream.  q: float32 CUDA [nq, dim].
        out = (scores f32 [nq,k], rows i64 [nq,k], counts i32 [nq]) CUDA tensors, allocated if None."""
        import torch
        if not (q.is_cuda and q.dtype == torch.float32 and q.is_contiguous() and q.dim() == 2 and q.shape[1] == self.dim):
            raise ValueError("search_device needs a contiguous float32 CUDA tensor [nq, dim]")
        nq = q.shape[0]
        if out is None:
            out = (torch.empty((nq, k), dtype=torch.float32, device=q.device),
                   torch.empty((nq, k), dtype=torch.int64, device=q.device),
                   torch.empty((nq,), dtype=torch.int32, device=q.device))
        opts = int(options) | N.OPT_DEVICE_IO | (0 if sync else N.OPT_NO_SYNC)
        stream = torch.cuda.current_stream(q.device).cuda_stream
        N.check(self._lib.mrag_search(self._h, q.data_ptr(), nq, int(k), flt.ref() if flt is not None else None,
                                      out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), opts, stream))
        return out

    def filter_mask(self, flt: Filter | None):
        """K2 alone: (row bitmap as a torch int32 CUDA tensor of ceil(n/32) words, rows passing)."""
        import torch
        words = (len(self) + 31) // 32
        mask = torch.zeros(max(words, 1), dtype=torch.int32, device=f"cuda:{self.device}")
        n_pass = C.c_int64(0)
        stream = torch.cuda.current_stream(mask.device).cuda_stream
        N.check(self._lib.mrag_filter_mask(self._h, flt.ref() if flt is not None else None, mask.data_ptr(),
                                           C.byref(n_pass), stream))
        return mask[:words], int(n_pass.value)

    # -- introspection ---------------------------------------------------------------------
    def last_kernel_ms(self, what: int) -> float:
        return float(self._lib.mrag_last_kernel_ms(int(what)))

    def last_scan_kind(self) -> str:
        return self._lib.mrag_last_scan_kind().decode()


def merge_topk(device: int, scores_in, rows_in, counts_in, n_lists: int, nq: int, k: int,
               strides: tuple[int, int, int], out=None):
    """K4 on torch CUDA tensors (see mrag_merge_topk)."""
    import torch
    lib = N.load()
    dev = torch.device(f"cuda:{device}")
    if out is None:
        out = (torch.empty((nq, k), dtype=torch.float32, device=dev),
               torch.empty((nq, k), dtype=torch.int64, device=dev),
               torch.empty((nq,), dtype=torch.int32, device=dev))
    stream = torch.cuda.current_stream(dev).cuda_stream
    N.check(lib.mrag_merge_topk(int(device), int(n_lists), int(nq), int(k), scores_in.data_ptr(), rows_in.data_ptr(),
                                counts_in.data_ptr(), int(strides[0]), int(strides[1]), int(strides[2]),
                                out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), stream))
    return out


def launch_count() -> int:
    return int(N.load().mrag_launch_count())


def profile_begin(n: int) -> None:
    N.check(N.load().mrag_profile_begin(int(n)))


def profile_read(what: int, max_n: int) -> list[float]:
    buf = (C.c_float * max_n)()
    n = N.load().mrag_profile_read(int(what), buf, max_n)
    return [float(buf[i]) for i in range(max(n, 0))]

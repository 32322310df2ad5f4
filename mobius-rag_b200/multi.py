"""MultiIndex: several row shards of ONE table owned by ONE process (a FastAPI worker that holds all GPUs of a box).

The SPMD path (sharded.ShardedSearcher, one process per GPU under torchrun) is what the benchmark scales; the
reference's callers, however, are a single Python process that calls ``store.asearch(...)`` / ``_vector_arm(...)``
(vector_store.py:181-226, corpus_search.py:1427-1438).  This class gives that process the same row-sharded scan:

  * every shard is an ``Index`` (its own ``mrag_index`` handle, device, streams);  ``devices`` may name a GPU more
    than once (two shards on one GPU), which is how the 1-GPU test box exercises the code;
  * a DOCUMENT lives on one shard (doc_idx filters and tombstones stay local); new documents go to the shard with
    the fewest rows;
  * a row's id is its position in the host table (``mrag_set_row_ids``), so the k-way merge of the shards' lists
    (``mrag_merge_topk``: score DESC, NaN last, id ASC) returns exactly what one unsharded index would;
  * one search = the queries copied to every device, ``mrag_search`` enqueued on one stream per shard (all GPUs
    scan concurrently, nothing blocks the host until the end), the per-shard lists peer-copied to the first
    device, the K4 merge kernel there, one device-to-host copy.

torch is plumbing here (device buffers, streams, peer copies); every kernel is libmrag.so's.
"""
from __future__ import annotations

import os
from typing import Sequence

import numpy as np

from . import _native as N
from .index import Filter, Index, merge_topk


class MultiIndex:
    def __init__(self, dim: int, dtype: str | int, devices: Sequence[int], capacity: int):
        if not devices:
            raise ValueError("MultiIndex needs at least one device")
        self.devices = [int(d) for d in devices]
        s = len(self.devices)
        per = (int(capacity) * 11 // 10 + s - 1) // s + 1024           # least-loaded routing keeps shards within a document of each other
        self.shards = [Index(dim, dtype, d, per) for d in self.devices]
        self.dim, self.dtype, self.device, self.capacity = int(dim), self.shards[0].dtype, self.devices[0], int(capacity)
        self._n = 0
        self.pos_shard = np.zeros(1024, dtype=np.uint8)                 # host position -> shard
        self.pos_local = np.zeros(1024, dtype=np.int64)                 # host position -> row inside that shard
        self.doc_home: dict[int, int] = {}                              # doc_idx -> shard
        self.shard_ids: list[list[np.ndarray]] = [[] for _ in self.devices]   # per shard: host positions of its rows, in row order
        self._last_kind = "none"

    # -- lifecycle ---------------------------------------------------------------------------
    def close(self) -> None:
        for ix in self.shards:
            ix.close()

    def __len__(self) -> int:
        return self._n

    def shard_sizes(self) -> list[int]:
        return [len(ix) for ix in self.shards]

    # -- write side --------------------------------------------------------------------------
    def append(self, X: np.ndarray, meta: np.ndarray | None = None) -> int:
        X = np.ascontiguousarray(X, dtype=np.float32)
        n = X.shape[0]
        first = self._n
        if n == 0:
            return first
        if meta is None:
            from .index import make_meta
            meta = make_meta(n, doc_idx=np.arange(first, first + n, dtype=np.uint32))
        docs = np.asarray(meta["doc_idx"], dtype=np.int64)
        sizes = self.shard_sizes()
        target = np.empty(n, dtype=np.int64)
        uniq, inv, counts = np.unique(docs, return_inverse=True, return_counts=True)
        home = np.empty(uniq.shape[0], dtype=np.int64)
        for j in np.argsort(np.unique(docs, return_index=True)[1]):   # documents in order of first appearance
            d = int(uniq[j])
            s = self.doc_home.get(d)
            if s is None:
                s = int(np.argmin(sizes))
                self.doc_home[d] = s
            sizes[s] += int(counts[j])
            home[j] = s
        target = home[inv]
        need = first + n
        if need > self.pos_shard.shape[0]:
            cap = max(need, 2 * self.pos_shard.shape[0])
            self.pos_shard = np.concatenate([self.pos_shard, np.zeros(cap - self.pos_shard.shape[0], dtype=np.uint8)])
            self.pos_local = np.concatenate([self.pos_local, np.zeros(cap - self.pos_local.shape[0], dtype=np.int64)])
        for s, ix in enumerate(self.shards):
            sel = np.flatnonzero(target == s)
            if sel.size == 0:
                continue
            base = len(ix)
            ids = (first + sel).astype(np.int64)
            N.check(ix._lib.mrag_set_row_ids(ix._h, base, ids.ctypes.data, ids.shape[0]))    # ids first: rows become searchable with ids in place
            got = ix.append(X[sel], np.ascontiguousarray(meta[sel]))
            assert got == base
            self.pos_shard[first + sel] = s
            self.pos_local[first + sel] = base + np.arange(sel.size)
            self.shard_ids[s].append(ids)
        self._n = need
        return first

    def append_device_shard(self, s: int, X, meta: np.ndarray) -> int:
        """Bulk load: rows already resident on shard s's GPU (torch float32 CUDA [m, dim]) go to that shard as a block;
        their ids are the next m host positions.  The caller keeps a document's rows together (a loader that streams whole
        documents per shard, publish.py:310-313)."""
        ix = self.shards[s]
        m = int(X.shape[0])
        first, base = self._n, len(ix)
        ids = np.arange(first, first + m, dtype=np.int64)
        N.check(ix._lib.mrag_set_row_ids(ix._h, base, ids.ctypes.data, m))
        got = ix.append_device(X, meta)
        assert got == base
        need = first + m
        if need > self.pos_shard.shape[0]:
            cap = max(need, 2 * self.pos_shard.shape[0])
            self.pos_shard = np.concatenate([self.pos_shard, np.zeros(cap - self.pos_shard.shape[0], dtype=np.uint8)])
            self.pos_local = np.concatenate([self.pos_local, np.zeros(cap - self.pos_local.shape[0], dtype=np.int64)])
        self.pos_shard[first:need] = s
        self.pos_local[first:need] = base + np.arange(m)
        self.shard_ids[s].append(ids)
        for d in np.unique(meta["doc_idx"]):
            self.doc_home.setdefault(int(d), s)
        self._n = need
        return first

    def set_doc_tags(self, first_doc: int, bits: np.ndarray) -> None:
        for ix in self.shards:                                         # the per-document tables are small: every shard holds all of them
            ix.set_doc_tags(first_doc, bits)

    def set_doc_jtags(self, first_doc: int, bits: np.ndarray) -> None:
        for ix in self.shards:
            ix.set_doc_jtags(first_doc, bits)

    def set_chunk_features(self, first_row: int, feat: np.ndarray) -> None:
        n = feat.shape[0]
        sh, lo = self.pos_shard[first_row:first_row + n], self.pos_local[first_row:first_row + n]
        for s, ix in enumerate(self.shards):
            sel = np.flatnonzero(sh == s)
            if sel.size == 0:
                continue
            loc = lo[sel]
            # rows of one shard inside a host range are consecutive there as well (both orders are append order)
            assert (np.diff(loc) == 1).all()
            ix.set_chunk_features(int(loc[0]), np.ascontiguousarray(feat[sel]))

    def _ids_of(self, s: int) -> np.ndarray:
        if len(self.shard_ids[s]) > 1:
            self.shard_ids[s] = [np.concatenate(self.shard_ids[s])]
        return self.shard_ids[s][0] if self.shard_ids[s] else np.zeros(0, dtype=np.int64)

    def set_dtag_overflow(self, rows: np.ndarray, codes: np.ndarray) -> None:
        rows = np.asarray(rows, dtype=np.int64)
        codes = np.asarray(codes, dtype=np.uint16)
        sh, lo = self.pos_shard[rows], self.pos_local[rows]
        for s, ix in enumerate(self.shards):
            sel = np.flatnonzero(sh == s)
            order = sel[np.argsort(lo[sel], kind="stable")]
            ix.set_dtag_overflow(lo[order].astype(np.uint32), codes[order])

    def dtag_rows(self, flt, codes):
        """as Index.dtag_rows over every shard; rows come back as host positions"""
        rows, n_total, per_code = [], 0, None
        for s, ix in enumerate(self.shards):
            r, nt, pc = ix.dtag_rows(flt, codes)
            rows.append(self._ids_of(s)[r])
            n_total += nt
            per_code = pc if per_code is None else [a + b for a, b in zip(per_code, pc)]
        return np.sort(np.concatenate(rows)) if rows else np.zeros(0, dtype=np.int64), n_total, per_code or []

    def rerank_candidates(self, cands, n: int, hq):
        return self.shards[0].rerank_candidates(cands, n, hq)       # every shard holds all per-document tag sets

    def search_hybrid(self, Q: np.ndarray, k: int, hq, flt: Filter | None = None):
        """the fused hybrid rerank on every shard, then a k-way merge of the <= 8 short lists (score DESC, id ASC)"""
        parts = [ix.search_hybrid(Q, k, hq, flt) for ix in self.shards]
        nq = parts[0][0].shape[0]
        scores = np.full((nq, k), np.nan, dtype=np.float32)
        cos = np.full((nq, k), np.nan, dtype=np.float32)
        rows = np.full((nq, k), -1, dtype=np.int64)
        counts = np.zeros(nq, dtype=np.int32)
        for q in range(nq):
            s_all = np.concatenate([p[0][q, :int(p[3][q])] for p in parts])
            c_all = np.concatenate([p[1][q, :int(p[3][q])] for p in parts])
            r_all = np.concatenate([p[2][q, :int(p[3][q])] for p in parts])
            order = np.lexsort((r_all, -s_all))[:k]
            m = order.shape[0]
            scores[q, :m], cos[q, :m], rows[q, :m], counts[q] = s_all[order], c_all[order], r_all[order], m
        return scores, cos, rows, counts

    def tombstone_doc(self, doc_idx: int) -> int:
        s = self.doc_home.get(int(doc_idx))
        if s is None:
            return 0
        return self.shards[s].tombstone_doc(doc_idx)

    # -- read side ---------------------------------------------------------------------------
    def search(self, Q: np.ndarray, k: int, flt: Filter | None = None, options: int = 0):
        """Host-buffer search over all shards.  Returns (scores f32 [nq,k], ids i64 [nq,k], counts i32 [nq])."""
        import torch
        Q = np.ascontiguousarray(np.atleast_2d(np.asarray(Q, dtype=np.float32)))
        if Q.shape[1] != self.dim:
            raise ValueError(f"query dim {Q.shape[1]} != index dim {self.dim}")
        nq, k, S = Q.shape[0], int(k), len(self.shards)
        options = int(options) & ~N.OPT_COALESCE                        # coalescing is per handle; the fan-out has its own batching
        Qh = torch.from_numpy(Q)
        dev0 = torch.device(f"cuda:{self.devices[0]}")
        parts, events = [], []
        for ix in self.shards:
            dev = torch.device(f"cuda:{ix.device}")
            st = torch.cuda.Stream(device=dev)
            with torch.cuda.device(dev), torch.cuda.stream(st):
                qd = Qh.to(dev, non_blocking=False)
                out = ix.search_device(qd, k, flt, sync=False, options=options)
                ev = torch.cuda.Event()
                ev.record(st)
            parts.append(out)
            events.append(ev)
        self._last_kind = self.shards[0].last_scan_kind()
        with torch.cuda.device(dev0):
            st0 = torch.cuda.Stream(device=dev0)
            with torch.cuda.stream(st0):
                for ev in events:
                    st0.wait_event(ev)
                sc = torch.empty((S, nq, k), dtype=torch.float32, device=dev0)
                ro = torch.empty((S, nq, k), dtype=torch.int64, device=dev0)
                co = torch.empty((S, nq), dtype=torch.int32, device=dev0)
                for s, (a, b, c) in enumerate(parts):                  # peer copies (device-to-device over NVLink)
                    sc[s].copy_(a, non_blocking=True)
                    ro[s].copy_(b, non_blocking=True)
                    co[s].copy_(c, non_blocking=True)
                if S * k <= 16384:
                    m = merge_topk(self.devices[0], sc, ro, co, S, nq, k, (nq * k, nq * k, nq))
                else:                                                  # very wide LIMITs: merge pairwise
                    m = (sc[0], ro[0], co[0])
                    for s in range(1, S):
                        a = torch.stack([m[0], sc[s]]).contiguous()
                        b = torch.stack([m[1], ro[s]]).contiguous()
                        c = torch.stack([m[2], co[s]]).contiguous()
                        m = merge_topk(self.devices[0], a, b, c, 2, nq, k, (nq * k, nq * k, nq))
                res = tuple(t.cpu() for t in m)
            st0.synchronize()
        for (a, b, c) in parts:
            del a, b, c
        return res[0].numpy(), res[1].numpy(), res[2].numpy()

    def last_scan_kind(self) -> str:
        return self._last_kind

    def last_kernel_ms(self, what: int) -> float:
        return float("nan")

    # -- host position <-> shard rows --------------------------------------------------------
    def locate(self, rows: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
        rows = np.asarray(rows, dtype=np.int64)
        return self.pos_shard[rows], self.pos_local[rows]

    # -- snapshot ----------------------------------------------------------------------------
    def save(self, path: str, version: int = 0) -> None:
        for s, ix in enumerate(self.shards):
            ix.save(f"{path}.{s}", version)
        np.savez(path + ".routing.npz", pos_shard=self.pos_shard[:self._n], pos_local=self.pos_local[:self._n],
                 **{f"ids{s}": self._ids_of(s) for s in range(len(self.shards))},
                 doc_home=np.asarray(sorted(self.doc_home.items()), dtype=np.int64).reshape(-1, 2), n=np.int64(self._n))

    @classmethod
    def load(cls, path: str, devices: Sequence[int], capacity: int = 0) -> tuple["MultiIndex", int]:
        self = cls.__new__(cls)
        self.devices = [int(d) for d in devices]
        z = np.load(path + ".routing.npz")
        self._n = int(z["n"])
        self.pos_shard, self.pos_local = z["pos_shard"].copy(), z["pos_local"].copy()
        self.doc_home = {int(a): int(b) for a, b in z["doc_home"]}
        self.shard_ids = [[z[f"ids{s}"].copy()] for s in range(len(self.devices))]
        s = len(self.devices)
        per = 0 if capacity <= 0 else (int(capacity) * 11 // 10 + s - 1) // s + 1024
        self.shards, ver = [], 0
        for i, d in enumerate(self.devices):
            if not os.path.exists(f"{path}.{i}"):
                raise ValueError(f"snapshot has no shard {i}: it was written with fewer shards")
            ix, ver = Index.load(f"{path}.{i}", d, per)
            self.shards.append(ix)
        self.dim, self.dtype, self.device = self.shards[0].dim, self.shards[0].dtype, self.devices[0]
        self.capacity = sum(ix.capacity for ix in self.shards)
        self._last_kind = "none"
        return self, ver

"""Row-sharded search across the GPUs of one box.

Every rank holds a contiguous, document-aligned block of rows in its own Index and sees every
query.  One search = local scan + select on each rank, ONE allgather of the packed per-rank
top-k lists (k*12 + 4 bytes per query) over NCCL / NVLink, then the k-way merge kernel (K4) on
every rank.  No all-reduce: top-k merge is not a sum.

The collective goes through ``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests); the
local search and the merge are injectable so the host logic can be exercised without a GPU.
"""
from __future__ import annotations

from typing import Callable

import numpy as np


def shard_bounds(doc_of_row: np.ndarray, world: int) -> list[tuple[int, int]]:
    """Contiguous [lo, hi) row blocks, cut at document boundaries nearest to r*n/world so a
    document's chunks stay on one rank (doc_idx filters and tombstones stay local)."""
    n = int(doc_of_row.shape[0])
    cuts = [0]
    for r in range(1, world):
        t = (n * r) // world
        t = max(t, cuts[-1])
        if 0 < t < n:
            # move forward to the first row of the next document
            d = doc_of_row[t - 1]
            while t < n and doc_of_row[t] == d:
                t += 1
        cuts.append(min(t, n))
    cuts.append(n)
    return [(cuts[i], cuts[i + 1]) for i in range(world)]


def packed_layout(nq: int, k: int) -> dict:
    """Byte layout of one rank's slot in the allgather buffer: rows i64 | scores f32 | counts i32,
    padded to 8 bytes."""
    rows_off = 0
    scores_off = nq * k * 8
    counts_off = scores_off + nq * k * 4
    size = counts_off + nq * 4
    size = (size + 7) // 8 * 8
    return {"rows_off": rows_off, "scores_off": scores_off, "counts_off": counts_off, "size": size}


class ShardedSearcher:
    """One per rank.  ``local_search(q, k, flt, out)`` must write the local top-k (global row
    ids) into the three views of this rank's slot; ``merge(gathered, world, nq, k, layout)``
    must return (scores, rows, counts) of the global top-k."""

    def __init__(self, index=None, group=None, local_search: Callable | None = None, merge: Callable | None = None,
                 device=None, exchange: str = "auto"):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.index = index
        self.device = device
        self._local = local_search or self._cuda_local
        self._merge = merge or self._cuda_merge
        self._buf_key = None
        # "p2p": results travel by peer stores over NVLink inside the merge kernel (mrag_exchange_merge) -- no
        # collective call; "nccl": all_gather_into_tensor + merge kernel; "auto": p2p when symmetric memory is there
        self.exchange = exchange
        self._p2p = None
        self._epoch = 0
        # search_async: hold a search's exchange back until the next search's prepare phase has run (see there)
        self.defer_exchange = True

    # -- buffers -----------------------------------------------------------------------------
    def _buffers(self, nq: int, k: int):
        import torch
        key = (nq, k)
        if self._buf_key != key:
            lay = packed_layout(nq, k)
            dev = self.device if self.device is not None else (f"cuda:{self.index.device}" if self.index is not None else "cpu")
            self._lay = lay
            self._gathered = torch.zeros(self.world * lay["size"], dtype=torch.uint8, device=dev)
            self._slot = self._gathered[self.rank * lay["size"]:(self.rank + 1) * lay["size"]]
            self._buf_key = key
        return self._lay, self._slot, self._gathered

    @staticmethod
    def slot_views(slot, nq: int, k: int, lay: dict):
        import torch
        rows = slot[lay["rows_off"]:lay["rows_off"] + nq * k * 8].view(torch.int64).view(nq, k)
        scores = slot[lay["scores_off"]:lay["scores_off"] + nq * k * 4].view(torch.float32).view(nq, k)
        counts = slot[lay["counts_off"]:lay["counts_off"] + nq * 4].view(torch.int32)
        return scores, rows, counts

    # -- default (CUDA) implementations ----------------------------------------------------------
    def _cuda_local(self, q, k, flt, out):
        self.index.search_device(q, k, flt, out=out, sync=False)

    def _cuda_merge(self, gathered, world, nq, k, lay):
        from .index import merge_topk
        scores0, rows0, counts0 = self.slot_views(gathered[:lay["size"]], nq, k, lay)
        return merge_topk(self.index.device, scores0, rows0, counts0, world, nq, k,
                          (lay["size"] // 4, lay["size"] // 8, lay["size"] // 4))

    # -- peer-to-peer exchange -------------------------------------------------------------------
    def _p2p_state(self, nq: int, k: int):
        """Symmetric (peer-mapped) buffer for this (nq, k): 2 gather areas + flags; None if unavailable."""
        if self.exchange == "nccl" or self.world == 1 or self.world > 8 or self.index is None:
            return None
        key = (nq, k)
        if self._p2p is not None and self._p2p["key"] == key:
            return self._p2p
        if self._p2p is not None and "xstream" in self._p2p:
            self._flush_deferred(self._p2p)
            self._p2p["xstream"].synchronize()          # exchanges of the old shape still in flight on the side stream
        try:
            import ctypes as C
            import torch
            import torch.distributed._symmetric_memory as symm_mem
            lay = packed_layout(nq, k)
            nbytes = 2 * self.world * lay["size"] + self.world * nq * 4
            dev = torch.device(f"cuda:{self.index.device}")
            buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=dev)
            buf.zero_()
            gname = self.group.group_name if self.group is not None else self.dist.group.WORLD.group_name
            hdl = symm_mem.rendezvous(buf, gname)
            torch.cuda.synchronize(dev)
            hdl.barrier()
            ptrs = (C.c_void_p * self.world)(*[int(p) for p in hdl.buffer_ptrs])
            self._p2p = {"key": key, "lay": lay, "buf": buf, "hdl": hdl, "ptrs": ptrs}
            self._epoch = 0
        except Exception as exc:                    # no symmetric memory on this system: NCCL path
            if self.exchange == "p2p":
                raise
            import logging
            logging.getLogger(__name__).info("sharded search: symmetric memory unavailable (%s); using NCCL allgather", exc)
            self.exchange = "nccl"
            self._p2p = None
        return self._p2p

    def _search_p2p(self, st, q, k: int, flt, events=None):
        import torch
        from . import _native as N
        nq = int(q.shape[0])
        lay, buf = st["lay"], st["buf"]
        # exchanges issued by search_async run on a side stream: this rank's exchange kernels must stay in epoch order
        self._flush_deferred(st)
        for ev in st.get("done", ()):
            if ev is not None:
                torch.cuda.current_stream(buf.device).wait_event(ev)
        self._epoch += 1
        area = (self._epoch & 1) * self.world * lay["size"]
        slot = buf[area + self.rank * lay["size"]: area + (self.rank + 1) * lay["size"]]
        if events:
            events[0].record()
        self._local(q, k, flt, self.slot_views(slot, nq, k, lay))
        if events:
            events[1].record()
        out = (torch.empty((nq, k), dtype=torch.float32, device=buf.device),
               torch.empty((nq, k), dtype=torch.int64, device=buf.device),
               torch.empty((nq,), dtype=torch.int32, device=buf.device))
        stream = torch.cuda.current_stream(buf.device).cuda_stream
        N.check(N.load().mrag_exchange_merge(self.index.device, self.world, self.rank, nq, int(k), st["ptrs"],
                                             lay["size"], lay["scores_off"], lay["counts_off"], self._epoch,
                                             out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), stream))
        if events:
            events[2].record()
        return out

    # -- pipelined search --------------------------------------------------------------------------
    def _enqueue_exchange(self, st, rec, after_event) -> None:
        """Launch the exchange + k-way merge of the search `rec` on the side stream, behind `after_event`."""
        import torch
        from . import _native as N
        lay, xs = st["lay"], st["xstream"]
        xs.wait_event(rec["local_done"])              # (already implied by after_event unless another thread moved the hook's event)
        xs.wait_event(after_event)
        out = rec["out"]
        N.check(N.load().mrag_exchange_merge(self.index.device, self.world, self.rank, rec["nq"], rec["k"], st["ptrs"],
                                             lay["size"], lay["scores_off"], lay["counts_off"], rec["epoch"],
                                             out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), xs.cuda_stream))
        done = torch.cuda.Event()
        done.record(xs)
        rec["done"] = done
        st["done"][rec["j"] % 3] = done

    def _flush_deferred(self, st) -> None:
        """Enqueue the exchange of the last search_async if it is still waiting for a successor."""
        rec = st.get("deferred") if st is not None else None
        if rec is not None:
            st["deferred"] = None
            self._enqueue_exchange(st, rec, rec["local_done"])

    def search_async(self, q, k: int, flt=None) -> "PendingSearch":
        """Like `search`, but the cross-rank exchange + k-way merge runs on a side stream, so it overlaps the local scan
        of the NEXT search issued on the caller's stream (the exchange block is small enough to be resident beside a scan
        CTA).  Returns a handle; `.result()` orders the caller's stream after the exchange and hands out the tensors.

        The exchange of search j is enqueued when search j + 1 is issued, behind the event the library records after
        j + 1's prepare phase (mrag_set_prepared_event): by then scan j + 1 is already queued behind its prepare kernel, so
        the exchange launch does not compete with it at the moment search j ends.  (`.result()` enqueues it at once if no
        successor came.)  At most two searches are in flight: issuing search j first waits (on the device) for the
        exchange of search j - 2, whose gather area and output buffers it reuses; the tensors of a result stay valid
        until three more searches have been issued.  Without peer-mapped memory this degrades to `search`."""
        import torch
        nq = int(q.shape[0])
        st = None
        if self._local == self._cuda_local and self._merge == self._cuda_merge:
            st = self._p2p_state(nq, k)
        if st is None:
            return PendingSearch(self.search(q, k, flt), None, None, None)
        from . import _native as N
        dev = st["buf"].device
        cur = torch.cuda.current_stream(dev)
        if "xstream" not in st:
            st["xstream"] = torch.cuda.Stream(device=dev)
            st["ring"] = [(torch.empty((nq, k), dtype=torch.float32, device=dev),
                           torch.empty((nq, k), dtype=torch.int64, device=dev),
                           torch.empty((nq,), dtype=torch.int32, device=dev)) for _ in range(3)]
            st["done"] = [None, None, None]
            st["issued"] = 0
            st["deferred"] = None
            st["prepared"] = []
            for _ in range(3):                          # persistent events (torch creates the CUDA event at its first record)
                ev = torch.cuda.Event()
                ev.record(cur)
                st["prepared"].append(ev)
        j = st["issued"]
        st["issued"] = j + 1
        prev2 = st["done"][(j - 2) % 3] if j >= 2 else None
        if prev2 is not None:
            cur.wait_event(prev2)                      # gather area (epoch parity) and slot of search j - 2 are free again
        lay, buf = st["lay"], st["buf"]
        self._epoch += 1
        area = (self._epoch & 1) * self.world * lay["size"]
        slot = buf[area + self.rank * lay["size"]: area + (self.rank + 1) * lay["size"]]
        prev = st["deferred"]
        prepared = st["prepared"][j % 3]
        lib = N.load()
        if prev is not None and self.defer_exchange:
            N.check(lib.mrag_set_prepared_event(self.index._h, prepared.cuda_event))
        try:
            self._local(q, k, flt, self.slot_views(slot, nq, k, lay))
        finally:
            if prev is not None and self.defer_exchange:
                lib.mrag_set_prepared_event(self.index._h, None)
        local_done = torch.cuda.Event()
        local_done.record(cur)
        rec = {"j": j, "nq": nq, "k": int(k), "epoch": self._epoch, "out": st["ring"][j % 3], "local_done": local_done, "done": None}
        if prev is not None:
            # the previous search's exchange: behind this search's prepare phase (or, without the hook, its own end)
            st["deferred"] = None
            self._enqueue_exchange(st, prev, prepared if self.defer_exchange else prev["local_done"])
        if self.defer_exchange:
            st["deferred"] = rec
        else:
            self._enqueue_exchange(st, rec, local_done)
        return PendingSearch(rec["out"], rec, st, self)

    def phase_names(self) -> tuple:
        """Names of the intervals between the events `search(..., events=)` records, for the exchange in use."""
        if self.exchange != "nccl" and self._p2p is not None:
            return ("local_search", "exchange+kway_merge (one kernel, peer stores)")
        return ("local_search", "allgather", "kway_merge")

    # -- the search --------------------------------------------------------------------------
    def search(self, q, k: int, flt=None, events=None):
        """q: [nq, dim] float32 on this rank's device, identical on every rank.
        Returns (scores [nq,k], rows [nq,k] global ids, counts [nq]) on every rank.
        events: optional list of 4 torch.cuda.Event (timing enabled) recorded around the phases (see phase_names)."""
        nq = int(q.shape[0])
        if self._local == self._cuda_local and self._merge == self._cuda_merge:
            st = self._p2p_state(nq, k)
            if st is not None:
                return self._search_p2p(st, q, k, flt, events)
        lay, slot, gathered = self._buffers(nq, k)
        if events:
            events[0].record()
        self._local(q, k, flt, self.slot_views(slot, nq, k, lay))
        if events:
            events[1].record()
        if self.world > 1:
            self.dist.all_gather_into_tensor(gathered, slot, group=self.group)
        if events:
            events[2].record()
        out = self._merge(gathered, self.world, nq, k, lay)
        if events:
            events[3].record()
        return out


class PendingSearch:
    """Handle of `ShardedSearcher.search_async`."""

    def __init__(self, out, rec, st, searcher):
        self._out = out
        self._rec = rec
        self._st = st
        self._searcher = searcher

    def _done_event(self):
        if self._rec is None:
            return None
        if self._rec["done"] is None:                 # no successor was issued: enqueue the exchange now
            self._searcher._flush_deferred(self._st)
        return self._rec["done"]

    def copy_to_host(self, scores, rows, counts):
        """Device -> (pinned) host copies of the result, enqueued behind the exchange on ITS stream, so the caller's
        stream -- and the next search's scan on it -- does not wait for them.  Returns the event to synchronize on."""
        import torch
        dev = self._out[0].device
        done = self._done_event()
        stream = self._st["xstream"] if done is not None else torch.cuda.current_stream(dev)
        with torch.cuda.stream(stream):
            scores.copy_(self._out[0], non_blocking=True)
            rows.copy_(self._out[1], non_blocking=True)
            counts.copy_(self._out[2], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(stream)
        return ev

    def result(self, host_sync: bool = False):
        """(scores, rows, counts); the caller's current stream is ordered after the exchange (host_sync: the host too)."""
        done = self._done_event()
        if done is not None:
            import torch
            if host_sync:
                done.synchronize()
            else:
                torch.cuda.current_stream(self._out[0].device).wait_event(done)
        return self._out

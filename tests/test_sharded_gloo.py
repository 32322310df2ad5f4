"""world_size-2 (and 3) gloo test of the row-sharded search's host logic on CPU: document-aligned
partition, packed allgather slot layout, global row ids, and the merge contract.  The local scan
and the merge are the oracle here (this is a test); on GPUs they are the CUDA kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mrag_b200 import sharded, synth


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _np_merge(gathered, world, nq, k, lay):
    """Reference merge for the test: order = score desc, NaN last, row asc."""
    out_s = np.full((nq, k), np.nan, np.float32)
    out_r = np.full((nq, k), -1, np.int64)
    out_c = np.zeros(nq, np.int32)
    views = [sharded.ShardedSearcher.slot_views(gathered[r * lay["size"]:(r + 1) * lay["size"]], nq, k, lay)
             for r in range(world)]
    for q in range(nq):
        cand = []
        for s, r, c in views:
            for j in range(int(c[q])):
                cand.append((float(s[q, j]), int(r[q, j])))
        cand.sort(key=lambda t: (np.isnan(t[0]), -t[0] if not np.isnan(t[0]) else 0.0, t[1]))
        cand = cand[:k]
        out_c[q] = len(cand)
        for j, (s, r) in enumerate(cand):
            out_s[q, j], out_r[q, j] = s, r
    return torch.from_numpy(out_s), torch.from_numpy(out_r), torch.from_numpy(out_c)


def _worker(rank, world, port, n, dim, k, nq, result_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle
        X, valid = synth.make_corpus(n, dim, seed=21, null_frac=1e-2)
        meta, _, info = synth.make_metadata(n, seed=22, rows_per_doc=16, valid=valid)
        Q = synth.make_queries(X, nq, seed=23)
        lo, hi = sharded.shard_bounds(info["doc_of_row"], world)[rank]

        def local_search(q, kk, flt, out):
            rows, sims, counts = oracle.search(X[lo:hi], q.numpy(), kk, valid[lo:hi].astype(bool))
            rows = np.where(rows >= 0, rows + lo, -1)
            out[0].copy_(torch.from_numpy(sims.astype(np.float32)))
            out[1].copy_(torch.from_numpy(rows))
            out[2].copy_(torch.from_numpy(counts.astype(np.int32)))

        ss = sharded.ShardedSearcher(index=None, local_search=local_search, merge=_np_merge, device="cpu")
        scores, rows, counts = ss.search(torch.from_numpy(Q), k)
        # without peer-mapped device memory search_async is a completed search behind the same handle
        p = ss.search_async(torch.from_numpy(Q), k)
        s2, r2, c2 = p.result(host_sync=True)
        assert torch.equal(r2, rows) and torch.equal(c2, counts)
        if rank == 0:
            np.savez(os.path.join(result_dir, "out.npz"), scores=scores.numpy(), rows=rows.numpy(), counts=counts.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_search_gloo(tmp_path, oracle, world):
    n, dim, k, nq = 3000, 32, 12, 5
    mp.spawn(_worker, args=(world, _free_port(), n, dim, k, nq, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "out.npz")
    X, valid = synth.make_corpus(n, dim, seed=21, null_frac=1e-2)
    Q = synth.make_queries(X, nq, seed=23)
    mask = valid.astype(bool)
    for i in range(nq):
        sim_all = oracle.all_similarities(X, Q[i])
        oracle.check_topk(got["rows"][i], got["scores"][i], int(got["counts"][i]), sim_all, mask, k, rtol=1e-5)

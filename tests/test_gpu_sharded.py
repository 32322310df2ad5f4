"""Row-sharded search on GPUs: (1) emulated on one GPU (several shards = several Index objects on
cuda:0, the packed allgather slots filled by hand) so the path is covered on a single-GPU box;
(2) real NCCL allgather across 2 ranks when the box has at least 2 GPUs."""
import os
import socket

import numpy as np
import pytest

from mrag_b200 import sharded, synth
from mrag_b200.index import Filter, Index

pytestmark = pytest.mark.gpu


def _check(oracle, Xs, Q, mask, k, s, r, c, rtol):
    for i in range(Q.shape[0]):
        oracle.check_topk(r[i], s[i], int(c[i]), oracle.all_similarities(Xs, Q[i]), mask, k, rtol=rtol)


@pytest.mark.parametrize("dtype,world,nq,k", [("bf16", 3, 9, 10), ("f32", 2, 3, 100), ("bf16", 8, 64, 10), ("bf16", 4, 2, 300)])
def test_sharded_emulated_on_one_gpu(oracle, dtype, world, nq, k):
    import torch
    n, dim = 60000, 256
    X, valid = synth.make_corpus(n, dim, seed=91, null_frac=1e-3)
    meta, doc_tags, info = synth.make_metadata(n, seed=92, rows_per_doc=64, valid=valid)
    Q = synth.make_queries(X, nq, seed=93)
    bounds = sharded.shard_bounds(info["doc_of_row"], world)
    shards = []
    for lo, hi in bounds:
        idx = Index(dim, dtype, 0, max(hi - lo, 1))
        if hi > lo:
            idx.append(X[lo:hi], meta[lo:hi])
        idx.set_doc_tags(0, doc_tags)
        idx.set_row_base(lo)
        shards.append(idx)
    lay = sharded.packed_layout(nq, k)
    gathered = torch.zeros(world * lay["size"], dtype=torch.uint8, device="cuda:0")
    qd = torch.from_numpy(Q).cuda()
    pool = np.random.default_rng(1).choice(info["n_docs"], size=60, replace=False)
    for flt, mask in [(None, valid.astype(bool)),
                      (Filter().doc_pool(pool), np.isin(meta["doc_idx"], pool) & valid.astype(bool))]:
        for r, idx in enumerate(shards):
            out = sharded.ShardedSearcher.slot_views(gathered[r * lay["size"]:(r + 1) * lay["size"]], nq, k, lay)
            idx.search_device(qd, k, flt, out=out, sync=False)
        from mrag_b200.index import merge_topk
        s0, r0, c0 = sharded.ShardedSearcher.slot_views(gathered[:lay["size"]], nq, k, lay)
        s, r, c = merge_topk(0, s0, r0, c0, world, nq, k, (lay["size"] // 4, lay["size"] // 8, lay["size"] // 4))
        torch.cuda.synchronize()
        Xs = oracle.round_bf16(X) if dtype == "bf16" else X
        _check(oracle, Xs, Q, mask, k, s.cpu().numpy(), r.cpu().numpy(), c.cpu().numpy(), 1e-2 if dtype == "bf16" else 1e-4)
    for idx in shards:
        idx.close()


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _nccl_worker(rank, world, port, n, dim, nq, k, outdir, exchange="nccl", pipelined=0):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    try:
        X, valid = synth.make_corpus(n, dim, seed=191, null_frac=1e-3)
        meta, _, info = synth.make_metadata(n, seed=192, rows_per_doc=64, valid=valid)
        Q = synth.make_queries(X, nq, seed=193)
        lo, hi = sharded.shard_bounds(info["doc_of_row"], world)[rank]
        idx = Index(dim, "bf16", rank, hi - lo)
        idx.append(X[lo:hi], meta[lo:hi])
        idx.set_row_base(lo)
        ss = sharded.ShardedSearcher(index=idx, exchange=exchange)
        qd = torch.from_numpy(Q).cuda()
        if pipelined:
            # a different batch every step, two searches in flight, then a plain search on the same searcher
            def lagged():
                pend, outs = [], []
                for i in range(pipelined):
                    pend.append(ss.search_async(torch.from_numpy(synth.make_queries(X, nq, seed=193 + i)).cuda(), k))
                    if i >= 1:
                        outs.append(tuple(t.clone() for t in pend[i - 1].result()))
                outs.append(tuple(t.clone() for t in pend[-1].result(host_sync=True)))
                return outs
            outs = lagged()                              # exchange of step i released by step i + 1's prepare phase
            s, r, c = ss.search(qd, k)
            torch.cuda.synchronize()
            assert torch.equal(r, outs[0][1]) and torch.equal(c, outs[0][2])
            # the result taken at once (no successor: the handle enqueues the exchange itself)
            for i in range(3):
                o = ss.search_async(torch.from_numpy(synth.make_queries(X, nq, seed=193 + i)).cuda(), k).result(host_sync=True)
                assert torch.equal(o[1], outs[i][1]) and torch.equal(o[0], outs[i][0])
            ss.defer_exchange = False                    # exchange enqueued right behind its own search
            outs2 = lagged()
            for a, b in zip(outs, outs2):
                assert torch.equal(a[1], b[1]) and torch.equal(a[0], b[0]) and torch.equal(a[2], b[2])
            np.savez(os.path.join(outdir, f"out{rank}.npz"), s=np.stack([o[0].cpu().numpy() for o in outs]),
                     r=np.stack([o[1].cpu().numpy() for o in outs]), c=np.stack([o[2].cpu().numpy() for o in outs]))
            idx.close()
            return
        for _ in range(5):                             # several epochs: both gather areas and the flags are reused
            s, r, c = ss.search(qd, k)
        assert ss.exchange == exchange
        torch.cuda.synchronize()
        np.savez(os.path.join(outdir, f"out{rank}.npz"), s=s.cpu().numpy(), r=r.cpu().numpy(), c=c.cpu().numpy())
        idx.close()
    finally:
        dist.destroy_process_group()


def test_sharded_nccl_two_ranks(oracle, tmp_path):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    n, dim, nq, k = 80000, 768, 7, 10
    mp.spawn(_nccl_worker, args=(2, _free_port(), n, dim, nq, k, str(tmp_path)), nprocs=2, join=True)
    X, valid = synth.make_corpus(n, dim, seed=191, null_frac=1e-3)
    Q = synth.make_queries(X, nq, seed=193)
    outs = [np.load(tmp_path / f"out{r}.npz") for r in range(2)]
    assert (outs[0]["r"] == outs[1]["r"]).all() and (outs[0]["c"] == outs[1]["c"]).all()     # every rank gets the same answer
    _check(oracle, oracle.round_bf16(X), Q, valid.astype(bool), k, outs[0]["s"], outs[0]["r"], outs[0]["c"], 1e-2)


def test_sharded_p2p_exchange_two_ranks(oracle, tmp_path):
    """The fused exchange + merge kernel (peer stores over NVLink, no collective call) gives what NCCL + K4 gives."""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    n, dim, nq, k = 80000, 768, 7, 10
    mp.spawn(_nccl_worker, args=(2, _free_port(), n, dim, nq, k, str(tmp_path), "p2p"), nprocs=2, join=True)
    X, valid = synth.make_corpus(n, dim, seed=191, null_frac=1e-3)
    Q = synth.make_queries(X, nq, seed=193)
    outs = [np.load(tmp_path / f"out{r}.npz") for r in range(2)]
    assert (outs[0]["r"] == outs[1]["r"]).all() and (outs[0]["c"] == outs[1]["c"]).all()
    _check(oracle, oracle.round_bf16(X), Q, valid.astype(bool), k, outs[0]["s"], outs[0]["r"], outs[0]["c"], 1e-2)


def test_sharded_pipelined_exchange_two_ranks(oracle, tmp_path):
    """search_async: the exchange of search i runs on a side stream while search i + 1 scans; every step returns its own
    batch's answer on both ranks."""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    n, dim, nq, k, steps = 80000, 768, 7, 10, 7
    mp.spawn(_nccl_worker, args=(2, _free_port(), n, dim, nq, k, str(tmp_path), "p2p", steps), nprocs=2, join=True)
    X, valid = synth.make_corpus(n, dim, seed=191, null_frac=1e-3)
    outs = [np.load(tmp_path / f"out{r}.npz") for r in range(2)]
    assert (outs[0]["r"] == outs[1]["r"]).all() and (outs[0]["c"] == outs[1]["c"]).all()
    Xb = oracle.round_bf16(X)
    for i in range(steps):
        _check(oracle, Xb, synth.make_queries(X, nq, seed=193 + i), valid.astype(bool), k, outs[0]["s"][i], outs[0]["r"][i], outs[0]["c"][i], 1e-2)

"""bench.py's reference arm (`--impl reference`) runs on host cores only, so its JSON contract is checked here,
without a GPU: one line, the keys the driver reads, the same metric / unit / config as the product arm, and a
cpu_baseline block that describes the run.  (The product arm needs a B200; its line is checked by the driver.)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*extra, env=None):
    base = ["--steps", "1", "--warmup", "1", "--full-scan", "0"]
    for flag in ("--steps", "--warmup", "--full-scan"):
        if flag in extra:
            i = base.index(flag)
            del base[i:i + 2]
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *base, "--cpu-rows", "20000", "--cpu-queries", "2", *extra]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run()
    assert d["impl"] == "reference"
    assert d["unit"] == "queries/s" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["steps"] == 1 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["vs_baseline"] is None and d["data"] == "synthetic"
    cfg = d["config"]
    assert "10000000x768" in cfg["workload"] and cfg["k"] == 10 and cfg["batch"] == 64
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"]
    assert "20000 rows" in cb["sample"] and cb["unit"] == d["unit"]
    e2e = d["e2e"]
    assert e2e["value"] == d["value"] and e2e["unit"] == d["unit"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    assert d["single_thread"]["cores"] == 1 and d["single_thread"]["value"] > 0


def test_reference_arm_honours_steps_and_validates_the_extrapolation():
    d = _run("--rows", "60000", "--full-scan", "1", "--steps", "3", "--warmup", "2")
    assert d["steps"] == 3 and d["warmup"] == 2
    v = d["full_scan_validation"]
    assert v["rows"] == 60000 and 0.2 < v["ratio_measured_over_extrapolated"] < 5.0


def test_reference_arm_other_ranks_exit_quietly():
    """Under torchrun only rank 0 runs the CPU arm; the other ranks print nothing and exit 0."""
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29999")
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1",
           "--cpu-rows", "20000", "--cpu-queries", "2"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    assert not [ln for ln in out.stdout.splitlines() if ln.startswith("{")]


def test_reference_arm_follows_workload_flags():
    d = _run("--workload", "c3")
    assert d["config"]["k"] == 100 and "top-100" in d["config"]["workload"]


# ---------------------------------------------------------------------------------------------
# host-side pieces of the product arm that need no GPU: document layout, shard cuts, the WHERE of a workload, the digest
# ---------------------------------------------------------------------------------------------
def test_ragged_documents_and_document_aligned_shards():
    import numpy as np
    sys.path.insert(0, ROOT)
    import bench
    rows = 1_000_003
    ends = bench.doc_layout(rows)
    lens = np.diff(np.concatenate([[0], ends]))
    assert ends[-1] == rows and (lens >= 1).all() and lens[:-1].min() >= 8 and lens.max() <= 120 and 60 < lens.mean() < 68
    docs = bench.docs_of(ends, 0, rows)
    assert docs[0] == 0 and docs[-1] == len(ends) - 1 and (np.diff(docs.astype(np.int64)) >= 0).all()
    assert (bench.docs_of(ends, 12345, 1000) == docs[12345:13345]).all()
    for world in (1, 2, 4, 8):
        cuts = bench.shard_cuts(ends, rows, world)
        assert cuts[0] == 0 and cuts[-1] == rows and len(cuts) == world + 1
        sizes = np.diff(cuts)
        assert sizes.max() - sizes.min() < 2 * 120                     # balanced to within a document
        for c in cuts[1:-1]:
            assert docs[c] != docs[c - 1], "a shard boundary must fall between two documents"


def test_workload_filters_and_digest():
    import argparse
    import numpy as np
    import torch
    import bench
    ends = bench.doc_layout(100_000)
    docs = bench.docs_of(ends, 0, 100_000)
    pool, passes = bench.filter_spec(argparse.Namespace(doc_pool=40, tag_filter=0, payer_filter=0), ends)
    ok = passes(docs)
    assert ok.any() and np.isin(docs[ok], pool).all() and len(set(docs[ok].tolist())) == 40
    _, passes = bench.filter_spec(argparse.Namespace(doc_pool=0, tag_filter=10, payer_filter=13), ends)
    ok = passes(docs)
    assert ok.any() and ((docs[ok] % 10) == 0).all() and ((docs[ok] % 13) == 3).all()
    assert not passes(docs[(docs % 10) != 0]).any()
    args = argparse.Namespace(doc_pool=0, tag_filter=0, payer_filter=0)
    assert bench.filter_spec(args, ends)[1](docs).all()
    r = torch.arange(20, dtype=torch.int64).reshape(2, 10)
    c = torch.tensor([10, 10], dtype=torch.int32)
    assert bench.result_digest(r, c) == bench.result_digest(r.clone(), c.clone()) != bench.result_digest(r.flip(1), c)


def test_stream_check_equals_one_shot_check(oracle):
    import numpy as np
    from mrag_b200 import synth
    X, valid = synth.make_corpus(5000, 48, seed=5)
    Q = synth.make_queries(X, 3, seed=6)
    rows, sims, counts = oracle.search(X, Q, 10, valid.astype(bool))
    sc = oracle.StreamCheck(5000, Q)
    for lo in range(0, 5000, 1234):
        sc.feed(lo, X[lo:lo + 1234])
    for i in range(3):
        sc.check(i, rows[i], sims[i].astype(np.float32), int(counts[i]), valid.astype(bool), 10, rtol=1e-4)
    bad = rows[0].copy()
    bad[3], bad[7] = bad[7], bad[3]
    with __import__("pytest").raises(AssertionError):
        sc.check(0, bad, sims[0].astype(np.float32), int(counts[0]), valid.astype(bool), 10, rtol=1e-4)

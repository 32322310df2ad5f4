#!/usr/bin/env python
"""bench.py -- QPS of exact cosine top-k over a row-sharded synthetic corpus (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A "step" is one batch of `--batch` queries scanned against the whole corpus (each rank scans its
row shard, one allgather of the per-rank top-k lists, k-way merge).  Total corpus size is fixed as
N grows (strong scaling: the metric is quoted on ONE 10M x 768 corpus at 1/2/4/8 GPUs).

Prints ONE JSON line: value = whole-job QPS with inputs resident in HBM (CUDA events, max over
ranks); e2e = the same through the host-buffer API (pinned host queries in, host results out);
roofline = the scan kernel against the measured HBM peak; cpu_baseline = the CPU oracle
(restatement of the reference's pgvector path) timed on this box's cores on a bounded sample.
`--impl reference` times that CPU path alone (rank 0 only) and prints the same line shape.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "QPS exact cosine top-10, 10Mx768 corpus"
DEFAULT_BATCH = 64         # queries per step (see DESIGN.md "Measurement")
CHUNK = 1 << 16            # rows generated per chunk (the generator's unit; shard cuts fall on DOCUMENT boundaries, not here)
PARITY_QUERIES = 4         # queries of the timed batch re-checked against the CPU oracle over the whole corpus, at every N


def doc_layout(rows: int, seed: int = 77) -> np.ndarray:
    """Ragged documents: 8..120 chunks each (mean 64), rows doc-contiguous as publish writes them (publish.py:310-313).
    Returns ends[d] = first row after document d."""
    rng = np.random.default_rng(seed)
    lens = rng.integers(8, 121, size=rows // 8 + 2)
    ends = np.cumsum(lens)
    n_docs = int(np.searchsorted(ends, rows, side="left")) + 1
    ends = ends[:n_docs].copy()
    ends[-1] = rows
    return ends


def docs_of(ends: np.ndarray, first: int, m: int) -> np.ndarray:
    return np.searchsorted(ends, np.arange(first, first + m), side="right").astype(np.uint32)


def shard_cuts(ends: np.ndarray, rows: int, world: int) -> list[int]:
    """Row-shard boundaries: the document boundary at or after r * rows / world (a document stays on one rank; the
    shards differ by less than one document, < 120 rows)."""
    cuts = [0]
    for r in range(1, world):
        t = (rows * r) // world
        cuts.append(int(ends[np.searchsorted(ends, t, side="left")]) if t > 0 else 0)
    cuts.append(rows)
    return cuts


def filter_spec(args, ends: np.ndarray):
    """The WHERE of the workload as (pool array or None, callable doc ids -> bool pass array)."""
    n_docs = len(ends)
    pool = None
    if args.doc_pool:
        pool = np.random.default_rng(5).choice(n_docs, size=min(args.doc_pool, n_docs), replace=False).astype(np.uint32)

    def passes(docs: np.ndarray) -> np.ndarray:
        ok = np.ones(docs.shape[0], dtype=bool)
        if args.tag_filter:
            ok &= (docs % args.tag_filter) == 0
        if pool is not None:
            ok &= np.isin(docs, pool)
        if args.payer_filter:
            ok &= (docs % args.payer_filter) == (3 % args.payer_filter)
        return ok
    return pool, passes


def result_digest(rows_t, counts_t) -> str:
    """sha256 over the returned row ids (+ counts) of the timed batch: must be equal at every N (ties are broken by
    ascending global row, so the answer does not depend on the sharding)."""
    import hashlib
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(rows_t.cpu().numpy()).tobytes())
    h.update(np.ascontiguousarray(counts_t.cpu().numpy()).tobytes())
    return h.hexdigest()[:16]


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--batch", type=int, default=DEFAULT_BATCH)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--sweep", default="1,4,16,256,1024", help="extra batch sizes measured briefly at N=1 ('' = none)")
    ap.add_argument("--cpu-rows", type=int, default=200_000, help="rows of the CPU baseline sample")
    ap.add_argument("--cpu-queries", type=int, default=8, help="queries of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--full-scan", type=int, default=1, help="--impl reference: also scan the full row count once, streamed, to validate the extrapolation (0 = skip)")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle re-check of the timed batch over the whole corpus")
    ap.add_argument("--also-f32", type=int, default=-1,
                    help="1: after the bf16 run, store the same corpus as float4 and measure the same batch (default: on for the headline workload)")
    ap.add_argument("--threads", type=int, default=32, help="host threads of the concurrent single-query measurement (0 = skip)")
    ap.add_argument("--tag-filter", type=int, default=0, metavar="M",
                    help="document-tag filter passing every M-th document (0 = no filter); C2 uses 10")
    ap.add_argument("--payer-filter", type=int, default=0, metavar="M",
                    help="payor bitset filter passing the documents of 1 payer out of M (0 = none); C4 uses 13")
    ap.add_argument("--doc-pool", type=int, default=0, metavar="D",
                    help="document_id = ANY(pool) filter with D random documents (64 rows each); the pinned-pool case of corpus_search.py:414-420")
    ap.add_argument("--workload", default="", choices=["", "c2", "c3", "c4", "c5", "pool"],
                    help="shortcut: c2 = 1Mx768 fp32, batch 256, top-10, tag filter 10%%; c3 = 10Mx768 bf16 top-100; "
                         "c5 = hybrid rerank over 10M chunks, 22-query bank, top-50")
    return ap.parse_args()


def load_traffic(kernel: str, args, n_local: int, q_per_launch: int):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed
    `ncu --set full` capture (profiles/ncu_traffic.json), if one exists for this shape."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        d = json.load(open(p))
    except Exception:
        return None
    key = f"{kernel}:{n_local}x{args.dim}:{args.dtype}:q{q_per_launch}:k{args.k}"
    e = d.get(key)
    return float(e["dram_bytes"]) if e else None


def load_tensor_peak():
    """Dense bf16 TFLOP/s a kernel inside a long step can sustain (cuBLAS, back to back): MEASURED_PEAKS.json, else the
    profiling recipe's fallback."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d.get("bf16_tflops_sustained") or d["bf16_tflops"]), "measured sustained (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 1400.0, "fallback sustained (B200_PROFILING.md)"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region: NVML polled every 2 ms from a thread (a timed region
    can be shorter than one nvidia-smi period); nvidia-smi -lms 20 if NVML cannot be loaded."""
    Q = ("index,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.lines = gpu_index, None, []
        self.nvml, self.h, self.samples, self._stop = None, None, [], False

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.gpu]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else self.gpu
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
            self._stop = False
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self._stop:
            try:
                sm = n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)
                mx = n.nvmlDeviceGetMaxClockInfo(self.h, n.NVML_CLOCK_SM)
                try:
                    rs = n.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((sm, mx, rs))
            except Exception:
                pass
            time.sleep(0.002)

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.nvml is not None:
            self._stop = True
            self.t.join(timeout=1)
            n = self.nvml
            bits = {"hw_slowdown": getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                    "hw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                    "sw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                    "sw_power_cap": getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
            reasons = sorted({name for _, _, rs in self.samples for name, b in bits.items() if rs & b})
            sm = [x[0] for x in self.samples]
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(x[1] for x in self.samples)) if sm else None,
                    "reasons": reasons, "samples": len(sm), "source": "nvml, 2 ms period"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 20"}


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle (restatement of the reference's pgvector path) on this box's cores
# ---------------------------------------------------------------------------------------------
def cpu_reference_qps(args, steps: int, warmup: int, extras: bool = False) -> dict:
    """Times oracle.search (C restatement of pgvector cosine_distance + ORDER BY/LIMIT, all host
    threads) on a bounded sample: `cpu_rows` rows x `cpu_queries` queries per step, and scales the
    per-query time linearly in the row count to the full corpus (the scan is linear in rows).
    extras: also (a) the single-thread "pgvector model" figure (one backend process per statement) and (b) ONE pass over
    the FULL row count, streamed chunk by chunk, to validate the linear extrapolation."""
    from oracle import oracle
    from mrag_b200 import synth
    cores = os.cpu_count() or 1
    oracle.set_threads(cores)
    n_s, nq_s = min(args.cpu_rows, args.rows), args.cpu_queries
    X, valid = synth.make_corpus(n_s, args.dim, seed=1234)
    if args.dtype == "bf16":
        X = oracle.round_bf16(X)
    Q = synth.make_queries(X, nq_s, seed=4321)
    mask = valid.astype(bool)
    for _ in range(max(0, warmup)):
        oracle.search(X, Q, args.k, mask)
    t0 = time.perf_counter()
    for _ in range(steps):
        oracle.search(X, Q, args.k, mask)
    dt = time.perf_counter() - t0
    per_query_sample = dt / (steps * nq_s)
    scale = args.rows / n_s
    per_query_full = per_query_sample * scale
    out = {
        "value": 1.0 / per_query_full, "unit": "queries/s", "cores": cores, "kind": "port",
        "sample": f"{n_s} rows x {args.dim} ({args.dtype} rows upcast to fp32, as pgvector stores float4) x {nq_s} queries x "
                  f"{steps} steps, {cores} scan threads; per-query time scaled x{scale:.0f} to {args.rows} rows",
        "seconds": dt, "ms_per_query_sample": per_query_sample * 1e3,
    }
    if extras:
        # (a) one Postgres backend executes one statement: the same scan on ONE thread
        oracle.set_threads(1)
        oracle.search(X, Q[:1], args.k, mask)
        t1 = time.perf_counter()
        reps = 2
        for _ in range(reps):
            oracle.search(X, Q[:2], args.k, mask)
        st = (time.perf_counter() - t1) / (reps * 2) * scale
        out["single_thread"] = {"value": 1.0 / st, "unit": "queries/s", "cores": 1,
                                "note": "one scan thread = one Postgres backend per statement (pgvector runs a query on one process)"}
        oracle.set_threads(cores)
        # (b) the full row count once, streamed: same kernel over `rows` rows in chunks of the sample size (fresh random
        #     rows per chunk, generation not timed), per-chunk top-k merged on the host like a parallel seq scan's gather
        import torch
        g = torch.Generator().manual_seed(99)
        scan_s, done = 0.0, 0
        while done < args.rows and args.full_scan:
            m_rows = min(n_s, args.rows - done)
            Xc = torch.randn((m_rows, args.dim), generator=g, dtype=torch.float32).numpy()
            t2 = time.perf_counter()
            r, sm, c = oracle.search(Xc, Q, args.k, None)
            scan_s += time.perf_counter() - t2
            done += m_rows
        if args.full_scan:
            out["full_scan_validation"] = {"rows": args.rows, "queries": nq_s, "ms_per_query": scan_s / nq_s * 1e3,
                                       "extrapolated_ms_per_query": per_query_full * 1e3,
                                       "ratio_measured_over_extrapolated": (scan_s / nq_s) / per_query_full}
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    base = cpu_reference_qps(args, steps, warmup, extras=True)
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": "queries/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": base["seconds"] / steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.batch),
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "single_thread": base.get("single_thread"), "full_scan_validation": base.get("full_scan_validation"),
        "note": "reference = Postgres+pgvector, not runnable here (no Postgres, extension not vendored): this arm "
                "is the CPU oracle port of its exact-scan plan, all host threads; each step = the bounded sample in cpu_baseline.sample",
    }
    print(json.dumps(line), flush=True)


def workload_config(args, batch):
    return {
        "workload": f"{args.rows}x{args.dim} {args.dtype} corpus, top-{args.k}, query batch {batch}, row-sharded",
        "rows": args.rows, "dim": args.dim, "corpus_dtype": args.dtype, "accumulate": "f32", "k": args.k, "batch": batch,
        "filter": (f"document_id = ANY(pool of {args.doc_pool} documents)" if args.doc_pool else
                   f"payor bitset passing 1/{args.payer_filter} of the documents" if args.payer_filter else
                   "embedding_vec IS NOT NULL only" if not args.tag_filter else
                   f"document tag filter (relaxed, 1 tag) passing 1/{args.tag_filter} of the documents"), "l2": "inputs larger than L2 (no flush needed)",
        "parallelism": f"rowshard{args.gpus}",
    }


# ---------------------------------------------------------------------------------------------
# config 5: hybrid rerank fused with the scan (single GPU; host-buffer API, so value == e2e)
# ---------------------------------------------------------------------------------------------
def c5_cpu_baseline(n_rows: int, k: int) -> dict:
    """The reference's `_rerank` (oracle restatement, pure Python like the original) over a bounded candidate sample,
    scaled linearly in the candidate count to the rows the fused GPU path scores per query."""
    from oracle import oracle
    rng = np.random.default_rng(7)
    pool = ["prior authorization", "timely filing", "appeal", "medical records", "provider services", "claims submission",
            "credentialing", "behavioral health"]
    filler = ("the plan requires that providers follow the documented process for each covered service and retain supporting "
              "notes for review by the health plan within the stated period of time").split()
    n_s = 4000
    cands = []
    for i in range(n_s):
        words = [pool[int(rng.integers(0, len(pool)))] if rng.random() < 0.1 else filler[int(rng.integers(0, len(filler)))] for _ in range(60)]
        sim = float(rng.random())
        cands.append({"id": f"c{i}", "text": " ".join(words), "document_name": f"Provider Manual {i % 50}", "document_id": f"d{i % 50}",
                      "source_type": "hierarchical", "authority_level": "payer_policy", "payer": "Sunshine Health", "state": "FL",
                      "similarity": sim, "arm_scores": {"vector": sim}, "_arm": "vector", "chunk_d_tags": {}})
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        oracle.rerank([dict(c) for c in cands], "timely filing appeal deadline", ["timely filing", "appeal"], [0.9, 0.7], [None, None])
    per_query = (time.perf_counter() - t0) / reps * (n_rows / n_s)
    return {"value": 1.0 / per_query, "unit": "queries/s", "cores": 1, "kind": "port",
            "sample": f"`_rerank` restatement (pure Python, as the reference) over {n_s} candidates x {reps} queries, 2 required phrases; per-query "
                      f"time scaled x{n_rows / n_s:.0f} to the {n_rows} rows the fused GPU path scores per query (the reference itself only ever "
                      f"reranks <= ~3k RRF candidates)"}


def run_c5(args, idx, plant, dev, n_local, world, rank, hbm_peak, peak_src, t_build, clocks):
    import ctypes as C
    import torch
    from mrag_b200 import _native as N
    from mrag_b200 import index as mi
    from mrag_b200 import synth
    if world != 1:
        raise SystemExit("bench.py --workload c5 runs on one GPU")
    n, nq, k = n_local, args.batch, args.k
    rng = np.random.default_rng(55)
    # ---- synthetic text features: 64 dictionary phrases present in 0.1 % .. 5 % of the chunks,
    #      5 % of the chunks carry a chunk d-tag, sparse JPD hits, documents of 64 rows with random j-tags
    t0 = time.perf_counter()
    feat = np.zeros(n, dtype=mi.FEAT_DTYPE)
    dens = np.geomspace(0.001, 0.05, 64)
    for p in range(64):
        hit = rng.choice(n, size=int(n * dens[p]), replace=False)
        feat["phrase_bits"][hit, 0] |= np.uint64(1 << p)
    for c in range(N.MRAG_JPD_CATS):
        hit = rng.choice(n, size=n // 20, replace=False)
        feat["jpd_hits"][hit, c] = rng.integers(1, 4, size=hit.shape[0])
    feat["length_score"] = rng.random(n, dtype=np.float32)
    feat["flags"] = (rng.random(n) < 0.3).astype(np.uint8) * N.CF_SHORT_TEXT | (rng.random(n) < 0.01).astype(np.uint8) * N.CF_CONTACT_VALUE
    tagged = rng.choice(n, size=n // 20, replace=False)
    feat["dtags"][tagged, 0] = rng.integers(1, 33, size=tagged.shape[0])
    idx.set_chunk_features(0, feat)
    n_docs = len(doc_layout(n))                          # ragged documents, as main() numbered them
    jt = np.zeros((n_docs, N.MRAG_JTAG_WORDS), dtype=np.uint64)
    jt[:, 0] = rng.integers(0, 256, size=n_docs).astype(np.uint64) & rng.integers(0, 256, size=n_docs).astype(np.uint64)   # 8 j-codes, 25 % each
    idx.set_doc_jtags(0, jt)
    t_feat = time.perf_counter() - t0
    # ---- the query bank: 22 queries shaped like eval/queries.yaml entries, 1-4 required phrases each
    # (its own generator: the bank does not depend on the corpus layout.  Shape of eval/queries.yaml entries: a payer / state
    #  phrase that carries a j: code -- binary credit for every chunk of a tagged document -- only next to topical phrases,
    #  which are substring-tested; some topical phrases carry a d: code)
    qrng = np.random.default_rng(56)
    hq = (N.HybridQuery * nq)()
    for i in range(nq):
        h = hq[i]
        npz = int(qrng.integers(1, 5))
        h.n_phrases = npz
        for j in range(npz):
            h.phrase_weight[j] = float(qrng.uniform(0.65, 1.0))
            h.phrase_bit[j] = int(qrng.integers(0, 64))
            h.phrase_jbit[j] = int(qrng.integers(0, 8)) if (j == 0 and npz >= 2 and qrng.random() < 0.5) else -1
            h.phrase_dcode[j] = int(qrng.integers(1, 33)) if (h.phrase_jbit[j] < 0 and qrng.random() < 0.2) else 0
        for c in range(N.MRAG_JPD_CATS):
            h.qcat[c] = float(qrng.random() < 0.2) * 0.4
        for a in range(32):
            h.auth_score[a] = 0.1
        h.w_sim, h.w_auth, h.w_len, h.w_cov, h.boost, h.floor = 0.25, 0.10, 0.05, 0.55, 1.5, 1.0
        h.w_jpd = 0.20 if any(h.qcat[c] > 0 for c in range(N.MRAG_JPD_CATS)) else 0.0
    Q = synth.cuda_queries(plant, nq, args.dim, dev, seed=4321).cpu().numpy()
    for _ in range(max(args.warmup, 3)):
        res = idx.search_hybrid(Q, k, hq)
    clocks.start()
    mi.profile_begin(args.steps)
    l0 = mi.launch_count()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = idx.search_hybrid(Q, k, hq)
    dt = time.perf_counter() - t0
    launches = mi.launch_count() - l0
    dev_ms = [x for x in mi.profile_read(3, args.steps) if x >= 0]
    prep_ms = [x for x in mi.profile_read(0, args.steps) if x >= 0]
    scan_ms = [x for x in mi.profile_read(1, args.steps) if x >= 0]
    mi.profile_begin(0)
    clk = clocks.stop()
    qps_dev = nq / (float(np.mean(dev_ms)) * 1e-3)
    survivors = float(np.mean(res[3]))
    # algorithmic bytes of a step: the 40-byte feature record + doc_idx + source_type of every row once (floor / mask pass),
    # the per-query row bitmaps written and read, and the vector of every row that passes some query's floor (scan passes)
    mask_bytes = n * (40 + 4 + 1) + (n // 8) * (nq + 1)
    mask_ms = float(np.mean(prep_ms))
    need = np.zeros(nq, dtype=np.uint64)
    for i in range(nq):
        for j in range(hq[i].n_phrases):
            if hq[i].phrase_jbit[j] < 0:
                need[i] |= np.uint64(1 << hq[i].phrase_bit[j])
    group = 4                                                # queries per scan pass (scan_gemv, hybrid mode)
    kind = idx.last_scan_kind()                              # "pairs_hybrid": one warp per surviving (query, row) pair
    pass_rows = 0
    elem = 2 if args.dtype == "bf16" else 4
    if kind == "pairs_hybrid":
        for i in range(nq):
            pass_rows += int(((feat["phrase_bits"][:, 0] & need[i]) == need[i]).sum())
        # per pair: the vector, the feature record + doc_idx + authority, the row id written and read, the key written and read;
        # the bitmaps are read twice more (count, fill)
        scan_bytes = pass_rows * (args.dim * elem + 40 + 4 + 1 + 8 + 16) + 2 * (n // 8) * nq
    else:
        for g0 in range(0, nq, group):
            u = np.zeros(n, dtype=bool)
            for i in range(g0, min(g0 + group, nq)):
                u |= (feat["phrase_bits"][:, 0] & need[i]) == need[i]
            pass_rows += int(u.sum())
        scan_bytes = pass_rows * args.dim * elem + 2 * (n // 8) * nq
    scan_ms_mean = float(np.mean(scan_ms))
    dominant = "scan" if scan_ms_mean >= mask_ms else "mask"
    dom_bytes, dom_ms = (scan_bytes, scan_ms_mean) if dominant == "scan" else (mask_bytes, mask_ms)
    cpu = c5_cpu_baseline(n, k) if not args.no_cpu_baseline else None
    line = {
        "metric": "QPS hybrid rerank (coverage floor + weighted signals fused with cosine), top-50, 10Mx768 corpus",
        "value": qps_dev, "unit": "queries/s", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": float(np.mean(dev_ms)), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": f"{n}x{args.dim} {args.dtype} corpus, hybrid rerank, {nq}-query bank (1-4 required phrases each), top-{k}",
                   "rows": n, "dim": args.dim, "k": k, "batch": nq, "l2": "inputs larger than L2 (no flush needed)",
                   "mean_rows_returned": survivors},
        "e2e": {"value": nq * args.steps / dt, "unit": "queries/s", "h2d_bytes_per_step": nq * args.dim * 4 + nq * C.sizeof(N.HybridQuery),
                "d2h_bytes_per_step": nq * k * 16 + nq * 4},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": dom_bytes / (dom_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                     "frac": dom_bytes / (dom_ms * 1e-3) / 1e9 / hbm_peak, "traffic": None,
                     "kernel": (("hybrid_count + hybrid_fill + hybrid_pair_score (one warp per surviving pair)" if kind == "pairs_hybrid"
                                 else "scan_gemv<hybrid> x %d passes" % ((nq + group - 1) // group)) if dominant == "scan"
                                else "hybrid_mask_kernel (+ query prep)"),
                     "ms_per_launch": dom_ms, "algorithmic_bytes_per_launch": dom_bytes, "peak_source": peak_src,
                     "note": "dominant phase of the step; both phases in `phases`",
                     "phases": {"mask": {"ms": mask_ms, "bytes": mask_bytes, "frac": mask_bytes / (mask_ms * 1e-3) / 1e9 / hbm_peak},
                                "scan": {"ms": scan_ms_mean, "bytes": scan_bytes, "frac": scan_bytes / (scan_ms_mean * 1e-3) / 1e9 / hbm_peak,
                                         "rows_scored_per_step": pass_rows}}},
        "cpu_baseline": cpu, "clocks": clk,
        "phases_ms": {"prepare+floor_mask": mask_ms, "scan": float(np.mean(scan_ms)), "total": float(np.mean(dev_ms))},
        "build_s": t_build, "feature_build_s": t_feat,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# parity of the timed batch against the CPU oracle over the whole corpus (rank 0, every N)
# ---------------------------------------------------------------------------------------------
def parity_check(args, dev, ends, passes, checks) -> dict:
    """checks: [(corpus dtype, Q cuda [B, dim], (scores, rows, counts) cuda = the GLOBAL result of the timed batch)].
    Re-generates the seeded corpus chunk by chunk on the GPU (the generator IS the corpus definition), copies each chunk
    to the host and runs the C restatement of pgvector's cosine_distance for PARITY_QUERIES queries of every check; then
    oracle.check_topk (ids and order identical up to ties within 1e-6; scores within 1e-4 relative for float4 rows,
    1e-2 for the bf16 storage mode, whose oracle sees the bf16-rounded rows)."""
    import torch
    from oracle import oracle
    from mrag_b200 import synth
    t0 = time.perf_counter()
    oracle.set_threads(os.cpu_count() or 1)
    B = int(checks[0][1].shape[0])
    # two of the random-direction half and two of the planted-neighbour half of the batch
    sel = sorted({0, B // 2 - 1, B // 2, B - 1} & set(range(B))) if B >= PARITY_QUERIES else list(range(B))
    streams = [oracle.StreamCheck(args.rows, Q[sel].cpu().numpy()) for _, Q, _ in checks]
    for first, X in synth.cuda_corpus_chunks(args.rows, args.dim, dev, seed=1234, chunk=CHUNK):
        Xf = X.cpu().numpy() if any(dt == "f32" for dt, _, _ in checks) else None
        Xb = X.to(torch.bfloat16).to(torch.float32).cpu().numpy() if any(dt == "bf16" for dt, _, _ in checks) else None
        for (dt, _, _), sc in zip(checks, streams):
            sc.feed(first, Xb if dt == "bf16" else Xf)
    mask = passes(docs_of(ends, 0, args.rows)) if (args.tag_filter or args.doc_pool or args.payer_filter) else None
    out = {"status": "ok", "queries_checked": len(sel) * len(checks), "query_indices": sel, "rows": args.rows,
           "oracle": "oracle/pgv_oracle.c (restated pgvector cosine_distance) + ORDER BY/LIMIT, tie-aware check_topk over the whole corpus",
           "rtol": {"f32": 1e-4, "bf16": 1e-2}, "checked": [dt for dt, _, _ in checks]}
    for (dt, _, res), sc in zip(checks, streams):
        s, r, c = (t.cpu().numpy() for t in res)
        for j, qi in enumerate(sel):
            try:
                sc.check(j, r[qi], s[qi], int(c[qi]), mask, args.k, rtol=1e-4 if dt == "f32" else 1e-2)
            except AssertionError as e:
                out["status"] = "FAILED"
                out.setdefault("failures", []).append(f"{dt} query {qi}: {e}")
    out["seconds"] = time.perf_counter() - t0
    return out


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    if args.workload == "c2":
        args.rows, args.dim, args.dtype, args.batch, args.k, args.tag_filter = 1_000_000, 768, "f32", 256, 10, 10
        args.sweep = ""
    elif args.workload == "c3":
        args.rows, args.dim, args.dtype, args.k = 10_000_000, 768, "bf16", 100
    elif args.workload == "c4":
        # 50M x 1536 bf16 sharded over 8 GPUs = 6.25M rows (19.2 GB) per GPU, per-payor bitset, top-10: the corpus grows with
        # the world size (8 ranks = the stated 50M rows; fewer ranks run that many shares of it)
        world_env = int(os.environ.get("WORLD_SIZE", "1"))
        args.rows, args.dim, args.dtype, args.k, args.payer_filter = 6_250_000 * world_env, 1536, "bf16", 10, 13
        if args.batch == DEFAULT_BATCH and "--batch" not in sys.argv:
            args.batch = 1
        if "--sweep" not in sys.argv:
            args.sweep = "4,64" if world_env == 1 else ""
    elif args.workload == "pool":
        # the case the reference works around (exact cosine sort over a pinned pool: 130 ms warm .. 14 s cold on Cloud SQL,
        # corpus_search.py:414-420): production-shaped corpus, one query, a 50-document pool
        args.rows, args.dim, args.dtype, args.k, args.batch, args.sweep = 1_900_000, 1536, "f32", 10, 1, ""
        args.doc_pool = args.doc_pool or 50
    elif args.workload == "c5":
        args.rows, args.dim, args.dtype, args.k, args.batch, args.sweep = 10_000_000, 768, "bf16", 50, 22, ""
    if args.also_f32 < 0:
        args.also_f32 = 1 if (args.workload == "" and args.dtype == "bf16" and args.dim == 768 and not (args.tag_filter or args.payer_filter or args.doc_pool)) else 0
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    import mrag_b200
    from mrag_b200 import index as mi
    from mrag_b200 import sharded, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the scan has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if world != args.gpus and rank == 0:
        print(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; using {world}", file=sys.stderr)

    # ---- build this rank's shard: contiguous row block cut at document boundaries, ragged documents
    ends = doc_layout(args.rows)
    n_docs_all = len(ends)
    cuts = shard_cuts(ends, args.rows, world)
    lo, hi = cuts[rank], cuts[rank + 1]
    n_local = hi - lo
    pool, passes = filter_spec(args, ends)

    def build_shard(dtype: str):
        """(index, plant, seconds): this rank's rows of the seeded corpus, appended from device chunks."""
        t0 = time.perf_counter()
        ix = mi.Index(args.dim, dtype, local_rank, max(n_local, 1))
        ix.set_row_base(lo)
        pl = None
        for first, X in synth.cuda_corpus_chunks(args.rows, args.dim, dev, seed=1234, chunk=CHUNK):
            if first == 0:
                pl = X[:4096].clone()                     # queries are planted near rows of chunk 0 on every rank
            a, b = max(first, lo), min(first + X.shape[0], hi)
            if first >= hi:
                break
            if a >= b:
                continue
            docs = docs_of(ends, a, b - a)
            meta = mi.make_meta(b - a, doc_idx=docs,
                                payer=(docs % args.payer_filter).astype(np.uint16) if args.payer_filter else None)
            ix.append_device(X[a - first:b - first], meta)
        if args.tag_filter:
            bits = np.zeros((n_docs_all, 8), dtype=np.uint64)
            bits[::args.tag_filter, 0] = 1                   # tag bit 0 on every M-th document
            ix.set_doc_tags(0, bits)
        torch.cuda.synchronize()
        assert len(ix) == n_local
        return ix, pl, time.perf_counter() - t0

    idx, plant, t_build = build_shard(args.dtype)
    flt = None
    if args.tag_filter:
        flt = mi.Filter().tag_relaxed([0])
    if pool is not None:
        flt = (flt or mi.Filter()).doc_pool(pool)
    if args.payer_filter:
        flt = (flt or mi.Filter()).payer_in([3 % args.payer_filter])
    local_pass = passes(docs_of(ends, lo, n_local)) if flt is not None else None
    n_pass_local = int(local_pass.sum()) if local_pass is not None else n_local
    tile_rows_local = n_local
    if local_pass is not None:                               # rows of the 64-row tiles that hold a passing row
        padded = np.zeros((n_local + 63) // 64 * 64, dtype=bool)
        padded[:n_local] = local_pass
        tile_rows_local = int(padded.reshape(-1, 64).any(axis=1).sum()) * 64

    elem = 2 if args.dtype == "bf16" else 4
    hbm_peak, peak_src = load_peaks()
    if args.workload == "c5":
        run_c5(args, idx, plant, dev, n_local, world, rank, hbm_peak, peak_src, t_build, ClockSampler(local_rank))
        idx.close()
        return

    # N > 1: overlap every step's exchange with the next step's scan (MRAG_PIPELINE=0: one search at a time)
    pipelined = world > 1 and os.environ.get("MRAG_PIPELINE", "1") != "0"

    class Arm:
        """One resident shard + (N > 1) its cross-rank searcher; everything measured below goes through it."""
        def __init__(self, ix, dtype):
            self.idx, self.dtype = ix, dtype
            self.ss = sharded.ShardedSearcher(index=ix, exchange=os.environ.get("MRAG_EXCHANGE", "auto")) if world > 1 else None

        def step(self, qd, k, out=None):
            if self.ss is not None:
                return self.ss.search(qd, k, flt)
            return self.idx.search_device(qd, k, flt, out=out, sync=False)

        def run(self, qd, k, steps, out=None):
            """`steps` searches back to back; returns the last result.  N > 1: two searches in flight (search_async) -- the
            exchange + k-way merge of step i runs on a side stream under the scan of step i + 1; every step's result is
            taken (the caller's stream is ordered after its exchange) before step i + 2 is issued."""
            res = None
            if self.ss is not None and pipelined:
                pend = None
                for _ in range(steps):
                    p = self.ss.search_async(qd, k, flt)
                    if pend is not None:
                        res = pend.result()
                    pend = p
                return pend.result() if pend is not None else res
            for _ in range(steps):
                res = self.step(qd, k, out)
            return res

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def measure(arm, batch: int, steps: int, warmup: int, sample_clocks: bool):
        Q = synth.cuda_queries(plant, batch, args.dim, dev, seed=4321)
        out = None
        if arm.ss is None:
            out = (torch.empty((batch, args.k), dtype=torch.float32, device=dev),
                   torch.empty((batch, args.k), dtype=torch.int64, device=dev),
                   torch.empty((batch,), dtype=torch.int32, device=dev))
        if arm.ss is not None:
            arm.ss.defer_exchange = os.environ.get("MRAG_DEFER_EXCHANGE", "1") != "0"
        res = arm.run(Q, args.k, max(warmup, 3), out)
        barrier()
        clocks = ClockSampler(local_rank) if sample_clocks else None
        if clocks:
            clocks.start()
        mi.profile_begin(steps)
        launches0 = mi.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        res = arm.run(Q, args.k, steps, out)
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1))
        launches = mi.launch_count() - launches0
        scan_ms = [x for x in mi.profile_read(1, steps) if x >= 0]
        prep_ms = [x for x in mi.profile_read(0, steps) if x >= 0]
        merge_ms = [x for x in mi.profile_read(2, steps) if x >= 0]
        mi.profile_begin(0)
        clk = clocks.stop() if clocks else None
        return {"batch": batch, "ms": ms, "steps": steps, "launches": launches, "scan_ms": scan_ms, "prep_ms": prep_ms,
                "merge_ms": merge_ms, "clocks": clk, "result": tuple(t.clone() for t in res), "Q": Q}

    def roofline_of(arm, m, batch):
        kind = arm.idx.last_scan_kind()
        group = {"gemv": 4, "gemv_shadow": 4, "mma": 64, "mma_ks": 64, "mma128": 128}[kind]      # queries per scan launch
        if kind == "mma128" and batch > 128:
            group = 256                                                             # CTA pairs: 256 queries per pass
        scan_launches = (batch + group - 1) // group
        if not m["scan_ms"]:
            return None
        per_launch_ms = float(np.mean(m["scan_ms"])) / scan_launches
        q_per_launch = min(batch, group)
        # bytes = n_pass * D * s + ceil(N/8) + N*4 (1/|x|, tensor-core paths) + q*D*4 + q*k*12, with
        # n_pass = rows passing the filter and s = the element size of the storage the kernel streams (the tensor-core
        # kernels stream bf16 rows -- for an fp32 corpus its bf16 shadow).  The tensor-core kernels fetch whole 64-row
        # tiles, so with a filter they read `tile_rows_streamed` >= n_pass rows: that shows up as frac < 1, not here.
        tensor_kind = kind in ("mma", "mma_ks", "mma128")
        scan_elem = 2 if (tensor_kind or kind == "gemv_shadow") else (2 if arm.dtype == "bf16" else 4)
        bytes_launch = (n_pass_local * args.dim * scan_elem + (n_local + 7) // 8 + (n_pass_local * 4 if tensor_kind else 0)
                        + q_per_launch * args.dim * 4 + q_per_launch * args.k * 12)
        ach = bytes_launch / (per_launch_ms * 1e-3) / 1e9
        traffic = load_traffic(f"scan_{kind}", args, n_local, q_per_launch) if arm.dtype == args.dtype else None
        out = {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": traffic,
               "kernel": f"scan_{kind}", "launches_per_step": scan_launches,
               "ms_per_launch": per_launch_ms, "algorithmic_bytes_per_launch": bytes_launch, "peak_source": peak_src,
               "rows_passing_filter": n_pass_local, "tile_rows_streamed": tile_rows_local if tensor_kind else n_pass_local}
        if tensor_kind:
            # the tensor-core view of the same launch: 2 flops per (query, row, dimension); the exact 64-query kernels
            # multiply the hi and lo halves of every query (128 tensor-memory lanes)
            lanes = 2 * q_per_launch if kind in ("mma", "mma_ks") else q_per_launch
            tf = 2.0 * lanes * n_pass_local * args.dim / (per_launch_ms * 1e-3) / 1e12
            tpeak, tsrc = load_tensor_peak()
            out["tensor"] = {"achieved": tf, "peak": tpeak, "unit": "TFLOP/s", "frac": tf / tpeak, "peak_source": tsrc}
        return out

    # ---- headline
    arm = Arm(idx, args.dtype)
    ss = arm.ss
    exchange_name = None
    if ss is not None:
        ss.search(synth.cuda_queries(plant, args.batch, args.dim, dev, seed=4321), args.k, flt)     # decides p2p vs NCCL
        exchange_name = ("peer stores inside the merge kernel (symmetric memory)" if ss.exchange != "nccl" and ss._p2p
                         else "NCCL all_gather_into_tensor + merge kernel")
    m = measure(arm, args.batch, args.steps, args.warmup, sample_clocks=True)
    qps = args.batch * args.steps / (m["ms"] * 1e-3)
    roof = roofline_of(arm, m, args.batch)
    # whole-step fractions (everything between the step's first and last kernel, not the scan launch alone)
    step_s = m["ms"] / args.steps * 1e-3
    step_view = None
    if roof:
        step_view = {"hbm_frac": roof["algorithmic_bytes_per_launch"] * roof["launches_per_step"] / step_s / 1e9 / hbm_peak}
        if "tensor" in roof:
            tpeak, _ = load_tensor_peak()
            lanes = 2 if roof["kernel"] in ("scan_mma", "scan_mma_ks") else 1
            step_view["tensor_frac"] = 2.0 * lanes * args.batch * n_pass_local * args.dim / step_s / 1e12 / tpeak

    # ---- end to end through the host-buffer API: pinned host queries in, host results out, every step
    Qh = m["Q"].cpu().pin_memory()
    hs = torch.empty((args.batch, args.k), dtype=torch.float32).pin_memory()
    hr = torch.empty((args.batch, args.k), dtype=torch.int64).pin_memory()
    hc = torch.empty((args.batch,), dtype=torch.int32).pin_memory()
    qd = torch.empty_like(m["Q"])

    def e2e_step():
        if ss is None:
            idx.search_pinned(Qh, args.k, hs, hr, hc, flt)     # H2D + scan + select + D2H + sync inside the C ABI
        else:
            qd.copy_(Qh, non_blocking=True)
            s, r, c = ss.search(qd, args.k, flt)
            hs.copy_(s, non_blocking=True); hr.copy_(r, non_blocking=True); hc.copy_(c, non_blocking=True)
            torch.cuda.current_stream().synchronize()

    def e2e_run(steps):
        if ss is None or not pipelined:
            for _ in range(steps):
                e2e_step()
            return
        # two requests in flight: H2D + scan of step i on the main stream, exchange + D2H of step i on the side stream;
        # the host takes step i - 1's results (event) before it issues step i + 1
        prev = None
        for i in range(steps):
            qd2[i & 1].copy_(Qh, non_blocking=True)
            p = ss.search_async(qd2[i & 1], args.k, flt)
            if prev is not None:                      # step i - 1: its exchange was released by issuing step i
                prev.copy_to_host(hs2[(i - 1) & 1], hr2[(i - 1) & 1], hc2[(i - 1) & 1]).synchronize()
            prev = p
        j = (steps - 1) & 1
        prev.copy_to_host(hs2[j], hr2[j], hc2[j]).synchronize()
        hs.copy_(hs2[j]); hr.copy_(hr2[j]); hc.copy_(hc2[j])
    if ss is not None and pipelined:
        qd2 = [torch.empty_like(m["Q"]) for _ in range(2)]
        hs2 = [torch.empty_like(hs).pin_memory() for _ in range(2)]
        hr2 = [torch.empty_like(hr).pin_memory() for _ in range(2)]
        hc2 = [torch.empty_like(hc).pin_memory() for _ in range(2)]
    e2e_run(3)
    barrier()
    t0 = time.perf_counter()
    e2e_run(args.steps)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_qps = args.batch * args.steps / e2e_s
    # the e2e path must return what the device path returned
    assert torch.equal(hr, m["result"][1].cpu()), "e2e result differs from device-resident result"

    # ---- where a sharded step spends its time on the path that was timed (rank 0's view)
    shard_phases = None
    if ss is not None:
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
        barrier()
        for i in range(args.steps):
            ss.search(m["Q"], args.k, flt, events=evs[i])
        barrier()
        names = ss.phase_names()
        shard_phases = {name: float(np.mean([e[j].elapsed_time(e[j + 1]) for e in evs])) for j, name in enumerate(names)}

    # ---- brief sweep over other batch sizes (N=1 only, not the headline)
    sweep = []
    if world == 1 and args.sweep:
        for b in [int(x) for x in args.sweep.split(",") if x.strip()]:
            if b == args.batch:
                continue
            mm = measure(arm, b, 5, 3, sample_clocks=False)
            rr = roofline_of(arm, mm, b)
            st_s = mm["ms"] / mm["steps"] * 1e-3
            ent = {"batch": b, "qps": b * mm["steps"] / (mm["ms"] * 1e-3), "ms_per_step": mm["ms"] / mm["steps"],
                   "scan_frac_of_hbm_peak": rr["frac"] if rr else None,
                   "scan_frac_of_tensor_peak": rr["tensor"]["frac"] if rr and "tensor" in rr else None,
                   "kernel": f"scan_{idx.last_scan_kind()}"}
            if rr and "tensor" in rr:
                lanes = 2 if rr["kernel"] in ("scan_mma", "scan_mma_ks") else 1
                ent["step_frac_of_tensor_peak"] = 2.0 * lanes * b * n_pass_local * args.dim / st_s / 1e12 / load_tensor_peak()[0]
            if rr:
                ent["step_frac_of_hbm_peak"] = rr["algorithmic_bytes_per_launch"] * rr["launches_per_step"] / st_s / 1e9 / hbm_peak
            sweep.append(ent)

    # ---- serving view: T host threads issuing single-query searches through the host-buffer C ABI, with and
    #      without request coalescing (concurrent requests share one pass over the corpus)
    concurrent = None
    if world == 1 and args.threads > 0:
        from mrag_b200 import _native as N
        Qc = synth.cuda_queries(plant, 256, args.dim, dev, seed=99).cpu().numpy()

        def serve(opts, seconds=1.0):
            done = [0] * args.threads
            stop = time.perf_counter() + seconds

            def worker(t):
                i = t
                while time.perf_counter() < stop:
                    idx.search(Qc[i % 256:i % 256 + 1], args.k, flt, options=opts)
                    done[t] += 1
                    i += args.threads
            th = [threading.Thread(target=worker, args=(t,)) for t in range(args.threads)]
            t0 = time.perf_counter()
            [x.start() for x in th]
            [x.join() for x in th]
            return sum(done) / (time.perf_counter() - t0)
        serve(0, 0.2)
        concurrent = {"threads": args.threads, "qps_independent_calls": serve(0),
                      "qps_coalesced_calls": serve(N.OPT_COALESCE) if flt is None else None,
                      "note": "single-query mrag_search calls from host threads, host buffers; coalesced = MRAG_OPT_COALESCE"}

    # ---- the same corpus stored as float4 (the reference's own precision, add_pgvector_columns.py:49), same batch
    checks = [(args.dtype, m["Q"], m["result"])]
    f32_line = None
    if args.also_f32 and args.dtype == "bf16":
        idx.close()
        arm = ss = None
        torch.cuda.empty_cache()
        idx32, _, t_build32 = build_shard("f32")
        arm32 = Arm(idx32, "f32")
        m32 = measure(arm32, args.batch, max(3, args.steps // 2), 3, sample_clocks=False)
        r32 = roofline_of(arm32, m32, args.batch)
        f32_line = {"value": args.batch * m32["steps"] / (m32["ms"] * 1e-3), "unit": "queries/s", "ms_per_step": m32["ms"] / m32["steps"],
                    "corpus_dtype": "f32", "kernel": f"scan_{idx32.last_scan_kind()}", "steps": m32["steps"],
                    "scan_frac_of_hbm_peak": r32["frac"] if r32 else None, "build_s": t_build32,
                    "result_digest": result_digest(m32["result"][1], m32["result"][2]),
                    "note": "same rows stored as float4 (30.7 GB + a 15.4 GB bf16 shadow for candidate generation); results are "
                            "exact fp32 (nominees rescored from the float4 rows, certificate, exact rescan on failure)"}
        checks.append(("f32", m32["Q"], m32["result"]))
        idx32.close()
    else:
        idx.close()

    # ---- parity: rank 0 re-checks PARITY_QUERIES queries of each timed batch against the CPU oracle (C restatement of
    #      pgvector's `<=>` + ORDER BY / LIMIT) over the WHOLE corpus, streamed chunk by chunk from the same seeded generator
    parity = None
    if rank == 0 and not args.no_parity:
        parity = parity_check(args, dev, ends, passes, checks)

    # ---- CPU baseline beside it (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            b = cpu_reference_qps(args, steps=3, warmup=1)
            cpu = {k: b[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as e:   # the GPU number stands even if the checker cannot be built here
            cpu = {"value": None, "unit": "queries/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": m["ms"] / args.steps, "higher_is_better": True,
            "scaling": "weak" if args.workload == "c4" else "strong", "vs_baseline": None, "dtype": args.dtype,
            "data": "synthetic", "config": workload_config(args, args.batch),
            "e2e": {"value": e2e_qps, "unit": "queries/s",
                    "h2d_bytes_per_step": args.batch * args.dim * 4,
                    "d2h_bytes_per_step": args.batch * args.k * 12 + args.batch * 4},
            "gpu_launches": int(m["launches"]),
            "roofline": roof, "cpu_baseline": cpu, "clocks": m["clocks"],
            "parity": parity, "result_digest": result_digest(m["result"][1], m["result"][2]),
            "phases_ms": {"prepare": float(np.mean(m["prep_ms"])) if m["prep_ms"] else None,
                          "scan": float(np.mean(m["scan_ms"])) if m["scan_ms"] else None,
                          "merge": float(np.mean(m["merge_ms"])) if m["merge_ms"] else None},
            "step_frac": step_view,
            "rows_per_gpu": n_local, "build_s": t_build, "sweep": sweep,
        }
        if f32_line:
            line["f32_corpus"] = f32_line
        if shard_phases:
            line["shard_phases_ms"] = shard_phases
        if world > 1:
            line["config"]["exchange"] = exchange_name
            line["config"]["pipeline"] = ("two searches in flight: exchange + k-way merge of step i on a side stream under the scan of step i + 1 "
                                          "(shard_phases_ms is the one-search-at-a-time latency view)") if pipelined else "one search at a time"
        if concurrent:
            line["concurrent_single_query"] = concurrent
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if parity and parity.get("status") != "ok":
        sys.exit(1)


if __name__ == "__main__":
    main()

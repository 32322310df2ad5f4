"""CPU tests of the host side: the C-ABI library loads and exports every symbol include/mrag.h
declares (no compute calls), struct layouts, vocabularies / filter builders, the plugin
boundary's error behaviour, and the shard arithmetic."""
import asyncio
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import mrag_b200
from mrag_b200 import _native as N
from mrag_b200 import sharded, synth
from mrag_b200.index import Filter, META_DTYPE
from mrag_b200.vocab import Vocab

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions() -> list[str]:
    src = open(os.path.join(ROOT, "include", "mrag.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mrag_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = N.load()
    names = header_functions()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"libmrag.so does not export {name}"
    assert sorted(N.EXPORTS) == names, "python binding list and include/mrag.h disagree"
    assert b"sm_100a" in lib.mrag_version()


def test_library_is_sm100a_cubin():
    out = subprocess.run(["cuobjdump", "-lelf", N.lib_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_struct_layouts_match_c(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include "mrag.h"\n#include <stdio.h>\n#include <stddef.h>\n'
                   'int main(){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(mrag_filter), sizeof(mrag_rowmeta),'
                   'offsetof(mrag_filter, doc_pool), offsetof(mrag_filter, tag_any), offsetof(mrag_filter, alt_state),'
                   'offsetof(mrag_rowmeta, valid));return 0;}')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert got == [C.sizeof(N.FilterStruct), C.sizeof(N.RowMeta), N.FilterStruct.doc_pool.offset,
                   N.FilterStruct.tag_any.offset, N.FilterStruct.alt_state.offset, N.RowMeta.valid.offset]
    assert META_DTYPE.itemsize == 12 and META_DTYPE.fields["valid"][1] == N.RowMeta.valid.offset
    src.write_text('#include "mrag.h"\n#include <stdio.h>\n#include <stddef.h>\n'
                   'int main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(mrag_chunkfeat), sizeof(mrag_hybrid_query),'
                   'offsetof(mrag_chunkfeat, length_score), offsetof(mrag_chunkfeat, dtags), offsetof(mrag_hybrid_query, qcat),'
                   'offsetof(mrag_hybrid_query, auth_score), offsetof(mrag_hybrid_query, contact_query),'
                   'offsetof(mrag_hybrid_query, source_type_any));return 0;}')
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    from mrag_b200.index import FEAT_DTYPE
    assert got == [C.sizeof(N.ChunkFeat), C.sizeof(N.HybridQuery), N.ChunkFeat.length_score.offset, N.ChunkFeat.dtags.offset,
                   N.HybridQuery.qcat.offset, N.HybridQuery.auth_score.offset, N.HybridQuery.contact_query.offset,
                   N.HybridQuery.source_type_any.offset]
    assert FEAT_DTYPE.itemsize == 40 and FEAT_DTYPE.fields["dtags"][1] == N.ChunkFeat.dtags.offset


def test_no_device_fails_loudly_not_silently():
    """On this CPU-only box every compute entry point must refuse (there is no CPU path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(N.MragError) as e:
        mrag_b200.Index(64, "f32", 0, 1024)
    assert e.value.code == N.MRAG_ERR_CUDA
    store = mrag_b200.B200VectorStore(dim=8)
    with pytest.raises(N.MragError):
        store.search([0.0] * 8, 3)


def test_product_does_not_import_oracle():
    """The product path must never route through the CPU oracle."""
    pkg = os.path.join(ROOT, "mobius-rag_b200")
    bad = re.compile(r"(^\s*(import|from)\s+oracle\b)|libpgv_oracle|pgv_oracle|oracle\.oracle", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not bad.search(text), f"{f} references the oracle"


def test_vocab_and_ilike():
    v = Vocab()
    codes = [v.payer.encode(p) for p in synth.PAYERS]
    assert codes == list(range(len(synth.PAYERS)))
    assert v.payer.encode(None) == N.MRAG_CODE_NONE
    assert v.payer.lookup("Nobody") == 0xFFFE
    assert v.payer.lookup("Aetna") == synth.PAYERS.index("Aetna")
    got = {synth.PAYERS[c] for c in v.payer.ilike("%molina healthcare%")}
    assert got == {"Molina Healthcare", "Molina Healthcare of Florida"}
    assert {synth.PAYERS[c] for c in v.payer.ilike("%AHCA%")} == {"AHCA", "Ahca.myflorida"}
    assert v.tag_bit("d", "x", True) == 0 and v.tag_bit("p", "x", True) == 1 and v.tag_bit("d", "x", False) == 0
    assert v.tag_bit("d", "never", False) is None
    s = Vocab().state
    for i in range(254):
        s.encode(f"S{i}")
    with pytest.raises(ValueError):
        s.encode("overflow")


def test_filter_builder_bits():
    f = Filter().payer_in([0, 65], [3], alt_state=7).state_eq(2).doc_pool([5, 9, 9]).tag_relaxed([0, 64, 511])
    s = f.s
    assert s.flags == N.F_PAYER | N.F_STATE | N.F_DOC_POOL | N.F_TAG_RELAXED
    assert s.payer_any[0] == 1 and s.payer_any[1] == 2 and s.payer_alt_any[0] == 8 and s.alt_state == 7
    assert s.n_doc_pool == 3 and s.tag_any[0] == 1 and s.tag_any[1] == 1 and s.tag_any[7] == 1 << 63
    assert Filter().active is False
    empty_pool = Filter().doc_pool([])
    assert empty_pool.s.n_doc_pool == 0 and empty_pool.active


def test_store_contract_without_gpu():
    with pytest.raises(ValueError):
        mrag_b200.B200VectorStore(table_name="users")                   # vector_store.py:163-164
    store = mrag_b200.B200VectorStore(dim=8)

    async def inside_loop():
        with pytest.raises(RuntimeError, match="asearch"):
            store.search([0.0] * 8)                                     # vector_store.py:209-216
    asyncio.run(inside_loop())
    assert mrag_b200.NoopVectorStore().search([1.0]) == []
    old = os.environ.pop("VECTOR_STORE", None)
    try:
        assert isinstance(mrag_b200.get_vector_store(), mrag_b200.NoopVectorStore)
        os.environ["VECTOR_STORE"] = "B200 "
        assert isinstance(mrag_b200.get_vector_store(), mrag_b200.B200VectorStore)
    finally:
        os.environ.pop("VECTOR_STORE", None)
        if old is not None:
            os.environ["VECTOR_STORE"] = old


def test_vector_arm_is_fail_soft_without_gpu(caplog):
    """_vector_arm swallows every exception and returns [] (corpus_search.py:1561-1563)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")

    class Broken:
        class index:
            dim = 4
        def filter_corpus(self, *a):
            raise RuntimeError("boom")
        vocab = Vocab()
    assert mrag_b200.vector_arm(Broken(), [0.0] * 4, 5, None, None) == []
    assert asyncio.run(mrag_b200._vector_arm(Broken(), [0.0] * 4, 5, None, None, search_id="s1")) == []


def test_shard_bounds_are_doc_aligned():
    _, _, info = synth.make_metadata(10000, seed=3, rows_per_doc=64)
    doc = info["doc_of_row"]
    for world in (1, 2, 3, 4, 8):
        b = sharded.shard_bounds(doc, world)
        assert b[0][0] == 0 and b[-1][1] == 10000 and len(b) == world
        for (lo, hi), (lo2, _) in zip(b, b[1:]):
            assert hi == lo2 and lo <= hi
            if 0 < hi < 10000:
                assert doc[hi - 1] != doc[hi], "cut inside a document"
        sizes = [hi - lo for lo, hi in b]
        assert max(sizes) - min(sizes) <= 2 * 64 * 8
    # degenerate: fewer documents than ranks
    b = sharded.shard_bounds(np.zeros(10, dtype=np.uint32), 4)
    assert b[0] == (0, 10) or sum(hi - lo for lo, hi in b) == 10


def test_packed_layout_alignment():
    for nq, k in [(1, 10), (3, 7), (64, 100), (5, 1)]:
        lay = sharded.packed_layout(nq, k)
        assert lay["size"] % 8 == 0 and lay["scores_off"] % 4 == 0 and lay["counts_off"] % 4 == 0
        assert lay["size"] >= nq * k * 12 + nq * 4


def test_synth_is_deterministic_and_shaped():
    X1, v1 = synth.make_corpus(2000, 48, seed=9, null_frac=1e-2)
    X2, v2 = synth.make_corpus(2000, 48, seed=9, null_frac=1e-2)
    assert (X1 == X2).all() and (v1 == v2).all() and v1.sum() == 1980
    assert (np.abs(X1).sum(axis=1) == 0).sum() >= 1
    # duplicates exist
    _, inv, cnt = np.unique(X1, axis=0, return_inverse=True, return_counts=True)
    assert (cnt > 1).any()
    meta, tags, info = synth.make_metadata(2000, seed=1, valid=v1)
    assert (np.diff(meta["doc_idx"].astype(np.int64)) >= 0).all()
    assert tags.shape == (info["n_docs"], N.MRAG_TAG_WORDS)

"""Host columns of the table (mobius-rag_b200/columns.py) and the bulk insert path -- CPU only."""
import numpy as np
import pytest

import mrag_b200  # noqa: F401
from mrag_b200.columns import CodeCol, IntCol, JsonCol, StrCol


def test_strcol_values_nulls_growth_and_round_trip(tmp_path):
    c = StrCol()
    vals = [None if i % 7 == 0 else ("" if i % 11 == 0 else f"väl-{i}" * (i % 5)) for i in range(5000)]
    c.extend(vals[:3])
    c.extend(vals[3:4000])
    c.extend(iter(vals[4000:]))
    assert len(c) == 5000
    assert [c[i] for i in (0, 1, 7, 11, 77, 4999, -1)] == [vals[i] for i in (0, 1, 7, 11, 77, 4999, -1)]
    assert c[11] == "" and c[7] is None                    # '' and NULL stay distinct
    with pytest.raises(IndexError):
        c[5000]
    np.savez(tmp_path / "c.npz", **c.arrays("x"))
    d = StrCol.from_arrays(np.load(tmp_path / "c.npz", allow_pickle=False), "x")
    assert len(d) == 5000 and all(d[i] == vals[i] for i in range(0, 5000, 37))
    d.extend(["tail"])
    assert d[5000] == "tail"
    c.truncate(10)
    assert len(c) == 10 and c[9] == vals[9]
    c.extend(["again"])
    assert c[10] == "again"


def test_intcol_jsoncol_codecol():
    i = IntCol()
    i.extend([1, None, 3])
    i.extend(np.arange(2000))
    assert (i[0], i[1], i[2], i[3], i[2002]) == (1, None, 3, 0, 1999)
    j = JsonCol()
    j.extend([{"a.b": 1, "c": 2}, None, {}, [1, 2]])
    assert j[0] == {"a.b": 1, "c": 2} and list(j[0]) == ["a.b", "c"] and j[1] is None and j[2] == {} and j[3] == [1, 2]
    values = ["x", "y"]
    c = CodeCol(values, none_code=0xFF, dtype=np.uint8)
    c.extend([0, 1, 0xFF])
    values.append("z")                                     # the vocabulary grows behind the column
    c.extend(np.full(3000, 2))
    assert (c[0], c[1], c[2], c[3], c[-1]) == ("x", "y", None, "z", "z") and len(c) == 3003

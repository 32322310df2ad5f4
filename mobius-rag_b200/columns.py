"""Columnar host storage for the hydration half of ``rag_published_embeddings``.

The reference hydrates a result row in the same SELECT that ranks it (`_BM25_COLS`, corpus_search.py:621-640).
Here the ranking happens on the GPU and returns row numbers; the strings those rows carry live in these columns:
append-only numpy buffers (amortised doubling), one per table column, so that a 10M-50M row corpus is a handful of
arrays -- not tens of millions of Python objects -- and a snapshot is ``np.savez`` of those arrays (no pickle).

  StrCol    nullable variable-length UTF-8 strings: offsets int64[n+1] + bytes uint8[...] + null bool[n]
  IntCol    nullable integers: int64[n] + null bool[n]
  JsonCol   nullable JSON values (the JSONB columns chunk_d_tags / chunk_p_tags / chunk_j_tags) over a StrCol
  CodeCol   dictionary-coded strings: integer codes [n] + the vocabulary that decodes them

Every column supports ``len``, ``col[i]`` (Python value or None), ``extend(values)``, ``truncate(n)`` and
``arrays()`` / ``from_arrays()`` for the snapshot.
"""
from __future__ import annotations

import json
from typing import Any, Iterable, Sequence

import numpy as np


def _grow(a: np.ndarray, need: int) -> np.ndarray:
    if need <= a.shape[0]:
        return a
    cap = max(need, a.shape[0] * 2, 1024)
    b = np.zeros(cap, dtype=a.dtype)
    b[:a.shape[0]] = a
    return b


class StrCol:
    def __init__(self):
        self.off = np.zeros(1025, dtype=np.int64)
        self.buf = np.zeros(1 << 14, dtype=np.uint8)
        self.null = np.zeros(1024, dtype=bool)
        self.n = 0

    def __len__(self) -> int:
        return self.n

    def __getitem__(self, i: int):
        i = int(i)
        if i < 0:
            i += self.n
        if not 0 <= i < self.n:
            raise IndexError(i)
        if self.null[i]:
            return None
        return self.buf[self.off[i]:self.off[i + 1]].tobytes().decode("utf-8")

    def extend(self, values: Iterable[Any]) -> None:
        enc = [None if v is None else (v if isinstance(v, bytes) else str(v).encode("utf-8")) for v in values]
        m = len(enc)
        if m == 0:
            return
        lens = np.fromiter((0 if e is None else len(e) for e in enc), dtype=np.int64, count=m)
        null = np.fromiter((e is None for e in enc), dtype=bool, count=m)
        blob = b"".join(e for e in enc if e is not None)
        self.extend_raw(lens, np.frombuffer(blob, dtype=np.uint8), null)

    def extend_raw(self, lens: np.ndarray, data: np.ndarray, null: np.ndarray | None = None) -> None:
        """Bulk append: per-value byte lengths, the concatenated bytes, and the null flags."""
        m = int(lens.shape[0])
        used = int(self.off[self.n])
        self.off = _grow(self.off, self.n + m + 1)
        self.null = _grow(self.null, self.n + m)
        self.buf = _grow(self.buf, used + int(data.shape[0]))
        self.off[self.n + 1:self.n + m + 1] = used + np.cumsum(lens)
        self.buf[used:used + data.shape[0]] = data
        self.null[self.n:self.n + m] = False if null is None else null
        self.n += m

    def truncate(self, n: int) -> None:
        self.n = min(self.n, int(n))

    def arrays(self, prefix: str) -> dict[str, np.ndarray]:
        used = int(self.off[self.n])
        return {prefix + ".off": self.off[:self.n + 1].copy(), prefix + ".buf": self.buf[:used].copy(),
                prefix + ".null": self.null[:self.n].copy()}

    @classmethod
    def from_arrays(cls, z, prefix: str) -> "StrCol":
        c = cls()
        off, buf, null = z[prefix + ".off"], z[prefix + ".buf"], z[prefix + ".null"]
        c.n = int(null.shape[0])
        c.off = np.ascontiguousarray(off, dtype=np.int64)
        c.buf = np.ascontiguousarray(buf, dtype=np.uint8)
        c.null = np.ascontiguousarray(null, dtype=bool)
        return c


class IntCol:
    def __init__(self):
        self.val = np.zeros(1024, dtype=np.int64)
        self.null = np.zeros(1024, dtype=bool)
        self.n = 0

    def __len__(self) -> int:
        return self.n

    def __getitem__(self, i: int):
        i = int(i)
        if i < 0:
            i += self.n
        if not 0 <= i < self.n:
            raise IndexError(i)
        return None if self.null[i] else int(self.val[i])

    def extend(self, values: Iterable[Any]) -> None:
        if isinstance(values, np.ndarray) and values.dtype.kind in "iu":
            vals, null = values.astype(np.int64), np.zeros(values.shape[0], dtype=bool)
        else:
            values = list(values)
            null = np.fromiter((v is None for v in values), dtype=bool, count=len(values))
            vals = np.fromiter((0 if v is None else int(v) for v in values), dtype=np.int64, count=len(values))
        m = vals.shape[0]
        self.val = _grow(self.val, self.n + m)
        self.null = _grow(self.null, self.n + m)
        self.val[self.n:self.n + m] = vals
        self.null[self.n:self.n + m] = null
        self.n += m

    def truncate(self, n: int) -> None:
        self.n = min(self.n, int(n))

    def arrays(self, prefix: str) -> dict[str, np.ndarray]:
        return {prefix + ".val": self.val[:self.n].copy(), prefix + ".null": self.null[:self.n].copy()}

    @classmethod
    def from_arrays(cls, z, prefix: str) -> "IntCol":
        c = cls()
        c.val = np.ascontiguousarray(z[prefix + ".val"], dtype=np.int64)
        c.null = np.ascontiguousarray(z[prefix + ".null"], dtype=bool)
        c.n = int(c.null.shape[0])
        return c


class JsonCol:
    """JSONB column: the serialised value per row (None = SQL NULL)."""

    def __init__(self, s: StrCol | None = None):
        self.s = s or StrCol()

    def __len__(self) -> int:
        return len(self.s)

    def __getitem__(self, i: int):
        raw = self.s[i]
        return None if raw is None else json.loads(raw)

    def extend(self, values: Iterable[Any]) -> None:
        self.s.extend(None if v is None else json.dumps(v, separators=(",", ":")) for v in values)

    def truncate(self, n: int) -> None:
        self.s.truncate(n)

    def arrays(self, prefix: str) -> dict[str, np.ndarray]:
        return self.s.arrays(prefix)

    @classmethod
    def from_arrays(cls, z, prefix: str) -> "JsonCol":
        return cls(StrCol.from_arrays(z, prefix))


class CodeCol:
    """Dictionary-coded string column: ``codes[i]`` decoded through ``values`` (``none_code`` = SQL NULL)."""

    def __init__(self, values: list, none_code: int, dtype):
        self.values, self.none_code = values, int(none_code)
        self.codes = np.zeros(1024, dtype=dtype)
        self.n = 0

    def __len__(self) -> int:
        return self.n

    def __getitem__(self, i: int):
        i = int(i)
        if i < 0:
            i += self.n
        if not 0 <= i < self.n:
            raise IndexError(i)
        c = int(self.codes[i])
        return None if c == self.none_code else self.values[c]

    def extend(self, codes: Sequence[int]) -> None:
        codes = np.asarray(codes, dtype=self.codes.dtype)
        m = codes.shape[0]
        self.codes = _grow(self.codes, self.n + m)
        self.codes[self.n:self.n + m] = codes
        self.n += m

    def truncate(self, n: int) -> None:
        self.n = min(self.n, int(n))

    def view(self) -> np.ndarray:
        return self.codes[:self.n]

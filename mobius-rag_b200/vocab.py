"""Dictionary coding of the string columns the WHERE clauses test.

The table denormalises document_payer / document_state / document_program /
document_authority_level / source_type on every row (app/models.py:242-280); the GPU sees one
small integer per column per row.  String equality becomes code equality; ``ILIKE '%v%'``
(corpus_search.py:1489-1494) becomes "code in the set of vocabulary entries that match", which
the host evaluates over the (tiny) vocabulary, not over rows.
"""
from __future__ import annotations

import re

from . import _native as N


class ColumnVocab:
    """value <-> code for one column.  None (SQL NULL) gets ``none_code`` and matches nothing;
    ``unknown_code`` stands for "a value no row has"."""

    def __init__(self, name: str, max_codes: int, none_code: int, unknown_code: int):
        self.name, self.max_codes, self.none_code, self.unknown_code = name, max_codes, none_code, unknown_code
        self.values: list[str] = []
        self._code: dict[str, int] = {}

    def encode(self, value: str | None) -> int:
        """Code for a value being WRITTEN (allocates)."""
        if value is None:
            return self.none_code
        c = self._code.get(value)
        if c is None:
            if len(self.values) >= self.max_codes:
                raise ValueError(f"too many distinct values in column {self.name!r} (max {self.max_codes})")
            c = len(self.values)
            self.values.append(value)
            self._code[value] = c
        return c

    def lookup(self, value: str) -> int:
        """Code for a value being SEARCHED (never allocates)."""
        return self._code.get(value, self.unknown_code)

    def ilike(self, pattern: str) -> list[int]:
        """Codes whose value matches Postgres ``ILIKE pattern`` (% = any run, _ = any one char)."""
        rx = re.compile("".join(".*" if ch == "%" else "." if ch == "_" else re.escape(ch) for ch in pattern),
                        re.IGNORECASE | re.DOTALL)
        return [c for c, v in enumerate(self.values) if rx.fullmatch(v) is not None]


class Vocab:
    def __init__(self):
        self.payer = ColumnVocab("document_payer", N.MRAG_PAYER_WORDS * 64 - 2, N.MRAG_CODE_NONE, 0xFFFE)
        self.state = ColumnVocab("document_state", 254, 0xFF, 0xFE)
        self.program = ColumnVocab("document_program", 254, 0xFF, 0xFE)
        self.authority = ColumnVocab("document_authority_level", 254, 0xFF, 0xFE)
        self.source_type = ColumnVocab("source_type", 254, 0xFF, 0xFE)
        # document_tags.d_tags / p_tags keys -> bit index (d and p keys are tested against their
        # own column, corpus_search.py:1501-1508, so they get separate bits)
        self._tag_bit: dict[tuple[str, str], int] = {}
        self._jtag_bit: dict[str, int] = {}

    def jtag_bit(self, key: str, allocate: bool) -> int | None:
        """bit of a document j-tag key (document_tags.j_tags, app/models.py:535-537) in the j-tag sets"""
        b = self._jtag_bit.get(key)
        if b is None and allocate:
            if len(self._jtag_bit) >= N.MRAG_JTAG_WORDS * 64:
                raise ValueError(f"more than {N.MRAG_JTAG_WORDS * 64} distinct document j-tag keys")
            b = len(self._jtag_bit)
            self._jtag_bit[key] = b
        return b

    def tag_bit(self, kind: str, key: str, allocate: bool) -> int | None:
        b = self._tag_bit.get((kind, key))
        if b is None and allocate:
            if len(self._tag_bit) >= N.MRAG_TAG_WORDS * 64:
                raise ValueError(f"more than {N.MRAG_TAG_WORDS * 64} distinct document tag keys")
            b = len(self._tag_bit)
            self._tag_bit[(kind, key)] = b
        return b

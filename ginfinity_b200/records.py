"""RNA input records for the B200 encoder path.

Host-side mirror of the reference input contract
(reference: src/ginfinity/_validation.py:91-258).  The behaviour that the
encoder path depends on is kept: ``T`` normalises to ``U``, the alphabet is
``ACGU`` / ``.()``, brackets must balance, at most 4096 nt, optional
0-based half-open ``[start, end)`` window.  Error type and message
fragments match the reference so its tests read the same against this
package.  Nothing here runs on the device.
"""
from __future__ import annotations

import numbers
from typing import Iterable, Mapping, Optional

MAXIMUM_LENGTH_NT = 4096
_BASES = frozenset("ACGU")
_BRACKETS = frozenset(".()")


class InputValidationError(ValueError):
    """An RNA record is outside the supported input contract."""


def _integer(value, label: str) -> int:
    if isinstance(value, bool) or not isinstance(value, numbers.Integral):
        raise InputValidationError(f"{label} must be an integer")
    return int(value)


def parse_position_list(value, *, name: str) -> list[int]:
    """``None``/blank -> ``[]``; int -> ``[int]``; ``"1, 2"`` -> ``[1, 2]``.

    Reference: _validation.py:19-45.
    """
    if value is None:
        return []
    if isinstance(value, bool):
        raise InputValidationError(f"{name} must be an integer")
    if isinstance(value, numbers.Integral):
        return [int(value)]
    if isinstance(value, float) and value.is_integer():
        return [int(value)]
    if not isinstance(value, str):
        raise InputValidationError(f"{name} must be an integer or integer list")
    stripped = value.strip()
    if stripped == "":
        return []
    out = []
    for token in stripped.split(","):
        token = token.strip()
        if token == "":
            raise InputValidationError(f"empty value in {name} list")
        try:
            out.append(int(token, 10))
        except ValueError as exc:
            raise InputValidationError(
                f"invalid integer {token!r} in {name}") from exc
    return out


def parse_slice_bounds(start, end) -> list[tuple[int, int]]:
    """Parallel start/end lists -> list of half-open windows."""
    lo = parse_position_list(start, name="start")
    hi = parse_position_list(end, name="end")
    if len(lo) != len(hi):
        raise InputValidationError(
            f"start has {len(lo)} value(s) but end has {len(hi)}")
    return [(a, b) for a, b in zip(lo, hi)]


def sliced_identifier(identifier: str, start: int, end: int) -> str:
    return f"{identifier}:{start}-{end}"


def check_column_names(*columns: Optional[str]) -> tuple[str, ...]:
    given = tuple(c for c in columns if c)
    if len(set(given)) != len(given):
        raise ValueError("RNA column names must differ")
    return given


def validate_and_normalize(sequence: str, structure: str, *,
                           maximum_length: int = MAXIMUM_LENGTH_NT
                           ) -> tuple[str, str]:
    """Normalise and check one sequence/structure pair.

    Reference: _validation.py:224-258 (same order of checks, so the first
    error reported for a bad record is the same).
    """
    seq = sequence.strip().upper().replace("T", "U")
    dbn = structure.strip()
    if seq == "":
        raise InputValidationError("empty sequence")
    if len(seq) > maximum_length:
        raise InputValidationError(
            f"sequence length {len(seq)} exceeds maximum {maximum_length}")
    if len(dbn) != len(seq):
        raise InputValidationError(
            f"structure is {len(dbn)} characters against a "
            f"{len(seq)} nt sequence")
    bad = sorted(set(seq) - _BASES)
    if bad:
        raise InputValidationError(
            "unsupported sequence character(s): " + " ".join(bad))
    bad = sorted(set(dbn) - _BRACKETS)
    if bad:
        raise InputValidationError(
            "unsupported structure character(s): " + " ".join(bad))
    depth = 0
    first_open_at_depth: list[int] = []
    for position, symbol in enumerate(dbn):
        if symbol == "(":
            first_open_at_depth.append(position)
            depth += 1
        elif symbol == ")":
            if depth == 0:
                raise InputValidationError(
                    f"unmatched ')' at 0-based position {position}")
            depth -= 1
            first_open_at_depth.pop()
    if depth:
        raise InputValidationError(
            f"unmatched '(' at 0-based position {first_open_at_depth[0]}")
    return seq, dbn


class RNA:
    """One RNA: identifier, sequence, dot-bracket structure, optional window.

    Immutable after construction (reference: _validation.py:91-150 is a
    frozen dataclass; attribute assignment raises here as well).
    """

    __slots__ = ("identifier", "sequence", "structure", "start", "end")

    def __init__(self, identifier: str, sequence: str, structure: str,
                 start: Optional[int] = None, end: Optional[int] = None):
        ident = identifier.strip()
        seq, dbn = validate_and_normalize(sequence, structure)
        if ident == "":
            raise InputValidationError("empty identifier")
        if any(c in ident for c in "\t\r\n"):
            raise InputValidationError(
                "identifier must not contain tabs or line breaks")
        if (start is None) != (end is None):
            raise InputValidationError("start and end must both be provided")
        if start is not None:
            start = _integer(start, "start")
            end = _integer(end, "end")
            if start < 0 or end > len(seq) or start >= end:
                raise InputValidationError(
                    f"invalid slice [{start}, {end}) for a {len(seq)} nt sequence")
        set_ = object.__setattr__
        set_(self, "identifier", ident)
        set_(self, "sequence", seq)
        set_(self, "structure", dbn)
        set_(self, "start", start)
        set_(self, "end", end)

    def __setattr__(self, key, value):
        raise AttributeError(f"cannot assign to field {key!r}")

    def __delattr__(self, key):
        raise AttributeError(f"cannot delete field {key!r}")

    def _astuple(self):
        return (self.identifier, self.sequence, self.structure,
                self.start, self.end)

    def __eq__(self, other):
        return isinstance(other, RNA) and self._astuple() == other._astuple()

    def __hash__(self):
        return hash(self._astuple())

    def __repr__(self):
        return ("RNA(identifier={!r}, sequence={!r}, structure={!r}, "
                "start={!r}, end={!r})".format(*self._astuple()))

    @property
    def length(self) -> int:
        return len(self.sequence)

    @property
    def sliced(self) -> bool:
        return self.start is not None

    @property
    def core_length(self) -> int:
        return self.length if self.start is None else self.end - self.start

    # -- mapping constructors (reference: _validation.py:152-221) ----------
    @classmethod
    def many_from_mapping(cls, row: Mapping[str, object], *,
                          identifier_column: str = "transcript_id",
                          sequence_column: str = "sequence",
                          structure_column: str = "secondary_structure",
                          start_column: Optional[str] = "start",
                          end_column: Optional[str] = "end",
                          suffix_identifier: bool = True) -> list["RNA"]:
        if (start_column is None) != (end_column is None):
            raise ValueError("start and end columns must both be provided")
        check_column_names(identifier_column, sequence_column,
                           structure_column, start_column, end_column)
        needed = (identifier_column, sequence_column, structure_column)
        absent = [c for c in needed if c not in row]
        if absent:
            raise InputValidationError(
                "missing RNA column(s): " + ", ".join(absent))
        ident, seq, dbn = (row[c] for c in needed)
        if not (isinstance(ident, str) and isinstance(seq, str)
                and isinstance(dbn, str)):
            raise InputValidationError(
                "RNA identifier, sequence, and structure must be strings")
        windows: list[tuple[int, int]] = []
        if start_column is not None and (start_column in row
                                         or end_column in row):
            absent = [c for c in (start_column, end_column) if c not in row]
            if absent:
                raise InputValidationError(
                    "missing RNA column(s): " + ", ".join(absent))
            windows = parse_slice_bounds(row[start_column], row[end_column])
        if not windows:
            return [cls(ident, seq, dbn)]
        suffix = suffix_identifier or len(windows) > 1
        return [cls(sliced_identifier(ident, a, b) if suffix else ident,
                    seq, dbn, start=a, end=b) for a, b in windows]

    @classmethod
    def from_mapping(cls, row: Mapping[str, object], *,
                     identifier_column: str = "transcript_id",
                     sequence_column: str = "sequence",
                     structure_column: str = "secondary_structure",
                     start_column: Optional[str] = None,
                     end_column: Optional[str] = None) -> "RNA":
        found = cls.many_from_mapping(
            row, identifier_column=identifier_column,
            sequence_column=sequence_column,
            structure_column=structure_column,
            start_column=start_column, end_column=end_column,
            suffix_identifier=False)
        if len(found) != 1:
            raise InputValidationError(
                "mapping defines multiple slices; use RNA.many_from_mapping()")
        return found[0]


def read_rna_table(path, *, identifier_column: str = "transcript_id",
                   sequence_column: str = "sequence",
                   structure_column: str = "secondary_structure",
                   start_column: Optional[str] = "start",
                   end_column: Optional[str] = "end",
                   delimiter: str = "\t") -> list[RNA]:
    """Delimited table -> validated records in file order.

    Reference: src/ginfinity/table.py:84-114.  Host I/O only; it exists so
    BASELINE config 1 (rouskin_sample_6k.tsv) can be fed to the encoder.
    """
    import csv
    from pathlib import Path

    if len(delimiter) != 1:
        raise ValueError("delimiter must be exactly one character")
    if (start_column is None) != (end_column is None):
        raise ValueError("start and end columns must both be provided")
    needed = check_column_names(identifier_column, sequence_column,
                                structure_column)
    check_column_names(identifier_column, sequence_column, structure_column,
                       start_column, end_column)
    path = Path(path)
    records: list[RNA] = []
    seen: set[str] = set()
    try:
        with path.open(newline="") as handle:
            reader = csv.DictReader(handle, delimiter=delimiter)
            header = reader.fieldnames
            if header is None:
                raise ValueError(f"empty RNA table: {path}")
            if len(set(header)) != len(header):
                raise ValueError(f"duplicate column name in RNA table: {path}")
            absent = [c for c in needed if c not in header]
            if absent:
                raise ValueError(f"RNA table {path} is missing column(s): "
                                 + ", ".join(absent))
            if start_column:
                have = [c in header for c in (start_column, end_column)]
                if any(have) and not all(have):
                    absent = [c for c in (start_column, end_column)
                              if c not in header]
                    raise ValueError(
                        f"RNA table {path} is missing column(s): "
                        + ", ".join(absent))
                if not any(have):
                    start_column = end_column = None
            for row in reader:
                where = f"RNA table {path} line {reader.line_num}"
                if None in row:
                    raise ValueError(f"{where} has extra fields")
                try:
                    expanded = RNA.many_from_mapping(
                        row, identifier_column=identifier_column,
                        sequence_column=sequence_column,
                        structure_column=structure_column,
                        start_column=start_column, end_column=end_column,
                        suffix_identifier=True)
                except InputValidationError as exc:
                    raise InputValidationError(f"{where}: {exc}") from exc
                for record in expanded:
                    if record.identifier in seen:
                        raise InputValidationError(
                            f"{where}: duplicate identifier "
                            f"{record.identifier!r}")
                    seen.add(record.identifier)
                    records.append(record)
    except UnicodeDecodeError as exc:
        raise ValueError(f"RNA table is not valid text: {path}") from exc
    if not records:
        raise ValueError(f"RNA table contains no records: {path}")
    return records


def iter_lengths(records: Iterable[RNA]) -> list[int]:
    return [r.length for r in records]

"""The reference's input-contract tests (tests/test_table.py, tests/test_validation.py) restated
against ginfinity_b200: same calls, same expected values, same message substrings.  Host only."""
import pytest

from ginfinity_b200 import InputValidationError, RNA, read_rna_table


# ---- tests/test_validation.py ---------------------------------------------------------------
def test_record_normalizes_t_and_whitespace():
    record = RNA(" id ", " acgt ", " (()) ")
    assert (record.identifier, record.sequence, record.structure) == ("id", "ACGU", "(())")


@pytest.mark.parametrize(("sequence", "structure", "message"), [
    ("", "", "empty sequence"),
    ("ACGN", "....", "unsupported sequence"),
    ("ACGU", "...", "characters against"),
    ("ACGU", "[..]", "unsupported structure"),
    ("ACGU", "((.)", "unmatched"),
])
def test_bad_input_is_rejected(sequence, structure, message):
    with pytest.raises(InputValidationError, match=message):
        RNA("id", sequence, structure)


def test_bad_identifier_is_rejected():
    with pytest.raises(InputValidationError, match="identifier"):
        RNA("bad\tid", "ACGU", "....")


@pytest.mark.parametrize(("start", "end", "message"), [
    (1, None, "both be provided"),
    (None, 3, "both be provided"),
    (-1, 2, "invalid slice"),
    (2, 2, "invalid slice"),
    (3, 2, "invalid slice"),
    (0, 5, "invalid slice"),
])
def test_invalid_slice_is_rejected(start, end, message):
    with pytest.raises(InputValidationError, match=message):
        RNA("id", "ACGU", "....", start=start, end=end)


def test_valid_slice_is_stored():
    record = RNA("id", "ACGUACGU", "((....))", start=2, end=6)
    assert (record.start, record.end, record.core_length) == (2, 6, 4)
    assert record.sliced


# ---- tests/test_table.py: the same cases, table-driven ----------------------------------------
HEAD3 = "transcript_id\tsequence\tsecondary_structure\n"
HEAD5 = "transcript_id\tsequence\tsecondary_structure\tstart\tend\n"
WINDOW_ROW = {"transcript_id": "rna-1", "sequence": "ACGUACGU", "secondary_structure": "((....))"}


def test_from_mapping_column_names_and_single_windows():
    custom = dict(identifier_column="name", sequence_column="bases", structure_column="dot_bracket")
    got = RNA.from_mapping({"name": "rna-1", "bases": "ACGT", "dot_bracket": "(())"}, **custom)
    assert got == RNA("rna-1", "ACGU", "(())")
    one = RNA.from_mapping({**WINDOW_ROW, "start": "2", "end": "6"}, start_column="start", end_column="end")
    assert (one.identifier, one.start, one.end) == ("rna-1", 2, 6)      # a single window keeps the name
    with pytest.raises(InputValidationError, match="multiple slices"):
        RNA.from_mapping({**WINDOW_ROW, "start": "2,4", "end": "6,8"}, start_column="start", end_column="end")


GOOD_TABLES = [
    # text, reader options, expected (identifier, structure, start, end) per record
    ("dot_bracket,source,rna_name,bases\n(()),example,first,ACGU\n....,example,second,GGAA\n",
     dict(identifier_column="rna_name", sequence_column="bases", structure_column="dot_bracket", delimiter=","),
     [("first", "(())", None, None), ("second", "....", None, None)]),
    (HEAD5 + "rna-1\tACGUACGU\t((....))\t2,4\t6, 8\n", {},
     [("rna-1:2-6", "((....))", 2, 6), ("rna-1:4-8", "((....))", 4, 8)]),
    (HEAD3 + "rna-1\tACGU\t(())\n", {}, [("rna-1", "(())", None, None)]),
]


@pytest.mark.parametrize("text,options,expected", GOOD_TABLES)
def test_tables_that_load(tmp_path, text, options, expected):
    path = tmp_path / "table.txt"
    path.write_text(text)
    records = read_rna_table(path, **options)
    assert [(r.identifier, r.structure, r.start, r.end) for r in records] == expected
    assert len({r.sequence for r in records}) <= len(records)


BAD_TABLES = [
    # text, reader options, exception, message fragment
    ("name\tbases\nfirst\tACGU\n",
     dict(identifier_column="name", sequence_column="bases", structure_column="dot_bracket"), ValueError, "dot_bracket"),
    (HEAD3 + "bad\tACGN\t....\n", {}, InputValidationError, "line 2"),
    (HEAD5 + "rna-1\tACGUACGU\t((....))\t2,4\t6\n", {}, InputValidationError, "start has 2"),
    (HEAD3 + "a\tACGU\t....\na\tGGAA\t....\n", {}, InputValidationError, "duplicate"),
    (HEAD3, {}, ValueError, "no records"),
    ("", {}, ValueError, "empty RNA table"),
    (HEAD3 + "a\tACGU\t....\tsurplus\n", {}, ValueError, "extra fields"),
    (HEAD3 + "a\tACGU\t....\n", dict(delimiter="::"), ValueError, "exactly one character"),
    (HEAD3 + "a\tACGU\t....\n", dict(start_column="start", end_column=None), ValueError, "both be provided"),
]


@pytest.mark.parametrize("text,options,error,fragment", BAD_TABLES)
def test_tables_that_are_refused(tmp_path, text, options, error, fragment):
    """The reference's three refusal tests plus the error paths of src/ginfinity/table.py:29-81 that
    its tests leave implicit (duplicates, empty files, surplus fields, reader options)."""
    path = tmp_path / "table.txt"
    path.write_text(text)
    with pytest.raises(error, match=fragment):
        read_rna_table(path, **options)

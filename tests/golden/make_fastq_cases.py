"""Writes tests/golden/fastq_cases.json: hand-written FASTQ texts with the answer a reader gives for each (sequence
line offsets / lengths, or the first fault).  The reference has no FASTQ parser (README.md:160-180 only shows the
caller's loop), so these cases -- not reference outputs -- pin the format definition the oracle and the CUDA path share.
The expected values are written by hand below, NOT computed, so that both restatements are checked against them.
Run: python tests/golden/make_fastq_cases.py"""
import json
from pathlib import Path

CASES = [
    # (name, text, expected)   expected = {"reads": [[start, len], ...]} or {"fault": [record, kind]}
    ("one_record", "@r\nACGT\n+\nIIII\n", {"reads": [[3, 4]]}),
    ("no_final_newline", "@r\nACGT\n+\nIIII", {"reads": [[3, 4]]}),
    ("two_records_plus_repeats_name", "@a\nAC\n+a\n!!\n@b desc\nGGT\n+\n###\n", {"reads": [[3, 2], [20, 3]]}),
    ("crlf", "@a\r\nACG\r\n+\r\nIII\r\n", {"reads": [[4, 3]]}),
    ("crlf_no_final_newline", "@a\r\nACG\r\n+\r\nIII", {"reads": [[4, 3]]}),
    ("empty_sequence", "@a\n\n+\n\n@b\nA\n+\nI\n", {"reads": [[3, 0], [10, 1]]}),
    ("quality_starts_with_at", "@a\nAC\n+\n@@\n@b\nG\n+\n+\n", {"reads": [[3, 2], [14, 1]]}),
    ("lower_case_and_n", "@a\nacgtn\n+\nIIIII\n", {"reads": [[3, 5]]}),
    ("empty_text", "", {"reads": []}),
    ("bad_header", "a\nAC\n+\nII\n", {"fault": [0, 1]}),
    ("bad_second_header", "@a\nAC\n+\nII\nb\nAC\n+\nII\n", {"fault": [1, 1]}),
    ("bad_separator", "@a\nAC\n-\nII\n", {"fault": [0, 2]}),
    ("quality_too_short", "@a\nACG\n+\nII\n", {"fault": [0, 3]}),
    ("quality_too_long", "@a\nAC\n+\nIII\n@b\nA\n+\nI\n", {"fault": [0, 3]}),
    ("truncated_after_separator", "@a\nAC\n+\n", {"fault": [0, 4]}),
    ("truncated_after_sequence", "@a\nAC\n+\nII\n@b\nAC\n", {"fault": [1, 4]}),
    ("truncated_header_only", "@a\nAC\n+\nII\n@b", {"fault": [1, 4]}),
    ("trailing_blank_line", "@a\nAC\n+\nII\n\n", {"fault": [1, 1]}),
    ("only_newline", "\n", {"fault": [0, 1]}),
    ("separator_fault_in_partial_record", "@a\nAC\n+\nII\n@b\nAC\nx\n", {"fault": [1, 2]}),
    ("earlier_record_wins", "@a\nAC\n+\nI\n@b\nAC\n-\nII\n", {"fault": [0, 3]}),
]

FASTA_CASES = [   # one sequence line per record
    ("fa_one_record", ">r\nACGT\n", {"reads": [[3, 4]]}),
    ("fa_no_final_newline", ">r desc\nACGT", {"reads": [[8, 4]]}),
    ("fa_crlf_and_empty", ">a\r\nACG\r\n>b\r\n\r\n>c\r\nT\r\n", {"reads": [[4, 3], [13, 0], [19, 1]]}),
    ("fa_lower_case_and_n", ">a\nacgtn\n", {"reads": [[3, 5]]}),
    ("fa_empty_text", "", {"reads": []}),
    ("fa_bad_header", "r\nACGT\n", {"fault": [0, 1]}),
    ("fa_wrapped_sequence_is_not_this_format", ">a\nACGT\nACGT\n>b\nAC\n", {"fault": [1, 1]}),
    ("fa_truncated", ">a\nAC\n>b\n", {"fault": [1, 4]}),
    ("fa_header_only", ">a", {"fault": [0, 4]}),
    ("fa_fastq_is_not_fasta", "@a\nAC\n+\nII\n", {"fault": [0, 1]}),
]

if __name__ == "__main__":
    out = [{"name": n, "text": t, **e} for n, t, e in CASES] + [{"name": n, "text": t, "fasta": True, **e} for n, t, e in FASTA_CASES]
    Path(__file__).with_name("fastq_cases.json").write_text(json.dumps(out, indent=1) + "\n")
    print(len(out), "cases")

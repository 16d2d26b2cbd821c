"""The product's host FASTQ / FASTA reader (katome_b200/csrc/host_reader.h) on the CPU, through the test
hook ktg_host_parse_file: against a plain-Python restatement of the rust-bio 0.10 record semantics the
reference relies on (builder.rs:118-165) and against the oracle's own reader (accepted reads / bytes of
files that hold only ACGT reads; tests/build.rs:27-28 pins those totals for the fixtures)."""
import ctypes as C
import os

import numpy as np
import pytest

from katome_b200 import _lib
from oracle import oracle as O
from tests import helpers as H

WS = b"\n\r \t\v\f"


def host_parse(path, fasta=False, batch_bytes=0):
    n, b, h = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
    rc = _lib.lib().ktg_host_parse_file(os.fsencode(str(path)), 1 if fasta else 0, batch_bytes, C.byref(n), C.byref(b), C.byref(h))
    return rc, n.value, b.value, h.value


def fnv(seqs):
    h = 0xcbf29ce484222325
    for s in seqs:
        for c in s + b"\n":
            h = ((h ^ c) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    return h


def py_fastq(data: bytes):
    """4-line records, '@' header, sequence line trimmed on the right, '+' line ignored, a record that
    ends before its quality line is an error, an empty line where a header should be ends the file"""
    lines = data.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()  # the text after the last newline, if empty, is not a line
    lines = [ln + b"\n" for ln in lines]
    seqs, i = [], 0
    while i < len(lines):
        if lines[i] == b"":
            break
        if not lines[i].startswith(b"@"):
            return None
        seqs.append(lines[i + 1].rstrip(WS) if i + 1 < len(lines) else b"")
        if i + 3 >= len(lines):
            return None
        i += 4
    return seqs


def py_fasta(data: bytes):
    lines = data.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    seqs, i = [], 0
    while i < len(lines):
        if not lines[i].startswith(b">"):
            return None
        i += 1
        s = b""
        while i < len(lines) and not lines[i].startswith(b">"):
            s += lines[i].rstrip(WS)
            i += 1
        seqs.append(s)
    return seqs


@pytest.mark.parametrize("name", ["data1", "data2", "data3"])
def test_fixtures_fastq_and_fasta(tmp_path, name):
    seqs = [s.encode() for s in H.golden_seqs(name)]
    fq = H.write_fastq(tmp_path / "a.fastq", H.golden_seqs(name))
    rc, n, b, h = host_parse(fq)
    assert (rc, n, b, h) == (0, len(seqs), sum(map(len, seqs)), fnv(seqs))
    assert py_fastq(open(fq, "rb").read()) == seqs
    fa = H.write_fasta(tmp_path / "a.fasta", H.golden_seqs(name), width=37)  # wrapped lines with trailing blanks
    rc, n, b, h = host_parse(fa, fasta=True)
    assert (rc, n, b, h) == (0, len(seqs), sum(map(len, seqs)), fnv(seqs))
    assert py_fasta(open(fa, "rb").read()) == seqs
    # small internal batches: records never straddle or get lost at a batch boundary
    assert host_parse(fq, batch_bytes=1000)[1:] == (len(seqs), sum(map(len, seqs)), fnv(seqs))
    assert host_parse(fa, fasta=True, batch_bytes=777)[1:] == (len(seqs), sum(map(len, seqs)), fnv(seqs))


def test_agrees_with_the_oracle_reader_on_acgt_only_files(tmp_path):
    rng = np.random.default_rng(11)
    seqs = H.random_reads(rng, 500, 50, 180)
    fq = H.write_fastq(tmp_path / "r.fastq", seqs)
    g, nbytes = O.OracleGIR.create(40, [fq], "fastq", False)
    rc, n, b, _ = host_parse(fq)
    assert rc == 0 and (n, b) == (g.accepted_reads, g.accepted_bytes) == (len(seqs), nbytes)
    fa = H.write_fasta(tmp_path / "r.fasta", seqs)
    g, nbytes = O.OracleGIR.create(40, [fa], "fasta", False)
    rc, n, b, _ = host_parse(fa, fasta=True)
    assert rc == 0 and (n, b) == (g.accepted_reads, g.accepted_bytes)


@pytest.mark.parametrize("text,ok", [
    (b"@r\nACGT\n+\nIIII\n", True),
    (b"@r\nACGT\n+\nIIII", True),                      # no newline at the end of the file
    (b"@r\r\nACGT  \r\n+\r\nIIII\r\n", True),           # CRLF and trailing blanks are trimmed
    (b"@r\nACGT\n+\nII\n@s\nGG\n+\nIIIIIII\n", True),   # quality length is not checked (data_too_short_read.txt)
    (b"@r\n\n+\n\n", True),                            # an empty sequence is a record
    (b"", True),
    (b"r\nACGT\n+\nIIII\n", False),                    # header without '@'
    (b"@r\nACGT\n+\n", False),                         # record ends before its quality line
    (b"@r\nACGT\n", False),
    (b"@r\nACGT\n+\nIIII\n@s\n", False),
])
def test_fastq_record_semantics(tmp_path, text, ok):
    p = tmp_path / "t.fastq"
    p.write_bytes(text)
    rc, n, b, h = host_parse(p)
    exp = py_fastq(text)
    assert (exp is not None) == ok
    if ok:
        assert (rc, n, b, h) == (0, len(exp), sum(map(len, exp)), fnv(exp))
    else:
        assert rc == _lib.KTG_ERR_BAD_RECORD
    # the oracle's reader takes the same decision
    try:
        g, _ = O.OracleGIR.create(3, [str(p)], "fastq", False)
        oracle_ok = True
    except O.OracleError as e:
        oracle_ok = e.code != O.KO_ERR_BAD_RECORD
    assert oracle_ok == ok


def test_file_errors(tmp_path):
    assert host_parse(tmp_path / "missing.fastq")[0] == _lib.KTG_ERR_IO   # builder.rs:57-77
    assert host_parse(tmp_path)[0] == _lib.KTG_ERR_IO                     # a directory
    p = tmp_path / "x.fasta"
    p.write_bytes(b"ACGT\n>r\nAC\n")
    assert host_parse(p, fasta=True)[0] == _lib.KTG_ERR_BAD_RECORD        # no '>' at the start

"""Host-only tests of the CLI's FASTA/Q reader (reference: bio/seqio/fastx as used by search.go
S:793-1000): `kmcp-gpu parse` runs the reader alone — no GPU — and prints id, length and CRC-32 of every record."""
import gzip
import os
import random
import subprocess
import zlib

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "kmcp_b200", "kmcp-gpu")


def _parse(args):
    if not os.path.exists(EXE):
        pytest.fail("kmcp_b200/kmcp-gpu is not built: run __graft_entry__.build()")
    p = subprocess.run([EXE, "parse"] + args, capture_output=True, timeout=120)
    assert p.returncode == 0, p.stderr.decode()
    return p.stdout.decode()


def _line(i, s):
    return "%s\t%d\t%08x\n" % (i.decode(), len(s), zlib.crc32(s))


@pytest.fixture(scope="module")
def files(tmp_path_factory):
    d = tmp_path_factory.mktemp("reads")
    rnd = random.Random(3)

    def seq(n):
        return bytes(rnd.choice(b"ACGTNacgt") for _ in range(n))

    recs = [(b"r%d" % i, seq(rnd.choice([0, 1, 30, 150, 151, 1000, 70000 if i % 997 == 0 else 150]))) for i in range(6000)]
    p = {"recs": recs, "fq_gz": str(d / "a.fq.gz"), "fq2_gz": str(d / "b.fq.gz"), "fa": str(d / "c.fa"), "fq_plain": str(d / "d.fq")}
    with gzip.open(p["fq_gz"], "wb") as f:
        for i, s in recs:
            f.write(b"@" + i + b" some description\n" + s + b"\n+\n" + b"I" * len(s) + b"\n")
    with gzip.open(p["fq2_gz"], "wb") as f:                      # CRLF line ends, tab after the ID, fewer records than mate 1
        for i, s in recs[:4000]:
            f.write(b"@" + i + b"/2\tx\r\n" + s[::-1] + b"\r\n+\r\n" + b"I" * len(s) + b"\r\n")
    with open(p["fa"], "wb") as f:                               # multi-line FASTA with blank lines, no final newline
        body = []
        for i, s in recs[:1500]:
            body.append(b">" + i + b" d\n" + b"\n".join(s[j:j + 60] for j in range(0, len(s), 60)) + b"\n")
        f.write(b"\n".join(body).rstrip(b"\n"))
    with open(p["fq_plain"], "wb") as f:                         # multi-line FASTQ whose quality lines start with '@' and '+'
        for i, s in recs[:500]:
            q = (b"@+" * len(s))[:len(s)]
            f.write(b"@" + i + b"\n" + b"\n".join(s[j:j + 70] for j in range(0, len(s), 70)) + b"\n+" + i + b"\n" +
                    b"\n".join(q[j:j + 70] for j in range(0, len(q), 70)) + b"\n")
    return p


@pytest.mark.parametrize("ahead", [[], ["--ahead"]])
def test_reader_single_and_several_files(files, ahead):
    recs = files["recs"]
    assert _parse(ahead + [files["fq_gz"]]) == "".join(_line(i, s) for i, s in recs)
    exp = "".join(_line(i, s) for i, s in recs[:1500]) + "".join(_line(i, s) for i, s in recs[:500]) + "".join(_line(i, s) for i, s in recs)
    assert _parse(ahead + [files["fa"], files["fq_plain"], files["fq_gz"]]) == exp


@pytest.mark.parametrize("ahead", [[], ["--ahead"]])
def test_reader_paired_files_stop_at_the_shorter_mate(files, ahead):
    recs = files["recs"]
    exp = ""
    for i, s in recs[:4000]:
        exp += _line(i, s) + _line(i + b"/2", s[::-1])
    exp += _line(*recs[4000])                                    # mate 1 was read before mate 2 ran out (S:806-867)
    assert _parse(ahead + ["-1", files["fq_gz"], "-2", files["fq2_gz"]]) == exp

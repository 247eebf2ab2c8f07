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


def _batches(args):
    """records of `kmcp-gpu parse --batches`: [(id, [(len, crc), ...])], and the batch sizes"""
    out = _parse(["--batches"] + args)
    recs, sizes = [], []
    for l in out.splitlines():
        if l.startswith("# batch of "):
            sizes.append(int(l.split()[-1]))
            continue
        f = l.split("\t")
        recs.append((f[0], [(int(f[i]), f[i + 1]) for i in range(1, len(f), 2)]))
    return recs, sizes


def _py_records(data):
    """what the reader must return for a FASTA/Q byte string (bio/seqio/fastx semantics as used by search.go: ID up to the
    first blank, sequence lines joined, FASTQ quality read until it is as long as the sequence)"""
    lines = [l.rstrip(b"\r\n") for l in data.split(b"\n")]
    if lines and lines[-1] == b"" and data.endswith(b"\n"):
        lines.pop()
    out, i = [], 0
    while i < len(lines):
        if lines[i] == b"":
            i += 1
            continue
        h = lines[i]
        assert h[:1] in (b">", b"@")
        rid = h[1:].split(b" ")[0].split(b"\t")[0]
        i += 1
        if h[:1] == b"@":
            seq = lines[i] if i < len(lines) else b""
            i += 1
            while i < len(lines) and not lines[i].startswith(b"+"):
                seq += lines[i]
                i += 1
            i += 1                                    # the '+' line
            got = 0
            while got < len(seq) and i < len(lines):
                got += len(lines[i])
                i += 1
        else:
            seq = b""
            while i < len(lines) and not lines[i].startswith(b">"):
                seq += lines[i]
                i += 1
        out.append((rid, seq))
    return out


def test_reader_random_mixtures_of_record_styles(tmp_path):
    """four-line records (the table-driven fast path), wrapped FASTQ, blank lines, CRLF, '@'/'+' at the start of quality
    lines, a missing final newline — in random order, so the fast path and the general reader hand over to each other at
    every kind of boundary; plain, gzip and chunk-parallel gzip input"""
    rnd = random.Random(17)

    def seq(n):
        return bytes(rnd.choice(b"ACGTN") for _ in range(n))

    for trial in range(12):
        eol = b"\r\n" if trial % 4 == 3 else b"\n"
        parts = []
        n = rnd.choice([1, 5, 300, 3000])
        for i in range(n):
            s = seq(rnd.choice([0, 1, 20, 150, 151, 400]))
            rid = b"x%d_%d" % (trial, i)
            style = rnd.choice(["four", "four", "four", "four", "wrapped", "blank", "atqual"])
            q = bytes(rnd.choice(b"FI#5") for _ in range(len(s)))
            if style == "four" or (style != "blank" and len(s) == 0):
                parts.append(b"@" + rid + rnd.choice([b"", b" desc", b"\tdesc 2"]) + eol + s + eol + b"+" + eol + q + eol)
            elif style == "atqual":
                q = (b"@" + q[1:]) if rnd.random() < 0.5 else (b"+" + q[1:])
                parts.append(b"@" + rid + eol + s + eol + b"+" + rid + eol + q + eol)
            elif style == "blank":
                parts.append(eol + b"@" + rid + eol + s + eol + b"+" + eol + q + eol + (eol if rnd.random() < 0.5 else b""))
            else:
                w = rnd.choice([7, 60])
                sl = [s[j:j + w] for j in range(0, len(s), w)]
                ql = [q[j:j + w] for j in range(0, len(q), w)]
                parts.append(b"@" + rid + eol + eol.join(sl) + eol + b"+" + eol + eol.join(ql) + eol)
        data = b"".join(parts)
        if trial % 3 == 1:
            data = data.rstrip(b"\r\n")
        exp = "".join(_line(i, s) for i, s in _py_records(data))
        p = str(tmp_path / ("t%d.fq" % trial))
        open(p, "wb").write(data)
        assert _parse([p]) == exp, trial
        assert _parse(["--ahead", p]) == exp, trial
        with gzip.open(p + ".gz", "wb") as f:
            f.write(data)
        assert _parse([p + ".gz"]) == exp, trial
        assert _parse(["--ahead", "--inflate-threads", "3", "--inflate-chunk", "65536", p + ".gz"]) == exp, trial
        # the batch builder with several parser threads per file: four-line pieces in parallel, anything else by the general reader
        want = [(i.decode(), [(len(s), "%08x" % zlib.crc32(s))]) for i, s in _py_records(data)]
        for extra in (["--parse-threads", "3"], ["--parse-threads", "2", "--inflate-threads", "2", "--inflate-chunk", "65536"]):
            assert _batches(extra + [p + ".gz"])[0] == want, (trial, extra)
        assert _batches(["--parse-threads", "3", "--batch-reads", "100", p])[0] == want, trial
        for piece in ("64", "1000", "20000"):                      # pieces far smaller than the files: cuts at record boundaries everywhere
            assert _batches(["--parse-threads", "3", "--parse-piece", piece, p])[0] == want, (trial, piece)
    # FASTA with wrapped lines between FASTQ files, empty file, file of blank lines only
    fa = str(tmp_path / "g.fa")
    recs = [(b"c%d" % i, seq(rnd.choice([0, 59, 60, 61, 5000]))) for i in range(50)]
    open(fa, "wb").write(b"".join(b">" + i + b" x\n" + b"".join(s[j:j + 60] + b"\n" for j in range(0, len(s), 60)) for i, s in recs))
    empty, blank = str(tmp_path / "e.fq"), str(tmp_path / "b.fq")
    open(empty, "wb").close()
    open(blank, "wb").write(b"\n\n\n")
    assert _parse([empty, fa, blank, str(tmp_path / "t0.fq")]) == "".join(_line(i, s) for i, s in recs) + \
        "".join(_line(i, s) for i, s in _py_records(open(str(tmp_path / "t0.fq"), "rb").read()))


def test_batch_builder_single_paired_and_whole_file(files, tmp_path):
    """the search command's own batch builder (S:793-1000) behind `parse --batches`: a parser thread per input file, blocks of
    records appended to the batches (single-end) or zipped pair by pair (paired-end, ends with the shorter file), -g whole files"""
    recs = files["recs"]
    crc = lambda s: "%08x" % zlib.crc32(s)
    # single-end, three files, batches of 1000 reads: every record once, in order, batches full except the last
    got, sizes = _batches(["--batch-reads", "1000", files["fa"], files["fq_plain"], files["fq_gz"]])
    exp = [(i.decode(), [(len(s), crc(s))]) for i, s in recs[:1500] + recs[:500] + recs]
    assert got == exp
    assert sum(sizes) == len(exp) and all(x == 1000 for x in sizes[:-1]) and 0 < sizes[-1] <= 1000
    for piece in ("8388608", "50000", "3000"):
        got2, sizes2 = _batches(["--parse-threads", "3", "--parse-piece", piece, "--batch-reads", "1000", files["fa"], files["fq_plain"], files["fq_gz"]])
        assert got2 == exp and sizes2 == sizes, piece                        # FASTA and wrapped FASTQ fall back, the .gz file is cut into pieces
    for extra in ([], ["--inflate-threads", "3", "--inflate-chunk", "65536"], ["--parse-threads", "4"], ["--parse-threads", "3", "--parse-piece", "30000"]):
        got, sizes = _batches(extra + ["--batch-reads", "777", "-1", files["fq_gz"], "-2", files["fq2_gz"]])
        assert got == [(i.decode(), [(len(s), crc(s)), (len(s), crc(s[::-1]))]) for i, s in recs[:4000]]
        assert all(x == 777 for x in sizes[:-1]) and sum(sizes) == 4000
    # default batch size: one batch
    got, sizes = _batches([files["fq_gz"]])
    assert sizes == [len(recs)] and got[-1][0] == recs[-1][0].decode()
    # -g: a file is ONE query; as in the reference (S:899-913) the run of k-1 = 20 'N' FOLLOWS every record from the second on
    got, sizes = _batches(["-g", files["fa"], files["fq_plain"]])
    joined = [rs[0][1] + b"".join(s + b"N" * 20 for _, s in rs[1:]) for rs in (recs[:1500], recs[:500])]
    assert got == [(recs[0][0].decode(), [(len(j), crc(j))]) for j in joined] and sizes == [2]
    # files without records
    e = str(tmp_path / "e.fq")
    open(e, "wb").close()
    got, sizes = _batches([e, files["fq_plain"], e])
    assert len(got) == 500 and sizes == [500]
    assert _batches([e]) == ([], [])
    assert _batches(["-1", e, "-2", files["fq_gz"]]) == ([], [])


def test_reader_abi_from_python(files, tmp_path):
    """kmcpg_reader_* (include/kmcp_gpu.h), the reader stage behind the C ABI as a Go / C host would call it: batches in input
    order with running query numbers, paired input, -g, errors reported through kmcpg_reader_error, early close"""
    import numpy as np
    from kmcp_b200 import api
    recs = files["recs"]

    def flat(gen, step=1):
        out, sizes, nxt = [], [], 0
        for first, ids, seq, off in gen:
            assert first == nxt and len(off) == step * len(ids) + 1
            nxt += len(ids)
            sizes.append(len(ids))
            for q, i in enumerate(ids):
                out.append((i, [seq[int(off[q * step + m]):int(off[q * step + m + 1])].tobytes() for m in range(step)]))
        return out, sizes

    got, sizes = flat(api.read_batches([files["fa"], files["fq_plain"], files["fq_gz"]], batch_reads=1000))
    assert got == [(i, [s]) for i, s in recs[:1500] + recs[:500] + recs]
    assert all(x == 1000 for x in sizes[:-1]) and sum(sizes) == 8000
    for kw in (dict(), dict(inflate_threads=3, inflate_chunk=65536, parse_threads=3, parse_piece=50000)):
        got, sizes = flat(api.read_batches(read1=files["fq_gz"], read2=files["fq2_gz"], batch_reads=777, **kw), step=2)
        assert got == [(i, [s, s[::-1]]) for i, s in recs[:4000]]
    got, _ = flat(api.read_batches([files["fa"]], whole_file=1, k=31, query_id="genome"))
    assert got == [(b"genome", [recs[0][1] + b"".join(s + b"N" * 30 for _, s in recs[1:1500])])]
    got, _ = flat(api.read_batches([files["fa"]], whole_file=1, use_filename=1))
    assert got[0][0] == b"c"                                   # c.fa with the extension cut
    # errors: missing file, a file that is not FASTA/Q, only one of the two mates
    for bad in ([str(tmp_path / "missing.fq")], [files["fq_plain"], str(tmp_path / "missing.fq")]):
        with pytest.raises(api.KmcpGpuError) as e:
            list(api.read_batches(bad))
        assert "no such file" in str(e.value)
    junk = str(tmp_path / "junk.txt")
    open(junk, "wb").write(b"this is not a sequence file\n" * 10)
    with pytest.raises(api.KmcpGpuError) as e:
        list(api.read_batches([junk]))
    assert "invalid FASTA/Q record" in str(e.value)
    with pytest.raises(api.KmcpGpuError):
        list(api.read_batches(read1=files["fq_gz"]))
    with pytest.raises(api.KmcpGpuError):
        list(api.read_batches([]))
    # a damaged stream ends with the decoder's error, never with a short result
    blob = bytearray(open(files["fq_gz"], "rb").read())
    blob[len(blob) // 2] ^= 0x55
    dmg = str(tmp_path / "damaged.fq.gz")
    open(dmg, "wb").write(bytes(blob))
    n = 0
    with pytest.raises(api.KmcpGpuError) as e:
        for first, ids, seq, off in api.read_batches([dmg], batch_reads=500):
            n += len(ids)
    assert "read error" in str(e.value) and "CRC-32" in str(e.value) and n < len(recs)
    # closing a reader that still has most of its input in front of it does not hang
    g = api.read_batches([files["fq_gz"], files["fq_gz"], files["fq_gz"]], batch_reads=100)
    next(g)
    g.close()
    # the CLI reports reader errors the way the reference's checkError does: a message and a non-zero exit code
    p = subprocess.run([EXE, "parse", "--batches", junk], capture_output=True, timeout=60)
    assert p.returncode != 0 and b"invalid FASTA/Q record" in p.stderr


def test_reader_abi_from_plain_c(files, tmp_path):
    """examples/read_host.c: the reader stage from a strict C99 program (what cgo binds), no device involved"""
    exe = str(tmp_path / "read_host")
    lib_dir = os.path.join(ROOT, "kmcp_b200")
    c = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", os.path.join(ROOT, "examples", "read_host.c"),
                        "-I" + os.path.join(ROOT, "include"), "-L" + lib_dir, "-lkmcp_gpu", "-Wl,-rpath," + lib_dir, "-o", exe], capture_output=True)
    assert c.returncode == 0, c.stderr.decode()
    recs = files["recs"]
    p = subprocess.run([exe, files["fq_gz"]], capture_output=True, timeout=120)
    assert p.returncode == 0, p.stderr.decode()
    lines = p.stdout.decode().splitlines()
    assert lines[-1] == "queries: %d" % len(recs) and len(lines) == 7
    assert lines[0] == "batch first=0 queries=1000 bytes=%d id=r0 len=%d" % (sum(len(s) for _, s in recs[:1000]), len(recs[0][1]))
    assert lines[3].startswith("batch first=3000 queries=1000 ") and " id=r3000 " in lines[3]
    p = subprocess.run([exe, "-p", files["fq_gz"], files["fq2_gz"]], capture_output=True, timeout=120)
    assert p.returncode == 0 and p.stdout.decode().splitlines()[-1] == "queries: 4000"
    assert " id=r0 len=%d,%d" % (len(recs[0][1]), len(recs[0][1])) in p.stdout.decode().splitlines()[0]
    p = subprocess.run([exe, str(tmp_path / "nothing.fq")], capture_output=True, timeout=120)
    assert p.returncode == 3 and b"no such file" in p.stderr


def test_reader_xz_bzip2_zstd_inputs(files, tmp_path):
    """the reference's xopen also reads .xz / .zst / .bz2 by their magic bytes: here the system's decompressor writes into a pipe;
    a failing or missing decompressor is a read error, never a short result"""
    import bz2
    import lzma
    import shutil
    recs = files["recs"]
    text = b"".join(b"@" + i + b"\n" + s + b"\n+\n" + b"F" * len(s) + b"\n" for i, s in recs)
    exp = "".join(_line(i, s) for i, s in recs)
    made = []
    if shutil.which("xz"):
        p = str(tmp_path / "r.fq.xz")
        open(p, "wb").write(lzma.compress(text, preset=1))
        made.append(p)
    if shutil.which("bzip2"):
        p = str(tmp_path / "r.fastq.bz2")
        open(p, "wb").write(bz2.compress(text, 1))
        made.append(p)
    if not made:
        pytest.skip("neither xz nor bzip2 installed")
    for p in made:
        assert _parse([p]) == exp
        assert _parse(["--ahead", p]) == exp
        got, _ = _batches(["--parse-threads", "3", "--parse-piece", "100000", p])       # a pipe: the plain parser thread
        assert len(got) == len(recs)
        blob = open(p, "rb").read()
        open(p + ".cut", "wb").write(blob[:len(blob) // 2])
        q = subprocess.run([EXE, "parse", p + ".cut"], capture_output=True, timeout=60)
        assert q.returncode != 0 and b"could not decompress" in q.stderr
    z = str(tmp_path / "r.fq.zst")
    open(z, "wb").write(b"\x28\xb5\x2f\xfd" + b"not really zstd" * 10)
    q = subprocess.run([EXE, "parse", z], capture_output=True, timeout=60)
    assert q.returncode != 0 and (b"zstd is not installed" in q.stderr or b"could not decompress" in q.stderr)


def test_search_command_plumbing_without_a_device(files, tmp_path):
    """kmcp-gpu search --dry-run: the command's reader → engine thread → writer pipeline with no device and no database (every
    query comes out unmatched): IDs, lengths and running query numbers in input order over many batches, plain and .gz output
    (asynchronous gzip members), paired input, the trailer lines (S:1023-1025)"""
    recs = files["recs"]
    hdr = "#query\tqLen\tqKmers\tFPR\thits\ttarget\tchunkIdx\tchunks\ttLen\tkSize\tmKmers\tqCov\ttCov\tjacc\tqueryIdx\n"
    tail = "# input queries: %d\n# matched queries: 0\n# matched percentage: 0.0000%%\n"

    def row(i, ln, idx):
        return "%s\t%d\t0\t0\t0\t\t-1\t0\t0\t21\t0\t0\t0\t0\t%d\n" % (i.decode(), ln, idx)

    def run(args, out):
        p = subprocess.run([EXE, "search", "--dry-run", "-q", "-K", "-o", out] + args, capture_output=True, timeout=300)
        assert p.returncode == 0, p.stderr.decode()
        return gzip.open(out, "rb").read().decode() if out.endswith(".gz") else open(out).read()

    exp = hdr + "".join(row(i, len(s), n) for n, (i, s) in enumerate(recs)) + tail % len(recs)
    assert run([files["fq_gz"]], str(tmp_path / "a.tsv")) == exp
    assert run(["--batch-reads", "500", files["fq_gz"]], str(tmp_path / "b.tsv")) == exp
    assert run(["--batch-reads", "333", "--inflate-threads", "3", "--parse-threads", "2", files["fq_gz"]], str(tmp_path / "c.tsv.gz")) == exp
    # several files: query numbers run on
    many = recs[:1500] + recs[:500] + recs
    assert run(["--batch-reads", "1000", files["fa"], files["fq_plain"], files["fq_gz"]], str(tmp_path / "d.tsv.gz")) == \
        hdr + "".join(row(i, len(s), n) for n, (i, s) in enumerate(many)) + tail % len(many)
    # paired: the query length is the sum of the mates, the pairs end with the shorter file
    exp_pe = hdr + "".join(row(i, 2 * len(s), n) for n, (i, s) in enumerate(recs[:4000])) + tail % 4000
    assert run(["--batch-reads", "700", "-1", files["fq_gz"], "-2", files["fq2_gz"]], str(tmp_path / "e.tsv")) == exp_pe
    # without -K nothing but the header and the trailer; -H drops the header
    p = subprocess.run([EXE, "search", "--dry-run", "-q", "-H", "-o", str(tmp_path / "f.tsv"), files["fq_plain"]], capture_output=True, timeout=300)
    assert p.returncode == 0 and open(str(tmp_path / "f.tsv")).read() == tail % 500
    # stdin
    with open(files["fq_plain"], "rb") as f:
        p = subprocess.run([EXE, "search", "--dry-run", "-q", "-K", "-o", "-"], stdin=f, capture_output=True, timeout=300)
    assert p.returncode == 0 and p.stdout.decode() == hdr + "".join(row(i, len(s), n) for n, (i, s) in enumerate(recs[:500])) + tail % 500
    # -i/--infile-list: one file per line
    lst = str(tmp_path / "files.txt")
    open(lst, "w").write(files["fq_plain"] + "\n\n" + files["fa"] + "\n")
    got = run(["-i", lst], str(tmp_path / "h.tsv"))
    both = recs[:500] + recs[:1500]
    assert got == hdr + "".join(row(i, len(s), n) for n, (i, s) in enumerate(both)) + tail % len(both)
    # a missing input file ends the command with the reference's kind of message and a non-zero code
    p = subprocess.run([EXE, "search", "--dry-run", "-q", "-o", str(tmp_path / "g.tsv"), str(tmp_path / "nothing.fq")], capture_output=True, timeout=300)
    assert p.returncode != 0 and b"no such file" in p.stderr


@pytest.mark.reference_data
def test_search_command_has_every_flag_of_the_reference():
    """every flag `kmcp search` declares (search.go init(): name, shorthand, default) is offered by `kmcp-gpu search` with the same
    name, shorthand and default — read from the reference's source, so it only runs where /root/reference exists"""
    import re
    src = "/root/reference/kmcp/cmd/search.go"
    if not os.path.exists(src):
        pytest.skip("no reference checkout here")
    go = open(src).read()
    live = "\n".join(l for l in go.splitlines() if not l.strip().startswith("//"))          # two flags are commented out in the reference
    flags = re.findall(r'searchCmd\.Flags\(\)\.(\w+)P\("([\w-]+)",\s*"(\w?)",\s*([^,]+),', live)
    assert len(flags) >= 20
    p = subprocess.run([EXE, "search", "--help"], capture_output=True, timeout=60)
    helptext = (p.stdout + p.stderr).decode()
    root = open("/root/reference/kmcp/cmd/root.go").read()
    live_root = "\n".join(l for l in root.splitlines() if not l.strip().startswith("//"))
    for name, short in re.findall(r'RootCmd\.PersistentFlags\(\)\.\w+P\("([\w-]+)",\s*"(\w?)"', live_root):     # -j, -q, -i, --log
        assert "--" + name in helptext and (not short or "-" + short + "," in helptext), name
    for kind, name, short, default in flags:
        assert "--" + name in helptext, name
        if short:
            assert re.search(r"-%s, --%s\b|--%s\b.*-%s\b|-%s, --\S+ / --%s" % (short, name, name, short, short, name), helptext) or ("-" + short + ",") in helptext, (name, short)
        default = default.strip().strip('"')
        if kind in ("Int", "Float64") and default not in ("0", "0.0"):
            m = re.search(r"--%s\b[^\n]*\(default ([^)]+)\)" % re.escape(name), helptext)
            assert m and float(m.group(1)) == float(default), (name, default, m and m.group(1))


@pytest.mark.parametrize("threads", [1, 4])
def test_bytes_behind_the_last_gzip_member_are_a_read_error(tmp_path, threads):
    """a multi-member .gz that was damaged or overwritten behind a member boundary: the reference reads through Go's multistream gzip
    reader (xopen), which fails with "gzip: invalid header"; the reader stage reports a read error too instead of searching a silently
    shorter input (the stand-alone `kmcp-gpu gunzip` stays as lenient as gzread, tests/test_fastgz.py)"""
    import gzip
    from kmcp_b200 import api
    rec = b"".join(b"@r%d\n%s\n+\n%s\n" % (i, b"ACGT" * 30, b"I" * 120) for i in range(30000))
    good = gzip.compress(rec[:len(rec) // 2], 6) + gzip.compress(rec[len(rec) // 2:], 6)
    ok, bad, pad = str(tmp_path / "ok.fq.gz"), str(tmp_path / "bad.fq.gz"), str(tmp_path / "pad.fq.gz")
    open(ok, "wb").write(good)
    open(bad, "wb").write(good + b"overwritten tail that is not a gzip member")
    open(pad, "wb").write(good + b"\0" * 512)
    kw = dict(inflate_threads=threads, inflate_chunk=65536)
    assert sum(len(ids) for _f, ids, _s, _o in api.read_batches([ok], **kw)) == 30000
    for p in (bad, pad):
        with pytest.raises(api.KmcpGpuError) as e:
            for _ in api.read_batches([p], **kw):
                pass
        assert "gzip: invalid header" in str(e.value)


def test_tsv_number_formatting_is_what_printf_prints():
    """kmcp_b200/csrc/tsv_format.h (the hand-written %.4f / %.4e / %d of the result table) against snprintf: uniform values, ratios of
    small integers, every kind of rounding tie of the fourth decimal and its neighbours, tiny and large values"""
    for seed in (1, 2, 3):
        p = subprocess.run([EXE, "fmt-selftest", "1500000", str(seed)], capture_output=True, timeout=300)
        assert p.returncode == 0, p.stdout.decode() + p.stderr.decode()
        assert p.stdout.decode().startswith("0 mismatches")

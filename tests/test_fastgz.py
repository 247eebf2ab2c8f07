"""Host-only tests of kmcp_b200/csrc/fastgz.h and pargz.h, the gzip decoders of the CLI's read ingest (the
reference reads its inputs through xopen + pgzip, search.go S:793-1000).  `kmcp-gpu gunzip` runs the decoder alone.
zlib (through Python) is the checker: every stream zlib can produce must decode to the same bytes, and every damaged
stream must be refused — never accepted with different bytes, never crash (a sanitizer build of the decoder runs the
damaged variants in-process)."""
import gzip
import os
import random
import struct
import subprocess
import zlib

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "kmcp_b200", "kmcp-gpu")


def _gunzip(path, *args, ok=True):
    if not os.path.exists(EXE):
        pytest.fail("kmcp_b200/kmcp-gpu is not built: run __graft_entry__.build()")
    p = subprocess.run([EXE, "gunzip", *args, path], capture_output=True, timeout=300)
    if ok:
        assert p.returncode == 0, p.stderr.decode()
    return p


def _gz(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, mem=8):
    c = zlib.compressobj(level, zlib.DEFLATED, 31, mem, strategy)
    return c.compress(data) + c.flush()


def _corpus():
    rnd = random.Random(11)
    fq = b"".join(b"@read%d len=150\n%s\n+\n%s\n" % (i, bytes(rnd.choice(b"ACGT") for _ in range(150)),
                                                    bytes(rnd.choice(b"FFFFFFFF:,#") for _ in range(150))) for i in range(1500))
    far = os.urandom(300)
    return {
        "empty": b"",
        "one": b"x",
        "fastq": fq,
        "random": os.urandom(200_000),                                             # stored blocks at any level
        "zeros": bytes(3_000_000),                                                 # distance 1, length 258, window slides
        "bytes": bytes(range(256)) * 300,
        "far": (far + os.urandom(32768 - 300)) * 40,                               # matches at distance exactly 32768
        "text": b" ".join(rnd.choice([b"kmer", b"bloom", b"index", b"read", b"chunk", b"%d" % rnd.randrange(10 ** 6)]) for _ in range(200_000)),
        "skewed": bytes(rnd.choice(b"aaaaaaaaaaaaaaaaaaaaaaaaaaaaaaab") if rnd.random() < 0.999 else rnd.randrange(256) for _ in range(400_000)),
    }


@pytest.fixture(scope="module")
def corpus():
    return _corpus()


def test_every_zlib_stream_decodes_to_the_same_bytes(corpus, tmp_path):
    n = 0
    for name, data in corpus.items():
        for level in (0, 1, 6, 9):
            for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FILTERED):
                if level in (0, 9) and strategy not in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED):
                    continue
                p = str(tmp_path / "a.gz")
                open(p, "wb").write(_gz(data, level, strategy, mem=1 if level == 1 else 8))
                assert _gunzip(p).stdout == data, (name, level, strategy)
                n += 1
    assert n > 100


def test_small_reads_and_small_output_chunks(corpus, tmp_path):
    """input arriving a few bytes at a time (pipes) and callers asking for tiny pieces see the same stream"""
    for name in ("fastq", "far", "random", "skewed"):
        data = corpus[name][:150_000]
        p = str(tmp_path / "a.gz")
        open(p, "wb").write(_gz(data[:60_000], 6) + _gz(data[60_000:], 1, zlib.Z_FIXED))
        for rs, ck in ((1, 1 << 20), (7, 13), (4096, 1), (65536, 70_000)):
            if ck == 1 and name != "fastq":
                continue
            assert _gunzip(p, "--read-size", str(rs), "--chunk", str(ck)).stdout == data, (name, rs, ck)
    with open(str(tmp_path / "a.gz"), "rb") as f:                                # from a pipe
        q = subprocess.run([EXE, "gunzip", "-"], stdin=f, capture_output=True, timeout=60)
    assert q.returncode == 0 and q.stdout == corpus["skewed"][:150_000]


def test_members_headers_plain_input_and_trailing_garbage(corpus, tmp_path):
    fq = corpus["fastq"]
    p = str(tmp_path / "m.gz")
    # several members (pgzip / bgzip / `cat a.gz b.gz`), empty members in between, a member with a file name (gzip.GzipFile)
    import io
    named = io.BytesIO()
    with gzip.GzipFile(filename="reads.fq", mode="wb", fileobj=named, mtime=12345) as g:
        g.write(fq[5000:9000])
    blob = _gz(fq[:5000]) + _gz(b"") + named.getvalue() + _gz(b"") + _gz(fq[9000:], 9)
    open(p, "wb").write(blob)
    assert _gunzip(p).stdout == fq
    # all optional header fields at once: FEXTRA (a BGZF-style subfield), FNAME, FCOMMENT, FHCRC
    raw = zlib.compressobj(6, zlib.DEFLATED, -15)
    body = raw.compress(fq) + raw.flush()
    extra = b"BC" + struct.pack("<HH", 2, 0xBEEF)
    hdr = b"\x1f\x8b\x08" + bytes([2 | 4 | 8 | 16]) + struct.pack("<IBB", 0, 0, 255) + struct.pack("<H", len(extra)) + extra + b"name.fq\0" + b"a comment\0"
    hdr += struct.pack("<H", zlib.crc32(hdr) & 0xFFFF)
    tail = struct.pack("<II", zlib.crc32(fq), len(fq) & 0xFFFFFFFF)
    open(p, "wb").write(hdr + body + tail)
    assert _gunzip(p).stdout == fq
    # trailing garbage after a complete member is ignored (gzread does the same); zero padding too
    open(p, "wb").write(hdr + body + tail + b"\0" * 700)
    assert _gunzip(p).stdout == fq
    open(p, "wb").write(_gz(fq) + b"garbage that is not gzip")
    assert _gunzip(p).stdout == fq
    # input that is not gzip passes through unchanged
    for plain in (b"", b"A", b"\x1f", b">s\nACGT\n", fq, b"\x1f\x8c not gzip either"):
        open(p, "wb").write(plain)
        assert _gunzip(p).stdout == plain
    # wrong CRC, wrong length, reserved flag bits, unknown method
    for bad in (hdr + body + struct.pack("<II", zlib.crc32(fq) ^ 1, len(fq)), hdr + body + struct.pack("<II", zlib.crc32(fq), len(fq) + 1),
                b"\x1f\x8b\x08\x20" + hdr[4:] + body + tail, b"\x1f\x8b\x07" + hdr[3:] + body + tail, hdr + body + tail[:5], hdr[:7], hdr + body[:-3]):
        open(p, "wb").write(bad)
        r = _gunzip(p, ok=False)
        assert r.returncode == 1 and b"kmcp-gpu gunzip" in r.stderr, bad[:12]


def test_hand_made_deflate_streams(tmp_path):
    """block types and code shapes zlib never emits: an empty stored block between others, a distance code with one symbol,
    the longest code lengths, a match that reaches back before the start of the member (must be refused)"""
    class W:
        def __init__(self): self.acc, self.n, self.out = 0, 0, bytearray()
        def bits(self, v, n):
            self.acc |= v << self.n; self.n += n
            while self.n >= 8: self.out.append(self.acc & 255); self.acc >>= 8; self.n -= 8
        def code(self, v, n): self.bits(int(bin(v)[2:].zfill(n)[::-1], 2), n)       # Huffman codes go MSB first
        def align(self):
            if self.n: self.out.append(self.acc & 255); self.acc = 0; self.n = 0
    def member(raw, plain):
        return b"\x1f\x8b\x08\0\0\0\0\0\0\xff" + bytes(raw) + struct.pack("<II", zlib.crc32(plain), len(plain))
    p = str(tmp_path / "h.gz")
    # fixed-Huffman block "abc" + match(len 3, dist 3), empty stored block, fixed block with a 258-long run at distance 1
    w = W()
    w.bits(0, 1); w.bits(1, 2)
    for ch in b"abc": w.code(0x30 + ch, 8)
    w.code(1, 7); w.code(2, 5)                     # length 3 (symbol 257), distance 3 (symbol 2)
    w.code(0, 7)                                   # end of block
    w.bits(0, 1); w.bits(0, 2); w.align(); w.out += struct.pack("<HH", 0, 0xFFFF)
    w.bits(1, 1); w.bits(1, 2)
    w.code(0x30 + ord("z"), 8)
    w.code(0xC5, 8); w.code(0, 5)                  # length 258 (symbol 285), distance 1
    w.code(0, 7); w.align()
    plain = b"abcabc" + b"z" * 259
    open(p, "wb").write(member(w.out, plain))
    assert zlib.decompress(bytes(w.out), -15) == plain
    assert _gunzip(p).stdout == plain
    # a match before the start of the member
    w = W()
    w.bits(1, 1); w.bits(1, 2); w.code(0x30 + ord("a"), 8); w.code(1, 7); w.code(2, 5); w.code(0, 7); w.align()
    open(p, "wb").write(_gz(b"0123456789") + member(w.out, b"aaaa"))            # the previous member's bytes are not history
    r = _gunzip(p, ok=False)
    assert r.returncode == 1 and b"too far back" in r.stderr
    # reserved block type
    w = W(); w.bits(1, 1); w.bits(3, 2); w.align()
    open(p, "wb").write(member(w.out, b""))
    assert _gunzip(p, ok=False).returncode == 1


def test_big_stream_matches_zlib_and_reader_agrees(corpus, tmp_path):
    """tens of MB through the 1 MB window (many slides), then the same file through the FASTQ reader"""
    rnd = random.Random(5)
    recs = []
    for i in range(120_000):
        s = bytes(rnd.choice(b"ACGT") for _ in range(30)) * 5
        recs.append(b"@r%d\n%s\n+\n%s\n" % (i, s, b"F" * 150))
    data = b"".join(recs)
    p = str(tmp_path / "big.fq.gz")
    open(p, "wb").write(_gz(data[:len(data) // 2], 6) + _gz(data[len(data) // 2:], 1))
    r = _gunzip(p)
    assert zlib.crc32(r.stdout) == zlib.crc32(data) and len(r.stdout) == len(data)
    q = subprocess.run([EXE, "parse", "--ahead", p], capture_output=True, timeout=300)
    assert q.returncode == 0
    lines = q.stdout.split(b"\n")
    assert len(lines) == 120_001 and lines[0].startswith(b"r0\t150\t") and lines[119_999].startswith(b"r119999\t150\t")
    # a damaged file stops the reader with an error instead of a short result
    blob = bytearray(open(p, "rb").read())
    blob[len(blob) // 3] ^= 0x10
    open(p, "wb").write(bytes(blob))
    q = subprocess.run([EXE, "parse", p], capture_output=True, timeout=300)
    assert q.returncode != 0 and b"read error" in q.stderr


def test_damaged_streams_under_sanitizers(corpus, tmp_path):
    """truncations, bit flips, overwritten and deleted stretches: refused or decoded to the same bytes, and clean under ASan/UBSan"""
    exe = str(tmp_path / "fuzz")
    c = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
                        "-pthread", "-I", os.path.join(ROOT, "kmcp_b200", "csrc"), os.path.join(ROOT, "tests", "fastgz_fuzz.cpp"), "-o", exe],
                       capture_output=True, timeout=300)
    if c.returncode != 0:
        pytest.skip("no sanitizer runtime for g++ here: " + c.stderr.decode()[-200:])
    for name, level, strategy in (("fastq", 6, zlib.Z_DEFAULT_STRATEGY), ("text", 9, zlib.Z_DEFAULT_STRATEGY), ("skewed", 1, zlib.Z_FIXED),
                                  ("random", 6, zlib.Z_DEFAULT_STRATEGY), ("far", 6, zlib.Z_DEFAULT_STRATEGY), ("bytes", 6, zlib.Z_HUFFMAN_ONLY)):
        p = str(tmp_path / "f.gz")
        d = corpus[name][:120_000]
        open(p, "wb").write(_gz(d[:50_000], level, strategy) + _gz(d[50_000:], level, strategy))
        r = subprocess.run([exe, p, "150", "50000"], capture_output=True, timeout=600)
        assert r.returncode == 0, (name, r.stdout.decode(), r.stderr.decode()[-2000:])
        # the driver fails when a variant is accepted with bytes gzread would not have produced for it, or when the
        # chunk-parallel decoder disagrees with the sequential one
        assert b"variants 150" in r.stdout
    # the worker / consumer hand-offs of the chunk-parallel decoder under ThreadSanitizer
    tsan = str(tmp_path / "fuzz_tsan")
    c = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=thread", "-pthread", "-I", os.path.join(ROOT, "kmcp_b200", "csrc"),
                        os.path.join(ROOT, "tests", "fastgz_fuzz.cpp"), "-o", tsan], capture_output=True, timeout=300)
    if c.returncode == 0:
        d = corpus["fastq"]
        open(p, "wb").write(_gz(d[:300_000], 6) + _gz(d[300_000:], 6))
        r = subprocess.run([tsan, p, "25", "300000"], capture_output=True, timeout=600)
        if r.returncode != 0 and b"ThreadSanitizer" not in r.stderr and b"variant" not in r.stdout and b"failed" not in r.stdout:
            pytest.skip("ThreadSanitizer cannot run here")                      # e.g. an unsupported address-space layout
        assert r.returncode == 0 and b"variants 25" in r.stdout, (r.stdout.decode(), r.stderr.decode()[-3000:])


# ---- pargz.h: one gzip stream decoded by several threads ---------------------------------------------------------------
def _par(path, threads=3, par_chunk=65536, ok=True, chunk=None):
    args = ["--threads", str(threads), "--par-chunk", str(par_chunk), "--stats"] + (["--chunk", str(chunk)] if chunk else [])
    return _gunzip(path, *args, ok=ok)


def _stats(p):
    import re
    m = re.search(rb"chunks used (\d+), stretches decoded again in order (\d+), symbols resolved (\d+)", p.stderr)
    return tuple(int(x) for x in m.groups())


def test_parallel_decoder_gives_the_same_bytes(corpus, tmp_path):
    """every stream of the corpus cut into 64 KB chunks: block starts found by the workers, windows passed on, markers resolved"""
    p = str(tmp_path / "a.gz")
    used_total = 0
    for name, data in corpus.items():
        for level, strategy in ((6, zlib.Z_DEFAULT_STRATEGY), (1, zlib.Z_DEFAULT_STRATEGY), (9, zlib.Z_FILTERED), (6, zlib.Z_FIXED), (0, zlib.Z_DEFAULT_STRATEGY),
                                (6, zlib.Z_HUFFMAN_ONLY), (6, zlib.Z_RLE)):
            z = _gz(data, level, strategy)
            if len(z) < 18:
                continue
            open(p, "wb").write(z)
            r = _par(p, threads=1 + (len(z) % 4))
            assert r.stdout == data, (name, level, strategy)
            used, redone, _ = _stats(r)
            used_total += used
            if strategy == zlib.Z_DEFAULT_STRATEGY and level == 6 and name in ("fastq", "text", "skewed"):
                assert used >= max(1, len(z) // 65536 - 1) and redone <= 2, (name, used, redone)        # dynamic blocks: every chunk is used
    assert used_total > 100
    # incompressible data is stored blocks, which have no header to look for: after a few misses the decoder stops looking and
    # decodes in order (same bytes, no wasted work)
    big_random = os.urandom(3_000_000)
    open(p, "wb").write(_gz(big_random, 6))
    r = _par(p, threads=3)
    assert r.stdout == big_random and b"gave up looking for block starts" in r.stderr
    # a FASTQ stream big enough for many chunks at the default chunk size too, odd output piece sizes
    big = corpus["fastq"] * 40
    open(p, "wb").write(_gz(big, 6))
    for th, pc, ck in ((4, 65536, 1000), (2, 300_000, None), (8, 2 << 20, None)):
        r = _par(p, threads=th, par_chunk=pc, chunk=ck)
        assert zlib.crc32(r.stdout) == zlib.crc32(big) and len(r.stdout) == len(big), (th, pc)


def test_parallel_decoder_members_garbage_and_block_kinds(corpus, tmp_path):
    p = str(tmp_path / "m.gz")
    fq, rnd_bytes, text = corpus["fastq"], corpus["random"], corpus["text"]
    # many members of every kind back to back (bgzip / pgzip style and `cat`), empty ones, stored and fixed blocks in between
    parts = [fq[:100_000], b"", rnd_bytes[:150_000], text[:300_000], b"x", fq[100_000:], text[300_000:]]
    blob = b"".join(_gz(d, lv, st) for d, lv, st in zip(parts, (6, 6, 6, 9, 1, 6, 1), (0, 0, 0, 0, zlib.Z_FIXED, zlib.Z_FIXED, 0)))
    open(p, "wb").write(blob)
    whole = b"".join(parts)
    for th in (1, 3):
        assert _par(p, threads=th).stdout == whole
    open(p, "wb").write(blob + b"\0" * 100_000)                  # zero padding / garbage behind the last member is ignored
    assert _par(p).stdout == whole
    open(p, "wb").write(blob + b"not gzip " * 30_000)
    assert _par(p).stdout == whole
    # 64 KB members as bgzip writes them: every chunk holds several member starts
    bg = b"".join(_gz(text[i:i + 60_000]) for i in range(0, len(text), 60_000))
    open(p, "wb").write(bg)
    r = _par(p, par_chunk=65536)
    assert r.stdout == text
    used, redone, resolved = _stats(r)
    assert used >= len(bg) // 65536 - 1 and redone <= 1 and resolved == 0      # chunks start at member starts: plain bytes, nothing to resolve
    # the same with members whose only block is the final one (what bgzip writes) and a long FNAME in every header
    import io
    parts = []
    for i in range(0, len(text), 50_000):
        m = io.BytesIO()
        with gzip.GzipFile(filename="a_rather_long_file_name_%d.txt" % i, mode="wb", fileobj=m, compresslevel=1) as g:
            g.write(text[i:i + 50_000])
        parts.append(m.getvalue())
    open(p, "wb").write(b"".join(parts))
    r = _par(p, par_chunk=65536, threads=4)
    assert r.stdout == text
    used, redone, resolved = _stats(r)
    assert used >= len(b"".join(parts)) // 65536 - 1 and redone <= 1 and resolved == 0
    # one flush per 10 KB: empty stored blocks and byte-aligned block starts all over the stream
    c = zlib.compressobj(6, zlib.DEFLATED, 31)
    z = b"".join(c.compress(fq[i:i + 10_000]) + c.flush(zlib.Z_FULL_FLUSH if i % 30_000 == 0 else zlib.Z_SYNC_FLUSH) for i in range(0, len(fq), 10_000)) + c.flush()
    open(p, "wb").write(z)
    assert _par(p).stdout == fq
    # not gzip / too small / a pipe: refused up front (the reader falls back to the sequential decoder)
    open(p, "wb").write(b"@r\nACGT\n+\nIIII\n" * 10)
    assert _par(p, ok=False).returncode == 2


def test_parallel_decoder_refuses_damaged_streams(corpus, tmp_path):
    """a flipped bit anywhere: an error (CRC, length or a broken code), never other bytes, never a hang"""
    p = str(tmp_path / "d.gz")
    fq = corpus["fastq"] * 3
    good = _gz(fq[:700_000], 6) + _gz(fq[700_000:], 6)
    rnd = random.Random(9)
    for trial in range(40):
        blob = bytearray(good)
        kind = trial % 4
        if kind == 0:
            blob[rnd.randrange(12, len(blob) - 8)] ^= 1 << rnd.randrange(8)
        elif kind == 1:
            blob = blob[:rnd.randrange(20, len(blob) - 1)]
        elif kind == 2:
            a = rnd.randrange(12, len(blob) - 100)
            blob[a:a + rnd.randrange(1, 80)] = os.urandom(5)
        else:
            blob[-rnd.randrange(1, 9)] ^= 0x40                                           # CRC-32 / ISIZE of the last member
        open(p, "wb").write(bytes(blob))
        r = _par(p, threads=1 + trial % 4, par_chunk=65536, ok=False)
        s = _gunzip(p, ok=False)
        assert r.returncode in (0, 1), r.stderr
        assert (r.returncode == 0) == (s.returncode == 0), (trial, kind, r.stderr, s.stderr)      # the sequential decoder agrees
        if r.returncode == 0:
            assert r.stdout == s.stdout
        else:
            assert s.stdout.startswith(r.stdout) or r.stdout.startswith(s.stdout)            # what came out before the error is real data


def test_reader_uses_the_parallel_decoder(corpus, tmp_path):
    """kmcp-gpu parse --inflate-threads N: same records as the sequential reader, single and paired files"""
    rnd = random.Random(2)
    recs = [(b"q%d" % i, bytes(rnd.choice(b"ACGT") for _ in range(rnd.choice([50, 150, 151])))) for i in range(40_000)]
    a, b = str(tmp_path / "a.fq.gz"), str(tmp_path / "b.fq.gz")
    with gzip.open(a, "wb", compresslevel=6) as f:
        for i, s in recs:
            f.write(b"@" + i + b" d\n" + s + b"\n+\n" + b"F" * len(s) + b"\n")
    with gzip.open(b, "wb", compresslevel=1) as f:
        for i, s in recs:
            f.write(b"@" + i + b"/2\n" + s[::-1] + b"\n+\n" + b"#" * len(s) + b"\n")
    def parse(*args):
        q = subprocess.run([EXE, "parse", *args], capture_output=True, timeout=300)
        assert q.returncode == 0, q.stderr.decode()
        return q.stdout
    env_small = ["--inflate-threads", "3", "--inflate-chunk", "65536"]
    assert parse(*env_small, a) == parse("--inflate-threads", "1", a)
    assert parse(*env_small, "--ahead", "-1", a, "-2", b) == parse("--inflate-threads", "1", "-1", a, "-2", b)
    exp = b"".join(b"%s\t%d\t%08x\n" % (i, len(s), zlib.crc32(s)) for i, s in recs)
    assert parse(*env_small, "--ahead", a) == exp


def test_result_writer_gz_members_in_order(corpus, tmp_path):
    """kmcp-gpu gzip-write: the CLI's .gz result writer alone (1 MB members compressed side by side, written in order,
    queued asynchronously) — zlib, the sequential decoder and the chunk-parallel decoder all read the text back"""
    text = corpus["text"] * 3 + corpus["fastq"]
    src, out = str(tmp_path / "t.tsv"), str(tmp_path / "o.tsv.gz")
    open(src, "wb").write(text)
    for piece in (str(19 << 20), "700001", "1"[:1] + "000"):
        r = subprocess.run([EXE, "gzip-write", src, out, piece], capture_output=True, timeout=300)
        assert r.returncode == 0, r.stderr.decode()
        blob = open(out, "rb").read()
        assert gzip.decompress(blob) == text
        assert blob.count(b"\x1f\x8b\x08") >= len(text) // (1 << 20)            # many members
    assert _gunzip(out).stdout == text
    assert _par(out, threads=3, par_chunk=65536).stdout == text
    plain = str(tmp_path / "o.tsv")
    assert subprocess.run([EXE, "gzip-write", src, plain], capture_output=True, timeout=300).returncode == 0
    assert open(plain, "rb").read() == text


def test_parallel_decoder_memory_cap_and_reader_fallback(tmp_path):
    """a stream that expands a few hundred times: the chunk-parallel decoder stops at its per-chunk cap (what it hands out up to
    there is real data) and the reader carries on with the sequential decoder behind the bytes it already has"""
    rec = b"@r\n" + b"A" * 150 + b"\n+\n" + b"F" * 150 + b"\n"
    data = rec * 60_000
    p = str(tmp_path / "rep.fq.gz")
    open(p, "wb").write(_gz(data, 6))
    r = _gunzip(p, "--threads", "3", "--par-chunk", "65536", "--par-cap", "1000000", ok=False)
    assert r.returncode == 4 and b"cap of the parallel decoder" in r.stderr and data.startswith(r.stdout) and len(r.stdout) < len(data)
    assert _gunzip(p, "--threads", "3", "--par-chunk", "65536").stdout == data           # under the default cap: decoded in parallel
    q = subprocess.run([EXE, "parse", "--ahead", "--inflate-threads", "3", "--inflate-chunk", "65536", "--inflate-cap", "1000000", p],
                       capture_output=True, timeout=300)
    assert q.returncode == 0, q.stderr.decode()
    assert q.stdout == (b"r\t150\t%08x\n" % zlib.crc32(b"A" * 150)) * 60_000

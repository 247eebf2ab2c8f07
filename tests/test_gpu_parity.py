"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.
Bit-exact everywhere: codes, per-target counts, hit sets, float columns of the engine output.
None of these read /root/reference (it does not exist on the GPU box)."""
import os

import numpy as np
import pytest

import parity_helpers as helpers

pytestmark = pytest.mark.gpu

GSEED, RSEED = 11, 23


@pytest.fixture(scope="module")
def small_db(oracle, tmp_path_factory):
    """40 genomes x 30 kb, 5 chunks -> 200 targets; h=1; blocks of 64 (8-byte rows, G=1) → 4 blocks"""
    O = oracle
    sp = O.sketch_params(21)
    targets = helpers.make_synth_targets(O, sp, GSEED, 40, 30000, 5, 150)
    out = str(tmp_path_factory.mktemp("db_small"))
    r001 = O.build_db(targets, out, sp, num_hashes=1, fpr=0.3, block_size=64)
    return r001


@pytest.fixture(scope="module")
def wide_db(oracle, tmp_path_factory):
    """150 genomes x 12 kb, 10 chunks -> 1500 targets in ONE block: 188-byte rows → two 128-byte task chunks; h=3"""
    O = oracle
    sp = O.sketch_params(21)
    targets = helpers.make_synth_targets(O, sp, GSEED + 1, 150, 12000, 10, 100)
    out = str(tmp_path_factory.mktemp("db_wide"))
    r001 = O.build_db(targets, out, sp, num_hashes=3, fpr=0.1, block_size=1500)
    return r001


def _oracle_opts(O, **kw):
    o = O.default_opts()
    for k, v in kw.items():
        setattr(o, k, v)
    return o


# ---------------------------------------------------------------------------------------------------- kernel 1
@pytest.mark.parametrize("k", [21, 31, 11, 64])
def test_generate_kmers_matches_oracle(gpu_ctx, oracle, k):
    from kmcp_b200 import api
    O = oracle
    reads = helpers.edge_reads(k) + helpers.make_reads(O, RSEED, 300, 40, 30000, GSEED)
    buf, off = api.pack_seqs(reads)
    for scaled, scale in ((0, 1), (1, 8)):
        sp = api.SketchParams(k, 1, scaled, scale, 0, 0, 0, 0)
        codes, coff = gpu_ctx.generate_kmers(buf, off, sp)
        osp = O.sketch_params(k, scaled=bool(scaled), scale=scale)
        for i, r in enumerate(reads):
            exp = O.generate_kmers(r, osp)
            got = codes[int(coff[i]):int(coff[i + 1])]
            assert np.array_equal(got, exp), (k, scaled, i, len(r))


def test_generate_kmers_long_sequence(gpu_ctx, oracle):
    from kmcp_b200 import api
    O = oracle
    seq = O.synth_genome(5, 0, 200000)
    seq = seq[:70000] + b"N" * 500 + seq[70000:]
    buf, off = api.pack_seqs([seq, seq[:5000]])
    for scaled, scale in ((0, 1), (1, 100)):
        sp = api.SketchParams(31, 1, scaled, scale, 0, 0, 0, 0)
        codes, coff = gpu_ctx.generate_kmers(buf, off, sp)
        osp = O.sketch_params(31, scaled=bool(scaled), scale=scale)
        assert np.array_equal(codes[:int(coff[1])], O.generate_kmers(seq, osp))
        assert np.array_equal(codes[int(coff[1]):], O.generate_kmers(seq[:5000], osp))


# ---------------------------------------------------------------------------------------------------- kernel 2
@pytest.mark.parametrize("dbname", ["small_db", "wide_db"])
def test_count_codes_matches_oracle(gpu_ctx, oracle, request, dbname):
    O = oracle
    r001 = request.getfixturevalue(dbname)
    odb = O.DB(r001)
    gpu_ctx.open_db(r001)
    info = gpu_ctx.db_info()
    assert info.n_targets == odb.info.n_targets and info.n_blocks == odb.info.n_blocks
    assert info.sum_row_bytes == odb.info.sum_row_bytes and info.disk_bytes == odb.info.total_bytes
    sp = odb.sketch_params()
    ng, gl = (40, 30000) if dbname == "small_db" else (150, 12000)
    gs = GSEED if dbname == "small_db" else GSEED + 1
    for n_reads in (1, 7):
        reads = helpers.make_reads(O, RSEED + n_reads, n_reads, ng, gl, gs)
        codes = np.concatenate([O.generate_kmers(r, sp) for r in reads])
        exp = odb.count_codes(codes)
        got = gpu_ctx.count_codes(codes)
        assert np.array_equal(got, exp)
        assert exp.max() > 50
    # > 255 and > 65535 codes exercise the 16- and 24-plane counters
    long_codes = O.generate_kmers(O.synth_genome(gs, 3, gl), sp)
    for codes in (long_codes[:300], long_codes[:5000], np.tile(long_codes[:7000], 10)):
        assert np.array_equal(gpu_ctx.count_codes(codes), odb.count_codes(codes))
    assert np.array_equal(gpu_ctx.count_codes(np.zeros(0, np.uint64)), np.zeros(info.n_targets, np.uint32))


def test_targets_metadata(gpu_ctx, oracle, small_db):
    O = oracle
    odb = O.DB(small_db)
    gpu_ctx.open_db(small_db)
    for g in range(odb.info.n_targets):
        a, b = odb.target(g), gpu_ctx.target(g)
        assert (a.name, a.index, a.genome_size, a.n_kmers, a.block, a.col) == (b.name, b.index, b.genome_size, b.n_kmers, b.block, b.col)
        assert b.resident == 1


# ---------------------------------------------------------------------------------------------------- whole path
def _compare_engine(O, odb, ctx, reads, paired=False, **opts):
    from kmcp_b200 import api
    buf, off = api.pack_seqs(reads)
    oo = _oracle_opts(O, **opts)
    exp = odb.search(reads, paired=paired, opts=oo)
    eo = ctx.default_engine_opts(paired=int(paired), **opts)
    got = ctx.engine_search(buf, off, eo)
    assert np.array_equal(got.query_len, exp.query_len)
    assert np.array_equal(got.n_kmers, exp.n_kmers)
    assert np.array_equal(got.match_off, exp.hit_off)
    for f in ("query", "target", "count", "fpr", "qcov", "tcov", "jacc"):
        assert np.array_equal(got.matches[f], exp.hits[f]), f      # floats compared bit for bit
    return got


@pytest.mark.parametrize("dbname,ng,gl,gs", [("small_db", 40, 30000, GSEED), ("wide_db", 150, 12000, GSEED + 1)])
def test_search_batch_hits_match_oracle(gpu_ctx, oracle, request, dbname, ng, gl, gs):
    from kmcp_b200 import api
    O = oracle
    r001 = request.getfixturevalue(dbname)
    odb = O.DB(r001)
    gpu_ctx.open_db(r001)
    reads = helpers.make_reads(O, RSEED, 2000, ng, gl, gs) + helpers.edge_reads(odb.k)
    buf, off = api.pack_seqs(reads)
    got = gpu_ctx.search_batch(buf, off)
    # oracle with the post filters disabled = exactly the device-side contract
    exp = odb.search(reads, opts=_oracle_opts(O, max_fpr=1.0, min_target_cov=0.0))
    assert np.array_equal(got.n_kmers, exp.n_kmers)
    assert np.array_equal(got.query_len, exp.query_len)
    assert helpers.hits_to_set(got.hits) == helpers.hits_to_set(exp.hits)
    assert len(got.hits) == len(exp.hits) > 1000
    # canonical order
    key = got.hits["query"].astype(np.uint64) << np.uint64(32) | got.hits["target"].astype(np.uint64)
    assert np.all(key[1:] > key[:-1])
    assert got.kernel_launches > 0
    n_sum = int(got.n_kmers.astype(np.int64).sum())
    assert got.probe_row_bytes == n_sum * odb.info.num_hashes * odb.info.sum_row_bytes


@pytest.mark.parametrize("dbname,ng,gl,gs", [("small_db", 40, 30000, GSEED), ("wide_db", 150, 12000, GSEED + 1)])
def test_engine_matches_oracle_defaults(gpu_ctx, oracle, request, dbname, ng, gl, gs):
    O = oracle
    r001 = request.getfixturevalue(dbname)
    odb = O.DB(r001)
    gpu_ctx.open_db(r001)
    reads = helpers.make_reads(O, RSEED + 5, 3000, ng, gl, gs) + helpers.edge_reads(odb.k)
    got = _compare_engine(O, odb, gpu_ctx, reads)
    assert len(got.matches) > 1500


def test_engine_large_batch_threaded_postfilter(gpu_ctx, oracle, small_db):
    """> 4096 queries: the host post-filter runs on several threads"""
    O = oracle
    odb = O.DB(small_db)
    gpu_ctx.open_db(small_db)
    reads = helpers.make_reads(O, RSEED + 11, 12000, 40, 30000, GSEED)
    got = _compare_engine(O, odb, gpu_ctx, reads)
    assert len(got.matches) > 8000


def test_engine_option_variants(gpu_ctx, oracle, small_db):
    O = oracle
    odb = O.DB(small_db)
    gpu_ctx.open_db(small_db)
    reads = helpers.make_reads(O, RSEED + 9, 800, 40, 30000, GSEED) + helpers.edge_reads(odb.k)
    _compare_engine(O, odb, gpu_ctx, reads, min_query_cov=0.3, sort_by=1)
    _compare_engine(O, odb, gpu_ctx, reads, min_query_cov=0.7, sort_by=2, top_n_scores=1)
    _compare_engine(O, odb, gpu_ctx, reads, min_query_cov=0.2, min_matched=3, min_target_cov=0.002, max_fpr=1e-6)
    _compare_engine(O, odb, gpu_ctx, reads, do_not_sort=1, min_query_len=100)
    _compare_engine(O, odb, gpu_ctx, reads, dedup_threshold=50)       # 150 bp reads now take the sort+unique path
    _compare_engine(O, odb, gpu_ctx, reads, min_query_cov=0.0)
    _compare_engine(O, odb, gpu_ctx, reads, min_query_cov=1.0)


def test_engine_paired_and_try_se(gpu_ctx, oracle, small_db):
    O = oracle
    odb = O.DB(small_db)
    gpu_ctx.open_db(small_db)
    r1 = helpers.make_reads(O, RSEED + 1, 600, 40, 30000, GSEED)
    r2 = helpers.make_reads(O, RSEED + 2, 600, 40, 30000, GSEED)
    r2[5] = b"ACGT"            # short mate
    r1[6] = b"ACGTACGT"        # short Seq, long Seq2
    r1[7] = b""; r2[7] = b""
    reads = [x for p in zip(r1, r2) for x in p]
    _compare_engine(O, odb, gpu_ctx, reads, paired=True)
    _compare_engine(O, odb, gpu_ctx, reads, paired=True, try_se=1, min_query_cov=0.6)
    _compare_engine(O, odb, gpu_ctx, reads, paired=True, dedup_threshold=200)


def test_long_reads_dedup(gpu_ctx, oracle, small_db):
    """HiFi-like queries: n >> 256 → device sort+unique (U:874-908), 16-plane counters"""
    O = oracle
    odb = O.DB(small_db)
    gpu_ctx.open_db(small_db)
    rng = np.random.default_rng(3)
    reads = []
    for i in range(40):
        g = int(rng.integers(0, 40)); L = int(rng.integers(300, 12000)); p = int(rng.integers(0, 30000 - L))
        s = O.synth_genome(GSEED, g, 30000)[p:p + L]
        if i % 3 == 0:
            s = s + s[:L // 2]                     # duplicated k-mers
        reads.append(s)
    reads += helpers.make_reads(O, RSEED, 50, 40, 30000, GSEED)     # mixed with short reads in one batch
    got = _compare_engine(O, odb, gpu_ctx, reads)
    assert got.n_kmers.max() > 5000


def test_device_resident_batch(gpu_ctx, oracle, small_db):
    from kmcp_b200 import api
    O = oracle
    odb = O.DB(small_db)
    gpu_ctx.open_db(small_db)
    reads = helpers.make_reads(O, RSEED + 3, 500, 40, 30000, GSEED)
    buf, off = api.pack_seqs(reads)
    host = gpu_ctx.search_batch(buf, off)
    dseq = gpu_ctx.device_alloc(buf.nbytes); doff = gpu_ctx.device_alloc(off.nbytes)
    gpu_ctx.h2d(dseq, buf); gpu_ctx.h2d(doff, off)
    dev = gpu_ctx.search_batch_ptr(dseq, doff, len(reads), gpu_ctx.default_params(), device=True, seq_bytes=buf.nbytes)
    gpu_ctx.device_free(dseq); gpu_ctx.device_free(doff)
    assert np.array_equal(host.hits, dev.hits) and np.array_equal(host.n_kmers, dev.n_kmers)


def test_sharded_blocks_union_equals_whole(oracle, small_db):
    """blocks → shards: per-shard hit lists are disjoint by target and their union is the 1-GPU result (§8e)"""
    from kmcp_b200 import api
    O = oracle
    reads = helpers.make_reads(O, RSEED + 4, 1500, 40, 30000, GSEED)
    buf, off = api.pack_seqs(reads)
    with api.Context(0) as whole:
        whole.open_db(small_db)
        ref = whole.search_batch(buf, off)
    parts = []
    for world in (2, 3):
        parts = []
        nres = 0
        for rank in range(world):
            with api.Context(0) as c:
                c.open_db(small_db, shard_rank=rank, shard_world=world)
                nres += c.db_info().n_resident_blocks
                parts.append(c.search_batch(buf, off).hits)
        assert nres == 4
        allh = np.concatenate(parts)
        allh = allh[np.lexsort((allh["target"], allh["query"]))]
        assert np.array_equal(allh, ref.hits)


# ---------------------------------------------------------------------------------------------------- synthetic tooling
def test_synth_reads_match_oracle_generator(gpu_ctx, oracle):
    O = oracle
    n, L = 500, 150
    d = gpu_ctx.device_alloc(n * L)
    gpu_ctx.synth_reads(RSEED, 100, n, L, GSEED, 40, 30000, d)
    got = gpu_ctx.d2h(d, n * L).reshape(n, L)
    gpu_ctx.device_free(d)
    for r in range(n):
        assert got[r].tobytes() == O.synth_read(RSEED, 100 + r, 40, 30000, L, GSEED), r


@pytest.mark.parametrize("h,fpr,bs", [(1, 0.3, 16), (3, 0.05, 40)])
def test_device_index_builder_matches_oracle_builder(gpu_ctx, oracle, tmp_path, h, fpr, bs):
    """kmcpg_build_synth_db (GPU compute+index) writes byte-identical .uniki blocks to the oracle's builder"""
    O = oracle
    sp = O.sketch_params(21)
    ng, gl, nc, ov = 12, 20000, 5, 150
    targets = helpers.make_synth_targets(O, sp, 77, ng, gl, nc, ov)
    r001 = O.build_db(targets, str(tmp_path / "o"), sp, num_hashes=h, fpr=fpr, block_size=bs)
    gpu_ctx.build_synth_db(77, ng, gl, k=21, n_chunks=nc, overlap=ov, num_hashes=h, fpr=fpr, block_size=bs)
    info = gpu_ctx.db_info()
    assert info.n_targets == len(targets)
    for b in range(info.n_blocks):
        p = str(tmp_path / ("g%d.uniki" % b))
        gpu_ctx.write_block(b, p)
        with open(p, "rb") as f1, open(os.path.join(r001, "_block%03d.uniki" % (b + 1)), "rb") as f2:
            assert f1.read() == f2.read(), b
    # and the freshly built in-HBM DB answers queries like the oracle's copy
    odb = O.DB(r001)
    reads = helpers.make_reads(O, RSEED, 400, ng, gl, 77)
    _compare_engine(O, odb, gpu_ctx, reads)


def test_errors_are_reported_not_fatal(gpu_ctx, tmp_path):
    from kmcp_b200 import api
    with pytest.raises(api.KmcpGpuError) as e:
        gpu_ctx.open_db(str(tmp_path / "missing"))
    assert e.value.code == api.KMCPG_EIO
    d = tmp_path / "bad"; d.mkdir()
    (d / "__db.yml").write_text("version: 3\nfiles:\n- x.uniki\n")
    with pytest.raises(api.KmcpGpuError) as e:
        gpu_ctx.open_db(str(d))
    assert e.value.code == api.KMCPG_EFORMAT


# ---------------------------------------------------------------------------------------------------- CLI
def _write_fastq(path, ids, seqs, gz=True):
    import gzip
    op = gzip.open if gz else open
    with op(path, "wb") as f:
        for i, s in zip(ids, seqs):
            f.write(b"@" + i + b" some description\n" + s + b"\n+\n" + b"I" * len(s) + b"\n")


def _run_cli(args):
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "kmcp_b200", "kmcp-gpu")
    p = subprocess.run([exe, "search", "-q"] + args, capture_output=True)
    assert p.returncode == 0, p.stderr.decode()
    return p


def test_cli_tsv_is_byte_identical_to_the_oracle_formatter(oracle, small_db, tmp_path):
    """kmcp-gpu search: same flags as `kmcp search`, 15-column TSV + trailer identical to the oracle's S:437-575 restatement"""
    import gzip
    O = oracle
    odb = O.DB(small_db)
    dbdir = os.path.dirname(small_db)
    reads = helpers.make_reads(O, RSEED + 20, 1500, 40, 30000, GSEED) + [r for r in helpers.edge_reads(21) if len(r) > 0]
    ids = [b"read_%d/1" % i for i in range(len(reads))]
    fq = str(tmp_path / "q.fq.gz")
    _write_fastq(fq, ids, reads)
    # single-end, defaults, gz output
    out = str(tmp_path / "o.tsv.gz")
    _run_cli(["-d", dbdir, fq, "-o", out])
    exp = O.format_tsv(odb, ids, odb.search(reads))
    assert gzip.open(out, "rb").read().decode() == exp
    # --ref-counts: `kmcp profile` stage-1 counters from the result stream == the oracle's stage 1 over the expected TSV
    rcf = str(tmp_path / "refcounts.tsv")
    _run_cli(["-d", dbdir, fq, "-o", out, "--ref-counts", rcf, "--ref-counts-min-qcov", "0.6", "--ref-counts-max-fpr", "0.001"])
    assert gzip.open(out, "rb").read().decode() == exp
    exp_reads, exp_prof = O.profile_stage1(exp, min_qcov=0.6, max_fpr=0.001)
    got_prof, got_reads = {}, None
    for l in open(rcf).read().splitlines():
        if l.startswith("# reads: "):
            got_reads = int(l.split(": ")[1])
        if l.startswith("#"):
            continue
        name, c, n, gs, m, u, hc = l.split("\t")
        t = got_prof.setdefault(name, (int(gs), [0.0] * int(n), [0.0] * int(n), [0.0] * int(n)))
        t[1][int(c)], t[2][int(c)], t[3][int(c)] = float(m), float(u), float(hc)
    assert got_reads == exp_reads and got_prof == exp_prof and len(exp_prof) >= 30
    # -K keeps unmatched rows; other thresholds; sort by jacc; top score
    out2 = str(tmp_path / "o2.tsv")
    _run_cli(["-d", dbdir, fq, "-o", out2, "-K", "-t", "0.4", "-c", "5", "-s", "jacc", "-n", "1", "-f", "0.05"])
    oo = O.default_opts(); oo.min_query_cov = 0.4; oo.min_matched = 5; oo.sort_by = 2; oo.top_n_scores = 1; oo.max_fpr = 0.05
    exp2 = O.format_tsv(odb, ids, odb.search(reads, opts=oo), keep_unmatched=True)
    assert open(out2).read() == exp2
    # paired-end (-1/-2): IDs of read1, mates concatenated
    r2 = helpers.make_reads(O, RSEED + 21, len(reads), 40, 30000, GSEED)
    fq2 = str(tmp_path / "q2.fq")
    _write_fastq(fq2, [b"read_%d/2" % i for i in range(len(r2))], r2, gz=False)
    out3 = str(tmp_path / "o3.tsv")
    _run_cli(["-d", dbdir, "-1", fq, "-2", fq2, "-o", out3, "-H"])
    inter = [x for p in zip(reads, r2) for x in p]
    exp3 = O.format_tsv(odb, ids, odb.search(inter, paired=True), header=False)
    assert open(out3).read() == exp3
    # whole-file query (-g) from a multi-record FASTA, custom ID
    fa = str(tmp_path / "g.fa")
    recs = [O.synth_genome(GSEED, 7, 30000)[:9000], O.synth_genome(GSEED, 7, 30000)[9000:15000], O.synth_genome(GSEED, 8, 30000)[:5000]]
    with open(fa, "wb") as f:
        for i, s in enumerate(recs):
            f.write(b">rec%d desc\n" % i)
            for j in range(0, len(s), 70):
                f.write(s[j:j + 70] + b"\n")
    out4 = str(tmp_path / "o4.tsv")
    _run_cli(["-d", dbdir, "-g", "--query-id", "myquery", fa, "-o", out4, "-t", "0.1"])
    whole = recs[0] + recs[1] + b"N" * 20 + recs[2] + b"N" * 20          # S:905-913
    o4 = O.default_opts(); o4.min_query_cov = 0.1
    exp4 = O.format_tsv(odb, [b"myquery"], odb.search([whole], opts=o4))
    assert open(out4).read() == exp4
    # untidy input: CRLF line ends, multi-line FASTQ, blank lines between records, wrapped FASTA without a final newline
    sub, sid = reads[:40], ids[:40]
    messy = str(tmp_path / "messy.fq")
    with open(messy, "wb") as f:
        for i, (n, r) in enumerate(zip(sid, sub)):
            h = len(r) // 2
            if i % 3 == 0:
                f.write(b"@" + n + b" some description\r\n" + r + b"\r\n+\r\n" + b"I" * len(r) + b"\r\n")
            elif i % 3 == 1:
                f.write(b"\n@" + n + b"\tx\n" + r[:h] + b"\n" + r[h:] + b"\n+" + n + b"\n" + b"@" * h + b"\n" + b"+" * (len(r) - h) + b"\n")
            else:
                f.write(b"@" + n + b"\n" + r + b"\n+\n" + b"5" * len(r) + b"\n\n")
    out5 = str(tmp_path / "o5.tsv")
    _run_cli(["-d", dbdir, messy, "-o", out5])
    assert open(out5).read() == O.format_tsv(odb, sid, odb.search(sub))
    messy_fa = str(tmp_path / "messy.fa")
    with open(messy_fa, "wb") as f:
        for i, (n, r) in enumerate(zip(sid, sub)):
            f.write(b">" + n + b" d\n" + b"\n".join(r[j:j + 60] for j in range(0, len(r), 60)) + (b"\n" if i + 1 < len(sub) else b""))
    out6 = str(tmp_path / "o6.tsv")
    _run_cli(["-d", dbdir, messy_fa, "-o", out6])
    assert open(out6).read() == O.format_tsv(odb, sid, odb.search(sub))


def test_cli_several_databases_are_merged_like_kmcp_merge(oracle, small_db, tmp_path):
    """`-d A -d B`: every database searched, rows of a query united, re-sorted by score, `hits` rewritten (merge.go:190-256)"""
    O = oracle
    sp = O.sketch_params(21)
    targets = helpers.make_synth_targets(O, sp, GSEED, 40, 30000, 5, 150)
    # second database: other genomes plus ten genomes shared with the first (so queries hit both)
    t2 = helpers.make_synth_targets(O, sp, GSEED + 5, 20, 30000, 5, 150) + targets[150:]
    r2 = O.build_db(t2, str(tmp_path / "db2"), sp, num_hashes=1, fpr=0.3, block_size=64)
    dbs = [O.DB(small_db), O.DB(r2)]
    reads = helpers.make_reads(O, RSEED + 50, 1200, 40, 30000, GSEED) + helpers.make_reads(O, RSEED + 51, 500, 20, 30000, GSEED + 5)
    ids = [b"q%d" % i for i in range(len(reads))]
    fq = str(tmp_path / "q.fq")
    _write_fastq(fq, ids, reads, gz=False)
    out = str(tmp_path / "m.tsv")
    _run_cli(["-d", os.path.dirname(small_db), "-d", os.path.dirname(r2), fq, "-o", out, "-K"])
    res = [d.search(reads) for d in dbs]
    exp = [O.TSV_HEADER]
    matched = 0
    for q in range(len(ids)):
        rows = []
        for di, (d, r) in enumerate(zip(dbs, res)):
            for h in r.hits[int(r.hit_off[q]):int(r.hit_off[q + 1])]:
                rows.append((-float(h["qcov"]), -float(h["tcov"]), di, int(h["target"]), h, d, r))
        if not rows:
            r = res[0]
            exp.append("%s\t%d\t%d\t0\t0\t\t-1\t0\t0\t%d\t0\t0\t0\t0\t%d\n" % (ids[q].decode(), r.query_len[q], r.n_kmers[q], r.k_used[q], q))
            continue
        matched += 1
        rows.sort(key=lambda x: x[:4])
        for _, _, _, _, h, d, r in rows:
            t = d.target(int(h["target"]))
            exp.append("%s\t%d\t%d\t%s\t%d\t%s\t%d\t%d\t%d\t%d\t%d\t%.4f\t%.4f\t%.4f\t%d\n" % (
                ids[q].decode(), r.query_len[q], r.n_kmers[q], O.go_fmt_e4(float(h["fpr"])), len(rows), t.name.decode(), t.index & 0xFFFF, t.index >> 16,
                t.genome_size, r.k_used[q], int(h["count"]), float(h["qcov"]), float(h["tcov"]), float(h["jacc"]), q))
    exp.append("# input queries: %d\n# matched queries: %d\n# matched percentage: %.4f%%\n" % (len(ids), matched, matched / len(ids) * 100))
    got = open(out).read()
    assert got == "".join(exp)
    assert sum(1 for l in got.splitlines() if l.split("\t")[4:5] not in ([], ["0"], ["1"], ["hits"])) > 100      # queries with hits in both databases


# ---------------------------------------------------------------------------------------------------- sketches
@pytest.mark.parametrize("k,kw", [(31, dict(syncmer_s=15)), (31, dict(syncmer_s=15, scaled=True, scale=4)), (21, dict(syncmer_s=11)),
                                  (21, dict(syncmer_s=20)), (21, dict(syncmer_s=1)), (21, dict(minimizer_w=5)),
                                  (31, dict(minimizer_w=20, scaled=True, scale=3)), (21, dict(minimizer_w=1)), (15, dict(minimizer_w=100))])
def test_sketch_selection_matches_oracle(gpu_ctx, oracle, k, kw):
    """closed syncmer (windowed) and minimizer (A.5) selection kernels vs the oracle"""
    from kmcp_b200 import api
    O = oracle
    osp = O.sketch_params(k, **kw)
    reads = helpers.edge_reads(k) + helpers.make_reads(O, RSEED, 100, 40, 30000, GSEED) + [O.synth_genome(3, 1, 40000), b"ACGT" * 300, b"A" * 500]
    reads.append(O.synth_genome(3, 2, 3000)[:1500] + b"N" * 100 + O.synth_genome(3, 2, 3000)[1500:])
    buf, off = api.pack_seqs(reads)
    sp = api.SketchParams(k, 1, osp.scaled, osp.scale, osp.minimizer, osp.minimizer_w, osp.syncmer, osp.syncmer_s)
    codes, coff = gpu_ctx.generate_kmers(buf, off, sp)
    for i, r in enumerate(reads):
        exp = O.generate_kmers(r, osp)
        got = codes[int(coff[i]):int(coff[i + 1])]
        assert np.array_equal(got, exp), (k, kw, i, len(r), len(got), len(exp))


def test_golden_vectors_from_reference_data(gpu_ctx, oracle):
    """tests/golden/*.json were generated from the reference's demo files by the oracle after it reproduced G1-G5"""
    import json
    from kmcp_b200 import api
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    d = json.load(open(os.path.join(gold, "reads_k21.json")))
    reads = [r["seq"].encode() for r in d["reads"]]
    buf, off = api.pack_seqs(reads)
    codes, coff = gpu_ctx.generate_kmers(buf, off, api.SketchParams(21, 1, 0, 1, 0, 0, 0, 0))
    for i, r in enumerate(d["reads"]):
        assert [int(c) for c in codes[int(coff[i]):int(coff[i + 1])]] == r["codes"]
    sk = json.load(open(os.path.join(gold, "sketch_k31.json")))
    buf, off = api.pack_seqs([sk["seq"].encode()])
    for name, sp in (("scaled1000", api.SketchParams(31, 1, 1, 1000, 0, 0, 0, 0)), ("scaled50", api.SketchParams(31, 1, 1, 50, 0, 0, 0, 0)),
                     ("syncmer15_scaled62", api.SketchParams(31, 1, 1, 62, 0, 0, 1, 15)), ("syncmer15", api.SketchParams(31, 1, 0, 1, 0, 0, 1, 15)),
                     ("minimizer10", api.SketchParams(31, 1, 0, 1, 1, 10, 0, 0))):
        codes, _ = gpu_ctx.generate_kmers(buf, off, sp)
        assert [int(c) for c in codes] == sk["lists"][name], name
    L = api.load()
    for c in json.load(open(os.path.join(gold, "fpr.json"))):
        assert float(L.kmcpg_query_fpr(c["n"], c["c"], c["p"])).hex() == c["fpr_hex"]


@pytest.mark.parametrize("kw,h,fpr", [(dict(syncmer_s=15, scaled=True, scale=8), 3, 0.01), (dict(minimizer_w=8), 1, 0.3),
                                      (dict(scaled=True, scale=20), 3, 0.01)])
def test_sketch_database_search_end_to_end(gpu_ctx, oracle, tmp_path, kw, h, fpr):
    """genome-vs-genome search on sketch DBs (the demo-searching shape: -g queries, sort by jacc, -t 0.1)"""
    O = oracle
    k = 31
    sp = O.sketch_params(k, **kw)
    # related genomes: mutated copies of a common ancestor so containment varies
    base = O.synth_genome(21, 0, 60000)
    rng = np.random.default_rng(5)
    genomes = []
    for g in range(12):
        b = bytearray(base)
        for pos in rng.integers(0, len(b), int(len(b) * 0.004 * g)):
            b[pos] = b"ACGT"[int(rng.integers(0, 4))]
        genomes.append(bytes(b))
    targets = []
    for g, s in enumerate(genomes):
        targets += O.compute_targets([(b"g", b"g", s)], "genome_%02d" % g, sp)
    r001 = O.build_db(targets, str(tmp_path / "db"), sp, num_hashes=h, fpr=fpr, block_size=8)
    odb = O.DB(r001)
    gpu_ctx.open_db(r001)
    queries = [genomes[0], genomes[5], genomes[11], O.synth_genome(22, 3, 50000), genomes[3][:200]]
    got = _compare_engine(O, odb, gpu_ctx, queries, min_query_cov=0.1, sort_by=2)
    assert len(got.matches) >= 12 and got.n_kmers.max() > 256


# ---------------------------------------------------------------------------------------------------- GPU compute+index
def _write_fasta(path, records, gz=False):
    import gzip
    op = gzip.open if gz else open
    with op(path, "wb") as f:
        for hdr, s in records:
            f.write(b">" + hdr + b"\n")
            for j in range(0, len(s), 60):
                f.write(s[j:j + 60] + b"\n")


def _make_genome_files(O, tmp_path, n=10):
    rng = np.random.default_rng(11)
    files = []
    for g in range(n):
        L = int(rng.integers(15000, 40000))
        chrom = O.synth_genome(41, g, L)
        recs = [(b"chr1 Genome %d chromosome" % g, chrom[:L // 2].lower() if g % 3 == 0 else chrom[:L // 2]),
                (b"chr2 second contig", chrom[L // 2:]),
                (b"p1 Genome %d Plasmid pX" % g, O.synth_genome(42, g, 3000))]
        if g == 4:
            recs.append((b"tiny", b"ACGTACGTAC"))
        path = str(tmp_path / ("GCF_%06d.%d.fa%s" % (g, 1 + g % 2, ".gz" if g % 2 else "")))
        _write_fasta(path, recs, gz=bool(g % 2))
        files.append(path)
    return files


@pytest.mark.parametrize("mode", ["split_k21_h1", "nosplit_syncmer_h3", "nosplit_scaled_h2"])
def test_gpu_compute_index_matches_oracle_builder(gpu_ctx, oracle, tmp_path, mode):
    """kmcpg_index_fasta (GPU `kmcp compute` + `kmcp index`) writes the same database, byte for byte, as the oracle's
    restatement of compute/index on the same FASTA files, and answers queries identically"""
    import re
    O = oracle
    files = _make_genome_files(O, tmp_path)
    name_re = r"^([\w\.\_]+\.\d+)"
    if mode == "split_k21_h1":
        kw = dict(k=21, num_hashes=1, fpr=0.3, split_number=5, split_overlap=100, block_size=16)
        sp = O.sketch_params(21)
    elif mode == "nosplit_syncmer_h3":
        kw = dict(k=31, num_hashes=3, fpr=0.01, syncmer_s=15, scale=4, threads=2)
        sp = O.sketch_params(31, scaled=True, scale=4, syncmer_s=15)
    else:
        kw = dict(k=31, num_hashes=2, fpr=0.05, scale=10, block_size=8)
        sp = O.sketch_params(31, scaled=True, scale=10)
    # oracle side
    targets = []
    for f in files:
        name = re.match(name_re, os.path.basename(f)).group(1)
        targets += O.compute_targets(list(O.read_fastx(f)), name, sp, split_number=kw.get("split_number", 1), split_overlap=kw.get("split_overlap", 0),
                                     name_filters=["plasmid"])
    r001 = O.build_db(targets, str(tmp_path / "odb"), sp, num_hashes=kw["num_hashes"], fpr=kw["fpr"], block_size=kw.get("block_size", 0),
                      threads=kw.get("threads", 16))
    # device side
    out = str(tmp_path / "gdb")
    gpu_ctx.index_fasta(files, out, ref_name_regexp=name_re, seq_name_filters=["plasmid"], **kw)
    gfiles = sorted(f for f in os.listdir(os.path.join(out, "R001")) if f.endswith(".uniki"))
    ofiles = sorted(f for f in os.listdir(r001) if f.endswith(".uniki"))
    assert gfiles == ofiles and len(gfiles) >= 2
    for fn in gfiles:
        assert open(os.path.join(out, "R001", fn), "rb").read() == open(os.path.join(r001, fn), "rb").read(), fn
    # the written DB opens in the oracle and reports the same parameters
    odb_g = O.DB(os.path.join(out, "R001"))
    odb = O.DB(r001)
    for f in ("n_targets", "n_blocks", "num_hashes", "scaled", "scale", "syncmer", "syncmer_s", "fpr", "total_bytes"):
        assert getattr(odb_g.info, f) == getattr(odb.info, f), f
    # the context holds the new DB: search it straight away
    qs = [O.synth_genome(41, 2, 9000)[1000:8000], O.synth_genome(41, 7, 9000)[:3000]] + helpers.make_reads(O, 3, 50, 10, 15000, 41)
    _compare_engine(O, odb, gpu_ctx, qs, min_query_cov=0.2)
    # and it can be re-opened from disk
    gpu_ctx.open_db(os.path.join(out, "R001"))
    _compare_engine(O, odb, gpu_ctx, qs, min_query_cov=0.2)


def test_cli_index_then_search_roundtrip(oracle, tmp_path):
    """kmcp-gpu index (GPU compute+index) → kmcp-gpu search, against the oracle's compute/index/search/TSV on the same files"""
    import re
    import subprocess
    O = oracle
    files = _make_genome_files(O, tmp_path, n=8)
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "kmcp_b200", "kmcp-gpu")
    out = str(tmp_path / "cli.kmcp")
    name_re = r"^([\w\.\_]+\.\d+)"
    p = subprocess.run([exe, "index", "-q", "-O", out, "-k", "21", "-n", "4", "-l", "150", "-B", "plasmid", "-N", name_re, "-f", "0.3", "-b", "8"] + files,
                       capture_output=True)
    assert p.returncode == 0, p.stderr.decode()
    sp = O.sketch_params(21)
    targets = []
    for f in files:
        targets += O.compute_targets(list(O.read_fastx(f)), re.match(name_re, os.path.basename(f)).group(1), sp, split_number=4, split_overlap=150,
                                     name_filters=["plasmid"])
    r001 = O.build_db(targets, str(tmp_path / "odb"), sp, num_hashes=1, fpr=0.3, block_size=8)
    for fn in sorted(f for f in os.listdir(r001) if f.endswith(".uniki")):
        assert open(os.path.join(out, "R001", fn), "rb").read() == open(os.path.join(r001, fn), "rb").read(), fn
    reads = helpers.make_reads(O, 13, 600, 8, 15000, 41)
    ids = [b"r%d" % i for i in range(len(reads))]
    fq = str(tmp_path / "q.fq.gz")
    _write_fastq(fq, ids, reads)
    tsv = str(tmp_path / "o.tsv")
    _run_cli(["-d", out, fq, "-o", tsv])
    odb = O.DB(r001)
    assert open(tsv).read() == O.format_tsv(odb, ids, odb.search(reads))


def test_engine_big_round_filters_parts_while_the_gpu_works(gpu_ctx, oracle, small_db):
    """800 k queries = several parts: the engine filters every delivered part inside the executor's callback"""
    O = oracle
    odb = O.DB(small_db)
    gpu_ctx.open_db(small_db)
    base = helpers.make_reads(O, RSEED + 30, 4000, 40, 30000, GSEED) + helpers.edge_reads(21)
    reads = (base * 200)[:800000]
    got = _compare_engine(O, odb, gpu_ctx, reads)
    assert len(got.matches) > 400000


def test_streaming_parts_equal_the_batch_result(gpu_ctx, oracle, small_db):
    """kmcpg_search_batch_cb delivers the same hits, part by part and in query order"""
    from kmcp_b200 import api
    O = oracle
    gpu_ctx.open_db(small_db)
    reads = (helpers.make_reads(O, RSEED + 40, 3000, 40, 30000, GSEED) * 120)[:330000]      # several parts
    buf, off = api.pack_seqs(reads)
    whole = gpu_ctx.search_batch(buf, off)
    parts, summary = gpu_ctx.search_batch_streaming(buf, off)
    assert len(parts) >= 2 and parts[0][0] == 0
    assert sum(p[1] for p in parts) == len(reads) and all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(len(parts) - 1))
    assert np.array_equal(np.concatenate([p[3] for p in parts]), whole.hits)
    assert np.array_equal(np.concatenate([p[2] for p in parts]), whole.n_kmers)
    assert np.array_equal(summary.hits, whole.hits)


def test_degenerate_batches(gpu_ctx, oracle, small_db):
    """empty batch, batch of empty / too-short sequences only, a single read"""
    from kmcp_b200 import api
    O = oracle
    odb = O.DB(small_db)
    gpu_ctx.open_db(small_db)
    r = gpu_ctx.search_batch(np.zeros(1, np.uint8), np.zeros(1, np.uint64))
    assert len(r.hits) == 0 and len(r.n_kmers) == 0
    e = gpu_ctx.engine_search(np.zeros(1, np.uint8), np.zeros(1, np.uint64))
    assert len(e.matches) == 0 and len(e.match_off) == 1
    _compare_engine(O, odb, gpu_ctx, [b"", b"ACGT", b""])
    _compare_engine(O, odb, gpu_ctx, helpers.make_reads(O, 77, 1, 40, 30000, GSEED))


# ---------------------------------------------------------------------------------------------------- full size
def test_full_size_c2_properties_and_sampled_oracle_parity(oracle):
    """BASELINE configs[1] at its full size (10,000 targets, 1.4 GB index, 1 M x 150 bp reads): size-independent
    properties of the search (determinism, batch-split and order invariance, strand symmetry of canonical k-mers,
    threshold and ordering invariants) plus exact parity with the CPU oracle on a random sample of the same batch
    over the same (dumped) index."""
    import shutil
    import bench
    from kmcp_b200 import api
    O = oracle
    if bench.SCALE != "full":
        pytest.skip("KMCP_BENCH_SCALE is not 'full'")
    NR, RL = bench.READS_PER_STEP, bench.READ_LEN
    tmp = "/dev/shm/kmcp_full_parity" if os.path.isdir("/dev/shm") else "/tmp/kmcp_full_parity"
    with api.Context(0) as ctx:
        ctx.build_synth_db(bench.GENOME_SEED, bench.N_GENOMES, bench.GENOME_LEN, k=bench.K, n_chunks=bench.N_CHUNKS, overlap=bench.OVERLAP,
                           num_hashes=bench.H, fpr=bench.FPR, block_size=bench.BLOCK_SIZE)
        info = ctx.db_info()
        assert info.n_targets == 10000 and info.resident_bytes > 1.3e9
        d = ctx.device_alloc(NR * RL)
        ctx.synth_reads(bench.READ_SEED, 0, NR, RL, bench.GENOME_SEED, bench.N_GENOMES, bench.GENOME_LEN, d)
        reads = np.frombuffer(ctx.d2h(d, NR * RL), dtype=np.uint8).reshape(NR, RL).copy()
        ctx.device_free(d)
        off = np.arange(NR + 1, dtype=np.uint64) * np.uint64(RL)
        p = ctx.default_params()
        whole = ctx.search_batch(reads.reshape(-1), off, p)
        h = whole.hits
        # invariants: ordered by (query, target); MinMatched and the strict qCov threshold; every query with k-mers counted
        key = h["query"].astype(np.uint64) << np.uint64(32) | h["target"].astype(np.uint64)
        assert len(h) > 500_000 and np.all(key[1:] > key[:-1])
        n_of = whole.n_kmers[h["query"]].astype(np.float64)
        assert np.all(h["count"] >= p.min_matched) and np.all(h["count"].astype(np.float64) > n_of * p.min_query_cov) and np.all(h["count"] <= n_of)
        assert np.all(whole.n_kmers == RL - bench.K + 1)
        # determinism
        again = ctx.search_batch(reads.reshape(-1), off, p)
        assert np.array_equal(again.hits, h)
        # batch-split invariance
        half = NR // 2
        a = ctx.search_batch(reads[:half].reshape(-1), off[:half + 1], p)
        b = ctx.search_batch(reads[half:].reshape(-1), off[:NR - half + 1], p)
        hb = b.hits.copy(); hb["query"] += half
        assert np.array_equal(np.concatenate([a.hits, hb]), h)
        # order invariance: reversed batch
        rv = ctx.search_batch(reads[::-1].copy().reshape(-1), off, p)
        hr = rv.hits.copy(); hr["query"] = NR - 1 - hr["query"]
        hr = hr[np.lexsort((hr["target"], hr["query"]))]
        assert np.array_equal(hr, h)
        # strand symmetry: canonical k-mers make the reverse complement of every read an identical query
        comp = np.zeros(256, np.uint8); comp[:] = ord("N")
        for x, y in zip(b"ACGT", b"TGCA"):
            comp[x] = y
        rc = ctx.search_batch(comp[reads[:, ::-1]].reshape(-1), off, p)
        assert np.array_equal(rc.hits, h) and np.array_equal(rc.n_kmers, whole.n_kmers)
        # exact parity on a sample, CPU oracle over the same index bytes
        shutil.rmtree(tmp, ignore_errors=True)
        try:
            r001 = bench.dump_db_for_cpu(ctx, tmp)
            odb = O.DB(r001)
            idx = np.sort(np.random.default_rng(7).choice(NR, 3000, replace=False))
            sample = [reads[i].tobytes() for i in idx]
            ores = odb.search(sample, algo=1, threads=os.cpu_count())
            er = ctx.engine_search(reads[idx].reshape(-1), off[:len(idx) + 1])
            assert np.array_equal(er.match_off, ores.hit_off)
            for f in ("target", "count", "fpr", "qcov", "tcov", "jacc"):
                assert np.array_equal(er.matches[f], ores.hits[f]), f
            # and the integer hit list of the whole batch restricted to the sample agrees with the oracle's pre-float-filter hits
            assert len(ores.hits) > 2000
        finally:
            shutil.rmtree(tmp, ignore_errors=True)


def test_c4_shape_wide_rows_many_blocks_sampled_oracle_parity():
    """BASELINE configs[3] shape on one GPU: 852,050 targets, h=3, 32 blocks of 3,329-byte rows (genome length scaled
    down as SURVEY §8(d) C4 allows): engine output of a read sample identical to the CPU oracle over the same index"""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, NR="20000", NCHK="200", GL="50000")
    p = subprocess.run([sys.executable, os.path.join(root, "tools", "c4_shape.py")], capture_output=True, env=env, timeout=900)
    assert p.returncode == 0, p.stderr.decode()[-2000:]
    r = json.loads(p.stdout.decode().strip().splitlines()[-1])
    assert r["db"]["targets"] == 852050 and r["db"]["blocks"] == 32 and r["probe_launches"] >= 32
    assert r["probe_row_bytes_per_read"] == 130 * 3 * sum(-(-n // 8) for n in [26632] * 31 + [852050 - 31 * 26632])
    assert r["oracle_sample"]["identical"] and r["oracle_sample"]["hits"] > 100


def test_refcounts_from_engine_batches(gpu_ctx, oracle, small_db):
    """kmcpg_refcounts_add on the engine's own result batches == the oracle's profile stage 1 over the oracle's TSV"""
    from kmcp_b200 import api
    O = oracle
    odb = O.DB(small_db)
    gpu_ctx.open_db(small_db)
    reads = helpers.make_reads(O, RSEED + 60, 3000, 40, 30000, GSEED)
    ids = [b"r%d" % i for i in range(len(reads))]
    exp_reads, exp = O.profile_stage1(O.format_tsv(odb, ids, odb.search(reads)), min_qcov=0.6, max_fpr=0.005, hic_min_qcov=0.8)
    rc = gpu_ctx.refcounts_create(min_query_cov=0.6, max_fpr=0.005, hic_min_qcov=0.8)
    try:
        for lo in range(0, len(reads), 1100):                      # three engine batches, in input order
            buf, off = api.pack_seqs(reads[lo:lo + 1100])
            gpu_ctx.engine_search(buf, off, refcounts=rc)
        got_reads, got = gpu_ctx.refcounts_get(rc)
    finally:
        gpu_ctx.refcounts_free(rc)
    assert got_reads == exp_reads and got == exp and exp_reads > 1500

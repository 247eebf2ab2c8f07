"""One database sharded over several contexts of ONE process: whole blocks per shard, column ranges of a
block when the DB has fewer blocks than shards, and the engine that merges the shards' hit lists.  Every shard context
lives on cuda:0 here (a context is a context: the library never assumes one context per device), so these run on a
one-GPU box; bit-exact against the one-context result and against the oracle."""
import os

import numpy as np
import pytest

import parity_helpers as helpers

pytestmark = pytest.mark.gpu

GSEED, RSEED = 11, 23


@pytest.fixture(scope="module")
def small_db(oracle, tmp_path_factory):
    """40 genomes x 30 kb, 5 chunks -> 200 targets; h=1; 4 blocks of 64"""
    O = oracle
    sp = O.sketch_params(21)
    targets = helpers.make_synth_targets(O, sp, GSEED, 40, 30000, 5, 150)
    return O.build_db(targets, str(tmp_path_factory.mktemp("db_small")), sp, num_hashes=1, fpr=0.3, block_size=64)


@pytest.fixture(scope="module")
def wide_db(oracle, tmp_path_factory):
    """150 genomes x 12 kb, 10 chunks -> 1500 targets in ONE block (188-byte rows, 12 column units of 128 targets); h=3"""
    O = oracle
    sp = O.sketch_params(21)
    targets = helpers.make_synth_targets(O, sp, GSEED + 1, 150, 12000, 10, 100)
    return O.build_db(targets, str(tmp_path_factory.mktemp("db_wide")), sp, num_hashes=3, fpr=0.1, block_size=1500)


@pytest.fixture(scope="module")
def three_db(oracle, tmp_path_factory):
    """the same 1500 targets in three blocks (640 + 640 + 220 columns); h=2"""
    O = oracle
    sp = O.sketch_params(21)
    targets = helpers.make_synth_targets(O, sp, GSEED + 1, 150, 12000, 10, 100)
    return O.build_db(targets, str(tmp_path_factory.mktemp("db_three")), sp, num_hashes=2, fpr=0.2, block_size=640)


def _open_shards(api, r001, world):
    ctxs = []
    for rank in range(world):
        c = api.Context(0)
        c.open_db(r001, shard_rank=rank, shard_world=world)
        ctxs.append(c)
    return ctxs


def _close(ctxs):
    for c in ctxs:
        c.close()


@pytest.mark.parametrize("dbname,ng,gl,worlds", [("wide_db", 150, 12000, (2, 3, 5, 12)), ("three_db", 150, 12000, (4, 7))])
def test_column_range_shards_union_equals_whole(oracle, request, dbname, ng, gl, worlds):
    """fewer blocks than shards: every shard keeps a column range; hit lists stay disjoint by target and their union, the
    dense counts and the residency flags all add up to the one-context database"""
    from kmcp_b200 import api
    O = oracle
    r001 = request.getfixturevalue(dbname)
    odb = O.DB(r001)
    reads = helpers.make_reads(O, RSEED + 4, 1500, ng, gl, GSEED + 1) + helpers.edge_reads(21)
    buf, off = api.pack_seqs(reads)
    codes = np.concatenate([O.generate_kmers(r, odb.sketch_params()) for r in reads[:5]])
    with api.Context(0) as whole:
        whole.open_db(r001)
        ref = whole.search_batch(buf, off)
        ref_counts = whole.count_codes(codes)
        winfo = whole.db_info()
    assert len(ref.hits) > 1000
    assert np.array_equal(ref_counts, odb.count_codes(codes))
    for world in worlds:
        pieces = api.shard_pieces(r001, world)
        ctxs = _open_shards(api, r001, world)
        try:
            parts, counts, resident = [], np.zeros_like(ref_counts), np.zeros(winfo.n_targets, np.int32)
            row_bytes = disk = 0
            for rank, c in enumerate(ctxs):
                mine = [p for p in pieces if p[1] == rank]
                info = c.db_info()
                assert info.n_resident_blocks == len(mine) and info.n_targets == winfo.n_targets
                assert info.sum_row_bytes == sum((p[3] + 7) // 8 for p in mine)
                row_bytes += info.sum_row_bytes; disk += info.disk_bytes
                if not mine:
                    continue
                r = c.search_batch(buf, off)
                assert np.array_equal(r.n_kmers, ref.n_kmers) and np.array_equal(r.query_len, ref.query_len)
                n_sum = int(r.n_kmers.astype(np.int64).sum())
                assert r.probe_row_bytes == n_sum * odb.info.num_hashes * info.sum_row_bytes
                parts.append(r.hits)
                cc = c.count_codes(codes)
                counts += cc
                res = np.array([c.target(g).resident for g in range(winfo.n_targets)], np.int32)
                assert np.all(cc[res == 0] == 0)
                resident += res
            assert np.all(resident == 1)                              # every target resident in exactly one shard
            assert np.array_equal(counts, ref_counts)
            assert row_bytes >= winfo.sum_row_bytes and row_bytes <= winfo.sum_row_bytes + len(pieces)
            allh = np.concatenate(parts)
            assert len(allh) == len(ref.hits)
            assert np.array_equal(allh[np.lexsort((allh["target"], allh["query"]))], ref.hits)
            assert np.array_equal(api.merge_hit_lists(parts), ref.hits)
        finally:
            _close(ctxs)


def _same_results(a, b):
    assert np.array_equal(a.query_len, b.query_len) and np.array_equal(a.n_kmers, b.n_kmers) and np.array_equal(a.k_used, b.k_used)
    assert np.array_equal(a.match_off, b.match_off)
    assert np.array_equal(a.matches, b.matches)                       # every column, floats bit for bit, same order


@pytest.mark.parametrize("dbname,ng,gl,gs,world", [("small_db", 40, 30000, GSEED, 2), ("small_db", 40, 30000, GSEED, 3),
                                                   ("wide_db", 150, 12000, GSEED + 1, 4), ("three_db", 150, 12000, GSEED + 1, 5)])
def test_sharded_engine_equals_one_context_and_the_oracle(oracle, request, dbname, ng, gl, gs, world):
    from kmcp_b200 import api
    O = oracle
    r001 = request.getfixturevalue(dbname)
    odb = O.DB(r001)
    reads = helpers.make_reads(O, RSEED + 6, 2500, ng, gl, gs) + helpers.edge_reads(21)
    buf, off = api.pack_seqs(reads)
    r1 = helpers.make_reads(O, RSEED + 1, 600, ng, gl, gs)
    r2 = helpers.make_reads(O, RSEED + 2, 600, ng, gl, gs)
    r2[5] = b"ACGT"; r1[6] = b"ACGTACGT"; r1[7] = b""; r2[7] = b""
    pbuf, poff = api.pack_seqs([x for p in zip(r1, r2) for x in p])
    variants = [dict(), dict(min_query_cov=0.3, sort_by=1), dict(min_query_cov=0.7, sort_by=2, top_n_scores=1), dict(do_not_sort=1),
                dict(min_query_cov=0.2, min_matched=3, min_target_cov=0.002, max_fpr=1e-6), dict(dedup_threshold=50)]
    pvariants = [dict(paired=1), dict(paired=1, try_se=1, min_query_cov=0.6)]
    with api.Context(0) as whole:
        whole.open_db(r001)
        ctxs = _open_shards(api, r001, world)
        try:
            live = [c for c in ctxs if c.db_info().n_resident_blocks > 0]
            assert len(live) >= 2
            for kw in variants:
                one = whole.engine_search(buf, off, whole.default_engine_opts(**kw))
                many = live[0].engine_search(buf, off, live[0].default_engine_opts(**kw), shards=live[1:])
                _same_results(many, one)
                assert many.kernel_launches > one.kernel_launches and many.probe_row_bytes >= one.probe_row_bytes
            for kw in pvariants:
                one = whole.engine_search(pbuf, poff, whole.default_engine_opts(**kw))
                many = live[0].engine_search(pbuf, poff, live[0].default_engine_opts(**kw), shards=live[1:])
                _same_results(many, one)
            # and against the oracle directly (defaults)
            exp = odb.search(reads)
            got = live[0].engine_search(buf, off, shards=live[1:])
            assert np.array_equal(got.match_off, exp.hit_off) and len(got.matches) > 1200
            for f in ("query", "target", "count", "fpr", "qcov", "tcov", "jacc"):
                assert np.array_equal(got.matches[f], exp.hits[f]), f
            # degenerate batches
            e = live[0].engine_search(np.zeros(1, np.uint8), np.zeros(1, np.uint64), shards=live[1:])
            assert len(e.matches) == 0 and len(e.match_off) == 1
            sb, so = api.pack_seqs([b"", b"ACGT", b""])
            _same_results(live[0].engine_search(sb, so, shards=live[1:]), whole.engine_search(sb, so))
            # the same context twice is refused, not deadlocked
            with pytest.raises(api.KmcpGpuError) as err:
                live[0].engine_search(buf, off, shards=[live[0]])
            assert err.value.code == api.KMCPG_EINVAL
        finally:
            _close(ctxs)


def test_sharded_engine_several_parts(oracle, small_db):
    """400 k queries = several parts per shard: parts of all shards are merged one by one while the device calls run on"""
    from kmcp_b200 import api
    O = oracle
    base = helpers.make_reads(O, RSEED + 30, 4000, 40, 30000, GSEED) + helpers.edge_reads(21)
    reads = (base * 100)[:400000]
    buf, off = api.pack_seqs(reads)
    with api.Context(0) as whole:
        whole.open_db(small_db)
        one = whole.engine_search(buf, off)
        ctxs = _open_shards(api, small_db, 3)
        try:
            many = ctxs[0].engine_search(buf, off, shards=ctxs[1:])
            _same_results(many, one)
            assert len(many.matches) > 200000
        finally:
            _close(ctxs)


def test_replicas_engine_equals_one_context(oracle, small_db, wide_db):
    """the whole database on every context, the reads split between them: concatenated answers == the one-context answer"""
    from kmcp_b200 import api
    O = oracle
    for r001, ng, gl, gs, n_rep in ((small_db, 40, 30000, GSEED, 3), (wide_db, 150, 12000, GSEED + 1, 2)):
        rng = np.random.default_rng(9)
        reads = helpers.make_reads(O, RSEED + 12, 3000, ng, gl, gs) + helpers.edge_reads(21)
        long_read = O.synth_genome(gs, 1, gl)[:9000]                   # one long query among the short ones unbalances the ranges
        reads.insert(700, long_read)
        buf, off = api.pack_seqs(reads)
        r1 = helpers.make_reads(O, RSEED + 1, 500, ng, gl, gs)
        r2 = helpers.make_reads(O, RSEED + 2, 500, ng, gl, gs)
        r2[5] = b"ACGT"; r1[6] = b""; r2[6] = b""
        pbuf, poff = api.pack_seqs([x for p in zip(r1, r2) for x in p])
        ctxs = []
        try:
            for _ in range(n_rep):
                c = api.Context(0)
                c.open_db(r001)
                ctxs.append(c)
            for kw in (dict(), dict(min_query_cov=0.7, sort_by=2, top_n_scores=1), dict(do_not_sort=1), dict(dedup_threshold=50)):
                one = ctxs[0].engine_search(buf, off, ctxs[0].default_engine_opts(**kw))
                many = ctxs[0].engine_search(buf, off, ctxs[0].default_engine_opts(**kw), replicas=ctxs[1:])
                _same_results(many, one)
            assert len(one.matches) > 1000
            for kw in (dict(paired=1), dict(paired=1, try_se=1, min_query_cov=0.6)):
                one = ctxs[0].engine_search(pbuf, poff, ctxs[0].default_engine_opts(**kw))
                many = ctxs[0].engine_search(pbuf, poff, ctxs[0].default_engine_opts(**kw), replicas=ctxs[1:])
                _same_results(many, one)
            # fewer queries than replicas, and an empty batch
            sb, so = api.pack_seqs(reads[:2])
            _same_results(ctxs[0].engine_search(sb, so, replicas=ctxs[1:]), ctxs[0].engine_search(sb, so))
            e = ctxs[0].engine_search(np.zeros(1, np.uint8), np.zeros(1, np.uint64), replicas=ctxs[1:])
            assert len(e.matches) == 0 and len(e.match_off) == 1
            # a context that holds only a shard cannot be a replica
            with api.Context(0) as sh:
                sh.open_db(r001, shard_rank=0, shard_world=2)
                with pytest.raises(api.KmcpGpuError) as err:
                    ctxs[0].engine_search(buf, off, replicas=[sh])
                assert err.value.code == api.KMCPG_EINVAL
            free, total = ctxs[0].device_memory()
            assert 0 < free <= total
        finally:
            _close(ctxs)


def test_cli_gpus_flag_shards_or_replicates_the_database(oracle, small_db, wide_db, tmp_path):
    """kmcp-gpu search --gpus 0,0,0 (three contexts, here all on device 0) in every --gpu-mode gives the byte-identical TSV"""
    from test_gpu_parity import _run_cli, _write_fastq
    O = oracle
    for r001, ng, gl, gs in ((small_db, 40, 30000, GSEED), (wide_db, 150, 12000, GSEED + 1)):
        odb = O.DB(r001)
        reads = helpers.make_reads(O, RSEED + 20, 1500, ng, gl, gs) + [r for r in helpers.edge_reads(21) if len(r) > 0]
        ids = [b"read_%d/1" % i for i in range(len(reads))]
        fq = str(tmp_path / "q.fq")
        _write_fastq(fq, ids, reads, gz=False)
        exp = O.format_tsv(odb, ids, odb.search(reads), keep_unmatched=True)
        one = str(tmp_path / "one.tsv")
        _run_cli(["-d", os.path.dirname(r001), fq, "-o", one, "-K"])
        assert open(one).read() == exp
        for mode in ("shard", "replicate", "auto"):
            many = str(tmp_path / ("many_%s.tsv" % mode))
            _run_cli(["-d", os.path.dirname(r001), fq, "-o", many, "-K", "--gpus", "0,0,0", "--gpu-mode", mode])
            assert open(many).read() == exp, mode


def test_plain_c_host_prints_the_same_hits(oracle, small_db, tmp_path):
    """examples/search_host.c (strict C99, no Python in the loop) against the ctypes binding on the same database and reads"""
    import subprocess
    from test_abi import _build_c_host
    from kmcp_b200 import api
    O = oracle
    reads = helpers.make_reads(O, RSEED + 8, 40, 40, 30000, GSEED)
    buf, off = api.pack_seqs(reads)
    with api.Context(0) as ctx:
        ctx.open_db(small_db)
        r = ctx.search_batch(buf, off)
        exp = "".join("%d\t%s\t%d\t%d\t%d\n" % (h["query"], ctx.target(int(h["target"])).name.decode(), ctx.target(int(h["target"])).index & 0xFFFF,
                                                h["count"], r.n_kmers[int(h["query"])]) for h in r.hits)
    assert len(r.hits) > 20
    p = subprocess.run([_build_c_host(tmp_path), small_db] + [x.decode() for x in reads], capture_output=True)
    assert p.returncode == 0, p.stderr.decode()
    assert p.stdout.decode() == exp

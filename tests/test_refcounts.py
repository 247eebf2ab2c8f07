"""`kmcp profile` stage-1 counters: the C++ accumulator fed with match lists against the oracle's
restatement of profile.go:761-990 fed with the search TSV.  Host code only (no GPU): the accumulator reads the block
headers of the database directory; the match lists come from the oracle's search."""
import numpy as np
import pytest

import parity_helpers as helpers


@pytest.fixture(scope="module")
def related_db(oracle, tmp_path_factory):
    """8 references in 5 chunks; references 1..3 are mutated copies of reference 0 (2 %, 6 %, 12 % substitutions) and
    reference 5 shares a 6 kb segment with reference 4, so reads hit several references and several chunks"""
    O = oracle
    sp = O.sketch_params(21)
    rng = np.random.default_rng(5)
    base = [bytearray(O.synth_genome(77, g, 20000)) for g in range(8)]
    for j, rate in ((1, 0.02), (2, 0.06), (3, 0.12)):
        g = bytearray(base[0])
        for pos in np.nonzero(rng.random(len(g)) < rate)[0]:
            g[pos] = b"ACGT"[(b"ACGT".index(g[pos]) + 1 + int(rng.integers(3))) % 4]
        base[j] = g
    base[5][3000:9000] = base[4][10000:16000]
    targets = []
    for g, seq in enumerate(base):
        targets += O.compute_targets([(b"s%d" % g, b"s%d" % g, bytes(seq))], "ref_%02d" % g, sp, split_number=5, split_overlap=150)
    out = str(tmp_path_factory.mktemp("db_related"))
    r001 = O.build_db(targets, out, sp, num_hashes=1, fpr=0.3, block_size=16)
    reads = []
    for i in range(6000):
        g = int(rng.integers(8))
        p = int(rng.integers(0, 20000 - 150))
        r = bytearray(base[g][p:p + 150])
        for pos in np.nonzero(rng.random(150) < 0.01)[0]:
            r[pos] = b"ACGT"[(b"ACGT".index(r[pos]) + 1) % 4]
        reads.append(bytes(r))
    reads += [O.synth_read(9, i, 8, 20000, 150, 1234) for i in range(300)]          # unrelated reads
    return r001, reads


PARAM_SETS = [
    dict(),                                                     # profile defaults: -t 0.55 -f 0.01
    dict(min_query_cov=0.0, max_fpr=1.0),
    dict(min_query_cov=0.7, max_fpr=1e-3, hic_min_qcov=0.9),
    dict(top_n_scores=1), dict(top_n_scores=2, min_query_cov=0.6),
    dict(keep_perfect=1), dict(keep_main=1, max_qcov_gap=0.1), dict(keep_main=1, max_qcov_gap=0.4, top_n_scores=3),
]


@pytest.mark.parametrize("kw", PARAM_SETS)
def test_refcounts_equal_profile_stage1_on_the_tsv(oracle, related_db, kw):
    from kmcp_b200 import api
    O = oracle
    r001, reads = related_db
    odb = O.DB(r001)
    ids = [b"read%d" % i for i in range(len(reads))]
    oo = O.default_opts(); oo.min_query_cov = 0.4; oo.max_fpr = 1.0         # loose search, so that profile's own filters bite
    res = odb.search(reads, opts=oo)
    nh = np.diff(res.hit_off.astype(np.int64))
    assert (nh > 1).sum() > 1000 and (nh == 0).sum() > 100
    names = dict(min_qcov="min_query_cov", keep_perfect="keep_perfect")
    okw = dict(min_qcov=kw.get("min_query_cov", 0.55), max_fpr=kw.get("max_fpr", 0.01), top_n_scores=kw.get("top_n_scores", 0),
               keep_perfect=bool(kw.get("keep_perfect", 0)), keep_main=bool(kw.get("keep_main", 0)), max_qcov_gap=kw.get("max_qcov_gap", 0.4),
               hic_min_qcov=kw.get("hic_min_qcov", 0.75))
    exp_reads, exp = O.profile_stage1(O.format_tsv(odb, ids, res), **okw)
    rc = api.refcounts_create(None, r001, **kw)
    try:
        # two batches, cut at a query boundary, as the engine would deliver them
        cut = len(reads) // 3
        a = int(res.hit_off[cut])
        api.refcounts_add_matches(rc, res.hit_off[:cut + 1], res.hits[:a])
        api.refcounts_add_matches(rc, res.hit_off[cut:] - res.hit_off[cut], res.hits[a:])
        got_reads, got = api.refcounts_get(rc)
    finally:
        api.refcounts_free(rc)
    assert got_reads == exp_reads and exp_reads > 3000
    assert set(got) == set(exp)
    for name, (gsize, match, uniq, hic) in exp.items():
        assert got[name] == (gsize, match, uniq, hic), name             # doubles compared exactly: same additions in the same order
    assert sum(sum(t[2]) for t in got.values()) > 500


def test_refcounts_errors(oracle, related_db, tmp_path):
    from kmcp_b200 import api
    with pytest.raises(api.KmcpGpuError):
        api.refcounts_create(None, str(tmp_path / "nope"))
    r001, _ = related_db
    rc = api.refcounts_create(None, r001)
    try:
        bad = np.zeros(1, api.MATCH_DTYPE); bad["target"] = 10_000; bad["qcov"] = 1.0
        with pytest.raises(api.KmcpGpuError):
            api.refcounts_add_matches(rc, np.array([0, 1], np.uint64), bad)
    finally:
        api.refcounts_free(rc)

"""Host-only test of the engine's result handling (reference U:7466-7491 tCov / FPR / derived columns,
U:273-311 sort + top-N): `kmcpg_internal_engine_standin` runs kmcpg_engine_search's own code over a hit list that is handed in
instead of coming from the device — here the oracle's device-contract hits (count >= -c and count > n*t only) — and the result
must equal the oracle's full search bit for bit.  The GPU tests make the same comparison through the device (test_gpu_parity.py)."""
import ctypes as C

import numpy as np
import pytest

import parity_helpers as helpers

GSEED, RSEED = 4242, 99


@pytest.fixture(scope="module")
def db(oracle, tmp_path_factory):
    O = oracle
    sp = O.sketch_params(21)
    targets = helpers.make_synth_targets(O, sp, GSEED, 40, 30000, 5, 150)
    return O.build_db(targets, str(tmp_path_factory.mktemp("db_engine_host")), sp, num_hashes=1, fpr=0.3, block_size=64)


def _standin(O, odb, reads, part_queries, paired=False, threads=0, **opts):
    from kmcp_b200 import api
    L = api.load()
    f = L.kmcpg_internal_engine_standin
    f.argtypes = [C.POINTER(api.EngineOpts), C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int64, C.c_double,
                  C.c_int, C.c_uint32, C.POINTER(api.Results)]
    oo = O.default_opts()
    for k, v in opts.items():
        setattr(oo, k, v)
    exp = odb.search(reads, paired=paired, opts=oo)
    # what the device delivers: the integer thresholds only, every hit, sorted by (query, target)
    ro = O.default_opts()
    for k, v in opts.items():
        setattr(ro, k, v)
    ro.min_target_cov, ro.max_fpr, ro.do_not_sort, ro.top_n_scores = 0.0, 1.0, 1, 0
    raw = odb.search(reads, paired=paired, opts=ro)
    order = np.lexsort((raw.hits["target"], raw.hits["query"]))
    hits = np.zeros(len(order), dtype=api.HIT_DTYPE)
    for fld in ("query", "target", "count"):
        hits[fld] = raw.hits[fld][order]
    nt = odb.info.n_targets
    tsize = np.array([odb.target(g).n_kmers for g in range(nt)], dtype=np.float64)
    eo = api.EngineOpts()
    L.kmcpg_default_engine_opts(C.byref(eo))
    for k, v in opts.items():
        setattr(eo, k, v)
    eo.paired, eo.threads = int(paired), threads
    nk = np.ascontiguousarray(raw.n_kmers, dtype=np.int32)
    ql = np.ascontiguousarray(raw.query_len, dtype=np.int32)
    r = api.Results()
    rc = f(C.byref(eo), len(nk), nk.ctypes.data, ql.ctypes.data, hits.ctypes.data, len(hits), tsize.ctypes.data, nt, odb.info.fpr, odb.k,
           part_queries, C.byref(r))
    assert rc == 0
    nq = r.n_queries
    got_off = api._np_from(r.match_off, nq + 1, 8, np.uint64)
    got = api._np_from(r.matches, r.n_matches, C.sizeof(api.Match), api.MATCH_DTYPE)
    assert np.array_equal(api._np_from(r.query_len, nq, 4, np.int32), exp.query_len)
    assert np.array_equal(api._np_from(r.n_kmers, nq, 4, np.int32), exp.n_kmers)
    assert np.all(api._np_from(r.k_used, nq, 4, np.int32) == odb.k)
    L.kmcpg_free_results(C.byref(r))
    assert np.array_equal(got_off, exp.hit_off)
    for fld in ("query", "target", "count", "fpr", "qcov", "tcov", "jacc"):
        assert np.array_equal(got[fld], exp.hits[fld]), fld           # floats compared bit for bit
    return got


def test_engine_result_handling_equals_the_oracle_without_a_device(oracle, db):
    O = oracle
    odb = O.DB(db)
    reads = helpers.make_reads(O, RSEED, 6000, 40, 30000, GSEED) + helpers.edge_reads(odb.k)
    got = _standin(O, odb, reads, part_queries=1 << 20)
    assert len(got) > 3000
    # the same answer whatever the parts and the number of filter threads
    for pq, th in ((1, 1), (7, 3), (1000, 0), (4097, 16)):
        _standin(O, odb, reads, part_queries=pq, threads=th)
    few = reads[:700] + helpers.edge_reads(odb.k)
    _standin(O, odb, few, 250, min_query_cov=0.3, sort_by=1)
    _standin(O, odb, few, 250, min_query_cov=0.7, sort_by=2, top_n_scores=1)
    _standin(O, odb, few, 250, min_query_cov=0.2, sort_by=0, top_n_scores=2)
    _standin(O, odb, few, 250, min_query_cov=0.2, min_matched=3, min_target_cov=0.002, max_fpr=1e-6)
    _standin(O, odb, few, 250, do_not_sort=1, min_query_len=100)
    _standin(O, odb, few, 250, do_not_sort=1, top_n_scores=1)          # -n is ignored with -S
    _standin(O, odb, few, 250, min_query_cov=0.0)
    _standin(O, odb, few, 250, min_query_cov=1.0)
    r1 = helpers.make_reads(O, RSEED + 1, 500, 40, 30000, GSEED)
    r2 = helpers.make_reads(O, RSEED + 2, 500, 40, 30000, GSEED)
    r2[5] = b"ACGT"; r1[6] = b"ACGTACGT"; r1[7] = b""; r2[7] = b""
    _standin(O, odb, [x for p in zip(r1, r2) for x in p], 100, paired=True)
    # degenerate batches
    _standin(O, odb, [b""], 10)
    _standin(O, odb, [b"ACGT" * 5] * 3, 1)
    odb.close()

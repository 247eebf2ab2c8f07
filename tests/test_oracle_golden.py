"""Pins the CPU oracle to the reference's own published outputs.
Runs only where /root/reference exists (the build container); the GPU box uses tests/golden/ instead."""
import glob
import os
import re

import numpy as np
import pytest

REF = "/root/reference"
pytestmark = [pytest.mark.reference_data,
              pytest.mark.skipif(not os.path.isdir(REF + "/demo-profiling/refs"), reason="reference demo data not present")]

G5_SIZES = {4903501, 4395762, 4857450, 5545864, 4659463, 2062405, 2881440, 2485591, 1890645, 2422602, 2755072, 1868883, 2117637,
            3980848, 6588339}     # demo-profiling/mock.kmcp.gz.kmcp.profile, column refsize
G1_UNAMBIGUOUS = [63665, 57300, 27831, 17898, 17152, 3433, 2951, 2417, 2182, 1599, 1445, 391, 259, 208, 180]   # ANALYSIS.md:44-58


@pytest.fixture(scope="module")
def demo_db(oracle, tmp_path_factory):
    """kmcp compute -k 21 --split-number 10 --split-overlap 150 -B plasmid -N '^([\\w\\.\\_]+\\.\\d+)'; kmcp index -f 0.3 -n 1
    with block size 16 (demo-profiling/README.md:238-255; SURVEY A.7: the golden counts need `-j 9..16`)"""
    O = oracle
    sp = O.sketch_params(21)
    targets = []
    for f in sorted(glob.glob(REF + "/demo-profiling/refs/*.fa.gz")):
        name = re.match(r"^([\w\.\_]+\.\d+)", os.path.basename(f)).group(1)
        targets += O.compute_targets(list(O.read_fastx(f)), name, sp, split_number=10, split_overlap=150, name_filters=["plasmid"])
    out = str(tmp_path_factory.mktemp("demo_db"))
    r001 = O.build_db(targets, out, sp, num_hashes=1, fpr=0.3, block_size=16)
    return r001, targets


def test_g5_genome_sizes(demo_db):
    _, targets = demo_db
    assert {t.genome_size for t in targets} == G5_SIZES
    assert len(targets) == 150 and all(t.n_chunks == 10 for t in targets)


def test_g1_g2_demo_profiling(oracle, demo_db):
    O = oracle
    r001, _ = demo_db
    db = O.DB(r001)
    ids, seqs = [], []
    for f in ("mock_1.fastq.gz", "mock_2.fastq.gz"):
        for i, _h, s in O.read_fastx(REF + "/demo-profiling/" + f):
            ids.append(i); seqs.append(s)
    assert len(seqs) == 349084                                       # mock.kmcp.gz.log:22
    res = db.search(seqs)
    nh = np.diff(res.hit_off.astype(np.int64))
    assert int((nh > 0).sum()) == 308839                             # mock.kmcp.gz.log:23
    names = [db.target(g).name.decode() for g in range(db.info.n_targets)]
    per_ref = {}
    unamb = 0
    for q in np.nonzero(nh > 0)[0]:
        a, b = int(res.hit_off[q]), int(res.hit_off[q + 1])
        refs = {names[t] for t in res.hits["target"][a:b]}
        if len(refs) == 1:
            unamb += 1
            r = next(iter(refs))
            per_ref[r] = per_ref.get(r, 0) + 1
    assert unamb == 198911                                           # ANALYSIS.md:17,21
    assert sorted(per_ref.values(), reverse=True) == G1_UNAMBIGUOUS
    # the same counts through the restated `profile` stage 1 (profile.go:761-990) fed with the search TSV, filters open
    n_reads, prof = O.profile_stage1(O.format_tsv(db, ids, res), min_qcov=0.0, max_fpr=1.0)
    assert n_reads == 308839
    assert sorted((int(sum(t[2])) for t in prof.values()), reverse=True) == G1_UNAMBIGUOUS
    assert abs(sum(sum(t[1]) for t in prof.values()) - sum(len({names[t] for t in res.hits["target"][int(res.hit_off[q]):int(res.hit_off[q + 1])]})
                                                            for q in np.nonzero(nh > 0)[0])) < 1e-6
    # G2: the reference's own first rows (docs/tutorial/profiling/index.md:203-211), all 15 columns
    sub = O.SearchResult(res.query_len[:10], res.n_kmers[:10], res.k_used[:10], res.hit_off[:11], res.hits[:int(res.hit_off[10])])
    tsv = O.format_tsv(db, ids[:10], sub, trailer=False).splitlines()
    exp = """NC_003197.2-64416/1	150	130	7.4626e-15	1	GCF_000006945.2	9	10	4857450	21	90	0.6923	0.0002	0.0002	1
NC_003197.2-64414/1	150	130	7.4626e-15	1	GCF_000006945.2	6	10	4857450	21	130	1.0000	0.0003	0.0003	2
NC_003197.2-64412/1	150	130	7.4626e-15	1	GCF_000006945.2	6	10	4857450	21	121	0.9308	0.0002	0.0002	3
NC_003197.2-64410/1	150	130	7.4626e-15	1	GCF_000006945.2	1	10	4857450	21	101	0.7769	0.0002	0.0002	4
NC_003197.2-64408/1	150	130	7.8754e-15	1	GCF_000006945.2	9	10	4857450	21	83	0.6385	0.0002	0.0002	5
NC_003197.2-64406/1	150	130	7.4626e-15	1	GCF_000006945.2	2	10	4857450	21	103	0.7923	0.0002	0.0002	6
NC_003197.2-64404/1	150	130	7.4671e-15	1	GCF_000006945.2	5	10	4857450	21	86	0.6615	0.0002	0.0002	7
NC_003197.2-64402/1	150	130	7.5574e-15	1	GCF_000006945.2	3	10	4857450	21	84	0.6462	0.0002	0.0002	8
NC_003197.2-64400/1	150	130	7.4626e-15	1	GCF_000006945.2	1	10	4857450	21	89	0.6846	0.0002	0.0002	9""".splitlines()
    assert tsv[0].startswith("#query\tqLen\tqKmers\tFPR\thits\ttarget")
    assert tsv[1:] == exp
    # the reference-shaped probe (64-row buffer + transpose + positional popcount) gives the same answer
    r1 = db.search(seqs[:20000], algo=1)
    r0 = db.search(seqs[:20000], algo=0)
    assert np.array_equal(r1.hits, r0.hits) and np.array_equal(r1.hit_off, r0.hit_off)


@pytest.mark.parametrize("label,kw,expected", [
    ("G3 FracMinHash", dict(scaled=True, scale=1000),                      # demo-searching/README.md:102-109
     [(1.0000, 1.0000, 1.0000), (0.7499, 0.7234, 0.5828), (0.6064, 0.6833, 0.4734), (0.5965, 0.6893, 0.4701), (0.5852, 0.5958, 0.4189),
      (0.5527, 0.5383, 0.3750)]),
    ("G4 closed syncmer", dict(scaled=True, scale=62, syncmer_s=15),       # demo-searching/README.md:61-68
     [(1.0000, 1.0000, 1.0000), (0.7439, 0.7189, 0.5763), (0.6041, 0.6768, 0.4688), (0.5972, 0.6807, 0.4665), (0.5782, 0.5868, 0.4109),
      (0.5482, 0.5322, 0.3699)]),
])
def test_g3_g4_demo_searching(oracle, tmp_path, label, kw, expected):
    O = oracle
    sp = O.sketch_params(31, **kw)
    targets = []
    for f in sorted(glob.glob(REF + "/demo-searching/refs/*.fasta.gz")):
        targets += O.compute_targets(list(O.read_fastx(f)), os.path.basename(f).replace(".fasta.gz", ""), sp, name_filters=["plasmid"])
    assert len(targets) == 9
    r001 = O.build_db(targets, str(tmp_path), sp, num_hashes=3, fpr=0.01, threads=16)        # blocks of 8 + 1
    db = O.DB(r001)
    assert db.info.n_blocks == 2
    recs = list(O.read_fastx(REF + "/demo-searching/refs/NC_018658.1.fasta.gz"))
    q = recs[0][2]
    for r in recs[1:]:
        q += r[2] + b"N" * 30                                              # S:885-937 whole-file query
    o = O.default_opts(); o.min_query_cov = 0.5; o.sort_by = 2
    res = db.search([q], opts=o)
    got = [(round(float(h["qcov"]), 4), round(float(h["tcov"]), 4), round(float(h["jacc"]), 4)) for h in res.hits]
    fmt = [("%.4f" % h["qcov"], "%.4f" % h["tcov"], "%.4f" % h["jacc"]) for h in res.hits]
    assert fmt == [("%.4f" % a, "%.4f" % b, "%.4f" % c) for a, b, c in expected], (label, got)

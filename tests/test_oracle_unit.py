"""CPU unit tests of the oracle against independent small restatements and the committed golden vectors."""
import json
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

SEED = {"A": 0x3c8bfbb395c60474, "C": 0x3193c18562a02b4c, "G": 0x20323ed082572324, "T": 0x295549f54be24456}
COMP = {"A": "T", "C": "G", "G": "C", "T": "A"}
M = (1 << 64) - 1


def rol(x, r):
    r %= 64
    return ((x << r) | (x >> (64 - r))) & M if r else x


def nthash_py(s, k):
    """pure-Python loop straight from the definition — an independent second opinion"""
    out = []
    for i in range(len(s) - k + 1):
        f = r = 0
        for j in range(k):
            b = s[i + j].upper()
            f ^= rol(SEED.get(b, 0), k - 1 - j)
            r ^= rol(SEED.get(COMP.get(b, "N"), 0), j)
        out.append(min(f, r))
    return out


def test_nthash_matches_definition(oracle):
    rng = np.random.default_rng(1)
    for k in (1, 5, 21, 31, 63, 64):
        s = "".join(rng.choice(list("ACGTNacgt"), 200))
        got = oracle.nthash_all(s.encode(), k)
        assert [int(x) for x in got] == nthash_py(s, k)
    assert oracle.nthash_all(b"ACG", 21).size == 0


def test_golden_read_codes(oracle):
    d = json.load(open(os.path.join(GOLD, "reads_k21.json")))
    sp = oracle.sketch_params(d["k"])
    for r in d["reads"]:
        assert [int(c) for c in oracle.generate_kmers(r["seq"].encode(), sp)] == r["codes"]
    # and they agree with the definition-level Python loop on a few reads
    for r in d["reads"][:3]:
        assert [c for c in nthash_py(r["seq"], 21) if c > 0] == r["codes"]


def test_golden_sketches(oracle):
    d = json.load(open(os.path.join(GOLD, "sketch_k31.json")))
    seq = d["seq"].encode()
    sps = {"scaled1000": oracle.sketch_params(31, scaled=True, scale=1000), "scaled50": oracle.sketch_params(31, scaled=True, scale=50),
           "syncmer15_scaled62": oracle.sketch_params(31, scaled=True, scale=62, syncmer_s=15), "syncmer15": oracle.sketch_params(31, syncmer_s=15),
           "minimizer10": oracle.sketch_params(31, minimizer_w=10)}
    for name, sp in sps.items():
        assert [int(c) for c in oracle.generate_kmers(seq, sp)] == d["lists"][name], name


def test_scaled_is_a_filter_of_all_kmers(oracle):
    seq = json.load(open(os.path.join(GOLD, "sketch_k31.json")))["seq"].encode()
    allk = oracle.generate_kmers(seq, oracle.sketch_params(31))
    mx = int(18446744073709551616.0 / 50)
    assert [int(c) for c in allk if int(c) <= mx] == [int(c) for c in oracle.generate_kmers(seq, oracle.sketch_params(31, scaled=True, scale=50))]


def test_fpr_golden_and_properties(oracle):
    for c in json.load(open(os.path.join(GOLD, "fpr.json"))):
        v = oracle.query_fpr(c["n"], c["c"], c["p"])
        assert float(v).hex() == c["fpr_hex"] and oracle.go_fmt_e4(v) == c["fmt"]
    # the four values of the reference's own table (docs/tutorial/profiling/index.md:203-211)
    assert [oracle.go_fmt_e4(oracle.query_fpr(130, c, 0.3)) for c in (90, 83, 86, 84)] == ["7.4626e-15", "7.8754e-15", "7.4671e-15", "7.5574e-15"]
    assert oracle.query_fpr(130, 10, 0.3) > 0.99
    assert 0 <= oracle.query_fpr(5000, 2500, 0.3) < 1e-9
    # Go math.Pow special cases used on the path
    L = oracle.lib()
    assert L.ko_go_pow(0.3, 0.0) == 1.0 and L.ko_go_pow(0.3, 1.0) == 0.3 and L.ko_go_pow(0.7, 20000.0) == 0.0


def test_dedup_and_hash_values(oracle):
    rng = np.random.default_rng(2)
    c = rng.integers(1, 50, 300).astype(np.uint64)
    assert np.array_equal(oracle.dedup(c, 256), np.unique(c))
    assert np.array_equal(oracle.dedup(c[:256], 256), c[:256])          # strict >: 256 codes are left alone
    code = 0xFFFFFFFF00000003
    assert oracle.hash_values(code, 1) == [code]
    assert oracle.hash_values(code, 4) == [(0xFFFFFFFF + 3 * i) & 0xFFFFFFFF for i in range(4)]


def test_uniki_roundtrip_and_search_shapes(oracle, tmp_path):
    import parity_helpers as helpers
    O = oracle
    sp = O.sketch_params(21)
    targets = helpers.make_synth_targets(O, sp, 3, 6, 8000, 3, 100)
    r001 = O.build_db(targets, str(tmp_path), sp, num_hashes=2, fpr=0.2, block_size=8)
    b = O.read_uniki(os.path.join(r001, "_block001.uniki"))
    assert b.k == 21 and b.canonical and b.num_hashes == 2 and len(b.names) == 8 and b.rows.shape == (b.num_sigs, 1)
    p2 = str(tmp_path / "copy.uniki")
    O.write_uniki(p2, b)
    assert open(p2, "rb").read() == open(os.path.join(r001, "_block001.uniki"), "rb").read()
    db = O.DB(r001)
    reads = helpers.make_reads(O, 5, 300, 6, 8000, 3) + helpers.edge_reads(21)
    r0, r1 = db.search(reads, algo=0), db.search(reads, algo=1)
    assert np.array_equal(r0.hits, r1.hits) and len(r0.hits) > 100
    # every set bit of a target's own k-mers is found: a chunk queried against the DB matches itself completely
    t = targets[4]
    counts = db.count_codes(t.codes)
    g = [i for i in range(db.info.n_targets) if db.target(i).name.decode() == t.name and (db.target(i).index & 0xFFFF) == t.chunk_idx][0]
    assert counts[g] == t.codes.size


def test_synth_generators_are_deterministic(oracle):
    a = oracle.synth_genome(1, 5, 1000)
    assert a == oracle.synth_genome(1, 5, 1000) and a[100:200] == oracle.synth_genome(1, 5, 200)[100:200]
    assert set(a) <= set(b"ACGT")
    r = [oracle.synth_read(2, i, 10, 5000, 150, 1) for i in range(50)]
    assert len(set(r)) == 50 and all(len(x) == 150 for x in r)

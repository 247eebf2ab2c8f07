"""GPU parity tests added in round 2: the arithmetic edges of the row-index path (64-bit numSigs, h = 4), multi-k databases
through the engine's k loop, and the asynchronous submit / wait form of the batch call.  Everything goes through the C ABI."""
import ctypes as C
import os
import threading

import numpy as np
import pytest

import parity_helpers as helpers

pytestmark = pytest.mark.gpu

GSEED, RSEED = 31, 47


def _row_indices(ctx, codes, h, num_sigs):
    from kmcp_b200 import api
    L = api.load()
    L.kmcpg_internal_row_indices.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_uint64, C.c_void_p]
    codes = np.ascontiguousarray(codes, dtype=np.uint64)
    out = np.zeros(codes.size * h, dtype=np.uint64)
    rc = L.kmcpg_internal_row_indices(ctx._h, codes.ctypes.data, codes.size, h, num_sigs, out.ctypes.data)
    assert rc == 0, L.kmcpg_last_error(ctx._h)
    return out.reshape(-1, h)


def _expected_rows(codes, h, num_sigs):
    """hashValues (util-hash.go:125-141) + `% numSigs` (U:6811) with Python integers"""
    out = np.zeros((len(codes), h), dtype=np.uint64)
    for i, c in enumerate(codes.tolist()):
        if h == 1:
            out[i, 0] = c % num_sigs
        else:
            x, y = c >> 32, c & 0xFFFFFFFF
            for j in range(h):
                out[i, j] = ((x + j * y) & 0xFFFFFFFF) % num_sigs
    return out


@pytest.mark.parametrize("h", [1, 2, 3, 4])
def test_row_index_arithmetic_is_exact_at_the_edges(gpu_ctx, h):
    """fastdiv.Mod replacement (Barrett on the device): d around 2^31, 2^32, 2^40, 2^63, tiny, prime, powers of two; dividends
    at the multiples of d, the u64 limits and random (property test near 2^32)"""
    rng = np.random.default_rng(1234 + h)
    ds = [1, 2, 3, 7, 10, 255, 256, 65537, 1000003, 1122448, 2**31 - 1, 2**31, 2**31 + 1, 2**32 - 3, 2**32 - 2, 2**32 - 1, 2**32, 2**32 + 1,
          2**32 + 15, 2**33 - 1, 2**40 - 87, 2**40, 2**40 + 1, 2**53 + 5, 2**63 - 25, 2**63, 2**64 - 59, 2**64 - 1]
    ds += [int(x) for x in rng.integers(1, 2**32, 6, dtype=np.uint64)] + [int(x) for x in rng.integers(2**32, 2**63, 4, dtype=np.uint64)]
    for d in ds:
        cand = [0, 1, d - 1, d, d + 1, 2 * d - 1, 2 * d, 2 * d + 1, 2**32 - 1, 2**32, 2**32 + 1, 2**63, 2**64 - 2, 2**64 - 1,
                (2**64 - 1) // d * d, (2**64 - 1) // d * d - 1, ((2**64 - 1) // d * d + d - 1)]
        # h > 1 looks at the two 32-bit halves: values whose wrapped sum hits 0, d-1, d, 2^32-1
        cand += [((d & 0xFFFFFFFF) << 32) | 0, (0xFFFFFFFF << 32) | 1, (0xFFFFFFFF << 32) | 0xFFFFFFFF, (1 << 32) | 0xFFFFFFFF]
        codes = np.array([c & (2**64 - 1) for c in cand if c >= 0] + [int(x) for x in rng.integers(0, 2**64, 400, dtype=np.uint64)], dtype=np.uint64)
        got = _row_indices(gpu_ctx, codes, h, d)
        exp = _expected_rows(codes, h, d)
        assert np.array_equal(got, exp), (h, d, codes[np.nonzero((got != exp).any(axis=1))[0][:3]])


def test_block_with_more_than_2_pow_32_signatures(gpu_ctx, oracle):
    """numSigs is a uint64 in the block format (index/serialization.go:173) and in the probe (U:6811): a block with > 2^32 rows
    (one 1-byte-wide column group, fpr 1e-4: 2^32 x 16 B = 69 GB of HBM) goes through the 64-bit row-index kernels."""
    from kmcp_b200 import api
    O = oracle
    free, _total = gpu_ctx.device_memory()
    if free < 90 * 2**30:
        pytest.skip("needs 90 GB of free HBM")
    ng, gl, k = 3, 440_000, 21
    gpu_ctx.build_synth_db(GSEED, ng, gl, k=k, n_chunks=1, overlap=0, num_hashes=1, fpr=1e-4, block_size=8)
    try:
        info = gpu_ctx.db_info()
        assert info.n_blocks == 1 and info.n_targets == ng
        sp = O.sketch_params(k)
        gcodes = [np.unique(O.generate_kmers(O.synth_genome(GSEED, g, gl), sp)) for g in range(ng)]
        num_sigs = O.calc_signature_size(max(c.size for c in gcodes), 1, 1e-4)
        assert num_sigs > 2**32 and info.disk_bytes == num_sigs          # 1 byte per row
        sizes = sorted(range(ng), key=lambda g: gcodes[g].size)          # I:667: columns ascending by k-mer count (stable)
        locsets = [np.unique(gcodes[g] % np.uint64(num_sigs)) for g in sizes]
        reads = helpers.make_reads(O, RSEED, 400, ng, gl, GSEED)
        buf, off = api.pack_seqs(reads)
        got = gpu_ctx.search_batch(buf, off, gpu_ctx.default_params(min_query_cov=0.3))
        exp = []
        for q, r in enumerate(reads):
            codes = O.generate_kmers(r, sp)
            if len(r) < 30 or codes.size < 10:
                continue
            locs = codes % np.uint64(num_sigs)
            for col in range(ng):
                c = int(np.isin(locs, locsets[col]).sum())
                if c >= 10 and float(c) > float(codes.size) * 0.3:
                    exp.append((q, col, c))
        assert len(exp) > 200
        assert [(int(h["query"]), int(h["target"]), int(h["count"])) for h in got.hits] == exp
        # dense counts of one code list, duplicates included
        codes = np.concatenate([O.generate_kmers(r, sp) for r in reads[:5]])
        dense = gpu_ctx.count_codes(codes)
        locs = codes % np.uint64(num_sigs)
        assert [int(x) for x in dense] == [int(np.isin(locs, locsets[col]).sum()) for col in range(ng)]
    finally:
        gpu_ctx.build_synth_db(GSEED, 2, 2000, k=21, n_chunks=1, overlap=0, num_hashes=1, fpr=0.3, block_size=8)     # release the 69 GB


@pytest.fixture(scope="module")
def h4_db(oracle, tmp_path_factory):
    """60 genomes x 15 kb, 4 chunks -> 240 targets, h = 4, two blocks of 128 (16-byte rows) and 112"""
    O = oracle
    sp = O.sketch_params(21)
    targets = helpers.make_synth_targets(O, sp, GSEED + 2, 60, 15000, 4, 120)
    return O.build_db(targets, str(tmp_path_factory.mktemp("db_h4")), sp, num_hashes=4, fpr=0.2, block_size=128)


def test_four_hash_database(gpu_ctx, oracle, h4_db):
    from kmcp_b200 import api
    O = oracle
    odb = O.DB(h4_db)
    gpu_ctx.open_db(h4_db)
    assert gpu_ctx.db_info().num_hashes == 4
    sp = odb.sketch_params()
    reads = helpers.make_reads(O, RSEED + 1, 600, 60, 15000, GSEED + 2) + helpers.edge_reads(21)
    for n in (1, 9):
        codes = np.concatenate([O.generate_kmers(r, sp) for r in reads[:n]])
        assert np.array_equal(gpu_ctx.count_codes(codes), odb.count_codes(codes))
    long_codes = O.generate_kmers(O.synth_genome(GSEED + 2, 5, 15000), sp)
    for codes in (long_codes[:300], long_codes[:6000], np.tile(long_codes[:7000], 10)):      # 16- and 24-plane counters
        assert np.array_equal(gpu_ctx.count_codes(codes), odb.count_codes(codes))
    buf, off = api.pack_seqs(reads)
    got = gpu_ctx.engine_search(buf, off)
    exp = odb.search(reads)
    assert np.array_equal(got.n_kmers, exp.n_kmers) and np.array_equal(got.match_off, exp.hit_off)
    for f in ("query", "target", "count", "fpr", "qcov", "tcov", "jacc"):
        assert np.array_equal(got.matches[f], exp.hits[f]), f
    assert len(got.matches) > 300


def _mutated_reads(O, gseed, n_genomes, genome_len, n, rate, seed, read_len=150):
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    reads = []
    for _ in range(n):
        g = int(rng.integers(0, n_genomes))
        pos = int(rng.integers(0, genome_len - read_len))
        b = O.synth_genome_bases(gseed, g, pos, read_len).copy()
        sub = rng.random(read_len) < rate
        b[sub] = (b[sub] + rng.integers(1, 4, int(sub.sum())).astype(np.uint8)) & 3
        reads.append(acgt[b].tobytes())
    return reads


def test_multi_k_database_falls_back_to_the_smaller_k(gpu_ctx, oracle, tmp_path):
    """`kmcp compute -k 21 -k 31`: one Bloom filter holds the k-mers of both sizes, the header carries the largest k, __db.yml lists
    ks; a query is searched with the largest k first and with the next smaller one when nothing matched (U:752-759, U:1018-1023)"""
    from kmcp_b200 import api
    O = oracle
    ng, gl = 12, 20000
    sp31, sp21 = O.sketch_params(31), O.sketch_params(21)
    t31 = helpers.make_synth_targets(O, sp31, GSEED + 3, ng, gl, 4, 150)
    t21 = helpers.make_synth_targets(O, sp21, GSEED + 3, ng, gl, 4, 150)
    assert len(t31) == len(t21)
    for a, b in zip(t31, t21):
        assert (a.name, a.chunk_idx) == (b.name, b.chunk_idx)
        a.codes = np.union1d(a.codes, b.codes)
    r001 = O.build_db(t31, str(tmp_path), sp31, num_hashes=1, fpr=0.3, block_size=16)
    yml = os.path.join(r001, "__db.yml")
    txt = open(yml).read()
    assert "ks:\n- 31\n" in txt
    open(yml, "w").write(txt.replace("ks:\n- 31\n", "ks:\n- 31\n- 21\n"))
    odb = O.DB(r001)
    assert list(odb.info.ks[:odb.info.n_ks]) == [31, 21]
    gpu_ctx.open_db(r001)
    info = gpu_ctx.db_info()
    assert list(info.ks[:info.n_ks]) == [31, 21]
    # clean reads match with k = 31; reads with 3 % substitutions mostly only with k = 21; random reads with neither
    reads = (_mutated_reads(O, GSEED + 3, ng, gl, 150, 0.0, 1) + _mutated_reads(O, GSEED + 3, ng, gl, 400, 0.03, 2) +
             _mutated_reads(O, GSEED + 3, ng, gl, 150, 0.06, 3) + helpers.edge_reads(31))
    oo = O.default_opts()
    oo.min_query_cov = 0.4
    exp = odb.search(reads, opts=oo)
    buf, off = api.pack_seqs(reads)
    got = gpu_ctx.engine_search(buf, off, gpu_ctx.default_engine_opts(min_query_cov=0.4))
    used = got.k_used[np.diff(got.match_off.astype(np.int64)) > 0]
    assert (used == 31).sum() > 100 and (used == 21).sum() > 50, np.unique(used, return_counts=True)
    assert np.array_equal(got.k_used, exp.k_used)
    assert np.array_equal(got.n_kmers, exp.n_kmers) and np.array_equal(got.query_len, exp.query_len)
    assert np.array_equal(got.match_off, exp.hit_off)
    for f in ("query", "target", "count", "fpr", "qcov", "tcov", "jacc"):
        assert np.array_equal(got.matches[f], exp.hits[f]), f


# ---------------------------------------------------------------------------------------------------- submit / wait
@pytest.fixture(scope="module")
def mid_db(oracle, tmp_path_factory):
    O = oracle
    sp = O.sketch_params(21)
    targets = helpers.make_synth_targets(O, sp, GSEED + 4, 30, 25000, 5, 150)
    return O.build_db(targets, str(tmp_path_factory.mktemp("db_mid")), sp, num_hashes=1, fpr=0.3, block_size=64)


def test_jobs_in_flight_give_the_answers_of_the_blocking_call(gpu_ctx, oracle, mid_db):
    """kmcpg_search_submit / kmcpg_search_wait: several batches queued at once (host and device input, caller-owned hit buffer)
    return exactly what kmcpg_search_batch returns for each of them, in any waiting order"""
    from kmcp_b200 import api
    O = oracle
    gpu_ctx.open_db(mid_db)
    p = gpu_ctx.default_params()
    batches = []
    for i, n in enumerate((3000, 1, 700, 0, 5000, 41)):
        reads = helpers.make_reads(O, RSEED + 10 + i, n, 30, 25000, GSEED + 4)
        buf, off = api.pack_seqs(reads)
        batches.append((buf, off, n, gpu_ctx.search_batch(buf, off, p)))
    # all on the host, waited for in reverse order
    jobs = [gpu_ctx.submit(buf.ctypes.data, off.ctypes.data, n, p) for buf, off, n, _ in batches]
    for (buf, off, n, ref), job in reversed(list(zip(batches, jobs))):
        got = gpu_ctx.wait(job)
        assert np.array_equal(got.hits, ref.hits) and np.array_equal(got.n_kmers, ref.n_kmers) and np.array_equal(got.query_len, ref.query_len)
    # device input with a host copy of the offsets, hits straight into the caller's pinned buffer
    cap = 1 << 16
    dst, dst_ptr = api.pinned_array(cap * 12 * len(batches))
    dev, jobs = [], []
    for i, (buf, off, n, _ref) in enumerate(batches):
        dseq = gpu_ctx.device_alloc(max(buf.nbytes, 1)); doff = gpu_ctx.device_alloc(off.nbytes)
        gpu_ctx.h2d(dseq, buf); gpu_ctx.h2d(doff, off)
        dev.append((dseq, doff))
        jobs.append(gpu_ctx.submit(dseq, doff, n, p, device=True, host_off_ptr=off.ctypes.data if i % 2 == 0 else 0,
                                   hits_dst=dst_ptr + i * cap * 12, hits_cap=cap))
    for i, ((buf, off, n, ref), job) in enumerate(zip(batches, jobs)):
        got = gpu_ctx.wait(job, copy=False)
        assert got.n_hits == len(ref.hits)
        mine = dst[i * cap * 12:i * cap * 12 + got.n_hits * 12].view(api.HIT_DTYPE)
        assert np.array_equal(mine, ref.hits)
    # a caller buffer that is too small is an error of that job only
    small = gpu_ctx.submit(dev[4][0], dev[4][1], batches[4][2], p, device=True, hits_dst=dst_ptr, hits_cap=5)
    ok = gpu_ctx.submit(dev[0][0], dev[0][1], batches[0][2], p, device=True)
    with pytest.raises(api.KmcpGpuError) as e:
        gpu_ctx.wait(small)
    assert e.value.code == api.KMCPG_ENOMEM
    assert np.array_equal(gpu_ctx.wait(ok).hits, batches[0][3].hits)
    for dseq, doff in dev:
        gpu_ctx.device_free(dseq); gpu_ctx.device_free(doff)
    api.host_free(dst_ptr)


def test_two_host_threads_share_one_context(gpu_ctx, oracle, mid_db):
    """the engine call from two threads at once (what a Go host does from goroutines): the executor interleaves their batches"""
    from kmcp_b200 import api
    O = oracle
    odb = O.DB(mid_db)
    gpu_ctx.open_db(mid_db)
    work = []
    for i in range(6):
        reads = helpers.make_reads(O, RSEED + 30 + i, 1500 + 100 * i, 30, 25000, GSEED + 4)
        work.append((api.pack_seqs(reads), odb.search(reads)))
    out = [None] * len(work)

    def run(t):
        for i in range(t, len(work), 2):
            (buf, off), _exp = work[i]
            out[i] = gpu_ctx.engine_search(buf, off)

    th = [threading.Thread(target=run, args=(t,)) for t in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for got, (_b, exp) in zip(out, work):
        assert np.array_equal(got.match_off, exp.hit_off)
        for f in ("query", "target", "count", "fpr", "qcov", "tcov", "jacc"):
            assert np.array_equal(got.matches[f], exp.hits[f]), f


# ---------------------------------------------------------------------------------------------------- reference-held golden
G2_ROWS = """NC_003197.2-64416/1	150	130	7.4626e-15	1	GCF_000006945.2	9	10	4857450	21	90	0.6923	0.0002	0.0002	1
NC_003197.2-64414/1	150	130	7.4626e-15	1	GCF_000006945.2	6	10	4857450	21	130	1.0000	0.0003	0.0003	2
NC_003197.2-64412/1	150	130	7.4626e-15	1	GCF_000006945.2	6	10	4857450	21	121	0.9308	0.0002	0.0002	3
NC_003197.2-64410/1	150	130	7.4626e-15	1	GCF_000006945.2	1	10	4857450	21	101	0.7769	0.0002	0.0002	4
NC_003197.2-64408/1	150	130	7.8754e-15	1	GCF_000006945.2	9	10	4857450	21	83	0.6385	0.0002	0.0002	5
NC_003197.2-64406/1	150	130	7.4626e-15	1	GCF_000006945.2	2	10	4857450	21	103	0.7923	0.0002	0.0002	6
NC_003197.2-64404/1	150	130	7.4671e-15	1	GCF_000006945.2	5	10	4857450	21	86	0.6615	0.0002	0.0002	7
NC_003197.2-64402/1	150	130	7.5574e-15	1	GCF_000006945.2	3	10	4857450	21	84	0.6462	0.0002	0.0002	8
NC_003197.2-64400/1	150	130	7.4626e-15	1	GCF_000006945.2	1	10	4857450	21	89	0.6846	0.0002	0.0002	9"""


def test_demo_profiling_golden_through_the_cuda_path(tmp_path):
    """The reference's own demo: `kmcp compute -k 21 -n 10 -l 150 -B plasmid -N ...` + `kmcp index -f 0.3 -n 1` (block
    size 16) over the 15 genomes of demo-profiling/refs, then `kmcp search` of the mock reads — here through kmcp-gpu index +
    kmcp-gpu search, i.e. the device builder and the device search path.  The first nine matched rows must be the rows the
    reference publishes (docs/tutorial/profiling/index.md:203-211), all 15 columns; the whole TSV of the 40,000-read subset must
    be the file the pinned oracle produced (tests/golden/make_demo_fixture.py; the oracle reproduces 308,839 / 349,084 matched
    on the full read set, which is too large to ship)."""
    import glob
    import gzip
    import json
    import subprocess
    demo = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "demo")
    refs = sorted(glob.glob(os.path.join(demo, "refs", "*.fa.gz")))
    assert len(refs) == 15
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "kmcp_b200", "kmcp-gpu")
    db = str(tmp_path / "refs-k21-n10.kmcp")
    p = subprocess.run([exe, "index", "-q", "-O", db, "-k", "21", "-n", "10", "-l", "150", "-B", "plasmid", "-N", r"^([\w\.\_]+\.\d+)", "-f", "0.3", "-b", "16"] + refs,
                       capture_output=True)
    assert p.returncode == 0, p.stderr.decode()
    assert len(glob.glob(db + "/R001/*.uniki")) == 10                      # 150 targets in blocks of 16
    tsv = str(tmp_path / "mock.tsv")
    p = subprocess.run([exe, "search", "-d", db, os.path.join(demo, "mock_1.20k.fastq.gz"), os.path.join(demo, "mock_2.20k.fastq.gz"), "-o", tsv], capture_output=True)
    assert p.returncode == 0, p.stderr.decode()
    got = open(tsv).read()
    rows = [ln for ln in got.splitlines() if not ln.startswith("#")]
    assert rows[:9] == G2_ROWS.splitlines()                               # G2, byte for byte
    exp = gzip.open(os.path.join(demo, "expected.20k.tsv.gz"), "rt").read()
    assert got == exp
    summary = json.load(open(os.path.join(demo, "summary.json")))
    assert len({ln.split("\t")[14] for ln in rows}) == summary["subset"]["matched"] == 36448 and len(rows) == summary["subset"]["rows"]
    log = p.stderr.decode()
    assert "(36448/40000) queries matched" in log, log[-400:]


def test_sharded_search_pipeline_single_rank(gpu_ctx, oracle, mid_db):
    """kmcp_b200.multigpu.ShardedSearch with one rank: the step pipeline bench.py runs under torchrun — staged batch + ready event, two jobs in
    flight, the library's device→host copy straight into the CUDA-registered shared-memory slot, consumer thread, slot hand-back over more steps
    than slots — must hand the consumer exactly the hits of the blocking call, step by step"""
    import torch
    from kmcp_b200 import api, multigpu
    O = oracle
    gpu_ctx.open_db(mid_db)
    n, steps = 2000, 5
    p = gpu_ctx.default_params()
    batches, refs = [], []
    off = np.arange(n + 1, dtype=np.uint64) * np.uint64(150)
    for s in range(steps):
        reads = helpers.make_reads(O, RSEED + 50 + s, n, 30, 25000, GSEED + 4)
        buf, off2 = api.pack_seqs(reads)
        assert np.array_equal(off, off2)
        refs.append(gpu_ctx.search_batch(buf, off, p))
        batches.append(np.concatenate([off.view(np.uint8), buf]))
    bb = batches[0].size
    dev = torch.device("cuda", 0)
    d = torch.from_numpy(np.concatenate(batches)).to(dev)
    sh = multigpu.ShardedSearch(gpu_ctx, 0, 1, n, bb, hit_cap=1 << 16, name="kmcp_test_%d" % os.getpid(), device=dev, dist=None, params=p)
    got = {}

    def consume(s, lists, meta):
        assert len(lists) == 1
        got[s] = (lists[0].copy(), meta.n_kmers.copy(), multigpu.hits_digest(multigpu.merge_lists(lists, 0, n)))

    try:
        for rep in range(2):          # the step counter of the exchange carries on across run() calls
            got.clear()
            outs = sh.run(steps, lambda s: (d[s * bb:(s + 1) * bb], False), consume, host_off=off if rep == 0 else None)
            assert len(outs) == steps and sorted(got) == list(range(steps))
            for s in range(steps):
                hits, nk, dig = got[s]
                assert np.array_equal(hits, refs[s].hits) and np.array_equal(nk, refs[s].n_kmers)
                assert dig == multigpu.hits_digest(refs[s].hits) and outs[s].n_hits == len(refs[s].hits)
    finally:
        sh.close()


G34_ORDER = ["NC_018658.1", "NZ_CP028116.1", "NC_000913.3", "NC_012971.2", "NZ_CP007592.1", "NC_002695.2"]    # the README's rows (name.map resolved)


@pytest.mark.parametrize("label,flags,expected", [
    ("G3 FracMinHash", ["-D", "1000"],                                     # demo-searching/README.md:102-109
     [("1.0000", "1.0000", "1.0000"), ("0.7499", "0.7234", "0.5828"), ("0.6064", "0.6833", "0.4734"), ("0.5965", "0.6893", "0.4701"),
      ("0.5852", "0.5958", "0.4189"), ("0.5527", "0.5383", "0.3750")]),
    ("G4 closed syncmer", ["-S", "15", "-D", "62"],                        # demo-searching/README.md:61-68
     [("1.0000", "1.0000", "1.0000"), ("0.7439", "0.7189", "0.5763"), ("0.6041", "0.6768", "0.4688"), ("0.5972", "0.6807", "0.4665"),
      ("0.5782", "0.5868", "0.4109"), ("0.5482", "0.5322", "0.3699")]),
])
def test_demo_searching_golden_tables_through_the_cuda_path(tmp_path, label, flags, expected):
    """The reference's genome-similarity demo: `kmcp compute -k 31 -B plasmid [-D 1000 | -S 15 -D 62]`, `kmcp index -n 3 -f 0.01`,
    `kmcp search -g -t 0.5 -n 0 -s jacc refs/NC_018658.1.fasta.gz` — through kmcp-gpu index + kmcp-gpu search (device sketching with the FracMinHash
    cut / closed-syncmer selection, device builder, -g whole-file query through the tile hashing and the CTA-per-task probe).  The qCov, tCov and jacc
    columns and the order of the six targets must be the table the reference prints in demo-searching/README.md."""
    import glob
    import subprocess
    demo = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "demo_searching", "refs")
    refs = sorted(glob.glob(os.path.join(demo, "*.fasta.gz")))
    assert len(refs) == 9
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "kmcp_b200", "kmcp-gpu")
    db = str(tmp_path / "refs.kmcp")
    p = subprocess.run([exe, "index", "-q", "-O", db, "-k", "31", "-B", "plasmid", "--num-hash", "3", "-f", "0.01", "-j", "16"] + flags + refs, capture_output=True)
    assert p.returncode == 0, p.stderr.decode()
    assert len(glob.glob(db + "/R001/*.uniki")) == 2                       # 9 references, -j 16: blocks of 8 + 1 (I:670-682)
    tsv = str(tmp_path / "out.tsv")
    p = subprocess.run([exe, "search", "-q", "-d", db, "-g", "-t", "0.5", "-n", "0", "-s", "jacc", os.path.join(demo, "NC_018658.1.fasta.gz"), "-o", tsv], capture_output=True)
    assert p.returncode == 0, p.stderr.decode()
    rows = [ln.split("\t") for ln in open(tsv).read().splitlines() if not ln.startswith("#")]
    assert [r[5] for r in rows] == G34_ORDER, label
    assert [(r[11], r[12], r[13]) for r in rows] == expected, label

"""No-GPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/kmcp_gpu.h
declares, refuses to run without a device (no CPU fallback), and its host-only helpers agree with the oracle."""
import ctypes as C
import json
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    h = open(os.path.join(ROOT, "include", "kmcp_gpu.h")).read()
    return sorted(set(re.findall(r"\b(kmcpg_[a-z0-9_]+)\s*\(", h)))


def test_library_exports_every_declared_symbol():
    from kmcp_b200 import api
    L = api.load()
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(L, s), "libkmcp_gpu.so does not export " + s
    assert set(syms) == set(api.ABI_SYMBOLS)
    assert L.kmcpg_abi_version() == 2


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from kmcp_b200 import api
    with pytest.raises(api.KmcpGpuError) as e:
        api.Context(0)
    assert e.value.code == api.KMCPG_ECUDA and "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    """the oracle is test infrastructure: nothing under kmcp_b200/ may import, link, dlopen or call it"""
    pat = re.compile(r"(import\s+oracle|from\s+oracle|kmcp_oracle|libkmcp_oracle|\bko_[a-z_]+\s*\()")
    for dirpath, _d, files in os.walk(os.path.join(ROOT, "kmcp_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")) or f == "Makefile":
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not pat.search(src), os.path.join(dirpath, f)


def test_host_fpr_matches_oracle_and_golden(oracle):
    from kmcp_b200 import api
    L = api.load()
    for c in json.load(open(os.path.join(ROOT, "tests", "golden", "fpr.json"))):
        assert float(L.kmcpg_query_fpr(c["n"], c["c"], c["p"])).hex() == c["fpr_hex"]
    for n, c in ((130, 72), (77, 40), (300, 200), (1500, 900)):
        assert L.kmcpg_query_fpr(n, c, 0.3) == oracle.query_fpr(n, c, 0.3)


def test_default_params_are_the_reference_defaults():
    from kmcp_b200 import api
    L = api.load()
    p = api.SearchParams()
    L.kmcpg_default_params(C.byref(p))
    assert (p.min_query_len, p.min_matched, p.dedup_threshold, p.min_query_cov, p.paired) == (30, 10, 256, 0.55, 0)   # S:1055-1069
    o = api.EngineOpts()
    L.kmcpg_default_engine_opts(C.byref(o))
    assert (o.min_query_cov, o.min_target_cov, o.max_fpr, o.top_n_scores, o.sort_by) == (0.55, 0.0, 0.01, 0, 0)          # S:1066-1093


def test_shard_plan_reads_headers_and_reports_format_errors(oracle, tmp_path):
    """host-only part of kmcpg_open_db: __db.yml + .uniki header parsing, error codes of X:38-56 / util-db-info.go:118"""
    import parity_helpers as helpers
    from kmcp_b200 import api
    O = oracle
    sp = O.sketch_params(21)
    targets = helpers.make_synth_targets(O, sp, 5, 6, 6000, 2, 50)
    r001 = O.build_db(targets, str(tmp_path / "db"), sp, num_hashes=1, fpr=0.3, block_size=8)       # 12 targets → 2 blocks
    assert sorted(api.shard_plan(r001, 2)) == [0, 1]
    blk = os.path.join(r001, "_block001.uniki")
    raw = open(blk, "rb").read()
    cases = {"bad magic": (b"XXXXXXXX" + raw[8:], api.KMCPG_EFORMAT), "version": (raw[:8] + bytes([3]) + raw[9:], api.KMCPG_EFORMAT),
             "truncated rows": (raw[:-100], api.KMCPG_EIO), "truncated header": (raw[:30], api.KMCPG_EIO)}
    for name, (data, code) in cases.items():
        open(blk, "wb").write(data)
        with pytest.raises(api.KmcpGpuError) as e:
            api.shard_plan(r001, 2)
        assert e.value.code == code, name
    open(blk, "wb").write(raw)
    yml = os.path.join(r001, "__db.yml")
    txt = open(yml).read()
    open(yml, "w").write(txt.replace("version: 4", "version: 3", 1))
    with pytest.raises(api.KmcpGpuError) as e:
        api.shard_plan(r001, 2)
    assert e.value.code == api.KMCPG_EFORMAT
    open(yml, "w").write(txt.replace("hashes: 1", "hashes: 2"))          # blocks say 1 hash: incompatible (U:689-695)
    with pytest.raises(api.KmcpGpuError) as e:
        api.shard_plan(r001, 2)
    assert e.value.code == api.KMCPG_EFORMAT
    with pytest.raises(api.KmcpGpuError) as e:
        api.shard_plan(str(tmp_path / "nope"), 2)
    assert e.value.code == api.KMCPG_EIO


def _check_pieces(pieces, n_names, world):
    """every column of every block in exactly one piece, cuts on 128-target boundaries, shards in order"""
    by_block = {}
    for b, s, c0, nc, _bytes in pieces:
        assert 0 <= s < world and nc > 0 and c0 % 128 == 0
        by_block.setdefault(b, []).append((c0, nc, s))
    assert sorted(by_block) == list(range(len(n_names)))
    for b, lst in by_block.items():
        lst.sort()
        pos = 0
        for c0, nc, _s in lst:
            assert c0 == pos
            pos += nc
        assert pos == n_names[b]


def test_shard_pieces_whole_blocks_and_column_ranges(oracle, tmp_path):
    """kmcpg_shard_pieces (host only): whole blocks while blocks >= shards, balanced column ranges when a DB has fewer
    blocks than shards"""
    import parity_helpers as helpers
    from kmcp_b200 import api
    O = oracle
    sp = O.sketch_params(21)
    # 300 genomes x 5 chunks = 1500 targets; one block of 1500 and one DB of 3 blocks (640 + 640 + 220)
    targets = helpers.make_synth_targets(O, sp, 5, 300, 1200, 5, 30)
    one = O.build_db(targets, str(tmp_path / "one"), sp, num_hashes=1, fpr=0.3, block_size=1500)
    three = O.build_db(targets, str(tmp_path / "three"), sp, num_hashes=1, fpr=0.3, block_size=640)
    # whole blocks: identical to the block plan
    for world in (1, 2, 3):
        pcs = api.shard_pieces(three, world)
        assert [(p[0], p[2], p[3]) for p in pcs] == [(0, 0, 640), (1, 0, 640), (2, 0, 220)]
        assert [p[1] for p in pcs] == api.shard_plan(three, world)
        assert set(p[1] for p in pcs) == set(range(world))
    # fewer blocks than shards: column ranges
    for r001, n_names in ((one, [1500]), (three, [640, 640, 220])):
        for world in (2, 4, 5, 8, 12, 16):
            if world <= len(n_names):
                continue
            pcs = api.shard_pieces(r001, world)
            _check_pieces(pcs, n_names, world)
            shards = [p[1] for p in pcs]
            assert shards == sorted(shards)                       # a shard holds one contiguous stretch of the column space
            used = sorted(set(shards))
            units = sum((n + 127) // 128 for n in n_names)
            assert used == list(range(min(world, len(used))))
            if units >= world:
                assert len(used) == world
                load = {}
                for p in pcs:
                    load[p[1]] = load.get(p[1], 0) + p[4]
                assert max(load.values()) <= 2.1 * (sum(load.values()) / world) + 1     # within one 128-target unit of the mean
            assert api.shard_plan(r001, world)[0] == 0
    # 1500 targets over 4 shards: 12 units of 128 -> 3 units each
    assert [(p[2], p[3]) for p in api.shard_pieces(one, 4)] == [(0, 384), (384, 384), (768, 384), (1152, 348)]


def test_sharded_engine_hit_merge_is_the_canonical_order():
    """the k-way merge of per-shard hit lists (disjoint by target, each sorted by (query, target)) = the one-context order"""
    import numpy as np
    from kmcp_b200 import api
    rng = np.random.default_rng(5)
    for k in (1, 2, 3, 8):
        n = 5000
        allh = np.zeros(n, dtype=api.HIT_DTYPE)
        pairs = rng.choice(400 * 300, size=n, replace=False)
        allh["query"], allh["target"] = pairs // 300, pairs % 300
        allh["count"] = rng.integers(1, 200, n)
        allh = allh[np.lexsort((allh["target"], allh["query"]))]
        owner = rng.integers(0, k, 300)                          # targets → shards (interleaved, like greedy block plans)
        lists = [allh[owner[allh["target"]] == s] for s in range(k)]
        if k == 3:
            lists[1] = lists[1][:0]                              # an empty shard list
            allh = allh[owner[allh["target"]] != 1]
        got = api.merge_hit_lists(lists)
        assert np.array_equal(got, allh)
    # the threaded form: query ranges found by binary search, every thread writes its own stretch of the output
    for k, nq, n, base in ((2, 3000, 200000, 0), (8, 250000, 800000, 1000), (5, 7, 5000, 50), (3, 100000, 40000, 0)):
        allh = np.zeros(n, dtype=api.HIT_DTYPE)
        pairs = rng.choice(nq * 1000, size=n, replace=False)
        allh["query"], allh["target"] = pairs // 1000 + base, pairs % 1000
        allh["count"] = rng.integers(1, 200, n)
        allh = allh[np.lexsort((allh["target"], allh["query"]))]
        owner = rng.integers(0, k, 1000)
        lists = [allh[owner[allh["target"]] == s] for s in range(k)]
        for threads in (2, 5, 16):
            assert np.array_equal(api.merge_hit_lists(lists, base, nq, threads), allh), (k, nq, threads)
    assert len(api.merge_hit_lists([np.zeros(0, api.HIT_DTYPE), np.zeros(0, api.HIT_DTYPE)])) == 0


@pytest.mark.timeout(120)
def test_sharded_round_merger_with_stand_in_shards():
    """the threaded part-by-part merger of kmcpg_engine_search_sharded, driven on the host by stand-in shards that deliver
    seeded hit lists with random delays: every merged part equals the union, errors of a shard surface, nothing hangs"""
    from kmcp_b200 import api
    L = api.load()
    f = L.kmcpg_internal_sharded_selftest
    f.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64]
    for shards in (1, 2, 3, 8):
        for parts in (0, 1, 2, 7, 40):
            for seed in (1, 2):
                assert f(shards, parts, -1, 0, seed) == 0, (shards, parts, seed)
    # a failing shard: first part, a middle part, after its last part; the error code comes back and every thread is joined
    for shards, parts, fs, fp in ((2, 5, 0, 0), (2, 5, 1, 3), (3, 6, 2, 6), (8, 10, 5, 9), (1, 3, 0, 1), (4, 0, 1, 0)):
        assert f(shards, parts, fs, fp, 9) == api.KMCPG_ECUDA, (shards, parts, fs, fp)


def _build_c_host(tmp_path):
    """examples/search_host.c: a strict C99 host of the C ABI (what a cgo/JNI/ctypes binding does, without the runtime)"""
    import subprocess
    exe = str(tmp_path / "search_host")
    lib_dir = os.path.join(ROOT, "kmcp_b200")
    p = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", os.path.join(ROOT, "examples", "search_host.c"),
                        "-I" + os.path.join(ROOT, "include"), "-L" + lib_dir, "-lkmcp_gpu", "-Wl,-rpath," + lib_dir, "-o", exe], capture_output=True)
    assert p.returncode == 0, p.stderr.decode()
    return exe


def test_header_is_plain_c99_and_a_c_host_links_and_fails_loudly_without_a_device(tmp_path):
    import subprocess
    import torch
    exe = _build_c_host(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present (tests/test_zz_gpu_sharded.py runs the C host against a database)")
    p = subprocess.run([exe, str(tmp_path / "no_db"), "ACGT" * 40], capture_output=True)
    assert p.returncode == 2 and b"no CPU fallback" in p.stderr and p.stdout == b""


@pytest.mark.timeout(180)
def test_replicas_split_and_concatenation_with_stand_in_replicas():
    """kmcpg_engine_search_replicas' host logic (byte-balanced query ranges, one thread per replica, answers moved side by side
    into one result) with stand-in replicas whose answer is a function of the global query index: n ranges == one range"""
    from kmcp_b200 import api
    L = api.load()
    f = L.kmcpg_internal_replicas_selftest
    f.argtypes = [C.c_int, C.c_uint32, C.c_int, C.c_int, C.c_uint64]
    for n_rep in (1, 2, 3, 8, 20, 64):
        for nq in (0, 1, 2, 7, 1000, 30011):
            for paired in (0, 1):
                assert f(n_rep, nq, paired, -1, 3 + nq) == 0, (n_rep, nq, paired)
    for n_rep, nq, fail in ((2, 1000, 0), (2, 1000, 1), (8, 5000, 5), (3, 10, 2)):
        assert f(n_rep, nq, 0, fail, 11) == api.KMCPG_ECUDA, (n_rep, nq, fail)


def test_binding_copies_records_out_of_library_memory():
    """api._np_from: one byte copy of the records behind a ctypes pointer; the result must not alias library memory"""
    import numpy as np
    from kmcp_b200 import api
    src = np.zeros(1000, dtype=api.MATCH_DTYPE)
    src["query"] = np.arange(1000); src["fpr"] = np.linspace(0, 1, 1000); src["jacc"] = 0.25
    out = api._np_from(C.cast(src.ctypes.data, C.POINTER(api.Match)), 1000, C.sizeof(api.Match), api.MATCH_DTYPE)
    assert out.dtype == api.MATCH_DTYPE and np.array_equal(out, src)
    src["query"] = 0
    assert out["query"][999] == 999 and out.flags.writeable
    assert C.sizeof(api.Match) == api.MATCH_DTYPE.itemsize == 48 and C.sizeof(api.Hit) == api.HIT_DTYPE.itemsize == 12
    assert len(api._np_from(None, 0, 12, api.HIT_DTYPE)) == 0


def test_block_rows_are_read_by_several_streams(tmp_path):
    """pread_parallel (the chunk reader of kmcpg_open_db): any offset / size / stream count gives the file's bytes; a range that
    reaches past the end of the file is an error (a truncated index file), never a short read"""
    import numpy as np
    from kmcp_b200 import api
    L = api.load()
    f = L.kmcpg_internal_pread_selftest
    f.argtypes = [C.c_char_p, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p]
    data = np.random.default_rng(5).integers(0, 256, 23_456_789, dtype=np.uint8)
    p = str(tmp_path / "rows.bin")
    data.tofile(p)
    for off, n, th in ((0, len(data), 8), (1, len(data) - 1, 3), (12345, 9_000_001, 16), (7, 1, 4), (100, 4 << 20, 2), (5, (4 << 20) + 1, 1), (0, 0, 4)):
        dst = np.zeros(max(n, 1), dtype=np.uint8)
        assert f(p.encode(), off, n, th, dst.ctypes.data) == 0, (off, n, th)
        assert np.array_equal(dst[:n], data[off:off + n]), (off, n, th)
    dst = np.zeros(1 << 20, dtype=np.uint8)
    assert f(p.encode(), len(data) - 1000, 1001, 4, dst.ctypes.data) == api.KMCPG_EIO
    assert f(str(tmp_path / "missing").encode(), 0, 10, 2, dst.ctypes.data) == api.KMCPG_EIO


def test_index_builder_loads_its_input_files_on_several_threads(tmp_path):
    """GenomeLoader (index_build.cu, host part of kmcpg_index_fasta): files loaded ahead by several threads come out in order
    and identical to loading them one by one; a missing file is reported at its place"""
    import gzip
    import random
    from kmcp_b200 import api
    L = api.load()
    f = L.kmcpg_internal_genome_loader_selftest
    f.argtypes = [C.POINTER(C.c_char_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
    rnd = random.Random(4)
    paths = []
    for i in range(37):
        p = str(tmp_path / ("g%02d.fa%s" % (i, ".gz" if i % 3 == 0 else "")))
        recs = b"".join(b">c%d_%d plasmid\n" % (i, j) + b"\n".join(bytes(rnd.choice(b"ACGTN") for _ in range(60)) for _ in range(rnd.randrange(1, 400))) + b"\n"
                        for j in range(rnd.randrange(1, 4)))
        (gzip.open if p.endswith(".gz") else open)(p, "wb").write(recs)
        paths.append(p.encode())
    # the other shapes a sequence file comes in: FASTQ (four-line and wrapped), CRLF, blank lines, no final newline, lower case
    extra = {
        "q4.fq": b"".join(b"@r%d x\n%s\n+\n%s\n" % (j, bytes(rnd.choice(b"ACGT") for _ in range(200)), b"I" * 200) for j in range(300)),
        "qw.fastq.gz": b"".join(b"@w%d\n%s\n%s\n+w%d\n%s\n%s\n" % (j, b"ACGTAC" * 10, b"GGTTAA" * 5, j, b"@" * 60, b"+" * 30) for j in range(50)),
        "crlf.fa": b">a desc\r\nACGTACGTAC\r\nGGGG\r\n\r\n>b\r\nTTTTT\r\n",
        "blank.fna": b"\n\n>x\nACGT\n\nACGT\n>y\n\n>z\nacgtn\n",
        "nonl.fa": b">only\nACGTACGTACGTACGTACGTACGTACGT",
    }
    for name, data in extra.items():
        p = str(tmp_path / name)
        (gzip.open if name.endswith(".gz") else open)(p, "wb").write(data)
        paths.append(p.encode())
    arr = (C.c_char_p * len(paths))(*paths)
    digests = set()
    for threads, split in ((1, 1), (3, 1), (8, 10), (16, 10), (2, 5)):
        d = C.c_uint64()
        assert f(arr, len(paths), 21, split, -1, threads, C.byref(d)) == 0, (threads, split)
        digests.add((split, d.value))
    assert len(digests) == 3                     # one digest per split setting, whatever the thread count
    bad = (C.c_char_p * 3)(paths[0], str(tmp_path / "missing.fa").encode(), paths[1])
    assert f(bad, 3, 21, 1, -1, 4, None) == api.KMCPG_EIO
    junk = str(tmp_path / "junk.fa")
    open(junk, "wb").write(b"ACGT without a header line\n")
    assert f((C.c_char_p * 2)(paths[0], junk.encode()), 2, 21, 1, -1, 2, None) == api.KMCPG_EIO      # not FASTA/Q: an error, as seqio/fastx gives
    broken = str(tmp_path / "broken.fa.gz")
    blob = bytearray(gzip.compress(b">g\n" + b"ACGT" * 20000 + b"\n"))
    blob[len(blob) // 2] ^= 0x20
    open(broken, "wb").write(bytes(blob))
    assert f((C.c_char_p * 1)(broken.encode()), 1, 21, 1, -1, 1, None) == api.KMCPG_EIO             # a damaged .gz is not a short genome


def test_go_stub_only_uses_what_the_header_declares():
    """go/engine_gpu.go cannot be compiled here (no Go toolchain): at least every C.kmcpg_* / C.KMCPG_* name it uses must be
    declared by include/kmcp_gpu.h, and the struct fields it touches must exist"""
    import re
    go = open(os.path.join(ROOT, "go", "engine_gpu.go")).read()
    hdr = open(os.path.join(ROOT, "include", "kmcp_gpu.h")).read()
    names = set(re.findall(r"\bC\.((?:kmcpg|KMCPG)_\w+)", go))
    assert len(names) > 15
    missing = [n for n in sorted(names) if not re.search(r"\b%s\b" % re.escape(n), hdr)]
    assert not missing, missing
    # the INTEGRATION.md code blocks and the source file say the same thing
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    for block in re.findall(r"```go\n(.*?)```", doc, flags=re.S):
        for line in block.splitlines():
            if line.strip() and not line.strip().startswith("//") and '"fmt"' not in line:
                assert line in go, line

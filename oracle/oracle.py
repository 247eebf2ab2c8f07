"""CPU ORACLE — Python face.  TEST INFRASTRUCTURE ONLY (see oracle/kmcp_oracle.h).

ctypes binding of oracle/libkmcp_oracle.so plus the fixture builders the reference performs offline
(`kmcp compute` chunking, `kmcp index` block assembly, `.uniki` / `__db.yml` writers) and the 15-column
TSV formatter of `kmcp search`.  Nothing under kmcp_b200/ may import this module.

Reference citations: C: = kmcp/cmd/compute.go, I: = kmcp/cmd/index.go, X: = kmcp/cmd/index/serialization.go,
S: = kmcp/cmd/search.go, U: = kmcp/cmd/util-db-search.go (all under /root/reference).
"""
from __future__ import annotations

import ctypes as C
import gzip
import math
import os
import re
import struct
import subprocess
from dataclasses import dataclass, field
from typing import Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libkmcp_oracle.so")


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc, a second or two)."""
    src = os.path.join(_HERE, "kmcp_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "kmcp_oracle.h"))):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _LIB_PATH


class SketchParams(C.Structure):
    _fields_ = [("k", C.c_int32), ("canonical", C.c_int32), ("scaled", C.c_int32), ("scale", C.c_uint32),
                ("minimizer", C.c_int32), ("minimizer_w", C.c_uint32), ("syncmer", C.c_int32), ("syncmer_s", C.c_uint32)]


class DBInfo(C.Structure):
    _fields_ = [("n_ks", C.c_int32), ("ks", C.c_int32 * 8), ("canonical", C.c_int32), ("num_hashes", C.c_int32),
                ("scaled", C.c_int32), ("scale", C.c_uint32), ("minimizer", C.c_int32), ("minimizer_w", C.c_uint32),
                ("syncmer", C.c_int32), ("syncmer_s", C.c_uint32), ("fpr", C.c_double), ("n_blocks", C.c_int32),
                ("n_targets", C.c_int64), ("sum_row_bytes", C.c_int64), ("total_bytes", C.c_int64)]


class Target(C.Structure):
    _fields_ = [("name", C.c_char_p), ("index", C.c_uint32), ("genome_size", C.c_uint64), ("n_kmers", C.c_uint64),
                ("block", C.c_int32), ("col", C.c_int32)]


class SearchOpts(C.Structure):
    _fields_ = [("min_query_len", C.c_int32), ("min_matched", C.c_int32), ("dedup_threshold", C.c_int32),
                ("min_query_cov", C.c_double), ("min_target_cov", C.c_double), ("max_fpr", C.c_double),
                ("sort_by", C.c_int32), ("do_not_sort", C.c_int32), ("top_n_scores", C.c_int32), ("try_se", C.c_int32)]


class Hit(C.Structure):
    _fields_ = [("query", C.c_uint32), ("target", C.c_uint32), ("count", C.c_uint32), ("_pad", C.c_uint32),
                ("fpr", C.c_double), ("qcov", C.c_double), ("tcov", C.c_double), ("jacc", C.c_double)]


HIT_DTYPE = np.dtype([("query", "<u4"), ("target", "<u4"), ("count", "<u4"), ("_pad", "<u4"),
                      ("fpr", "<f8"), ("qcov", "<f8"), ("tcov", "<f8"), ("jacc", "<f8")])


class Results(C.Structure):
    _fields_ = [("n_queries", C.c_uint32), ("query_len", C.POINTER(C.c_int32)), ("n_kmers", C.POINTER(C.c_int32)),
                ("k_used", C.POINTER(C.c_int32)), ("hit_off", C.POINTER(C.c_uint64)), ("hits", C.POINTER(Hit)),
                ("n_hits", C.c_uint64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        u8p = C.POINTER(C.c_uint8)
        u64p = C.POINTER(C.c_uint64)
        L.ko_nthash_all.restype = C.c_int64
        L.ko_nthash_all.argtypes = [u8p, C.c_int64, C.c_int, C.c_int, u64p]
        L.ko_generate_kmers.restype = C.c_int64
        L.ko_generate_kmers.argtypes = [u8p, C.c_int64, C.POINTER(SketchParams), u64p]
        L.ko_dedup.restype = C.c_int64
        L.ko_dedup.argtypes = [u64p, C.c_int64, C.c_int64]
        L.ko_hash_values.argtypes = [C.c_uint64, C.c_int, u64p]
        L.ko_query_fpr.restype = C.c_double
        L.ko_query_fpr.argtypes = [C.c_int, C.c_int, C.c_double]
        L.ko_go_pow.restype = C.c_double
        L.ko_go_pow.argtypes = [C.c_double, C.c_double]
        L.ko_calc_signature_size.restype = C.c_uint64
        L.ko_calc_signature_size.argtypes = [C.c_uint64, C.c_int, C.c_double]
        L.ko_db_open.restype = C.c_void_p
        L.ko_db_open.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        L.ko_db_close.argtypes = [C.c_void_p]
        L.ko_db_get_info.argtypes = [C.c_void_p, C.POINTER(DBInfo)]
        L.ko_db_target.argtypes = [C.c_void_p, C.c_int64, C.POINTER(Target)]
        L.ko_db_block.argtypes = [C.c_void_p, C.c_int, u64p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(u8p)]
        L.ko_default_opts.argtypes = [C.POINTER(SearchOpts)]
        L.ko_search.argtypes = [C.c_void_p, C.POINTER(SearchOpts), u8p, u64p, C.c_uint32, C.c_int, C.c_int, C.c_int,
                                C.POINTER(Results)]
        L.ko_free_results.argtypes = [C.POINTER(Results)]
        L.ko_count_codes.argtypes = [C.c_void_p, u64p, C.c_int64, C.POINTER(C.c_uint32)]
        _lib = L
    return _lib


def _u8(buf) -> Tuple[np.ndarray, C.POINTER(C.c_uint8)]:
    a = np.frombuffer(bytes(buf), dtype=np.uint8) if not isinstance(buf, np.ndarray) else np.ascontiguousarray(buf, dtype=np.uint8)
    if a.size == 0:
        a = np.zeros(1, dtype=np.uint8)
    return a, a.ctypes.data_as(C.POINTER(C.c_uint8))


# ------------------------------------------------------------------------------------------------
# hashing / sketching
# ------------------------------------------------------------------------------------------------
def sketch_params(k: int, canonical: bool = True, scaled: bool = False, scale: int = 1, minimizer_w: int = 0,
                  syncmer_s: int = 0) -> SketchParams:
    return SketchParams(k, int(canonical), int(scaled), scale, int(minimizer_w > 0), minimizer_w,
                        int(syncmer_s > 0), syncmer_s)


def nthash_all(seq: bytes, k: int, canonical: bool = True) -> np.ndarray:
    n = len(seq) - k + 1
    if n <= 0:
        return np.zeros(0, dtype=np.uint64)
    a, p = _u8(seq)
    out = np.zeros(n, dtype=np.uint64)
    lib().ko_nthash_all(p, len(seq), k, int(canonical), out.ctypes.data_as(C.POINTER(C.c_uint64)))
    return out


def generate_kmers(seq: bytes, sp: SketchParams) -> np.ndarray:
    """U:1037-1107 generateKmers for one sequence."""
    n = len(seq) - sp.k + 1
    if n <= 0:
        return np.zeros(0, dtype=np.uint64)
    a, p = _u8(seq)
    out = np.zeros(n, dtype=np.uint64)
    m = lib().ko_generate_kmers(p, len(seq), C.byref(sp), out.ctypes.data_as(C.POINTER(C.c_uint64)))
    return out[:m].copy()


def dedup(codes: np.ndarray, threshold: int = 256) -> np.ndarray:
    c = np.ascontiguousarray(codes, dtype=np.uint64).copy()
    if c.size == 0:
        return c
    n = lib().ko_dedup(c.ctypes.data_as(C.POINTER(C.c_uint64)), c.size, threshold)
    return c[:n]


def hash_values(code: int, h: int) -> List[int]:
    out = (C.c_uint64 * 8)()
    lib().ko_hash_values(code, h, out)
    return [int(out[i]) for i in range(h)]


def query_fpr(n: int, c: int, p: float) -> float:
    return lib().ko_query_fpr(n, c, p)


def calc_signature_size(n_elements: int, num_hashes: int, fpr: float) -> int:
    return int(lib().ko_calc_signature_size(n_elements, num_hashes, fpr))


# ------------------------------------------------------------------------------------------------
# FASTA/Q reading (bio/seqio/fastx semantics: ID = header up to first whitespace; sequences joined over
# lines; FASTQ 4-line records)
# ------------------------------------------------------------------------------------------------
def read_fastx(path: str) -> Iterator[Tuple[bytes, bytes, bytes]]:
    """yields (id, full_header_name, seq)"""
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rb") as fh:
        data = fh.read()
    if not data:
        return
    if data[:1] == b">":
        for rec in data.split(b"\n>"):
            if rec[:1] == b">":
                rec = rec[1:]
            nl = rec.find(b"\n")
            if nl < 0:
                name, seq = rec, b""
            else:
                name, seq = rec[:nl], rec[nl + 1:].replace(b"\n", b"").replace(b"\r", b"")
            name = name.rstrip(b"\r")
            yield name.split(None, 1)[0] if name.strip() else b"", name, seq
    elif data[:1] == b"@":
        lines = data.split(b"\n")
        for i in range(0, len(lines) - 3, 4):
            name = lines[i][1:].rstrip(b"\r")
            yield name.split(None, 1)[0] if name.strip() else b"", name, lines[i + 1].rstrip(b"\r")
    else:
        raise ValueError("not FASTA/Q: " + path)


# ------------------------------------------------------------------------------------------------
# `kmcp compute` (fixture builder): genome → chunks → unique sorted code sets.  C:577-826
# ------------------------------------------------------------------------------------------------
@dataclass
class TargetSet:
    name: str
    chunk_idx: int
    n_chunks: int
    genome_size: int
    codes: np.ndarray  # sorted unique uint64


def compute_targets(records: Sequence[Tuple[bytes, bytes, bytes]], name: str, sp: SketchParams,
                    split_number: int = 1, split_overlap: int = 0, split_min_ref: int = 1000,
                    name_filters: Sequence[str] = ()) -> List[TargetSet]:
    """One genome file → list of TargetSet.  records = [(id, header, seq)]"""
    k = sp.k
    res = [re.compile(p, re.I) for p in name_filters]           # C:587-600 (case-insensitive)
    seqs = [s for (_i, hdr, s) in records if not any(r.search(hdr.decode("latin1")) for r in res)]
    split_seq = split_number > 1
    if not split_seq:
        # non-split mode: every kept record hashed on its own, codes pooled (C:676-681, 805-807, 905-914)
        if not seqs:
            return []
        allc = [generate_kmers(s, sp) for s in seqs]
        codes = np.unique(np.concatenate(allc)) if allc else np.zeros(0, np.uint64)
        if codes.size == 0:
            return []
        return [TargetSet(name, 0, 1, sum(len(s) for s in seqs), codes)]
    if sum(len(s) for s in seqs) == 0:
        return []
    big = seqs[0] if len(seqs) == 1 else (b"N" * (k - 1)).join(seqs)      # C:612-626
    gsize = len(big)
    L = len(big)
    if L < split_min_ref:                                                  # C:676
        size, step = L, L
    else:
        size = (L + (split_number - 1) * split_overlap + split_number - 1) // split_number   # C:691
        step = size - split_overlap
    windows = []
    i = 0
    while i < L:                                                           # seq.Slider(size, step, circular=false, greedy=true)
        w = big[i:i + size]
        if not (len(w) - 1 <= split_overlap or len(w) < k):                # C:713, 742
            windows.append(w)
        if i + size >= L:
            # the reference slider keeps sliding until the start passes the end; remaining windows are pure
            # overlap tails (len-1 <= overlap) and get skipped by the rule above
            pass
        i += step
    out = []
    for idx, w in enumerate(windows):
        codes = np.unique(generate_kmers(w, sp))
        out.append(TargetSet(name, idx, len(windows), gsize, codes))
    return out


# ------------------------------------------------------------------------------------------------
# `.uniki` block codec (X:153-304 / 383-593) and `kmcp index` block assembly (I:667-682, 936-948, 1023, 1157)
# ------------------------------------------------------------------------------------------------
@dataclass
class Block:
    k: int
    canonical: bool
    num_hashes: int
    num_sigs: int
    names: List[str]
    gsizes: List[int]
    indices: List[int]      # chunkIdx | nChunks<<16
    sizes: List[int]
    rows: np.ndarray        # uint8 [num_sigs, row_bytes]

    @property
    def row_bytes(self) -> int:
        return (len(self.names) + 7) // 8


def write_uniki(path: str, b: Block) -> None:
    with open(path, "wb") as f:
        f.write(b".kmcpidx")
        f.write(bytes([4, b.k, 1 if b.canonical else 0, b.num_hashes]))
        f.write(struct.pack(">Q", b.num_sigs))
        f.write(struct.pack(">I", len(b.names)))
        for nm in b.names:
            raw = nm.encode() + b"\n"
            f.write(struct.pack(">I", len(raw)) + raw)
        f.write(struct.pack(">I", len(b.gsizes)))
        for g in b.gsizes:
            f.write(struct.pack(">IQ", 1, g))
        f.write(struct.pack(">I", len(b.indices)))
        for ix in b.indices:
            f.write(struct.pack(">II", 1, ix))
        for s in b.sizes:
            f.write(struct.pack(">Q", s))
        assert b.rows.shape == (b.num_sigs, b.row_bytes) and b.rows.dtype == np.uint8
        f.write(np.ascontiguousarray(b.rows).tobytes())


def read_uniki(path: str) -> Block:
    with open(path, "rb") as f:
        data = f.read()
    assert data[:8] == b".kmcpidx", "kmcp: invalid index format"
    ver, k, flag, nh = data[8:12]
    assert ver == 4, "kmcp: version mismatch"
    num_sigs, = struct.unpack(">Q", data[12:20])
    p = 20
    n, = struct.unpack(">I", data[p:p + 4]); p += 4
    names = []
    for _ in range(n):
        l, = struct.unpack(">I", data[p:p + 4]); p += 4
        names.append(data[p:p + l].split(b"\n")[0].decode()); p += l
    ng, = struct.unpack(">I", data[p:p + 4]); p += 4
    gsizes = []
    for _ in range(ng):
        c, = struct.unpack(">I", data[p:p + 4]); p += 4
        vals = struct.unpack(">%dQ" % c, data[p:p + 8 * c]); p += 8 * c
        gsizes.append(vals[0] if vals else 0)
    ni, = struct.unpack(">I", data[p:p + 4]); p += 4
    indices = []
    for _ in range(ni):
        c, = struct.unpack(">I", data[p:p + 4]); p += 4
        vals = struct.unpack(">%dI" % c, data[p:p + 4 * c]); p += 4 * c
        indices.append(vals[0] if vals else 0)
    sizes = list(struct.unpack(">%dQ" % n, data[p:p + 8 * n])); p += 8 * n
    rb = (n + 7) // 8
    rows = np.frombuffer(data, dtype=np.uint8, count=num_sigs * rb, offset=p).reshape(num_sigs, rb)
    return Block(k, bool(flag & 1), nh, num_sigs, names, gsizes, indices, sizes, rows)


def block_size_for(n_files: int, threads: int) -> int:
    """I:670-682: sBlock = (int(nFiles/threads)+7)/8*8 clamped to [8, nFiles]"""
    sb = (int(n_files / threads) + 7) // 8 * 8
    sb = min(sb, n_files)
    return max(sb, 8)


def assemble_block(targets: Sequence[TargetSet], k: int, num_hashes: int, fpr: float) -> Block:
    max_el = max(int(t.codes.size) for t in targets)                       # I:936-948
    num_sigs = calc_signature_size(max_el, num_hashes, fpr)                # I:1023
    rb = (len(targets) + 7) // 8
    rows = np.zeros((num_sigs, rb), dtype=np.uint8)
    for j, t in enumerate(targets):
        bit = np.uint8(1 << (7 - (j & 7)))                                 # I:1157
        col = j >> 3
        if num_hashes == 1:
            locs = (t.codes % np.uint64(num_sigs)).astype(np.int64)
            rows[locs, col] |= bit
        else:
            a = (t.codes >> np.uint64(32)).astype(np.uint32)
            bb = (t.codes & np.uint64(0xFFFFFFFF)).astype(np.uint32)
            for i in range(num_hashes):                                    # I:1188, util-hash.go:85-102
                v = (a + bb * np.uint32(i)).astype(np.uint32).astype(np.uint64)
                locs = (v % np.uint64(num_sigs)).astype(np.int64)
                rows[locs, col] |= bit
    return Block(k, True, num_hashes, num_sigs, [t.name for t in targets], [t.genome_size for t in targets],
                 [t.chunk_idx | (t.n_chunks << 16) for t in targets], [int(t.codes.size) for t in targets], rows)


def build_db(targets: Sequence[TargetSet], out_dir: str, sp: SketchParams, num_hashes: int = 1, fpr: float = 0.3,
             block_size: int = 0, threads: int = 16, alias: str = "db") -> str:
    """Writes <out_dir>/R001/{__db.yml,_blockNNN.uniki,__name_mapping.tsv}; returns the R001 path (I:1283-1400)."""
    ts = sorted(targets, key=lambda t: int(t.codes.size))                  # I:667 (stable here; reference sort is unstable)
    ts = [t for t in ts if t.codes.size > 0]
    sb = block_size if block_size > 0 else block_size_for(len(ts), threads)
    sb = max(min(sb, len(ts)), 8) if block_size <= 0 else sb
    r001 = os.path.join(out_dir, "R001")
    os.makedirs(r001, exist_ok=True)
    files = []
    total = 0
    for bi in range(0, len(ts), sb):
        blk = assemble_block(ts[bi:bi + sb], sp.k, num_hashes, fpr)
        fn = "_block%03d.uniki" % (len(files) + 1)
        write_uniki(os.path.join(r001, fn), blk)
        files.append(fn)
        total += sum(blk.sizes)
    write_db_yml(os.path.join(r001, "__db.yml"), dict(
        version=4, unikiVersion=4, alias=alias, k=sp.k, ks=[sp.k], hashed=True, canonical=bool(sp.canonical),
        scaled=bool(sp.scaled), scale=int(sp.scale) if sp.scaled else 0, minimizer=bool(sp.minimizer),
        **{"minimizer-w": int(sp.minimizer_w) if sp.minimizer else 0}, syncmer=bool(sp.syncmer),
        **{"syncmer-s": int(sp.syncmer_s) if sp.syncmer else 0},
        **{"split-seq": False, "split-size": 0, "split-num": 0, "split-overlap": 0, "compact-size": False},
        hashes=num_hashes, fpr=fpr, numNameGroups=len(ts), blocksize=sb, totalKmers=total, files=files))
    with open(os.path.join(r001, "__name_mapping.tsv"), "w") as f:
        for nm in sorted({t.name for t in ts}):
            f.write("%s\t%s\n" % (nm, nm))
    return r001


_YML_ORDER = ["version", "unikiVersion", "alias", "k", "ks", "hashed", "canonical", "scaled", "scale", "minimizer",
              "minimizer-w", "syncmer", "syncmer-s", "split-seq", "split-size", "split-num", "split-overlap",
              "compact-size", "hashes", "fpr", "numNameGroups", "blocksize", "totalKmers", "files"]


def write_db_yml(path: str, d: dict) -> None:
    """util-db-info.go:46-79 key order, yaml.v2 style."""
    with open(path, "w") as f:
        for key in _YML_ORDER:
            v = d[key]
            if isinstance(v, bool):
                f.write("%s: %s\n" % (key, "true" if v else "false"))
            elif isinstance(v, list):
                f.write("%s:\n" % key)
                for x in v:
                    f.write("- %s\n" % x)
            elif isinstance(v, float):
                f.write("%s: %s\n" % (key, repr(v)))
            else:
                f.write("%s: %s\n" % (key, v))


# ------------------------------------------------------------------------------------------------
# search
# ------------------------------------------------------------------------------------------------
def pack_seqs(seqs: Sequence[bytes]) -> Tuple[np.ndarray, np.ndarray]:
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    if len(seqs):
        off[1:] = np.cumsum([len(s) for s in seqs], dtype=np.uint64)
    buf = np.frombuffer(b"".join(seqs), dtype=np.uint8) if len(seqs) and off[-1] else np.zeros(1, np.uint8)
    return np.ascontiguousarray(buf), off


@dataclass
class SearchResult:
    query_len: np.ndarray
    n_kmers: np.ndarray
    k_used: np.ndarray
    hit_off: np.ndarray
    hits: np.ndarray    # HIT_DTYPE, per query in final order


class DB:
    def __init__(self, r001_dir: str):
        err = C.create_string_buffer(512)
        self._h = lib().ko_db_open(r001_dir.encode(), err, 512)
        if not self._h:
            raise RuntimeError(err.value.decode())
        self.info = DBInfo()
        lib().ko_db_get_info(self._h, C.byref(self.info))
        self.path = r001_dir

    def close(self):
        if self._h:
            lib().ko_db_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def k(self) -> int:
        return self.info.ks[0]

    def sketch_params(self, k: Optional[int] = None) -> SketchParams:
        i = self.info
        return SketchParams(k or i.ks[0], i.canonical, i.scaled, i.scale, i.minimizer, i.minimizer_w, i.syncmer, i.syncmer_s)

    def target(self, g: int) -> Target:
        t = Target()
        if lib().ko_db_target(self._h, g, C.byref(t)):
            raise IndexError(g)
        return t

    def block(self, b: int):
        ns = C.c_uint64(); rb = C.c_int32(); nn = C.c_int32(); rows = C.POINTER(C.c_uint8)()
        if lib().ko_db_block(self._h, b, C.byref(ns), C.byref(rb), C.byref(nn), C.byref(rows)):
            raise IndexError(b)
        arr = np.ctypeslib.as_array(rows, shape=(ns.value, rb.value))
        return ns.value, rb.value, nn.value, arr

    def count_codes(self, codes: np.ndarray) -> np.ndarray:
        c = np.ascontiguousarray(codes, dtype=np.uint64)
        out = np.zeros(self.info.n_targets, dtype=np.uint32)
        cp = c.ctypes.data_as(C.POINTER(C.c_uint64)) if c.size else C.POINTER(C.c_uint64)()
        lib().ko_count_codes(self._h, cp, c.size, out.ctypes.data_as(C.POINTER(C.c_uint32)))
        return out

    def search(self, seqs: Sequence[bytes] = None, packed=None, paired: bool = False, opts: Optional[SearchOpts] = None,
               threads: int = 0, algo: int = 0) -> SearchResult:
        if opts is None:
            opts = default_opts()
        buf, off = packed if packed is not None else pack_seqs(seqs)
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        r = Results()
        lib().ko_search(self._h, C.byref(opts), buf.ctypes.data_as(C.POINTER(C.c_uint8)),
                        off.ctypes.data_as(C.POINTER(C.c_uint64)), len(off) - 1, int(paired), threads, algo, C.byref(r))
        nq = r.n_queries
        def arr(p, n, dt):
            return np.ctypeslib.as_array(p, shape=(n,)).astype(dt, copy=True) if n else np.zeros(0, dt)
        out = SearchResult(arr(r.query_len, nq, np.int32), arr(r.n_kmers, nq, np.int32), arr(r.k_used, nq, np.int32),
                           arr(r.hit_off, nq + 1, np.uint64),
                           np.frombuffer(C.string_at(r.hits, int(r.n_hits) * C.sizeof(Hit)), dtype=HIT_DTYPE).copy()
                           if r.n_hits else np.zeros(0, HIT_DTYPE))
        lib().ko_free_results(C.byref(r))
        return out


def default_opts() -> SearchOpts:
    o = SearchOpts()
    lib().ko_default_opts(C.byref(o))
    return o


def go_fmt_e4(x: float) -> str:
    """strconv.FormatFloat(x,'e',4,64) == C printf %.4e (two-digit exponent minimum)"""
    return "%.4e" % x


TSV_HEADER = "#query\tqLen\tqKmers\tFPR\thits\ttarget\tchunkIdx\tchunks\ttLen\tkSize\tmKmers\tqCov\ttCov\tjacc\tqueryIdx\n"


def format_tsv(db: DB, ids: Sequence[bytes], res: SearchResult, keep_unmatched: bool = False, header: bool = True,
               trailer: bool = True) -> str:
    """S:437, 460-575, 1023-1025"""
    out = [TSV_HEADER] if header else []
    matched = 0
    for q in range(len(ids)):
        a, b = int(res.hit_off[q]), int(res.hit_off[q + 1])
        qid = ids[q].decode("latin1")
        if a == b:
            if keep_unmatched:
                out.append("%s\t%d\t%d\t0\t0\t\t-1\t0\t0\t%d\t0\t0\t0\t0\t%d\n" % (qid, res.query_len[q], res.n_kmers[q], res.k_used[q], q))
            continue
        matched += 1
        for h in res.hits[a:b]:
            t = db.target(int(h["target"]))
            out.append("%s\t%d\t%d\t%s\t%d\t%s\t%d\t%d\t%d\t%d\t%d\t%.4f\t%.4f\t%.4f\t%d\n" % (
                qid, res.query_len[q], res.n_kmers[q], go_fmt_e4(float(h["fpr"])), b - a, t.name.decode(),
                t.index & 0xFFFF, t.index >> 16, t.genome_size, res.k_used[q], int(h["count"]),
                float(h["qcov"]), float(h["tcov"]), float(h["jacc"]), q))
    if trailer:
        n = len(ids)
        out.append("# input queries: %d\n# matched queries: %d\n# matched percentage: %.4f%%\n" % (
            n, matched, (matched / n * 100) if n else float("nan")))
    return "".join(out)


def profile_stage1(tsv: str, min_qcov: float = 0.55, max_fpr: float = 0.01, top_n_scores: int = 0, keep_perfect: bool = False,
                   keep_main: bool = False, max_qcov_gap: float = 0.4, hic_min_qcov: float = 0.75):
    """`kmcp profile` stage 1/4 (profile.go:761-990) on the text of a search result, without taxonomy (`--level species`
    off): per reference, per chunk, Match (every kept row adds 1/rows-of-that-reference), UniqMatch (reads whose kept
    rows name one reference) and UniqMatchHic (those with qCov >= -H).  Values are parsed from the TSV text exactly as
    parseMatchResult does (util-profile.go:94-182): rows with qCov < -t or FPR > -f never reach the loop.
    Returns (n_reads, {reference: (genome_size, match[], uniq[], uniq_hic[])})."""
    profile = {}
    n_reads = 0
    state = {"matches": {}}

    def flush():
        nonlocal n_reads
        matches = state["matches"]
        if not matches:
            return
        n_reads += 1
        for name, ms in matches.items():
            first = True
            for (frag, idx_num, gsize, qcov) in ms:
                t = profile.get(name)
                if t is None:
                    t = profile[name] = (gsize, [0.0] * idx_num, [0.0] * idx_num, [0.0] * idx_num)
                if first:
                    if len(matches) == 1:
                        t[2][frag] += 1
                        if qcov >= hic_min_qcov:
                            t[3][frag] += 1
                    first = False
                t[1][frag] += 1.0 / float(len(ms))
        state["matches"] = {}

    prev = None
    p_score, n_score, process = 1024.0, 0, True
    for line in tsv.split("\n"):
        if line == "" or line[0] == "#":
            continue
        it = line.split("\t")
        qcov = float(it[11])
        if qcov < min_qcov:
            continue
        if float(it[3]) > max_fpr:
            continue
        query = it[0]
        if prev != query:
            flush()
            p_score, n_score, process = 1024.0, 0, True
        elif keep_perfect:
            if not process:
                prev = query
                continue
            if p_score == 1 and qcov < 1:
                process = False
                prev = query
                continue
        elif keep_main and p_score <= 1:
            if not process:
                prev = query
                continue
            if p_score - qcov > max_qcov_gap:
                process = False
                prev = query
                continue
        if top_n_scores > 0:
            if not process:
                prev = query
                continue
            if qcov < p_score:
                n_score += 1
                if n_score > top_n_scores:
                    process = False
                    prev = query
                    continue
        state["matches"].setdefault(it[5], []).append((int(it[6]), int(it[7]), int(it[8]), qcov))
        prev = query
        p_score = qcov
    flush()
    return n_reads, profile


# ------------------------------------------------------------------------------------------------
# seeded synthetic data — the SAME pure functions are implemented in kmcp_b200/csrc/synth.cu so that
# bench-scale inputs can be made on the device; tests compare the two generators byte for byte.
# ------------------------------------------------------------------------------------------------
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = (x.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def synth_genome_bases(seed: int, genome: int, start: int, length: int) -> np.ndarray:
    """2-bit codes of genome `genome`, positions [start, start+length)"""
    with np.errstate(over="ignore"):
        gkey = splitmix64(np.array([(seed * 0x100000001B3 + genome) & 0xFFFFFFFFFFFFFFFF], dtype=np.uint64))[0]
        pos = np.arange(start, start + length, dtype=np.uint64)
        w = splitmix64(gkey + (pos >> np.uint64(5)))
        return ((w >> (np.uint64(2) * (pos & np.uint64(31)))) & np.uint64(3)).astype(np.uint8)


def synth_genome(seed: int, genome: int, length: int) -> bytes:
    return _ACGT[synth_genome_bases(seed, genome, 0, length)].tobytes()


def synth_read(seed: int, r: int, n_genomes: int, genome_len: int, read_len: int, gseed: int) -> bytes:
    """read r: 80 % sampled from a genome (random strand, 1 % substitutions), 20 % uniform random"""
    with np.errstate(over="ignore"):
        u = int(splitmix64(np.array([(seed * 0x100000001B3 + r) & 0xFFFFFFFFFFFFFFFF], dtype=np.uint64))[0])
        j = np.arange(read_len, dtype=np.uint64)
        if (u & 0xFF) < 204 and genome_len >= read_len:
            g = ((u >> 8) & 0x7FFFFF) % n_genomes
            pos = (u >> 32) % (genome_len - read_len + 1)
            b = synth_genome_bases(gseed, g, pos, read_len)
            v = splitmix64(np.uint64(u) ^ (j * np.uint64(0xD1342543DE82EF95)))
            sub = (v & np.uint64(0xFFFF)) < np.uint64(655)
            b = np.where(sub, (b + np.uint8(1) + ((v >> np.uint64(16)) % np.uint64(3)).astype(np.uint8)) & np.uint8(3), b).astype(np.uint8)
            if ((u >> 31) & 1) == 1:                      # reverse complement
                b = (np.uint8(3) - b)[::-1]
            return _ACGT[b].tobytes()
        w = splitmix64(np.uint64(u) + (j >> np.uint64(5)))
        b = ((w >> (np.uint64(2) * (j & np.uint64(31)))) & np.uint64(3)).astype(np.uint8)
        return _ACGT[b].tobytes()

"""ctypes binding of libkmcp_gpu.so (the C ABI in include/kmcp_gpu.h).

This module is plumbing for tests and bench.py: the product is the shared library.  There is no Python or
CPU implementation of the search path here — if the library is missing, or no CUDA device is present,
every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libkmcp_gpu.so")

KMCPG_OK, KMCPG_EINVAL, KMCPG_EIO, KMCPG_EFORMAT, KMCPG_ECUDA, KMCPG_ENOMEM, KMCPG_EUNSUPPORTED = 0, -1, -2, -3, -4, -5, -6


class KmcpGpuError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("libkmcp_gpu error %d: %s" % (code, msg))
        self.code = code


class DbOpts(C.Structure):
    _fields_ = [("shard_rank", C.c_int32), ("shard_world", C.c_int32), ("max_resident_bytes", C.c_int64)]


class ShardPiece(C.Structure):
    _fields_ = [("block", C.c_int32), ("shard", C.c_int32), ("col0", C.c_uint32), ("n_cols", C.c_uint32), ("resident_bytes", C.c_uint64)]


class DbInfo(C.Structure):
    _fields_ = [("n_ks", C.c_int32), ("ks", C.c_int32 * 8), ("canonical", C.c_int32), ("num_hashes", C.c_int32),
                ("scaled", C.c_int32), ("scale", C.c_uint32), ("minimizer", C.c_int32), ("minimizer_w", C.c_uint32),
                ("syncmer", C.c_int32), ("syncmer_s", C.c_uint32), ("fpr", C.c_double), ("n_blocks", C.c_int32),
                ("n_resident_blocks", C.c_int32), ("n_targets", C.c_int64), ("sum_row_bytes", C.c_int64),
                ("resident_bytes", C.c_int64), ("disk_bytes", C.c_int64)]


class TargetInfo(C.Structure):
    _fields_ = [("name", C.c_char_p), ("index", C.c_uint32), ("genome_size", C.c_uint64), ("n_kmers", C.c_uint64),
                ("block", C.c_int32), ("col", C.c_int32), ("resident", C.c_int32)]


class SearchParams(C.Structure):
    _fields_ = [("min_query_len", C.c_int32), ("min_matched", C.c_int32), ("dedup_threshold", C.c_int32),
                ("paired", C.c_int32), ("min_query_cov", C.c_double), ("k", C.c_int32), ("mate_select", C.c_int32)]


class Hit(C.Structure):
    _fields_ = [("query", C.c_uint32), ("target", C.c_uint32), ("count", C.c_uint32)]


HIT_DTYPE = np.dtype([("query", "<u4"), ("target", "<u4"), ("count", "<u4")])


class Hits(C.Structure):
    _fields_ = [("n_queries", C.c_uint32), ("n_hits", C.c_uint64), ("n_kmers", C.POINTER(C.c_int32)),
                ("query_len", C.POINTER(C.c_int32)), ("hits", C.POINTER(Hit)), ("ms_hash", C.c_float), ("ms_locs", C.c_float),
                ("ms_probe", C.c_float), ("ms_total", C.c_float), ("probe_launches", C.c_uint32), ("probe_row_bytes", C.c_uint64),
                ("kernel_launches", C.c_uint32), ("_priv", C.c_void_p)]


class Part(C.Structure):
    _fields_ = [("first_query", C.c_uint32), ("n_queries", C.c_uint32), ("n_kmers", C.POINTER(C.c_int32)), ("query_len", C.POINTER(C.c_int32)),
                ("hits", C.POINTER(Hit)), ("n_hits", C.c_uint64)]


PART_CB = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(Part))


class Batch(C.Structure):
    _fields_ = [("seq", C.c_void_p), ("off", C.c_void_p), ("n_seqs", C.c_uint32), ("on_device", C.c_int32), ("host_off", C.c_void_p),
                ("ready_event", C.c_void_p), ("hits_dst", C.c_void_p), ("hits_cap", C.c_uint64), ("cb", C.c_void_p), ("user", C.c_void_p),
                ("first_query", C.c_uint32), ("_pad", C.c_uint32)]


class SketchParams(C.Structure):
    _fields_ = [("k", C.c_int32), ("canonical", C.c_int32), ("scaled", C.c_int32), ("scale", C.c_uint32),
                ("minimizer", C.c_int32), ("minimizer_w", C.c_uint32), ("syncmer", C.c_int32), ("syncmer_s", C.c_uint32)]


class EngineOpts(C.Structure):
    _fields_ = [("min_query_len", C.c_int32), ("min_matched", C.c_int32), ("dedup_threshold", C.c_int32),
                ("min_query_cov", C.c_double), ("min_target_cov", C.c_double), ("max_fpr", C.c_double),
                ("sort_by", C.c_int32), ("do_not_sort", C.c_int32), ("top_n_scores", C.c_int32), ("try_se", C.c_int32),
                ("paired", C.c_int32), ("threads", C.c_int32)]


class Match(C.Structure):
    _fields_ = [("query", C.c_uint32), ("target", C.c_uint32), ("count", C.c_uint32), ("_pad", C.c_uint32),
                ("fpr", C.c_double), ("qcov", C.c_double), ("tcov", C.c_double), ("jacc", C.c_double)]


MATCH_DTYPE = np.dtype([("query", "<u4"), ("target", "<u4"), ("count", "<u4"), ("_pad", "<u4"),
                        ("fpr", "<f8"), ("qcov", "<f8"), ("tcov", "<f8"), ("jacc", "<f8")])


class Results(C.Structure):
    _fields_ = [("n_queries", C.c_uint32), ("n_matches", C.c_uint64), ("query_len", C.POINTER(C.c_int32)),
                ("n_kmers", C.POINTER(C.c_int32)), ("k_used", C.POINTER(C.c_int32)), ("match_off", C.POINTER(C.c_uint64)),
                ("matches", C.POINTER(Match)), ("ms_gpu_total", C.c_float), ("ms_post", C.c_float), ("ms_total", C.c_float), ("probe_row_bytes", C.c_uint64),
                ("kernel_launches", C.c_uint32), ("_priv", C.c_void_p)]


class IndexParams(C.Structure):
    _fields_ = [("k", C.c_int32), ("num_hashes", C.c_int32), ("fpr", C.c_double), ("split_number", C.c_int32), ("split_overlap", C.c_int32),
                ("split_min_ref", C.c_int32), ("scale", C.c_uint32), ("minimizer_w", C.c_uint32), ("syncmer_s", C.c_uint32), ("block_size", C.c_int32),
                ("threads", C.c_int32), ("ref_name_regexp", C.c_char_p), ("seq_name_filters", C.POINTER(C.c_char_p)), ("n_seq_name_filters", C.c_int32)]


class SynthDb(C.Structure):
    _fields_ = [("genome_seed", C.c_uint64), ("n_genomes", C.c_uint32), ("genome_len", C.c_uint32), ("k", C.c_int32),
                ("n_chunks", C.c_int32), ("overlap", C.c_int32), ("num_hashes", C.c_int32), ("fpr", C.c_double),
                ("block_size", C.c_int32), ("scale", C.c_uint32), ("shard_rank", C.c_int32), ("shard_world", C.c_int32)]


# every symbol include/kmcp_gpu.h declares (checked by tests/test_abi.py without a GPU)
class RefcountParams(C.Structure):
    _fields_ = [("min_query_cov", C.c_double), ("max_fpr", C.c_double), ("top_n_scores", C.c_int32), ("keep_perfect", C.c_int32),
                ("keep_main", C.c_int32), ("max_qcov_gap", C.c_double), ("hic_min_qcov", C.c_double)]


class RefcountRow(C.Structure):
    _fields_ = [("name", C.c_char_p), ("genome_size", C.c_uint64), ("n_chunks", C.c_uint32), ("_pad", C.c_uint32),
                ("match", C.POINTER(C.c_double)), ("uniq_match", C.POINTER(C.c_double)), ("uniq_match_hic", C.POINTER(C.c_double))]


class RefcountTable(C.Structure):
    _fields_ = [("n_reads", C.c_uint64), ("n_refs", C.c_uint32), ("_pad", C.c_uint32), ("rows", C.POINTER(RefcountRow))]


READER_LOG = C.CFUNCTYPE(None, C.c_void_p, C.c_char_p, C.c_char_p)


class ReaderOpts(C.Structure):
    _fields_ = [("read1", C.c_char_p), ("read2", C.c_char_p), ("files", C.POINTER(C.c_char_p)), ("n_files", C.c_int32), ("whole_file", C.c_int32),
                ("use_filename", C.c_int32), ("query_id", C.c_char_p), ("k", C.c_int32), ("batch_reads", C.c_uint32), ("batch_bytes", C.c_uint64),
                ("inflate_threads", C.c_int32), ("parse_threads", C.c_int32), ("log", READER_LOG), ("log_user", C.c_void_p),
                ("inflate_chunk", C.c_uint64), ("inflate_cap", C.c_uint64), ("parse_piece", C.c_uint64)]


class ReadBatch(C.Structure):
    _fields_ = [("n_queries", C.c_uint32), ("n_seqs", C.c_uint32), ("seq", C.POINTER(C.c_uint8)), ("off", C.POINTER(C.c_uint64)),
                ("ids", C.POINTER(C.c_char)), ("id_off", C.POINTER(C.c_uint64)), ("first_query", C.c_uint64), ("_priv", C.c_void_p)]


ABI_SYMBOLS = [
    "kmcpg_abi_version", "kmcpg_set_stream", "kmcpg_create", "kmcpg_close", "kmcpg_last_error", "kmcpg_shard_plan", "kmcpg_shard_pieces", "kmcpg_open_db", "kmcpg_db_info", "kmcpg_target",
    "kmcpg_target_sizes", "kmcpg_shm_open", "kmcpg_shm_close", "kmcpg_engine_postfilter", "kmcpg_merge_hits", "kmcpg_hits_digest",
    "kmcpg_default_params", "kmcpg_search_batch", "kmcpg_search_batch_device", "kmcpg_search_batch_cb", "kmcpg_search_submit", "kmcpg_search_wait",
    "kmcpg_free_hits", "kmcpg_host_alloc",
    "kmcpg_host_free", "kmcpg_device_memory", "kmcpg_device_alloc", "kmcpg_device_free", "kmcpg_memcpy_h2d", "kmcpg_memcpy_d2h",
    "kmcpg_generate_kmers", "kmcpg_count_codes", "kmcpg_free", "kmcpg_default_engine_opts", "kmcpg_engine_search", "kmcpg_engine_search_sharded", "kmcpg_engine_search_replicas",
    "kmcpg_free_results", "kmcpg_query_fpr", "kmcpg_default_index_params", "kmcpg_index_fasta", "kmcpg_synth_reads", "kmcpg_synth_genomes", "kmcpg_build_synth_db", "kmcpg_write_block",
    "kmcpg_default_refcount_params", "kmcpg_refcounts_create", "kmcpg_refcounts_add", "kmcpg_refcounts_get", "kmcpg_refcounts_free",
    "kmcpg_default_reader_opts", "kmcpg_reader_open", "kmcpg_reader_next", "kmcpg_reader_free_batch", "kmcpg_reader_error", "kmcpg_reader_close",
]

_lib = None


def load() -> C.CDLL:
    """Loads libkmcp_gpu.so; raises if it has not been built (there is no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("kmcp_b200/libkmcp_gpu.so is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(make -C kmcp_b200/csrc). The search path has no CPU or Python fallback.")
    L = C.CDLL(LIB_PATH)
    u8p, u64p, vp = C.POINTER(C.c_uint8), C.POINTER(C.c_uint64), C.c_void_p
    L.kmcpg_abi_version.restype = C.c_int
    L.kmcpg_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.kmcpg_close.argtypes = [vp]
    L.kmcpg_set_stream.argtypes = [vp, vp]
    L.kmcpg_last_error.restype = C.c_char_p
    L.kmcpg_last_error.argtypes = [vp]
    L.kmcpg_open_db.argtypes = [vp, C.c_char_p, C.POINTER(DbOpts)]
    L.kmcpg_shard_plan.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_int32), C.c_int32]
    L.kmcpg_shard_pieces.argtypes = [C.c_char_p, C.c_int, C.POINTER(ShardPiece), C.c_int32]
    L.kmcpg_db_info.argtypes = [vp, C.POINTER(DbInfo)]
    L.kmcpg_target.argtypes = [vp, C.c_int64, C.POINTER(TargetInfo)]
    L.kmcpg_target_sizes.argtypes = [vp, vp, C.c_int64]
    L.kmcpg_shm_open.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.POINTER(vp)]
    L.kmcpg_shm_close.argtypes = [C.c_char_p, vp, C.c_size_t, C.c_int, C.c_int]
    L.kmcpg_engine_postfilter.argtypes = [C.POINTER(EngineOpts), C.c_uint32, vp, vp, vp, C.c_uint64, vp, C.c_int64, C.c_double, C.c_int, C.c_uint32, C.POINTER(Results)]
    L.kmcpg_merge_hits.argtypes = [C.POINTER(vp), C.POINTER(C.c_uint64), C.c_int, C.c_uint32, C.c_uint32, C.c_int, vp]
    L.kmcpg_hits_digest.argtypes = [vp, C.c_uint64, C.c_uint64]
    L.kmcpg_hits_digest.restype = C.c_uint64
    L.kmcpg_default_params.argtypes = [C.POINTER(SearchParams)]
    L.kmcpg_default_params.restype = None
    L.kmcpg_search_batch.argtypes = [vp, C.POINTER(SearchParams), vp, vp, C.c_uint32, C.POINTER(Hits)]
    L.kmcpg_search_batch_device.argtypes = [vp, C.POINTER(SearchParams), vp, vp, C.c_uint32, C.c_uint64, C.POINTER(Hits)]
    L.kmcpg_search_batch_cb.argtypes = [vp, C.POINTER(SearchParams), vp, vp, C.c_uint32, PART_CB, vp, C.POINTER(Hits)]
    L.kmcpg_search_submit.argtypes = [vp, C.POINTER(SearchParams), C.POINTER(Batch), C.POINTER(vp)]
    L.kmcpg_search_wait.argtypes = [vp, C.POINTER(Hits)]
    L.kmcpg_free_hits.argtypes = [C.POINTER(Hits)]
    L.kmcpg_free_hits.restype = None
    L.kmcpg_host_alloc.argtypes = [C.POINTER(vp), C.c_size_t]
    L.kmcpg_host_free.argtypes = [vp]
    L.kmcpg_device_memory.argtypes = [vp, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    L.kmcpg_device_alloc.argtypes = [vp, C.POINTER(vp), C.c_size_t]
    L.kmcpg_device_free.argtypes = [vp, vp]
    L.kmcpg_memcpy_h2d.argtypes = [vp, vp, vp, C.c_size_t]
    L.kmcpg_memcpy_d2h.argtypes = [vp, vp, vp, C.c_size_t]
    L.kmcpg_generate_kmers.argtypes = [vp, C.POINTER(SketchParams), vp, vp, C.c_uint32, C.POINTER(u64p), C.POINTER(u64p)]
    L.kmcpg_count_codes.argtypes = [vp, vp, C.c_uint64, vp]
    L.kmcpg_free.argtypes = [vp]
    L.kmcpg_free.restype = None
    L.kmcpg_default_engine_opts.argtypes = [C.POINTER(EngineOpts)]
    L.kmcpg_default_engine_opts.restype = None
    L.kmcpg_engine_search.argtypes = [vp, C.POINTER(EngineOpts), vp, vp, C.c_uint32, C.POINTER(Results)]
    L.kmcpg_engine_search_sharded.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(EngineOpts), vp, vp, C.c_uint32, C.POINTER(Results)]
    L.kmcpg_engine_search_replicas.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(EngineOpts), vp, vp, C.c_uint32, C.POINTER(Results)]
    L.kmcpg_internal_merge_hits.argtypes = [C.POINTER(vp), C.POINTER(C.c_uint64), C.c_int, vp, C.c_uint32, C.c_uint32, C.c_int]     # test hook, not part of the ABI
    L.kmcpg_internal_merge_hits.restype = None
    L.kmcpg_free_results.argtypes = [C.POINTER(Results)]
    L.kmcpg_default_reader_opts.argtypes = [C.POINTER(ReaderOpts)]
    L.kmcpg_default_reader_opts.restype = None
    L.kmcpg_reader_open.argtypes = [C.POINTER(ReaderOpts), C.POINTER(vp)]
    L.kmcpg_reader_next.argtypes = [vp, C.POINTER(ReadBatch)]
    L.kmcpg_reader_free_batch.argtypes = [C.POINTER(ReadBatch)]
    L.kmcpg_reader_free_batch.restype = None
    L.kmcpg_reader_error.argtypes = [vp]
    L.kmcpg_reader_error.restype = C.c_char_p
    L.kmcpg_reader_close.argtypes = [vp]
    L.kmcpg_free_results.restype = None
    L.kmcpg_default_refcount_params.argtypes = [C.POINTER(RefcountParams)]
    L.kmcpg_default_refcount_params.restype = None
    L.kmcpg_refcounts_create.argtypes = [vp, C.c_char_p, C.POINTER(RefcountParams), C.POINTER(vp)]
    L.kmcpg_refcounts_add.argtypes = [vp, C.POINTER(Results)]
    L.kmcpg_refcounts_get.argtypes = [vp, C.POINTER(RefcountTable)]
    L.kmcpg_refcounts_free.argtypes = [vp]
    L.kmcpg_refcounts_free.restype = None
    L.kmcpg_query_fpr.restype = C.c_double
    L.kmcpg_query_fpr.argtypes = [C.c_int, C.c_int, C.c_double]
    L.kmcpg_default_index_params.argtypes = [C.POINTER(IndexParams)]
    L.kmcpg_default_index_params.restype = None
    L.kmcpg_index_fasta.argtypes = [vp, C.POINTER(IndexParams), C.POINTER(C.c_char_p), C.c_int, C.c_char_p]
    L.kmcpg_synth_reads.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, vp]
    L.kmcpg_build_synth_db.argtypes = [vp, C.POINTER(SynthDb)]
    L.kmcpg_synth_genomes.argtypes = [vp, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, vp]
    L.kmcpg_write_block.argtypes = [vp, C.c_int, C.c_char_p]
    _lib = L
    return L


def pack_seqs(seqs: Sequence[bytes]) -> Tuple[np.ndarray, np.ndarray]:
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    if len(seqs):
        off[1:] = np.cumsum([len(s) for s in seqs], dtype=np.uint64)
    joined = b"".join(seqs)
    buf = np.frombuffer(joined, dtype=np.uint8).copy() if joined else np.zeros(1, np.uint8)
    return buf, off


@dataclass
class BatchHits:
    n_kmers: np.ndarray
    query_len: np.ndarray
    hits: np.ndarray            # HIT_DTYPE sorted by (query, target)
    ms_hash: float
    ms_locs: float
    ms_probe: float
    ms_total: float
    probe_launches: int
    probe_row_bytes: int
    kernel_launches: int
    n_hits: int = 0


@dataclass
class EngineResults:
    query_len: np.ndarray
    n_kmers: np.ndarray
    k_used: np.ndarray
    match_off: np.ndarray
    matches: np.ndarray         # MATCH_DTYPE
    ms_gpu_total: float
    probe_row_bytes: int
    kernel_launches: int
    n_matches: int = 0
    ms_post: float = 0.0
    ms_total: float = 0.0


def _np_from(ptr, n, ctype_size, dtype):
    """one copy of n records behind a ctypes pointer into a numpy array the caller owns"""
    if not n:
        return np.zeros(0, dtype)
    addr = C.cast(ptr, C.c_void_p).value
    view = (C.c_char * (int(n) * ctype_size)).from_address(addr)
    return np.frombuffer(view, dtype=np.uint8).copy().view(dtype)      # a byte copy: ten times faster than copying records field by field


class Context:
    """One GPU context == one kmcpg_ctx (one device)."""

    def __init__(self, device: int = 0):
        self._L = load()
        h = C.c_void_p()
        rc = self._L.kmcpg_create(device, C.byref(h))
        if rc:
            raise KmcpGpuError(rc, self._L.kmcpg_last_error(None).decode())
        self._h = h
        self.device = device

    def _check(self, rc: int):
        if rc:
            raise KmcpGpuError(rc, self._L.kmcpg_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self._L.kmcpg_close(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream: int = 0):
        self._check(self._L.kmcpg_set_stream(self._h, cuda_stream or None))

    # ---- database ----
    def open_db(self, r001_dir: str, shard_rank: int = 0, shard_world: int = 1, max_resident_bytes: int = 0):
        o = DbOpts(shard_rank, shard_world, max_resident_bytes)
        self._check(self._L.kmcpg_open_db(self._h, r001_dir.encode(), C.byref(o)))

    def build_synth_db(self, genome_seed: int, n_genomes: int, genome_len: int, k: int = 21, n_chunks: int = 10,
                       overlap: int = 150, num_hashes: int = 1, fpr: float = 0.3, block_size: int = 0, scale: int = 1,
                       shard_rank: int = 0, shard_world: int = 1):
        s = SynthDb(genome_seed, n_genomes, genome_len, k, n_chunks, overlap, num_hashes, fpr, block_size, scale, shard_rank, shard_world)
        self._check(self._L.kmcpg_build_synth_db(self._h, C.byref(s)))

    def index_fasta(self, files, out_dir: str, k: int = 21, num_hashes: int = 1, fpr: float = 0.3, split_number: int = 1, split_overlap: int = -1,
                    split_min_ref: int = 1000, scale: int = 1, minimizer_w: int = 0, syncmer_s: int = 0, block_size: int = 0, threads: int = 16,
                    ref_name_regexp: Optional[str] = None, seq_name_filters: Sequence[str] = ()):
        """kmcp compute + kmcp index on the GPU; the new database stays open in this context"""
        p = IndexParams()
        self._L.kmcpg_default_index_params(C.byref(p))
        p.k, p.num_hashes, p.fpr, p.split_number, p.split_overlap, p.split_min_ref = k, num_hashes, fpr, split_number, split_overlap, split_min_ref
        p.scale, p.minimizer_w, p.syncmer_s, p.block_size, p.threads = scale, minimizer_w, syncmer_s, block_size, threads
        p.ref_name_regexp = ref_name_regexp.encode() if ref_name_regexp else None
        flt = (C.c_char_p * max(1, len(seq_name_filters)))(*[f.encode() for f in seq_name_filters])
        p.seq_name_filters = C.cast(flt, C.POINTER(C.c_char_p)); p.n_seq_name_filters = len(seq_name_filters)
        arr = (C.c_char_p * len(files))(*[f.encode() for f in files])
        self._check(self._L.kmcpg_index_fasta(self._h, C.byref(p), arr, len(files), out_dir.encode() if out_dir else None))

    def write_block(self, resident_block: int, path: str):
        self._check(self._L.kmcpg_write_block(self._h, resident_block, path.encode()))

    def db_info(self) -> DbInfo:
        i = DbInfo()
        self._check(self._L.kmcpg_db_info(self._h, C.byref(i)))
        return i

    def target_sizes(self) -> np.ndarray:
        out = np.zeros(int(self.db_info().n_targets), dtype=np.float64)
        self._check(self._L.kmcpg_target_sizes(self._h, out.ctypes.data, out.size))
        return out

    def target(self, g: int) -> TargetInfo:
        t = TargetInfo()
        self._check(self._L.kmcpg_target(self._h, g, C.byref(t)))
        return t

    # ---- hot path ----
    def default_params(self, **kw) -> SearchParams:
        p = SearchParams()
        self._L.kmcpg_default_params(C.byref(p))
        for k, v in kw.items():
            setattr(p, k, v)
        return p

    def _take_hits(self, h: Hits, copy: bool = True) -> BatchHits:
        if not copy:       # timing harness: only the summary, no Python-side copies of the result arrays
            out = BatchHits(np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(int(h.n_hits), np.uint8)[:0], h.ms_hash, h.ms_locs, h.ms_probe,
                            h.ms_total, int(h.probe_launches), int(h.probe_row_bytes), int(h.kernel_launches))
            out.n_hits = int(h.n_hits)
            self._L.kmcpg_free_hits(C.byref(h))
            return out
        out = BatchHits(_np_from(h.n_kmers, h.n_queries, 4, np.int32), _np_from(h.query_len, h.n_queries, 4, np.int32),
                        _np_from(h.hits, h.n_hits, C.sizeof(Hit), HIT_DTYPE), h.ms_hash, h.ms_locs, h.ms_probe, h.ms_total,
                        int(h.probe_launches), int(h.probe_row_bytes), int(h.kernel_launches))
        out.n_hits = int(h.n_hits)
        self._L.kmcpg_free_hits(C.byref(h))
        return out

    def search_batch(self, buf: np.ndarray, off: np.ndarray, params: Optional[SearchParams] = None) -> BatchHits:
        p = params or self.default_params()
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        h = Hits()
        self._check(self._L.kmcpg_search_batch(self._h, C.byref(p), buf.ctypes.data, off.ctypes.data, len(off) - 1, C.byref(h)))
        return self._take_hits(h)

    def search_batch_ptr(self, seq_ptr: int, off_ptr: int, n_seqs: int, params: SearchParams, device: bool, seq_bytes: int = 0,
                         copy: bool = True) -> BatchHits:
        """raw-pointer form (pinned host buffers or device buffers owned by the caller)"""
        h = Hits()
        if device:
            self._check(self._L.kmcpg_search_batch_device(self._h, C.byref(params), seq_ptr, off_ptr, n_seqs, seq_bytes, C.byref(h)))
        else:
            self._check(self._L.kmcpg_search_batch(self._h, C.byref(params), seq_ptr, off_ptr, n_seqs, C.byref(h)))
        return self._take_hits(h, copy)

    def submit(self, seq_ptr: int, off_ptr: int, n_seqs: int, params: SearchParams, device: bool = False, host_off_ptr: int = 0,
               hits_dst: int = 0, hits_cap: int = 0, ready_event: int = 0, first_query: int = 0) -> int:
        """kmcpg_search_submit with raw pointers (the caller keeps the buffers alive until wait()); returns the job handle"""
        b = Batch(seq_ptr, off_ptr, n_seqs, 1 if device else 0, host_off_ptr or None, ready_event or None, hits_dst or None, hits_cap, None, None, first_query, 0)
        job = C.c_void_p()
        self._check(self._L.kmcpg_search_submit(self._h, C.byref(params), C.byref(b), C.byref(job)))
        return job.value

    def wait(self, job: int, copy=True) -> BatchHits:
        """kmcpg_search_wait: blocks until the job is done; copy=False returns only the summary (hits stay in hits_dst, if one was given),
        copy="meta" the summary plus n_kmers / query_len"""
        h = Hits()
        self._check(self._L.kmcpg_search_wait(job, C.byref(h)))
        if copy == "meta":
            out = BatchHits(_np_from(h.n_kmers, h.n_queries, 4, np.int32), _np_from(h.query_len, h.n_queries, 4, np.int32), np.zeros(0, HIT_DTYPE),
                            h.ms_hash, h.ms_locs, h.ms_probe, h.ms_total, int(h.probe_launches), int(h.probe_row_bytes), int(h.kernel_launches))
            out.n_hits = int(h.n_hits)
            self._L.kmcpg_free_hits(C.byref(h))
            return out
        return self._take_hits(h, copy)

    def search_batch_streaming(self, buf: np.ndarray, off: np.ndarray, params: Optional[SearchParams] = None):
        """kmcpg_search_batch_cb: returns (list of per-part (first_query, n_queries, n_kmers, hits) copies, summary)"""
        p = params or self.default_params()
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        parts = []

        def cb(_user, part):
            pt = part.contents
            parts.append((pt.first_query, pt.n_queries, _np_from(pt.n_kmers, pt.n_queries, 4, np.int32),
                          _np_from(pt.hits, pt.n_hits, C.sizeof(Hit), HIT_DTYPE)))

        h = Hits()
        self._check(self._L.kmcpg_search_batch_cb(self._h, C.byref(p), buf.ctypes.data, off.ctypes.data, len(off) - 1, PART_CB(cb), None, C.byref(h)))
        return parts, self._take_hits(h)

    def generate_kmers(self, buf: np.ndarray, off: np.ndarray, sp: SketchParams) -> Tuple[np.ndarray, np.ndarray]:
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        n = len(off) - 1
        pc, po = C.POINTER(C.c_uint64)(), C.POINTER(C.c_uint64)()
        self._check(self._L.kmcpg_generate_kmers(self._h, C.byref(sp), buf.ctypes.data, off.ctypes.data, n, C.byref(pc), C.byref(po)))
        oo = _np_from(po, n + 1, 8, np.uint64)
        cc = _np_from(pc, int(oo[-1]) if n else 0, 8, np.uint64)
        self._L.kmcpg_free(pc)
        self._L.kmcpg_free(po)
        return cc, oo

    def count_codes(self, codes: np.ndarray) -> np.ndarray:
        c = np.ascontiguousarray(codes, dtype=np.uint64)
        out = np.zeros(self.db_info().n_targets, dtype=np.uint32)
        self._check(self._L.kmcpg_count_codes(self._h, c.ctypes.data if c.size else None, c.size, out.ctypes.data))
        return out

    # ---- engine ----
    def default_engine_opts(self, **kw) -> EngineOpts:
        o = EngineOpts()
        self._L.kmcpg_default_engine_opts(C.byref(o))
        for k, v in kw.items():
            setattr(o, k, v)
        return o

    def engine_search(self, buf: np.ndarray, off: np.ndarray, opts: Optional[EngineOpts] = None, refcounts: Optional[int] = None,
                      shards: Sequence["Context"] = (), replicas: Sequence["Context"] = ()) -> EngineResults:
        o = opts or self.default_engine_opts()
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        return self.engine_search_ptr(buf.ctypes.data, off.ctypes.data, len(off) - 1, o, refcounts=refcounts, shards=shards, replicas=replicas)

    def engine_search_ptr(self, seq_ptr: int, off_ptr: int, n_seqs: int, o: EngineOpts, copy: bool = True, refcounts: Optional[int] = None,
                          shards: Sequence["Context"] = (), replicas: Sequence["Context"] = ()) -> EngineResults:
        r = Results()
        if shards or replicas:
            # this context + `shards` hold one database between them (kmcpg_engine_search_sharded), or this context and
            # `replicas` each hold all of it and the reads are split (kmcpg_engine_search_replicas)
            others = list(shards or replicas)
            hs = (C.c_void_p * (1 + len(others)))(self._h, *[c._h for c in others])
            fn = self._L.kmcpg_engine_search_sharded if shards else self._L.kmcpg_engine_search_replicas
            rc = fn(hs, 1 + len(others), C.byref(o), seq_ptr, off_ptr, n_seqs, C.byref(r))
            if rc:
                msgs = [self._L.kmcpg_last_error(c._h).decode() for c in (self, *others)]
                raise KmcpGpuError(rc, "; ".join(m for m in msgs if m))
        else:
            self._check(self._L.kmcpg_engine_search(self._h, C.byref(o), seq_ptr, off_ptr, n_seqs, C.byref(r)))
        if refcounts is not None:       # `kmcp profile` stage-1 counters of this batch (kmcpg_refcounts_add)
            self._check(self._L.kmcpg_refcounts_add(refcounts, C.byref(r)))
        nq = r.n_queries
        if not copy:
            out = EngineResults(np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.uint64), np.zeros(0, MATCH_DTYPE),
                                r.ms_gpu_total, int(r.probe_row_bytes), int(r.kernel_launches))
            out.n_matches = int(r.n_matches); out.ms_post = r.ms_post; out.ms_total = r.ms_total
            self._L.kmcpg_free_results(C.byref(r))
            return out
        out = EngineResults(_np_from(r.query_len, nq, 4, np.int32), _np_from(r.n_kmers, nq, 4, np.int32),
                            _np_from(r.k_used, nq, 4, np.int32), _np_from(r.match_off, nq + 1, 8, np.uint64),
                            _np_from(r.matches, r.n_matches, C.sizeof(Match), MATCH_DTYPE), r.ms_gpu_total,
                            int(r.probe_row_bytes), int(r.kernel_launches))
        out.n_matches = int(r.n_matches); out.ms_post = r.ms_post; out.ms_total = r.ms_total
        self._L.kmcpg_free_results(C.byref(r))
        return out

    # ---- `kmcp profile` stage-1 counters (kmcpg_refcounts_*) ----
    def refcounts_create(self, **kw) -> int:
        return refcounts_create(self._h, None, **kw)

    def refcounts_get(self, rc: int):
        return refcounts_get(rc)

    def refcounts_free(self, rc: int):
        self._L.kmcpg_refcounts_free(rc)

    # ---- memory helpers ----
    def device_memory(self) -> Tuple[int, int]:
        """(free, total) bytes of this context's device"""
        f, t = C.c_size_t(), C.c_size_t()
        self._check(self._L.kmcpg_device_memory(self._h, C.byref(f), C.byref(t)))
        return int(f.value), int(t.value)

    def device_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        self._check(self._L.kmcpg_device_alloc(self._h, C.byref(p), nbytes))
        return p.value

    def device_free(self, ptr: int):
        self._check(self._L.kmcpg_device_free(self._h, ptr))

    def h2d(self, dptr: int, arr: np.ndarray):
        a = np.ascontiguousarray(arr)
        self._check(self._L.kmcpg_memcpy_h2d(self._h, dptr, a.ctypes.data, a.nbytes))

    def d2h(self, dptr: int, nbytes: int) -> np.ndarray:
        out = np.empty(nbytes, dtype=np.uint8)
        self._check(self._L.kmcpg_memcpy_d2h(self._h, out.ctypes.data, dptr, nbytes))
        return out

    def synth_genomes(self, genome_seed: int, first: int, n_genomes: int, genome_len: int, dptr: int):
        self._check(self._L.kmcpg_synth_genomes(self._h, genome_seed, first, n_genomes, genome_len, dptr))

    def synth_reads(self, seed: int, first: int, n_reads: int, read_len: int, genome_seed: int, n_genomes: int,
                    genome_len: int, dptr: int):
        self._check(self._L.kmcpg_synth_reads(self._h, seed, first, n_reads, read_len, genome_seed, n_genomes, genome_len, dptr))


def shard_plan(r001_dir: str, world: int):
    """owner shard of every block of the DB (host only; the plan kmcpg_open_db follows)"""
    L = load()
    buf = (C.c_int32 * 65536)()
    n = L.kmcpg_shard_plan(r001_dir.encode(), world, buf, 65536)
    if n < 0:
        raise KmcpGpuError(n, L.kmcpg_last_error(None).decode())
    return [int(buf[i]) for i in range(n)]


def shard_pieces(r001_dir: str, world: int):
    """the plan as pieces [(block, shard, col0, n_cols, resident_bytes)] (host only): whole blocks, or column ranges when
    the DB has fewer blocks than shards"""
    L = load()
    buf = (ShardPiece * 65536)()
    n = L.kmcpg_shard_pieces(r001_dir.encode(), world, buf, 65536)
    if n < 0:
        raise KmcpGpuError(n, L.kmcpg_last_error(None).decode())
    return [(int(b.block), int(b.shard), int(b.col0), int(b.n_cols), int(b.resident_bytes)) for b in buf[:n]]


def merge_hit_lists(lists: Sequence[np.ndarray], first_query: int = 0, n_queries: int = 0, threads: int = 1) -> np.ndarray:
    """the k-way (query, target) merge the sharded engine applies to per-shard hit lists (host only test hook); with
    threads > 1 the query range [first_query, first_query + n_queries) is split over worker threads"""
    L = load()
    arrs = [np.ascontiguousarray(a, dtype=HIT_DTYPE) for a in lists]
    ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
    ns = (C.c_uint64 * len(arrs))(*[len(a) for a in arrs])
    out = np.zeros(sum(len(a) for a in arrs), dtype=HIT_DTYPE)
    L.kmcpg_internal_merge_hits(ptrs, ns, len(arrs), out.ctypes.data, first_query, n_queries, threads)
    return out


def host_alloc(nbytes: int) -> int:
    p = C.c_void_p()
    rc = load().kmcpg_host_alloc(C.byref(p), nbytes)
    if rc:
        raise KmcpGpuError(rc, "pinned allocation failed")
    return p.value


def host_free(ptr: int):
    load().kmcpg_host_free(ptr)


def pinned_array(nbytes: int, dtype=np.uint8) -> Tuple[np.ndarray, int]:
    """numpy view over pinned host memory (caller keeps ptr and frees it with host_free)"""
    ptr = host_alloc(nbytes)
    buf = (C.c_uint8 * nbytes).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype), ptr


def refcounts_create(ctx_handle, db_dir: Optional[str], **kw) -> int:
    """accumulator over the database of a context, or (ctx_handle None) over the block headers in db_dir (host only)"""
    L = load()
    p = RefcountParams()
    L.kmcpg_default_refcount_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    h = C.c_void_p()
    rc = L.kmcpg_refcounts_create(ctx_handle, db_dir.encode() if db_dir else None, C.byref(p), C.byref(h))
    if rc:
        raise KmcpGpuError(rc, (L.kmcpg_last_error(None) or b"").decode())
    return h.value


def refcounts_add_matches(rc: int, match_off: np.ndarray, matches: np.ndarray):
    """feed per-query match lists (MATCH_DTYPE, offsets n_queries+1) that were produced elsewhere"""
    L = load()
    r = Results()
    off = np.ascontiguousarray(match_off, dtype=np.uint64)
    m = np.ascontiguousarray(matches, dtype=MATCH_DTYPE)
    r.n_queries = len(off) - 1
    r.n_matches = len(m)
    r.match_off = C.cast(off.ctypes.data, C.POINTER(C.c_uint64))
    r.matches = C.cast(m.ctypes.data, C.POINTER(Match))
    rcode = L.kmcpg_refcounts_add(rc, C.byref(r))
    if rcode:
        raise KmcpGpuError(rcode, "kmcpg_refcounts_add")


def refcounts_get(rc: int):
    """(n_reads, {reference: (genome_size, match[], uniq[], uniq_hic[])})"""
    L = load()
    t = RefcountTable()
    rcode = L.kmcpg_refcounts_get(rc, C.byref(t))
    if rcode:
        raise KmcpGpuError(rcode, "kmcpg_refcounts_get")
    out = {}
    for i in range(t.n_refs):
        r = t.rows[i]
        n = r.n_chunks
        out[r.name.decode()] = (int(r.genome_size), [r.match[j] for j in range(n)], [r.uniq_match[j] for j in range(n)], [r.uniq_match_hic[j] for j in range(n)])
    return int(t.n_reads), out


def refcounts_free(rc: int):
    load().kmcpg_refcounts_free(rc)


def read_batches(files: Sequence[str] = (), read1: Optional[str] = None, read2: Optional[str] = None, **kw):
    """the reader stage (kmcpg_reader_*, host only): yields (first_query, ids [bytes], seq uint8 array, off uint64 array) per batch;
    kw: whole_file, use_filename, query_id, k, batch_reads, batch_bytes, inflate_threads, parse_threads, inflate_chunk, parse_piece"""
    L = load()
    o = ReaderOpts()
    L.kmcpg_default_reader_opts(C.byref(o))
    keep = [f.encode() for f in files]
    arr = (C.c_char_p * max(1, len(keep)))(*keep)
    if read1 or read2:
        o.read1 = read1.encode() if read1 else None
        o.read2 = read2.encode() if read2 else None
    else:
        o.files, o.n_files = arr, len(keep)
    for k_, v in kw.items():
        setattr(o, k_, v.encode() if isinstance(v, str) else v)
    h = C.c_void_p()
    rc = L.kmcpg_reader_open(C.byref(o), C.byref(h))
    if rc:
        raise KmcpGpuError(rc, "kmcpg_reader_open")
    try:
        while True:
            b = ReadBatch()
            rc = L.kmcpg_reader_next(h, C.byref(b))
            if rc == 0:
                return
            if rc < 0:
                raise KmcpGpuError(rc, L.kmcpg_reader_error(h).decode())
            nq, ns = b.n_queries, b.n_seqs
            off = _np_from(b.off, ns + 1, 8, np.uint64)
            id_off = _np_from(b.id_off, nq + 1, 8, np.uint64)
            seq = _np_from(b.seq, int(off[-1]), 1, np.uint8)
            raw = C.string_at(b.ids, int(id_off[-1])) if int(id_off[-1]) else b""
            ids = [raw[int(id_off[q]):int(id_off[q + 1])] for q in range(nq)]
            first = int(b.first_query)
            L.kmcpg_reader_free_batch(C.byref(b))
            yield first, ids, seq, off
    finally:
        L.kmcpg_reader_close(h)

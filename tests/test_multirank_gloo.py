"""N>1 host logic on CPU: world_size-2 gloo processes play the ranks of a one-process-per-GPU run.  Each rank holds the hits whose
targets live in the blocks the library's shard plan gives it (what its GPU context would return), hands them to rank 0 through the
shared-memory hit exchange (kmcp_b200.multigpu.HitExchange over kmcpg_shm_open, here without CUDA registration), and rank 0's merge +
host post-filter (kmcpg_merge_hits, kmcpg_engine_postfilter) must reproduce the single-process oracle result bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest

import parity_helpers as helpers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, r001, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ctypes as C
    import torch.distributed as dist
    from kmcp_b200 import api, multigpu
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    owner = api.shard_plan(r001, world)
    odb = O.DB(r001)
    n_steps, per_step = 5, 400
    o = O.default_opts()
    hx = multigpu.HitExchange("kmcp_gloo_test_%d" % port, rank, world, cap_hits=1 << 16, barrier=lambda: dist.barrier(), cuda_register=False)
    blk = np.array([odb.target(int(t)).block for t in range(odb.info.n_targets)])
    tsizes = np.array([odb.target(int(t)).n_kmers for t in range(odb.info.n_targets)], dtype=np.float64)
    ok, why = True, ""
    digest_parts = 0
    for s in range(n_steps):           # more steps than slots: the slot hand-back is exercised too
        reads = helpers.make_reads(O, 31 + s, per_step, 20, 20000, 9)
        oo = O.default_opts(); oo.max_fpr = 1.0; oo.min_target_cov = 0.0
        raw = odb.search(reads, opts=oo)          # every (query, target, count) above the query-coverage threshold = what the probe kernel emits
        mine = np.array([owner[b] == rank for b in blk[raw.hits["target"]]], dtype=bool) if len(raw.hits) else np.zeros(0, bool)
        local = np.zeros(int(mine.sum()), dtype=api.HIT_DTYPE)
        for f in ("query", "target", "count"):
            local[f] = raw.hits[f][mine]
        local = local[np.lexsort((local["target"], local["query"]))]
        # the rank's library would copy device→host into its slot; here the bytes are written directly
        hx.wait_free(s)
        dst = np.frombuffer((C.c_uint8 * (len(local) * 12)).from_address(hx.hits_ptr(s)), dtype=np.uint8)
        dst[:] = local.view(np.uint8)
        hx.publish(s, len(local))
        if rank == 0:
            lists = hx.collect(s)
            ok = ok and len(lists) == world and 0 < len(lists[0]) and sum(len(x) for x in lists) == len(raw.hits)
            merged = multigpu.merge_lists(lists, 0, per_step, threads=3)
            full = np.zeros(len(raw.hits), dtype=api.HIT_DTYPE)
            for f in ("query", "target", "count"):
                full[f] = raw.hits[f]
            full = full[np.lexsort((full["target"], full["query"]))]
            if not np.array_equal(merged, full):
                ok, why = False, "merge step %d" % s
            if multigpu.hits_digest(merged) != multigpu.hits_digest(full) or multigpu.hits_digest(merged) == multigpu.hits_digest(merged[::-1].copy()):
                ok, why = False, "digest step %d" % s
            digest_parts += multigpu.hits_digest(merged)
            # host post-filter over the merged lists == the oracle's full engine answer (defaults: FPR <= 0.01 filter, sort by qcov)
            exp = odb.search(reads, opts=o)
            eo = api.EngineOpts()
            api.load().kmcpg_default_engine_opts(C.byref(eo))
            got = multigpu.postfilter(eo, raw.n_kmers, raw.query_len, merged, tsizes, odb.info.fpr, odb.info.ks[0])
            same = np.array_equal(got.match_off, exp.hit_off) and all(np.array_equal(got.matches[f], exp.hits[f]) for f in ("query", "target", "count", "fpr", "qcov", "tcov", "jacc"))
            if not same or len(exp.hits) < 100:
                ok, why = False, "postfilter step %d" % s
            hx.release(s)
    dist.barrier()
    hx.close()
    if rank == 0:
        open(out_path, "w").write("ok" if ok and digest_parts else "mismatch: " + why)
    dist.destroy_process_group()


def test_two_rank_hit_exchange_equals_single_process(oracle, tmp_path):
    import torch.multiprocessing as mp
    O = oracle
    sp = O.sketch_params(21)
    targets = helpers.make_synth_targets(O, sp, 9, 20, 20000, 4, 100)
    r001 = O.build_db(targets, str(tmp_path / "db"), sp, num_hashes=1, fpr=0.3, block_size=16)   # 80 targets → 5 blocks
    from kmcp_b200 import api
    for world in (1, 2, 3):
        owner = api.shard_plan(r001, world)
        assert len(owner) == 5 and set(owner) == set(range(world))
    out = str(tmp_path / "result.txt")
    mp.spawn(_worker, args=(2, _free_port(), r001, out), nprocs=2, join=True)
    assert open(out).read() == "ok"


def test_shared_segment_roundtrip_and_errors(tmp_path):
    """kmcpg_shm_open / kmcpg_shm_close without CUDA registration: two mappings of one name see each other's bytes"""
    import ctypes as C
    from kmcp_b200 import api
    L = api.load()
    name = ("kmcp_shm_test_%d" % os.getpid()).encode()
    a, b = C.c_void_p(), C.c_void_p()
    assert L.kmcpg_shm_open(name, 4096, 1, 0, C.byref(a)) == 0
    assert L.kmcpg_shm_open(name, 4096, 0, 0, C.byref(b)) == 0 and a.value != b.value
    va = np.frombuffer((C.c_uint8 * 4096).from_address(a.value), dtype=np.uint8)
    vb = np.frombuffer((C.c_uint8 * 4096).from_address(b.value), dtype=np.uint8)
    va[:] = np.arange(4096, dtype=np.uint64).astype(np.uint8)
    assert np.array_equal(va, vb)
    c = C.c_void_p()
    assert L.kmcpg_shm_open(name, 8192, 0, 0, C.byref(c)) == api.KMCPG_EIO          # smaller than asked for
    assert L.kmcpg_shm_open(b"kmcp_shm_test_missing", 4096, 0, 0, C.byref(c)) == api.KMCPG_EIO
    del va, vb
    assert L.kmcpg_shm_close(name, b, 4096, 0, 0) == 0 and L.kmcpg_shm_close(name, a, 4096, 0, 1) == 0
    assert L.kmcpg_shm_open(name, 4096, 0, 0, C.byref(c)) == api.KMCPG_EIO          # unlinked

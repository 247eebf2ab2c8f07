"""N>1 host logic on CPU: world_size-2 gloo processes shard the hit space by block owner (the plan the library
uses), gather on rank 0, and the union must equal the single-process result."""
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
    import torch.distributed as dist
    from kmcp_b200 import api, multigpu
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    owner = api.shard_plan(r001, world)
    odb = O.DB(r001)
    reads = helpers.make_reads(O, 31, 1200, 20, 20000, 9)
    o = O.default_opts(); o.max_fpr = 1.0
    res = odb.search(reads, opts=o)
    # this rank's shard = hits whose target lives in a block it owns (what its GPU context would return)
    blk = np.array([odb.target(int(t)).block for t in range(odb.info.n_targets)])
    mine = np.array([owner[b] == rank for b in blk[res.hits["target"]]], dtype=bool) if len(res.hits) else np.zeros(0, bool)
    dt = np.dtype([("query", "<u4"), ("target", "<u4"), ("count", "<u4")])
    local = np.zeros(int(mine.sum()), dtype=dt)
    for f in ("query", "target", "count"):
        local[f] = res.hits[f][mine]
    merged = multigpu.gather_hits(local, rank, world)
    padded = multigpu.gather_hits_padded(local, rank, world, "cpu")           # the NCCL-shaped gather, here on gloo
    shifted = multigpu.gather_hits_padded(local, rank, world, "cpu", target_base=1000 * rank)
    if rank == 0:
        full = np.zeros(len(res.hits), dtype=dt)
        for f in ("query", "target", "count"):
            full[f] = res.hits[f]
        full = full[np.lexsort((full["target"], full["query"]))]
        ok = np.array_equal(merged, full) and len(full) > 500 and 0 < len(local) < len(full)
        ok = ok and np.array_equal(padded[np.lexsort((padded["target"], padded["query"]))], full)
        ok = ok and np.array_equal(shifted[:len(local)], local) and int(shifted["target"].astype(np.int64).sum() - padded["target"].astype(np.int64).sum()) == 1000 * (len(full) - len(local))
        open(out_path, "w").write("ok" if ok else "mismatch %d %d %d" % (len(merged), len(full), len(local)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_equals_single_process(oracle, tmp_path):
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

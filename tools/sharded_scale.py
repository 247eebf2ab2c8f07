"""One process, several GPUs (dev tool): the C2 index (one block of 10,000 targets) sharded by column range over the
devices in GPUS — and, MODES=replicas, loaded whole on every device with the reads split instead — (default "0,1"; "0,0" puts both shards on one GPU for a functional run), timed through
kmcpg_engine_search_sharded from pinned host reads, next to the one-context engine on device GPUS[0].
Strong scaling: the same index and the same reads at every device count.  Prints one JSON line.

  GPUS=0,1 NR=1000000 python tools/sharded_scale.py
  GPUS=0,1,2,3,4,5,6,7 WORLDS=2,4,8 python tools/sharded_scale.py
"""
import json
import os
import shutil
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kmcp_b200 import api

GPUS = [int(x) for x in os.environ.get("GPUS", "0,1").split(",")]
WORLDS = [int(x) for x in os.environ.get("WORLDS", str(len(GPUS))).split(",")]       # shard counts to time, e.g. 2,4,8 (devices GPUS[:world])
MODES = os.environ.get("MODES", "shard,replicas").split(",")   # shard: index cut over the devices; replicas: whole index everywhere, reads split
NG, GL = int(os.environ.get("NG", 1000)), int(os.environ.get("GL", 4_000_000))
NR, REPS = int(os.environ.get("NR", 1_000_000)), int(os.environ.get("REPS", 3))
RL, K, NCH, OV = 150, 21, 10, 150
BS = NG * NCH

YML = """version: 4
unikiVersion: 4
alias: sharded
k: {k}
ks:
- {k}
hashed: true
canonical: true
scaled: false
scale: 0
minimizer: false
minimizer-w: 0
syncmer: false
syncmer-s: 0
split-seq: true
split-size: 0
split-num: {nch}
split-overlap: {ov}
compact-size: false
hashes: 1
fpr: 0.3
numNameGroups: {nt}
blocksize: {bs}
totalKmers: 0
files:
{files}"""


def timed(fn, reps):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        r = fn()
        ts.append(time.perf_counter() - t0)
    return min(ts), r


def main():
    tmp = "/dev/shm/kmcp_sharded" if os.path.isdir("/dev/shm") else "/tmp/kmcp_sharded"
    shutil.rmtree(tmp, ignore_errors=True)
    r001 = os.path.join(tmp, "R001")
    os.makedirs(r001)
    out = {"gpus": GPUS, "reads": NR}
    whole = api.Context(GPUS[0])
    try:
        t0 = time.perf_counter()
        whole.build_synth_db(1, NG, GL, k=K, n_chunks=NCH, overlap=OV, num_hashes=1, fpr=0.3, block_size=BS)
        info = whole.db_info()
        files = []
        for b in range(info.n_resident_blocks):
            fn = "_block%03d.uniki" % (b + 1)
            whole.write_block(b, os.path.join(r001, fn))
            files.append(fn)
        open(os.path.join(r001, "__db.yml"), "w").write(YML.format(k=K, nch=NCH, ov=OV, nt=int(info.n_targets), bs=BS,
                                                                   files="".join("- %s\n" % f for f in files)))
        out["db"] = {"targets": int(info.n_targets), "blocks": info.n_blocks, "index_GB": round(info.resident_bytes / 1e9, 2),
                     "build_and_dump_s": round(time.perf_counter() - t0, 1)}
        d = whole.device_alloc(NR * RL)
        whole.synth_reads(2, 0, NR, RL, 1, NG, GL, d)
        host = whole.d2h(d, NR * RL)
        whole.device_free(d)
        pin, pin_ptr = api.pinned_array(NR * RL)
        pin[:] = host
        off, off_ptr = api.pinned_array((NR + 1) * 8)
        off.view(np.uint64)[:] = np.arange(NR + 1, dtype=np.uint64) * np.uint64(RL)
        eo = whole.default_engine_opts()
        t1, r1 = timed(lambda: whole.engine_search_ptr(pin_ptr, off_ptr, NR, eo, copy=False), REPS)
        out["one_context"] = {"ms": round(t1 * 1e3, 2), "reads_per_s": round(NR / t1), "matches": r1.n_matches,
                              "probe_row_bytes_per_read": round(r1.probe_row_bytes / NR)}
        n_chk = min(NR, 50_000)
        a = whole.engine_search_ptr(pin_ptr, off_ptr, n_chk, eo)
        out["sharded"] = []
        for mode, world in [(m, w) for m in MODES for w in WORLDS]:
            devs = GPUS[:world] if len(GPUS) >= world else (GPUS * world)[:world]
            t0 = time.perf_counter()
            shards = [None] * world

            def load(rank):
                c = api.Context(devs[rank])
                if mode == "replicas":
                    c.open_db(r001)
                else:
                    c.open_db(r001, shard_rank=rank, shard_world=world)
                shards[rank] = c

            th = [threading.Thread(target=load, args=(r,)) for r in range(world)]       # ctypes releases the GIL: shards load side by side
            for t in th:
                t.start()
            for t in th:
                t.join()
            load_s = time.perf_counter() - t0
            live = [c for c in shards if c is not None and c.db_info().n_resident_blocks > 0]
            kw = {"replicas": live[1:]} if mode == "replicas" else {"shards": live[1:]}
            tn, rn = timed(lambda: live[0].engine_search_ptr(pin_ptr, off_ptr, NR, eo, copy=False, **kw), REPS)
            b = live[0].engine_search_ptr(pin_ptr, off_ptr, n_chk, eo, **kw)
            same = bool(np.array_equal(a.match_off, b.match_off) and np.array_equal(a.matches, b.matches) and np.array_equal(a.n_kmers, b.n_kmers))
            out["sharded"].append({"mode": mode, "world": world, "devices": devs, "row_bytes": [int(c.db_info().sum_row_bytes) for c in live],
                                   "load_s": round(load_s, 2), "ms": round(tn * 1e3, 2), "reads_per_s": round(NR / tn), "matches": rn.n_matches,
                                   "probe_row_bytes_per_read": round(rn.probe_row_bytes / NR), "gpu_ms_max_shard": round(rn.ms_gpu_total, 2),
                                   "post_ms": round(rn.ms_post, 2), "speedup_vs_one_context": round(t1 / tn, 3),
                                   "identical_on_first_%d" % n_chk: same, "matches_equal": r1.n_matches == rn.n_matches})
            for c in shards:
                if c is not None:
                    c.close()
            print(json.dumps(out["sharded"][-1]), file=sys.stderr, flush=True)       # progress: survives a cut-off call
        api.host_free(pin_ptr); api.host_free(off_ptr)
    finally:
        whole.close()
        shutil.rmtree(tmp, ignore_errors=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()

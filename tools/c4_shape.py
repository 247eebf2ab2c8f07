"""BASELINE configs[3] shape on ONE GPU (dev tool): 85,205 genomes x 10 chunks = 852,050 targets, k=21, h=3, fpr 0.3,
blocks of 26,632 targets (3,329-byte rows, 32 blocks) — the GTDB-scale row width and block count — with the genome
length scaled down (GL, default 100 kb) so the index is a few GB instead of ~100 GB; SURVEY §8(d) C4 says to scale
genome length, not target count.  Reports reads/s, algorithmic GB/s of the probe launches, and checks a sample of the
reads against the CPU oracle on the same (dumped) index."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kmcp_b200 import api

NG, GL = int(os.environ.get("NG", 85205)), int(os.environ.get("GL", 100000))
NR, BS, NCHK = int(os.environ.get("NR", 100000)), int(os.environ.get("BS", 26632)), int(os.environ.get("NCHK", 300))
RL = 150
with api.Context(0) as ctx:
    t0 = time.perf_counter()
    ctx.build_synth_db(3, NG, GL, k=21, n_chunks=10, overlap=150, num_hashes=3, fpr=0.3, block_size=BS)
    tb = time.perf_counter() - t0
    info = ctx.db_info()
    d = ctx.device_alloc(NR * RL)
    ctx.synth_reads(4, 0, NR, RL, 3, NG, GL, d)
    off = np.arange(NR + 1, dtype=np.uint64) * np.uint64(RL)
    doff = ctx.device_alloc(off.nbytes); ctx.h2d(doff, off)
    p = ctx.default_params()
    res = []
    for rep in range(4):
        o = ctx.search_batch_ptr(d, doff, NR, p, device=True, seq_bytes=NR * RL)
        if rep: res.append(o)
    ms = float(np.mean([x.ms_total for x in res])); pm = float(np.mean([x.ms_probe for x in res]))
    o = res[-1]
    out = {"db": {"targets": int(info.n_targets), "blocks": info.n_blocks, "index_GB": round(info.resident_bytes / 1e9, 2), "build_s": round(tb, 1),
                  "genome_len": GL, "num_hashes": 3}, "reads": NR, "hits": len(o.hits), "call_ms": round(ms, 1), "probe_ms": round(pm, 1),
           "hash_ms": round(float(np.mean([x.ms_hash for x in res])), 2), "reads_per_s": round(NR / (ms / 1e3)),
           "probe_row_bytes_per_read": round(o.probe_row_bytes / NR), "probe_GBps": round(o.probe_row_bytes / (pm / 1e3) / 1e9, 1),
           "probe_launches": int(o.probe_launches)}
    # end to end through the engine from pinned host memory
    host = ctx.d2h(d, NR * RL)
    pin, pin_ptr = api.pinned_array(NR * RL)
    pin[:] = np.frombuffer(host, np.uint8)
    eo = ctx.default_engine_opts()
    t = []
    for rep in range(3):
        t0 = time.perf_counter(); r = ctx.engine_search_ptr(pin_ptr, off.ctypes.data, NR, eo, copy=False); t.append(time.perf_counter() - t0)
    out["e2e_reads_per_s"] = round(NR / min(t[1:])); out["e2e_matches"] = r.n_matches
    # parity on a sample against the CPU oracle over the same index
    if NCHK:
        sys.path.insert(0, ROOT)
        import bench
        from oracle import oracle as O
        tmp = "/dev/shm/kmcp_c4"
        r001 = os.path.join(tmp, "R001"); os.makedirs(r001, exist_ok=True)
        files = []
        for b in range(info.n_resident_blocks):
            fn = "_block%03d.uniki" % (b + 1); ctx.write_block(b, os.path.join(r001, fn)); files.append(fn)
        O.write_db_yml(os.path.join(r001, "__db.yml"), {"version": 4, "unikiVersion": 4, "alias": "c4", "k": 21, "ks": [21], "hashed": True, "canonical": True,
            "scaled": False, "scale": 0, "minimizer": False, "minimizer-w": 0, "syncmer": False, "syncmer-s": 0, "split-seq": True, "split-size": 0, "split-num": 10,
            "split-overlap": 150, "compact-size": False, "hashes": 3, "fpr": 0.3, "numNameGroups": int(info.n_targets), "blocksize": BS, "totalKmers": 0, "files": files})
        odb = O.DB(r001)
        sample = [bytes(host[i * RL:(i + 1) * RL]) for i in range(NCHK)]
        t0 = time.perf_counter(); ores = odb.search(sample, algo=1, threads=os.cpu_count()); tc = time.perf_counter() - t0
        sb, so = api.pack_seqs(sample)
        er = ctx.engine_search(sb, so)
        oh, gm = ores.hits, er.matches
        same = (np.array_equal(er.match_off, ores.hit_off) and np.array_equal(gm["target"], oh["target"]) and np.array_equal(gm["count"], oh["count"])
                and np.array_equal(gm["fpr"], oh["fpr"]) and np.array_equal(gm["qcov"], oh["qcov"]) and np.array_equal(gm["tcov"], oh["tcov"]))
        out["oracle_sample"] = {"reads": NCHK, "hits": len(oh), "identical": bool(same), "cpu_reads_per_s": round(NCHK / tc, 1), "cores": os.cpu_count()}
        import shutil; shutil.rmtree(tmp)
    print(json.dumps(out))

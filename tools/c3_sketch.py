"""BASELINE configs[2] shape (dev tool): FracMinHash (scale 1000, k=21, h=3) genome-vs-genome search, 1,000 seeded 4 Mb
assemblies in the DB, whole genomes as queries (device resident).  Reports genomes/s and bases/s of the whole path."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kmcp_b200 import api

NG, GL, NQ = int(os.environ.get("NG", 1000)), int(os.environ.get("GL", 4000000)), int(os.environ.get("NQ", 250))
with api.Context(0) as ctx:
    t0 = time.perf_counter()
    ctx.build_synth_db(1, NG, GL, k=21, n_chunks=1, overlap=0, num_hashes=3, fpr=0.001, block_size=0, scale=1000)
    tb = time.perf_counter() - t0
    info = ctx.db_info()
    d = ctx.device_alloc(NQ * GL)
    ctx.synth_genomes(1, 0, NQ, GL, d)
    off = np.arange(NQ + 1, dtype=np.uint64) * np.uint64(GL)
    doff = ctx.device_alloc(off.nbytes); ctx.h2d(doff, off)
    p = ctx.default_params(min_query_cov=0.5)
    res = []
    for rep in range(4):
        o = ctx.search_batch_ptr(d, doff, NQ, p, device=True, seq_bytes=NQ * GL)
        if rep: res.append(o)
    ms = float(np.mean([o.ms_total for o in res]))
    o = res[-1]
    self_hits = int(np.sum(o.hits["query"] == o.hits["target"]))      # n_chunks=1: target g of the sorted DB is not genome g in general
    print(json.dumps({"db": {"targets": int(info.n_targets), "blocks": info.n_blocks, "index_MB": round(info.resident_bytes / 1e6, 1), "build_s": round(tb, 2)},
                      "queries": NQ, "kmers_per_query": int(np.mean(o.n_kmers)), "hits": len(o.hits), "call_ms": round(ms, 2),
                      "hash_ms": round(float(np.mean([x.ms_hash for x in res])), 2), "probe_ms": round(float(np.mean([x.ms_probe for x in res])), 3),
                      "genomes_per_s": round(NQ / (ms / 1e3), 1), "Gbases_per_s": round(NQ * GL / (ms / 1e3) / 1e9, 2)}))

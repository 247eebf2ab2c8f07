"""One measurement of the probe kernel on the C2 workload with the current KMCPG_PROBE_* knobs (dev tool)."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kmcp_b200 import api

NG, GL, NR, L = int(os.environ.get("NG", 1000)), int(os.environ.get("GL", 4000000)), int(os.environ.get("NR", 1000000)), int(os.environ.get("RL", 150))
BS = int(os.environ.get("BS", 0)) or NG * 10
FPR = float(os.environ.get("FPR", 0.3))
H = int(os.environ.get("H", 1))
with api.Context(0) as ctx:
    ctx.build_synth_db(1, NG, GL, k=21, n_chunks=10, overlap=150, num_hashes=H, fpr=FPR, block_size=BS)
    info = ctx.db_info()
    reps = int(os.environ.get("REPS", 5))
    d = ctx.device_alloc((reps + 2) * NR * L)
    for s in range(reps + 2):
        ctx.synth_reads(2, s * NR, NR, L, 1, NG, GL, d + s * NR * L)
    off = np.arange(NR + 1, dtype=np.uint64) * np.uint64(L)
    doff = ctx.device_alloc(off.nbytes); ctx.h2d(doff, off)
    p = ctx.default_params()
    res = []
    for s in range(reps + 2):
        o = ctx.search_batch_ptr(d + s * NR * L, doff, NR, p, device=True, seq_bytes=NR * L)
        if s >= 2:
            res.append((o.ms_probe, o.probe_row_bytes, o.ms_hash, o.ms_locs, o.ms_total, len(o.hits)))
    ms = np.mean([r[0] for r in res]); gb = np.mean([r[1] for r in res]) / 1e9
    print(json.dumps({"cfg": {k: v for k, v in os.environ.items() if k.startswith("KMCPG_") or k in ("NG", "GL", "NR", "RL", "BS", "H", "FPR")},
                      "blocks": info.n_blocks, "sum_row_bytes": info.sum_row_bytes, "index_GB": round(info.resident_bytes / 1e9, 2),
                      "reads_per_s": round(NR / (float(np.mean([r[4] for r in res])) / 1e3)), "probe_ms": round(float(ms), 3),
                      "GBps": round(gb / (ms / 1e3), 1), "hash_ms": round(float(np.mean([r[2] for r in res])), 3),
                      "locs_ms": round(float(np.mean([r[3] for r in res])), 3), "call_ms": round(float(np.mean([r[4] for r in res])), 2),
                      "hits": res[0][5]}))

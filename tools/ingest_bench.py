#!/usr/bin/env python
"""Host-side ingest / egress stages of `kmcp-gpu search`, timed alone (no GPU needed): inflate of one .gz stream (zlib through
Python as the yardstick, fastgz.h, pargz.h on 2/4/8 workers), FASTQ parse into packed batches, the .gz result writer.
Writes one JSON object to stdout.  usage: python tools/ingest_bench.py [n_reads]"""
import json
import os
import re
import subprocess
import sys
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "kmcp_b200", "kmcp-gpu")


def best(fn, reps=3):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts)


def run(args, **kw):
    p = subprocess.run([EXE] + args, capture_output=True, **kw)
    assert p.returncode == 0, p.stderr.decode()
    return p


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    tmp = "/dev/shm/kmcp_ingest_bench" if os.path.isdir("/dev/shm") else "/tmp/kmcp_ingest_bench"
    os.makedirs(tmp, exist_ok=True)
    rng = np.random.default_rng(1)
    L = 150
    bases = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, (n, L))]
    qual = np.frombuffer(b"FFFFFFFFFFFF::,#", dtype=np.uint8)[rng.integers(0, 16, (n, L))]
    fq = os.path.join(tmp, "r.fq")
    with open(fq, "wb") as f:
        for i in range(n):
            f.write(b"@A00123:45:HXXXXXXX:1:1101:%d:%d 1:N:0:ACGTACGT\n" % (1000 + i % 30000, 1000 + i // 30))
            f.write(bases[i].tobytes()); f.write(b"\n+\n"); f.write(qual[i].tobytes()); f.write(b"\n")
    text_bytes = os.path.getsize(fq)
    subprocess.run("gzip -6 -k -f %s" % fq, shell=True, check=True)
    gz = fq + ".gz"
    z = open(gz, "rb").read()
    out = {"reads": n, "text_bytes": text_bytes, "gz_bytes": len(z), "cpus": os.cpu_count(),
           "cpu": [l.split(":")[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][:1]}

    def rate(seconds):
        return {"s": round(seconds, 3), "GB_per_s": round(text_bytes / seconds / 1e9, 3), "M_reads_per_s": round(n / seconds / 1e6, 2)}

    out["inflate"] = {"zlib (python zlib.decompress)": rate(best(lambda: zlib.decompress(z, 31)))}
    out["inflate"]["fastgz.h (kmcp-gpu gunzip)"] = rate(best(lambda: run(["gunzip", gz], stdout=None) if False else subprocess.run([EXE, "gunzip", gz], stdout=subprocess.DEVNULL, check=True)))
    for t in (2, 4, 8):
        out["inflate"]["pargz.h %d workers" % t] = rate(best(lambda: subprocess.run([EXE, "gunzip", "--threads", str(t), gz], stdout=subprocess.DEVNULL, check=True)))
    st = run(["gunzip", "--threads", "4", "--stats", gz], ).stderr.decode()
    out["pargz_chunks"] = st.strip()

    def parse(args):
        s = run(["parse"] + args).stderr.decode()
        m = re.search(r"([\d.]+) s, ([\d.]+) M (?:records|queries)/s", s)
        return {"s": float(m.group(1)), "M_reads_per_s": float(m.group(2))}

    out["parse"] = {
        "plain FASTQ, reader alone": min((parse(["--count", fq]) for _ in range(3)), key=lambda d: d["s"]),
        "plain FASTQ, batch builder (parser thread)": min((parse(["--batches", "--count", fq]) for _ in range(3)), key=lambda d: d["s"]),
        "plain FASTQ, batch builder, 4 parse threads": min((parse(["--batches", "--count", "--parse-threads", "4", fq]) for _ in range(3)), key=lambda d: d["s"]),
        "gz, sequential inflate thread + parser thread": min((parse(["--batches", "--count", "--inflate-threads", "1", gz]) for _ in range(3)), key=lambda d: d["s"]),
        "gz, 8 inflate workers + parser thread": min((parse(["--batches", "--count", "--inflate-threads", "8", gz]) for _ in range(3)), key=lambda d: d["s"]),
        "paired gz (the same file twice), 4 inflate workers each": min((parse(["--batches", "--count", "--inflate-threads", "4", "-1", gz, "-2", gz]) for _ in range(3)), key=lambda d: d["s"]),
    }
    # the writer: TSV rows like the search output
    tsv = os.path.join(tmp, "t.tsv")
    with open(tsv, "w") as f:
        r = np.random.default_rng(2)
        q = r.integers(90, 131, 2_000_000)
        for i in range(2_000_000):
            f.write("r%08d\t150\t130\t%.4e\t1\tgenome_%04d\t%d\t10\t4000000\t21\t%d\t%.4f\t%.4f\t%.4f\t%d\n" %
                    (i, 1e-12 * (i % 97), i % 1000, i % 10, q[i], q[i] / 130, q[i] / 400000, q[i] / 400100, i))
    tb = os.path.getsize(tsv)
    out["writer"] = {}
    for name, dst in (("tsv.gz (level 4, asynchronous members)", "o.tsv.gz"), ("tsv (plain)", "o.tsv")):
        ts = []
        for _ in range(3):
            s = run(["gzip-write", tsv, os.path.join(tmp, dst)]).stderr.decode()
            ts.append(float(re.search(r"in ([\d.]+) s", s).group(1)))
        out["writer"][name] = {"s": min(ts), "MB_per_s": round(tb / min(ts) / 1e6, 1)}
    t0 = time.perf_counter()
    zlib.compress(open(tsv, "rb").read(50_000_000), 6)
    out["writer"]["zlib level 6, one thread (python)"] = {"MB_per_s": round(50.0 / (time.perf_counter() - t0), 1)}
    # the whole command without a device (--dry-run: every query unmatched): reader stage + engine thread + writer, -K prints a row per query
    out["search --dry-run"] = {}
    for name, args in (("fastq, no rows", [fq, "-o", os.path.join(tmp, "d.tsv")]), ("fastq -> tsv (-K)", ["-K", fq, "-o", os.path.join(tmp, "d.tsv")]),
                       ("fastq.gz -> tsv.gz (-K)", ["-K", gz, "-o", os.path.join(tmp, "d.tsv.gz")]),
                       ("paired fastq.gz -> tsv.gz (-K)", ["-K", "-1", gz, "-2", gz, "-o", os.path.join(tmp, "d.tsv.gz")])):
        best_v = 0.0
        for _ in range(3):
            s = run(["search", "--dry-run"] + args).stderr.decode()
            best_v = max(best_v, float(re.search(r"speed: ([\d.]+) million queries per minute", s).group(1)))
        out["search --dry-run"][name] = {"M_queries_per_s": round(best_v / 60, 2)}
    print(json.dumps(out, indent=1))
    subprocess.run(["rm", "-rf", tmp])


if __name__ == "__main__":
    main()

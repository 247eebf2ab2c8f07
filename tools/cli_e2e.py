#!/usr/bin/env python
"""kmcp-gpu search end to end from FASTQ files to TSV: stages the C2 synthetic index and reads on
/dev/shm, runs the CLI and prints what its own log reports (load time, queries per minute, wall).  GPU box only."""
import json, os, re, subprocess, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from kmcp_b200 import api

def main():
    n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
    tmp = "/dev/shm/kmcp_cli_e2e"
    os.makedirs(tmp, exist_ok=True)
    ctx = api.Context(0)
    ctx.build_synth_db(bench.GENOME_SEED, bench.N_GENOMES, bench.GENOME_LEN, k=bench.K, n_chunks=bench.N_CHUNKS, overlap=bench.OVERLAP,
                       num_hashes=bench.H, fpr=bench.FPR, block_size=bench.BLOCK_SIZE)
    bench.dump_db_for_cpu(ctx, tmp)
    L = bench.READ_LEN
    d = ctx.device_alloc(n_reads * L)
    ctx.synth_reads(bench.READ_SEED, 0, n_reads, L, bench.GENOME_SEED, bench.N_GENOMES, bench.GENOME_LEN, d)
    reads = np.frombuffer(ctx.d2h(d, n_reads * L), dtype=np.uint8).reshape(n_reads, L)
    ctx.device_free(d); ctx.close()
    rec = np.empty((n_reads, 11 + L + 3 + L + 1), dtype=np.uint8)
    ids = np.char.zfill(np.arange(n_reads).astype("U8"), 8).astype("S8")
    rec[:, 0:2] = np.frombuffer(b"@r", np.uint8); rec[:, 2:10] = np.frombuffer(ids.tobytes(), np.uint8).reshape(n_reads, 8); rec[:, 10] = 10
    rec[:, 11:11 + L] = reads; rec[:, 11 + L:14 + L] = np.frombuffer(b"\n+\n", np.uint8); rec[:, 14 + L:14 + 2 * L] = ord("I"); rec[:, -1] = 10
    fq = os.path.join(tmp, "reads.fq")
    rec.tofile(fq)
    n_gz = min(n_reads, 1_000_000)
    fqgz = os.path.join(tmp, "reads_1m.fq.gz")
    rec[:n_gz].tofile(os.path.join(tmp, "reads_1m.fq"))
    subprocess.run("gzip -6 -c %s/reads_1m.fq > %s" % (tmp, fqgz), shell=True, check=True)
    # the reader stages alone (no GPU involved): inflate and parse rates of this box
    stages = {}
    for nm, args in (("parse plain", ["parse", "--count", fq]), ("parse gz sequential", ["parse", "--count", "--ahead", "--inflate-threads", "1", fqgz]),
                     ("parse gz 8 threads", ["parse", "--count", "--ahead", "--inflate-threads", "8", fqgz]),
                     ("batches plain", ["parse", "--batches", "--count", fq]), ("batches plain 4 parse threads", ["parse", "--batches", "--count", "--parse-threads", "4", fq]),
                     ("batches plain 8 parse threads", ["parse", "--batches", "--count", "--parse-threads", "8", fq])):
        stages[nm] = subprocess.run([os.path.join(ROOT, "kmcp_b200", "kmcp-gpu")] + args, capture_output=True).stderr.decode().strip()
    exe = os.path.join(ROOT, "kmcp_b200", "kmcp-gpu")
    out = []
    for name, inp, n, outp, extra in [("fastq -> tsv", fq, n_reads, "o.tsv", []), ("fastq -> tsv.gz", fq, n_reads, "o.tsv.gz", []),
                                      ("fastq -> tsv.gz level 6", fq, n_reads, "o6.tsv.gz", ["--compression-level", "6"]),
                                      ("fastq.gz -> tsv.gz (sequential inflate)", fqgz, n_gz, "o1.tsv.gz", ["--inflate-threads", "1"]),
                                      ("fastq.gz -> tsv.gz (4 inflate threads)", fqgz, n_gz, "o4.tsv.gz", ["--inflate-threads", "4"]),
                                      ("fastq.gz -> tsv.gz (8 inflate threads)", fqgz, n_gz, "o8.tsv.gz", ["--inflate-threads", "8"]),
                                      ("fastq.gz -> tsv.gz (default)", fqgz, n_gz, "od.tsv.gz", []),
                                      ("fastq -> tsv, 4 parse threads", fq, n_reads, "op.tsv", ["--parse-threads", "4"]),
                                      ("fastq.gz -> tsv.gz, 8 inflate + 4 parse threads", fqgz, n_gz, "op.tsv.gz", ["--inflate-threads", "8", "--parse-threads", "4"]),
                                      ("paired fastq.gz (the same file as both mates) -> tsv.gz", None, n_gz, "pe.tsv.gz", ["-1", fqgz, "-2", fqgz]),
                                      ("paired fastq -> tsv", None, n_reads, "pe.tsv", ["-1", fq, "-2", fq])]:
        t0 = time.time()
        p = subprocess.run([exe, "search", "-d", tmp] + ([inp] if inp else []) + ["-o", os.path.join(tmp, outp)] + extra, capture_output=True)
        wall = time.time() - t0
        log = p.stderr.decode()
        assert p.returncode == 0, log
        load = re.search(r"in HBM, ([\d.]+) s\)", log)
        speed = re.search(r"speed: ([\d.]+) million queries per minute", log)
        out.append({"case": name, "reads": n, "wall_s": round(wall, 2), "db_load_s": float(load.group(1)) if load else None,
                    "search_Mq_per_min": float(speed.group(1)) if speed else None,
                    "search_reads_per_s": round(float(speed.group(1)) * 1e6 / 60) if speed else None,
                    "out_bytes": os.path.getsize(os.path.join(tmp, outp))})
    a = subprocess.run("zcat %s/o.tsv.gz | md5sum; md5sum < %s/o.tsv" % (tmp, tmp), shell=True, capture_output=True).stdout.decode().split()
    print(json.dumps({"cases": out, "reader_stages": stages, "gz_equals_plain": a[0] == a[2]}, indent=1))
    subprocess.run(["rm", "-rf", tmp])

if __name__ == "__main__":
    main()

"""Builds tests/golden/demo/ from the reference's own demo data (run in the build container, where /root/reference exists; the GPU
box has no /root/reference, so the files it needs travel as fixtures):

  refs/*.fa.gz            the 15 reference genomes of /root/reference/demo-profiling/refs, byte for byte (input data, not code)
  mock_1.20k.fastq.gz     the first 20,000 reads of demo-profiling/mock_1.fastq.gz
  mock_2.20k.fastq.gz     the first 20,000 reads of demo-profiling/mock_2.fastq.gz
  expected.20k.tsv.gz     what `kmcp search -d refs-k21-n10.kmcp mock_1.20k.fastq.gz mock_2.20k.fastq.gz` prints for them, produced by
                          the CPU oracle — which, on the FULL read set, reproduces the reference's golden numbers G1 (349,084 queries,
                          308,839 matched, demo-profiling/mock.kmcp.gz.log:22-23) and G2 (docs/tutorial/profiling/index.md:203-211);
                          tests/test_oracle_golden.py checks that in this container.  The first nine matched rows of the file ARE the
                          reference's published rows (G2).
  summary.json            counts of the subset and of the full set as the oracle sees them here
"""
import glob
import gzip
import json
import os
import re
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

REF = "/root/reference/demo-profiling"
OUT = os.path.join(HERE, "demo")
N_SUB = 20000


def head_fastq(src, dst, n):
    with gzip.open(src, "rb") as f, gzip.GzipFile(dst, "wb", compresslevel=9, mtime=0) as g:
        for _ in range(4 * n):
            g.write(f.readline())


def main():
    O.build()
    os.makedirs(os.path.join(OUT, "refs"), exist_ok=True)
    for f in sorted(glob.glob(REF + "/refs/*.fa.gz")):
        shutil.copyfile(f, os.path.join(OUT, "refs", os.path.basename(f)))
    for m in ("mock_1", "mock_2"):
        head_fastq("%s/%s.fastq.gz" % (REF, m), "%s/%s.20k.fastq.gz" % (OUT, m), N_SUB)
    sp = O.sketch_params(21)
    targets = []
    for f in sorted(glob.glob(OUT + "/refs/*.fa.gz")):
        name = re.match(r"^([\w\.\_]+\.\d+)", os.path.basename(f)).group(1)
        targets += O.compute_targets(list(O.read_fastx(f)), name, sp, split_number=10, split_overlap=150, name_filters=["plasmid"])
    tmp = tempfile.mkdtemp()
    db = O.DB(O.build_db(targets, tmp, sp, num_hashes=1, fpr=0.3, block_size=16))
    summary = {}
    for tag, files in (("subset", ["%s/mock_1.20k.fastq.gz" % OUT, "%s/mock_2.20k.fastq.gz" % OUT]), ("full", [REF + "/mock_1.fastq.gz", REF + "/mock_2.fastq.gz"])):
        ids, seqs = [], []
        for f in files:
            for i, _h, s in O.read_fastx(f):
                ids.append(i); seqs.append(s)
        res = db.search(seqs)
        nh = np.diff(res.hit_off.astype(np.int64))
        summary[tag] = {"queries": len(seqs), "matched": int((nh > 0).sum()), "rows": int(len(res.hits))}
        if tag == "subset":
            tsv = O.format_tsv(db, ids, res)
            with gzip.GzipFile(OUT + "/expected.20k.tsv.gz", "wb", compresslevel=9, mtime=0) as g:
                g.write(tsv.encode())
    assert summary["full"]["queries"] == 349084 and summary["full"]["matched"] == 308839, summary      # G1
    json.dump(summary, open(OUT + "/summary.json", "w"), indent=1)
    shutil.rmtree(tmp)
    print(summary)


if __name__ == "__main__":
    main()


# tests/golden/demo_searching/refs/: the nine genomes of /root/reference/demo-searching/refs, byte for byte (input data of the reference's
# genome-similarity demo; the expected qCov / tCov / jacc tables are the ones printed in demo-searching/README.md:61-68 and :102-109, G3 / G4):
#   mkdir -p tests/golden/demo_searching/refs && cp /root/reference/demo-searching/refs/*.fasta.gz tests/golden/demo_searching/refs/

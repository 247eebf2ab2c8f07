"""Generates the committed golden fixtures from the REFERENCE's own demo data (run in the build container,
where /root/reference exists):  python tests/golden/make_golden.py

The values come from the oracle AFTER it reproduced the reference's published outputs G1-G5 on the same
files (tests/test_oracle_golden.py), so they pin both the oracle (CPU tests) and the CUDA kernels (GPU tests)
on a box that has no /root/reference.

  reads_k21.json   : 64 real reads of demo-profiling/mock_1.fastq.gz (incl. the 10 reads of golden table G2),
                     their canonical ntHash1 codes (k=21) and the expected per-read rows of G2.
  sketch_k31.json  : a 30 kb slice of demo-searching NC_018658.1 with its FracMinHash (scale 1000 and 50),
                     closed-syncmer (s=15, scale 62 and 1) and minimizer (w=10) code lists, k=31.
  fpr.json         : QueryFPR(n, c, p) known answers as exact float64 hex, incl. the four G2 values.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    reads = []
    for i, (rid, _h, s) in enumerate(O.read_fastx(REF + "/demo-profiling/mock_1.fastq.gz")):
        if i >= 64:
            break
        reads.append((rid.decode(), s.decode()))
    sp = O.sketch_params(21)
    out = {"k": 21, "reads": [{"id": r, "seq": s, "codes": [int(c) for c in O.generate_kmers(s.encode(), sp)]} for r, s in reads],
           # docs/tutorial/profiling/index.md:203-211 (reference's own table): queryIdx → (target, chunkIdx, mKmers, FPR)
           "g2_rows": {"1": ["GCF_000006945.2", 9, 90, "7.4626e-15"], "2": ["GCF_000006945.2", 6, 130, "7.4626e-15"],
                       "3": ["GCF_000006945.2", 6, 121, "7.4626e-15"], "4": ["GCF_000006945.2", 1, 101, "7.4626e-15"],
                       "5": ["GCF_000006945.2", 9, 83, "7.8754e-15"], "6": ["GCF_000006945.2", 2, 103, "7.4626e-15"],
                       "7": ["GCF_000006945.2", 5, 86, "7.4671e-15"], "8": ["GCF_000006945.2", 3, 84, "7.5574e-15"],
                       "9": ["GCF_000006945.2", 1, 89, "7.4626e-15"]}}
    json.dump(out, open(os.path.join(HERE, "reads_k21.json"), "w"))

    rec = next(iter(O.read_fastx(REF + "/demo-searching/refs/NC_018658.1.fasta.gz")))
    seq = rec[2][100000:130000]
    sk = {"k": 31, "seq": seq.decode(), "lists": {}}
    for name, spx in (("scaled1000", O.sketch_params(31, scaled=True, scale=1000)), ("scaled50", O.sketch_params(31, scaled=True, scale=50)),
                      ("syncmer15_scaled62", O.sketch_params(31, scaled=True, scale=62, syncmer_s=15)),
                      ("syncmer15", O.sketch_params(31, syncmer_s=15)), ("minimizer10", O.sketch_params(31, minimizer_w=10))):
        sk["lists"][name] = [int(c) for c in O.generate_kmers(seq, spx)]
    json.dump(sk, open(os.path.join(HERE, "sketch_k31.json"), "w"))

    cases = [(130, 90, 0.3), (130, 83, 0.3), (130, 86, 0.3), (130, 84, 0.3), (130, 72, 0.3), (130, 130, 0.3), (130, 40, 0.3), (130, 10, 0.3),
             (100, 56, 0.3), (260, 150, 0.3), (20, 11, 0.3), (10071, 7552, 0.01), (500, 300, 0.05), (3000, 1700, 0.3), (130, 0, 0.3), (249, 140, 0.25)]
    json.dump([{"n": n, "c": c, "p": p, "fpr_hex": float(O.query_fpr(n, c, p)).hex(), "fmt": O.go_fmt_e4(O.query_fpr(n, c, p))} for n, c, p in cases],
              open(os.path.join(HERE, "fpr.json"), "w"), indent=0)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()

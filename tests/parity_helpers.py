"""Shared builders for the parity tests (oracle side only builds inputs and expected values)."""
import os

import numpy as np


def make_synth_targets(O, sp, gseed, n_genomes, genome_len, n_chunks, overlap):
    targets = []
    for g in range(n_genomes):
        seq = O.synth_genome(gseed, g, genome_len)
        targets += O.compute_targets([(b"g%d" % g, b"g%d" % g, seq)], "synth_%06d" % g, sp, split_number=n_chunks,
                                     split_overlap=overlap)
    return targets


def make_reads(O, rseed, n, n_genomes, genome_len, gseed, read_len=150):
    return [O.synth_read(rseed, r, n_genomes, genome_len, read_len, gseed) for r in range(n)]


def edge_reads(k):
    """the edge cases the survey lists: N runs, lowercase, len<k, len==k, empty, IUPAC, duplicates"""
    rng = np.random.default_rng(7)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    def rnd(n):
        return acgt[rng.integers(0, 4, n)].tobytes()
    r = [b"", b"A", rnd(k - 1), rnd(k), rnd(k + 1), b"N" * 150, rnd(60) + b"N" * 30 + rnd(60), rnd(150).lower(),
         rnd(40) + b"RYKMSW" + rnd(100), b"ACGT" * 40, b"A" * 150, rnd(29), rnd(30), rnd(31), rnd(300), rnd(1000)]
    return r


def hits_to_set(hits):
    return set((int(h["query"]), int(h["target"]), int(h["count"])) for h in hits)

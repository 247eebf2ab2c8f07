"""bench.py --impl reference on CPU: the JSON contract of the reference arm and its N>1 workload (every read against the
N blocks of the b200 arm's config), with the GPU index builder replaced by the oracle's builder at a small scale."""
import argparse
import importlib
import json

import numpy as np

import parity_helpers as helpers


def test_reference_arm_contract_and_multi_rank_workload(oracle, monkeypatch, capsys):
    O = oracle
    monkeypatch.setenv("KMCP_BENCH_SCALE", "quick")
    import bench
    importlib.reload(bench)
    monkeypatch.setattr(bench, "N_GENOMES", 20)
    monkeypatch.setattr(bench, "GENOME_LEN", 20000)
    monkeypatch.setattr(bench, "BLOCK_SIZE", 200)
    staged = {}

    def stage(tmp, world, n_total):
        # what bench.py's helper process leaves behind: ONE database of `world` blocks under tmp/R001, the reads in tmp/reads.u8
        sp = O.sketch_params(bench.K)
        targets = helpers.make_synth_targets(O, sp, bench.GENOME_SEED, 20 * world, 20000, bench.N_CHUNKS, bench.OVERLAP)
        r001 = O.build_db(targets, tmp, sp, num_hashes=bench.H, fpr=bench.FPR, block_size=200)
        assert len(O.DB(r001).info.__class__._fields_) and O.DB(r001).info.n_blocks == world
        reads = helpers.make_reads(O, bench.READ_SEED, n_total, 20, 20000, bench.GENOME_SEED, bench.READ_LEN)
        np.frombuffer(b"".join(reads), dtype=np.uint8).tofile(tmp + "/reads.u8")
        staged["world"], staged["n"] = world, n_total

    for world in (1, 2):
        args = argparse.Namespace(gpus=world, steps=2, warmup=1)
        bench.run_reference(args, 0, world, stage=stage)
        out = capsys.readouterr().out.strip().splitlines()
        line = json.loads(out[-1])
        assert line["impl"] == "reference" and line["metric"] == bench.METRIC and line["unit"] == "reads/s"
        assert line["n_gpus"] == world and line["steps"] == 2 and line["warmup"] == 1 and line["higher_is_better"] is True
        assert line["e2e"] == {"value": line["value"], "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
        assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["value"] == line["value"] and line["cpu_baseline"]["cores"] >= 1
        assert line["config"]["blocks_searched"] == world and line["gpu_launches"] == 0
        step_reads = line["config"]["reads_per_step_cpu_sample"]
        assert staged == {"world": world, "n": step_reads * 3}
        # value counts read x shard probes, like the b200 arm at N ranks
        assert abs(line["value"] - world * line["job_reads_per_s"]) < 1e-6 * line["value"]
        assert abs(line["job_reads_per_s"] - step_reads * 2 / (line["ms_per_step"] * 2 / 1e3)) < 1e-6 * line["job_reads_per_s"]
    # the other ranks exit without work or output
    bench.run_reference(argparse.Namespace(gpus=2, steps=2, warmup=1), 1, 2, stage=stage)
    assert capsys.readouterr().out == ""
    # without a GPU the stock staging reports "unavailable" on one line and does not raise
    import torch
    if not torch.cuda.is_available():
        bench.run_reference(argparse.Namespace(gpus=1, steps=1, warmup=1), 0, 1)
        line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
        assert line["impl"] == "reference" and "unavailable" in line


def test_hifi_batch_is_seeded_and_follows_the_genome_generator(oracle, monkeypatch):
    """bench.py's HiFi read generator (C5): deterministic in (seed, step), lengths in the stated range, and every read a
    lightly mutated window of the seeded genome the device generator produces (oracle.synth_genome)"""
    O = oracle
    monkeypatch.setenv("KMCP_BENCH_SCALE", "quick")
    import bench
    importlib.reload(bench)
    off, seq = bench.hifi_batch(5, 0, 40, 3, 7, 60000)
    off2, seq2 = bench.hifi_batch(5, 0, 40, 3, 7, 60000)
    assert np.array_equal(off, off2) and np.array_equal(seq, seq2)
    assert not np.array_equal(seq[:200], bench.hifi_batch(5, 1, 40, 3, 7, 60000)[1][:200])
    lens = np.diff(off.astype(np.int64))
    assert lens.min() >= 45 and lens.max() <= 45000 and 4000 < lens.mean() < 16000
    genomes = [O.synth_genome(3, g, 60000) for g in range(7)]
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    for i in range(10):
        r = seq[int(off[i]):int(off[i + 1])].tobytes()
        best = 1.0
        for cand in (r, r.translate(comp)[::-1]):
            probe = cand[:24] if len(cand) >= 24 else cand
            for g in genomes:
                # a 24-mer free of substitutions somewhere in the read locates it; compare the whole window then
                for start in range(0, max(1, len(cand) - 24), 24):
                    pos = g.find(cand[start:start + 24])
                    if pos >= 0 and pos - start >= 0 and pos - start + len(cand) <= len(g):
                        w = g[pos - start:pos - start + len(cand)]
                        best = min(best, sum(a != b for a, b in zip(w, cand)) / len(cand))
                        break
        assert best < 0.02, (i, best)

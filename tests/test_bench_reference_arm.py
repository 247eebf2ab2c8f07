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

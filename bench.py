#!/usr/bin/env python
"""bench.py — reads/s of the `kmcp search` hot path on B200 (BASELINE.json metric), one JSON line on rank 0.

Workload = BASELINE.json configs[1]: synthetic 10k-chunk COBS index (1,000 seeded random genomes x 4 Mb, 10 chunks, k=21,
h=1, fpr 0.3, ONE block of 10,000 targets ≈ 1.4 GB ≫ L2) per GPU, resident in HBM, 150 bp synthetic reads (80 % sampled from
the genomes with 1 % substitutions, 20 % random).  A step = one batch of READS_PER_STEP reads through the whole hot path
(hash → probe → hit sort → D2H).  Every step uses different reads; the index is far larger than L2, so no explicit L2 flush
is needed (stated in `config`).  Two batches are kept submitted (kmcpg_search_submit), so step s+1 starts on the GPU the
moment step s ends.

  value   : reads/s, inputs resident in HBM when the timed region starts.
  e2e     : reads/s from PINNED HOST buffers to matches in host memory: H2D of the reads, kernels, D2H of hits, host post-filter
            (tCov/FPR/sort) — the number to compare with the CPU arm.
  roofline: probe kernel only; achieved = algorithmic row bytes (n_kmers·h·Σ numRowBytes per read) of a launch ÷ its
            CUDA-event duration on the launching stream; peak = MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline: the oracle's restatement of the reference algorithm (64-row buffer, byte transpose, positional popcount) on all
            host threads, bounded sample — kind "port" (the Go reference cannot be built here: no Go toolchain).
  gtdb_scale: BASELINE.json configs[3] shape — ONE fixed index of 852,050 targets (85,205 genomes x 10 chunks), h=3, 32 blocks
            of 26,632 targets (3,329-byte rows), block-sharded over the N ranks exactly as kmcpg_open_db(shard_rank, shard_world)
            shards a database, the SAME reads at every N (strong scaling): job reads/s, per-rank probe GB/s, and a digest of the
            merged (query, target, count) list that must be equal at N = 1, 2, 4, 8.

N>1 (torchrun, one rank per GPU; kmcp_b200/multigpu.py): ONE database of N blocks of 10,000 targets, rank r holds the block the
shard plan gives it (the index grows with N = weak scaling).  Inside the timed region of `value`: the NCCL broadcast of every
batch from rank 0's HBM (NVLink; prefetched one step ahead), the search on every rank, and the return of every rank's hit list to
rank 0's host (device→host into per-rank shared-memory segments that rank 0 reads in place).  `e2e` adds rank 0's H2D of every batch
from pinned host memory, the merge of the per-rank lists and the host post-filter on rank 0.

`--impl reference` times the CPU port alone (rank 0), same metric/config/unit.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SCALE = os.environ.get("KMCP_BENCH_SCALE", "full")
if SCALE == "full":
    N_GENOMES, GENOME_LEN, READS_PER_STEP, CPU_SAMPLE0 = 1000, 4_000_000, 1_000_000, 20_000
    # GTDB-scale shape: genome length scaled ("scale down genome length, not target count — row width is what matters")
    GTDB_GENOMES, GTDB_GL, GTDB_BLOCK, GTDB_READS = 85_205, int(os.environ.get("KMCP_GTDB_GL", 875_000)), 26_632, int(os.environ.get("KMCP_GTDB_READS", 100_000))
else:  # quick functional check of the harness
    N_GENOMES, GENOME_LEN, READS_PER_STEP, CPU_SAMPLE0 = 100, 200_000, 100_000, 5_000
    GTDB_GENOMES, GTDB_GL, GTDB_BLOCK, GTDB_READS = 2_000, 60_000, 632, 20_000
N_CHUNKS, OVERLAP, K, H, FPR, READ_LEN = 10, 150, 21, 1, 0.3, 150
BLOCK_SIZE = N_GENOMES * N_CHUNKS
GENOME_SEED, READ_SEED = 1, 2
GTDB_SEED, GTDB_READ_SEED, GTDB_H, GTDB_STEPS, GTDB_WARMUP = 3, 4, 3, 3, 1
# BASELINE configs[4]: HiFi reads (benchmarks/mock-hifi-zymo/README.md:27-31: min 45, median 8.3 kb, mean 9.3 kb, max 45.8 kb), 0.5 % errors
C5_SEED, C5_READS, C5_STEPS, C5_WARMUP = 5, (2_000 if SCALE == "full" else 300), 2, 1
# BASELINE configs[2]: FracMinHash (scale 1000, k=21, h=3, fpr 0.001) genome-vs-genome search, 1,000 assemblies, one GPU
C3_GENOMES, C3_GL, C3_QUERIES = (1000, 4_000_000, 250) if SCALE == "full" else (60, 400_000, 30)
METRIC = "reads/sec (kmcp search, 150bp)"


def workload_name():
    return ("synthetic %d-chunk COBS index (%d genomes x %.1f Mb, k=%d, h=%d, fpr %.1f, 1 block/GPU) resident in HBM, "
            "%d x %d bp reads per step" % (BLOCK_SIZE, N_GENOMES, GENOME_LEN / 1e6, K, H, FPR, READS_PER_STEP, READ_LEN))


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for i, nm in enumerate(names):
                    if r[5 + i].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """dram bytes per probe launch from the committed ncu --set full summary, if one exists"""
    p = os.path.join(ROOT, "profiles", "probe_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


def db_yml(alias, n_targets, files, hashes=H, block_size=BLOCK_SIZE):
    return {"version": 4, "unikiVersion": 4, "alias": alias, "k": K, "ks": [K], "hashed": True, "canonical": True, "scaled": False,
            "scale": 0, "minimizer": False, "minimizer-w": 0, "syncmer": False, "syncmer-s": 0, "split-seq": True, "split-size": 0,
            "split-num": N_CHUNKS, "split-overlap": OVERLAP, "compact-size": False, "hashes": hashes, "fpr": FPR,
            "numNameGroups": int(n_targets), "blocksize": block_size, "totalKmers": 0, "files": files}


def dump_db_for_cpu(ctx, tmpdir):
    """HBM-resident synthetic DB → .uniki files + __db.yml that the CPU port can open"""
    from oracle import oracle as O
    r001 = os.path.join(tmpdir, "R001")
    os.makedirs(r001, exist_ok=True)
    info = ctx.db_info()
    files = []
    for b in range(info.n_resident_blocks):
        fn = "_block%03d.uniki" % (b + 1)
        ctx.write_block(b, os.path.join(r001, fn))
        files.append(fn)
    O.write_db_yml(os.path.join(r001, "__db.yml"), db_yml("bench", info.n_targets, files))
    return r001


def time_cpu_port(r001, reads_u8, n_reads, target_seconds=12.0):
    """reference-algorithm port on all host threads, bounded sample; returns (reads/s, cores, sample_n)"""
    from oracle import oracle as O
    odb = O.DB(r001)
    cores = os.cpu_count() or 1
    off_all = np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(READ_LEN)
    n0 = min(CPU_SAMPLE0, n_reads)
    t0 = time.perf_counter()
    odb.search(packed=(reads_u8[: n0 * READ_LEN], off_all[: n0 + 1]), threads=cores, algo=1)
    dt0 = time.perf_counter() - t0
    n1 = int(min(n_reads, max(n0, n0 * target_seconds / max(dt0, 1e-6))))
    t0 = time.perf_counter()
    odb.search(packed=(reads_u8[: n1 * READ_LEN], off_all[: n1 + 1]), threads=cores, algo=1)
    dt = time.perf_counter() - t0
    odb.close()
    return n1 / dt, cores, n1


def stage_main(tmp, world, n_reads_total):
    """helper process of the reference arm (`bench.py --stage-reference`): the index of this arm's config — ONE database of `world`
    blocks of 10,000 targets, as the b200 arm shards it — as .uniki files the CPU port can open, plus the seeded reads, written under
    `tmp`.  Built by the GPU index builder (byte-identical to the oracle's builder, tests/test_gpu_parity.py) because the 1.4 GB
    blocks take minutes on the CPU; the timing process itself never maps libkmcp_gpu.so."""
    from kmcp_b200 import api
    from oracle import oracle as O
    ctx = api.Context(0)
    try:
        ctx.build_synth_db(GENOME_SEED, N_GENOMES * world, GENOME_LEN, k=K, n_chunks=N_CHUNKS, overlap=OVERLAP, num_hashes=H, fpr=FPR, block_size=BLOCK_SIZE)
        r001 = os.path.join(tmp, "R001")
        os.makedirs(r001, exist_ok=True)
        info = ctx.db_info()
        files = []
        for b in range(info.n_resident_blocks):
            fn = "_block%03d.uniki" % (b + 1)
            ctx.write_block(b, os.path.join(r001, fn))
            files.append(fn)
        O.write_db_yml(os.path.join(r001, "__db.yml"), db_yml("bench", info.n_targets, files))
        d = ctx.device_alloc(n_reads_total * READ_LEN)
        ctx.synth_reads(READ_SEED, 0, n_reads_total, READ_LEN, GENOME_SEED, N_GENOMES, GENOME_LEN, d)
        ctx.d2h(d, n_reads_total * READ_LEN).tofile(os.path.join(tmp, "reads.u8"))
        ctx.device_free(d)
    finally:
        ctx.close()


def stage_in_helper_process(tmp, world, n_total):
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT")}
    p = subprocess.run([sys.executable, os.path.abspath(__file__), "--stage-reference", tmp, str(world), str(n_total)], env=env, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError((p.stderr.strip().splitlines() or ["helper process failed"])[-1][:160])


def run_reference(args, rank, world, stage=stage_in_helper_process):
    """--impl reference: the CPU port of the reference algorithm, rank 0 only, on this arm's config: at N ranks the index is N blocks
    of 10,000 targets (one per GPU in the b200 arm), every read is searched against all of them, and `value` counts read x shard
    probes exactly as the b200 arm does."""
    if rank != 0:
        return
    from oracle import oracle as O
    O.build()
    tmp = "/dev/shm/kmcp_bench_ref" if os.path.isdir("/dev/shm") else "/tmp/kmcp_bench_ref"
    shutil.rmtree(tmp, ignore_errors=True)
    os.makedirs(tmp)
    built_by = "gpu index builder in a helper process (byte-identical to the oracle builder, tests/test_gpu_parity.py)"
    step_reads = max(1000, (50_000 if SCALE == "full" else 5_000) // world)      # the CPU work per step stays the same at every N
    n_total = step_reads * (args.steps + args.warmup)
    try:
        stage(tmp, world, n_total)       # leaves tmp/R001 (the database) and tmp/reads.u8
    except Exception as e:  # no usable GPU: nothing to build the 1.4 GB index with in reasonable time
        shutil.rmtree(tmp, ignore_errors=True)
        print(json.dumps({"impl": "reference", "unavailable": "cannot stage the synthetic index without the GPU builder: %s" % str(e)[:160]}))
        return
    reads = np.fromfile(os.path.join(tmp, "reads.u8"), dtype=np.uint8)
    odb = O.DB(os.path.join(tmp, "R001"))
    if stage is stage_in_helper_process:
        assert "kmcp_b200" not in sys.modules, "the reference arm must not load the product library"
    cores = os.cpu_count() or 1
    off = np.arange(step_reads + 1, dtype=np.uint64) * np.uint64(READ_LEN)
    n_hits = [0]

    def step(i):      # explicit thread count: torchrun sets OMP_NUM_THREADS=1
        batch = reads[i * step_reads * READ_LEN:(i + 1) * step_reads * READ_LEN]
        n_hits[0] += len(odb.search(packed=(batch, off), threads=cores, algo=1).hits)

    for i in range(args.warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(args.warmup + i)
    dt = time.perf_counter() - t0
    v = world * step_reads * args.steps / dt
    odb.close()
    shutil.rmtree(tmp, ignore_errors=True)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "reads/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64/u8 bitset",
        "data": "synthetic", "config": {"workload": workload_name(), "reads_per_step": READS_PER_STEP, "targets_total": BLOCK_SIZE * world,
                                        "reads_per_step_cpu_sample": step_reads, "db_built_by": built_by, "blocks_searched": world,
                                        "multi_gpu_units": "value counts read×shard probes (every read against each of the %d 10k-target blocks)" % world if world > 1 else "reads"},
        "job_reads_per_s": step_reads * args.steps / dt,
        "cpu_baseline": {"value": v, "unit": "reads/s", "cores": cores, "kind": "port",
                         "sample": "%d reads per step against %d block(s), restated reference algorithm (oracle algo=1: 64-row buffer, byte transpose, "
                                   "positional popcount), OpenMP all threads" % (step_reads, world)},
        "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))


def _splitmix64(x):
    with np.errstate(over="ignore"):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def hifi_batch(seed, step, n_reads, gseed, n_genomes, genome_len):
    """one packed batch [offsets | bases] of HiFi-shaped reads sampled from the seeded genomes (the generator of synth.cu, in numpy):
    log-normal lengths (median 8.3 kb, sigma 0.5, clipped to 45 .. 45,000 and to the genome), 0.5 % substitutions, random strand"""
    rng = np.random.default_rng(seed * 1_000_003 + step)
    lens = np.clip(rng.lognormal(np.log(8348.0), 0.5, n_reads), 45, min(45_000, genome_len)).astype(np.int64)
    off = np.zeros(n_reads + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens).astype(np.uint64)
    seq = np.empty(int(off[-1]), dtype=np.uint8)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    for i in range(n_reads):
        g, ln = int(rng.integers(0, n_genomes)), int(lens[i])
        pos = int(rng.integers(0, genome_len - ln + 1))
        with np.errstate(over="ignore"):
            gkey = _splitmix64(np.array([(gseed * 0x100000001B3 + g) & 0xFFFFFFFFFFFFFFFF], dtype=np.uint64))[0]
            p = np.arange(pos, pos + ln, dtype=np.uint64)
            b = ((_splitmix64(gkey + (p >> np.uint64(5))) >> (np.uint64(2) * (p & np.uint64(31)))) & np.uint64(3)).astype(np.uint8)
        sub = rng.random(ln) < 0.005
        b[sub] = (b[sub] + rng.integers(1, 4, int(sub.sum())).astype(np.uint8)) & 3
        if rng.integers(0, 2):
            b = (3 - b)[::-1]
        seq[int(off[i]):int(off[i + 1])] = acgt[b]
    return off, seq


def pack_batches(torch, ctx, seed, first_read, n_steps, n_reads, gseed, n_genomes, genome_len, off_np):
    """device tensor of n_steps packed batches: per batch the (n_reads + 1) u64 offsets, then the read bytes"""
    off_bytes, seq_bytes = off_np.nbytes, n_reads * READ_LEN
    bb = off_bytes + seq_bytes
    t = torch.empty(n_steps * bb, dtype=torch.uint8, device="cuda")
    d_off = torch.from_numpy(off_np.view(np.uint8)).cuda()
    for s in range(n_steps):
        t[s * bb:s * bb + off_bytes].copy_(d_off)
        ctx.synth_reads(seed, first_read + s * n_reads, n_reads, READ_LEN, gseed, n_genomes, genome_len, t.data_ptr() + s * bb + off_bytes)
    torch.cuda.synchronize()
    return t, bb, off_bytes


def main():
    if len(sys.argv) >= 5 and sys.argv[1] == "--stage-reference":
        stage_main(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]))
        return
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gtdb", action="store_true", help="skip the GTDB-scale block (development runs)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist
    from kmcp_b200 import api, multigpu

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
        del os.environ["NCCL_DEBUG"]             # keeps NCCL's version banner out of stdout: one JSON line only
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream()
    ctx = api.Context(local_rank)
    ctx.set_stream(stream.cuda_stream)
    run_id = "kmcpb_%s_%d" % (os.environ.get("MASTER_PORT", "0"), os.getpid())        # names the shared-memory segments of this run
    if world > 1:
        box = [run_id]
        dist.broadcast_object_list(box, src=0)          # every rank uses rank 0's id, however the ranks were launched
        run_id = box[0]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def gather_objects(obj):
        if world == 1:
            return [obj]
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    # ONE database of `world` blocks of 10,000 targets; this rank keeps the block(s) the shard plan gives it
    t_build = time.perf_counter()
    ctx.build_synth_db(GENOME_SEED, N_GENOMES * world, GENOME_LEN, k=K, n_chunks=N_CHUNKS, overlap=OVERLAP, num_hashes=H, fpr=FPR, block_size=BLOCK_SIZE,
                       shard_rank=rank, shard_world=world)
    t_build = time.perf_counter() - t_build
    info = ctx.db_info()

    n_steps_total = args.warmup + args.steps
    off_np = np.arange(READS_PER_STEP + 1, dtype=np.uint64) * np.uint64(READ_LEN)
    h_off, h_off_ptr = api.pinned_array(off_np.nbytes)
    h_off[:] = off_np.view(np.uint8)
    step_bytes = READS_PER_STEP * READ_LEN
    params = ctx.default_params()
    # batches of every step resident in rank 0's HBM before any timing (reads sampled from the first N_GENOMES genomes)
    d_batches = bb = off_bytes = None
    if rank == 0:
        d_batches, bb, off_bytes = pack_batches(torch, ctx, READ_SEED, 0, n_steps_total, READS_PER_STEP, GENOME_SEED, N_GENOMES, GENOME_LEN, off_np)
    else:
        bb, off_bytes = off_np.nbytes + step_bytes, off_np.nbytes

    sampler = ClockSampler(local_rank)
    digests = []
    if world == 1:
        # ---------------- value (N = 1): device-resident inputs, two jobs in flight ----------------
        def submit(s):
            return ctx.submit(d_batches.data_ptr() + s * bb + off_bytes, d_batches.data_ptr() + s * bb, READS_PER_STEP, params, device=True, host_off_ptr=h_off_ptr)

        def run_steps(first, n):
            outs, jobs = [], []
            for s in range(first, first + n):
                jobs.append(submit(s))
                if len(jobs) == 2:
                    outs.append(ctx.wait(jobs.pop(0), copy=False))
            while jobs:
                outs.append(ctx.wait(jobs.pop(0), copy=False))
            return outs

        with torch.cuda.stream(stream):
            run_steps(0, args.warmup)
            barrier()
            sampler.start()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(stream)
            outs = run_steps(args.warmup, args.steps)
            ev1.record(stream)
            barrier()
            clocks = sampler.stop()
            ms_total = ev0.elapsed_time(ev1)
        sh = None
    else:
        # ---------------- value (N > 1): broadcast + search + hit return inside the timed region ----------------
        sh = multigpu.ShardedSearch(ctx, rank, world, READS_PER_STEP, bb, hit_cap=4 * READS_PER_STEP, name=run_id, device=dev, dist=dist, params=params)

        def feed_dev(first):
            return lambda s: (d_batches[(first + s) * bb:(first + s + 1) * bb], False)

        def consume_digest(s, lists, meta):
            merged = multigpu.merge_lists(lists, 0, READS_PER_STEP)
            digests.append((len(merged), multigpu.hits_digest(merged)))

        with torch.cuda.stream(stream):
            sh.run(args.warmup, feed_dev(0), consume_digest, host_off=off_np)
            digests.clear()
            barrier()
            if rank == 0:
                sampler.start()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(stream)
            outs = sh.run(args.steps, feed_dev(args.warmup), consume_digest, host_off=off_np)      # returns when rank 0 holds every hit list
            ev1.record(stream)
            barrier()
            clocks = sampler.stop() if rank == 0 else None
            ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    probe_ms = sum(o.ms_probe for o in outs)
    probe_launches = sum(o.probe_launches for o in outs)
    probe_bytes = sum(o.probe_row_bytes for o in outs)
    launches = sum(o.kernel_launches for o in outs)
    n_hits = sum(o.n_hits for o in outs)
    value = world * READS_PER_STEP * args.steps / (ms_total / 1e3)       # units all ranks processed ÷ time (§5: read×shard probes)
    value_hits = sum(n for n, _d in digests) if world > 1 else n_hits

    # ---------------- e2e: pinned host buffers → matches in host memory ----------------
    h_batches = h_ptr = None
    if rank == 0:
        h_batches, h_ptr = api.pinned_array(args.steps * bb)
        h_batches[:] = d_batches[args.warmup * bb:].cpu().numpy()
    eopts = ctx.default_engine_opts()
    e2e_matches = [0]
    e2e_break = {"search_call_ms": 0.0, "post_filter_ms": 0.0, "engine_call_ms": 0.0}
    with torch.cuda.stream(stream):
        if world == 1:
            # kmcpg_engine_search from two host threads (a Go host calls it from goroutines): the executor runs their batches back to back
            ctx.engine_search_ptr(h_ptr + off_bytes, h_off_ptr, READS_PER_STEP, eopts)          # warm the host-side caches once
            res = [None] * args.steps

            def worker(t):
                for s in range(t, args.steps, 2):
                    res[s] = ctx.engine_search_ptr(h_ptr + s * bb + off_bytes, h_off_ptr, READS_PER_STEP, eopts, copy=False)

            barrier()
            t0 = time.perf_counter()
            th = [threading.Thread(target=worker, args=(t,)) for t in range(2)]
            for t in th:
                t.start()
            for t in th:
                t.join()
            torch.cuda.synchronize()
            e2e_s = time.perf_counter() - t0
            for r in res:
                e2e_matches[0] += r.n_matches
                e2e_break["search_call_ms"] += r.ms_gpu_total / args.steps; e2e_break["post_filter_ms"] += r.ms_post / args.steps
                e2e_break["engine_call_ms"] += r.ms_total / args.steps
            e2e_path = "kmcpg_engine_search from two host threads (pinned host reads → H2D → kernels → D2H hits → host tCov/FPR/sort)"
        else:
            tsizes = ctx.target_sizes()

            def feed_host(s):
                return torch.from_numpy(h_batches[s * bb:(s + 1) * bb]), True

            def consume_full(s, lists, meta):
                merged = multigpu.merge_lists(lists, 0, READS_PER_STEP)
                r = multigpu.postfilter(eopts, meta.n_kmers, meta.query_len, merged, tsizes, FPR, K, copy=False)
                e2e_matches[0] += r.n_matches
                e2e_break["post_filter_ms"] += r.ms_post / args.steps

            barrier()
            t0 = time.perf_counter()
            sh.run(args.steps, feed_host, consume_full, host_off=off_np)
            barrier()
            e2e_s = max_over_ranks(time.perf_counter() - t0)
            e2e_path = ("rank 0 pinned reads → H2D → ncclBroadcast (prefetched one step ahead) → kmcpg_search_submit on every rank → hit lists D2H into per-rank "
                        "shared-memory segments → rank 0 merges them in place (kmcpg_merge_hits) → host tCov/FPR/sort (kmcpg_engine_postfilter)")
    d2h_bytes = int(12 * n_hits / args.steps + 8 * READS_PER_STEP)
    e2e_value = world * READS_PER_STEP * args.steps / e2e_s

    # ---------------- CPU baseline beside it (rank 0, N=1 only) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        tmp = "/dev/shm/kmcp_bench_cpu" if os.path.isdir("/dev/shm") else "/tmp/kmcp_bench_cpu"
        shutil.rmtree(tmp, ignore_errors=True)
        try:
            r001 = dump_db_for_cpu(ctx, tmp)
            reads_only = np.concatenate([h_batches[s * bb + off_bytes:(s + 1) * bb] for s in range(args.steps)])
            v, cores, n1 = time_cpu_port(r001, reads_only, args.steps * READS_PER_STEP)
            del reads_only
            cpu = {"value": v, "unit": "reads/s", "cores": cores, "kind": "port",
                   "sample": "%d reads of the timed set, restated reference algorithm (oracle algo=1: 64-row buffer, byte transpose, "
                             "positional popcount), OpenMP on all %d host threads, same index in RAM" % (n1, cores)}
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
    if rank == 0:
        api.host_free(h_ptr)
    del d_batches
    if sh is not None:
        sh.close()

    # ---------------- GTDB-scale block: fixed index sharded over the ranks, fixed reads, digest of the merged hit list ----------------
    gtdb = c5 = c3 = None
    if not args.no_gtdb:
        gtdb = gtdb_block(torch, dist, api, multigpu, ctx, stream, dev, rank, world, run_id, barrier, max_over_ranks, gather_objects)
        c5 = c5_block(torch, dist, api, multigpu, ctx, stream, dev, rank, world, run_id, barrier, max_over_ranks, gather_objects)      # against the same sharded index
        if world == 1:
            c3 = c3_block(torch, api, ctx)

    api.host_free(h_off_ptr)
    if rank == 0:
        peak, peak_src = measured_peak()
        achieved = (probe_bytes / 1e9) / (probe_ms / 1e3) if probe_ms > 0 else 0.0
        traffic = ncu_traffic()
        line = {
            "metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64 hash / u32 bit-sliced counters", "data": "synthetic",
            "config": {"workload": workload_name(), "reads_per_step": READS_PER_STEP, "index_bytes_per_gpu": int(info.resident_bytes),
                       "index_disk_bytes_per_gpu": int(info.disk_bytes), "targets_per_gpu": int(info.n_targets // world), "targets_total": int(info.n_targets),
                       "algorithmic_bytes_per_read": int(probe_bytes / max(1, READS_PER_STEP * args.steps)),
                       "l2": "index (%.2f GB) ≫ 126 MB L2 and every step probes different random rows: no explicit flush" % (info.resident_bytes / 1e9),
                       "parallelism": "one database of %d block(s) sharded over %d GPU(s), read batch broadcast, hit lists returned to rank 0" % (info.n_blocks, world),
                       "db_build_s": round(t_build, 2), "hits_per_step": int(value_hits / args.steps), "pipeline": "two batches submitted at a time (kmcpg_search_submit)",
                       "timed_region": "device-resident batches → search → hits in host memory" if world == 1 else
                                       "batches resident on rank 0 → ncclBroadcast → search on every rank → every rank's hits in rank 0's host memory (merged + digested there)",
                       "multi_gpu_units": "value counts read×shard probes (each rank probes every read against its own 10k-target block)" if world > 1 else "reads"},
            "job_reads_per_s": READS_PER_STEP * args.steps / (ms_total / 1e3),
            "roofline": {"bound": "hbm", "kernel": "probe_kernel<1,0,2,2,4,u32>", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_src, "launches": probe_launches, "avg_launch_ms": probe_ms / max(1, probe_launches),
                         "algorithmic_bytes_per_launch": probe_bytes / max(1, probe_launches),
                         "whole_step_frac": (probe_bytes / 1e9) / (ms_total / 1e3) / peak,
                         "traffic": (traffic["dram_over_algorithmic"] * probe_bytes / max(1, probe_launches)) if traffic else None,
                         "traffic_note": ("average launch of this run x the DRAM/algorithmic ratio of the committed ncu --set full capture: " + traffic["source"])
                         if traffic else "no ncu --set full capture committed yet",
                         "note": "launch durations are CUDA-event brackets on the probe stream; the query preparation of the next part runs beside the probe on its own stream"},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "reads/s", "h2d_bytes_per_step": step_bytes + off_np.nbytes, "d2h_bytes_per_step": d2h_bytes,
                    "matches_per_step": int(e2e_matches[0] / args.steps), "ms_per_step": e2e_s / args.steps * 1e3, "breakdown_ms_per_step": e2e_break,
                    "path": e2e_path},
            "gpu_launches": int(launches), "clocks": clocks,
            "stage_ms_per_step": {"hash (own stream, beside the probes)": sum(o.ms_hash for o in outs) / args.steps,
                                  "probe": probe_ms / args.steps, "call_wall": sum(o.ms_total for o in outs) / args.steps},
            "hit_list_digest": ("%016x" % (sum(d for _n, d in digests) & (2**64 - 1))) if digests else None,
            "gtdb_scale": gtdb, "c5_hifi": c5, "c3_fracminhash": c3,
        }
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def gtdb_block(torch, dist, api, multigpu, ctx, stream, dev, rank, world, run_id, barrier, max_over_ranks, gather_objects):
    """BASELINE.json configs[3] shape: a fixed index block-sharded over the ranks, fixed reads, strong scaling"""
    t0 = time.perf_counter()
    ctx.build_synth_db(GTDB_SEED, GTDB_GENOMES, GTDB_GL, k=K, n_chunks=N_CHUNKS, overlap=OVERLAP, num_hashes=GTDB_H, fpr=FPR, block_size=GTDB_BLOCK,
                       shard_rank=rank, shard_world=world)
    build_s = time.perf_counter() - t0
    info = ctx.db_info()
    n = GTDB_READS
    off_np = np.arange(n + 1, dtype=np.uint64) * np.uint64(READ_LEN)
    steps = GTDB_WARMUP + GTDB_STEPS
    d_batches = None
    bb = off_np.nbytes + n * READ_LEN
    if rank == 0:
        d_batches, bb, _ob = pack_batches(torch, ctx, GTDB_READ_SEED, 0, steps, n, GTDB_SEED, GTDB_GENOMES, GTDB_GL, off_np)
    sh = multigpu.ShardedSearch(ctx, rank, world, n, bb, hit_cap=max(1 << 20, 8 * n), name=run_id + "g", device=dev, dist=dist, params=ctx.default_params())
    digests = []

    def feed(first):
        return lambda s: (d_batches[(first + s) * bb:(first + s + 1) * bb], False)

    def consume(s, lists, meta):
        merged = multigpu.merge_lists(lists, 0, n)
        digests.append((len(merged), multigpu.hits_digest(merged)))

    with torch.cuda.stream(stream):
        sh.run(GTDB_WARMUP, feed(0), consume, host_off=off_np)
        digests.clear()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        outs = sh.run(GTDB_STEPS, feed(GTDB_WARMUP), consume, host_off=off_np)
        ev1.record(stream)
        barrier()
        ms = max_over_ranks(ev0.elapsed_time(ev1))
    sh.close()
    mine = {"rank": rank, "resident_blocks": int(info.n_resident_blocks), "index_bytes": int(info.resident_bytes), "sum_row_bytes": int(info.sum_row_bytes),
            "probe_ms": sum(o.ms_probe for o in outs), "probe_bytes": sum(o.probe_row_bytes for o in outs), "probe_launches": sum(o.probe_launches for o in outs),
            "launches": sum(o.kernel_launches for o in outs), "build_s": round(build_s, 1)}
    per_rank = gather_objects(mine)
    if rank != 0:
        return None
    peak, _src = measured_peak()
    for r in per_rank:
        r["probe_GBps"] = (r["probe_bytes"] / 1e9) / (r["probe_ms"] / 1e3) if r["probe_ms"] > 0 else 0.0
        r["probe_frac"] = r["probe_GBps"] / peak
    total_bytes = sum(r["probe_bytes"] for r in per_rank)
    job = n * GTDB_STEPS / (ms / 1e3)
    return {
        "workload": "GTDB-scale shape: %d genomes x %d chunks = %d targets, k=%d, h=%d, fpr %.1f, %d blocks of %d targets (%d-byte rows), genome length %d (index %.1f GB; "
                    "full GTDB would be 3.5 Mb / ~96 GB: SURVEY §8d C4 scales genome length, not target count), %d x %d bp reads per step, %d timed steps"
                    % (GTDB_GENOMES, N_CHUNKS, int(info.n_targets), K, GTDB_H, FPR, info.n_blocks, GTDB_BLOCK, (GTDB_BLOCK + 7) // 8, GTDB_GL,
                       sum(r["index_bytes"] for r in per_rank) / 1e9, n, READ_LEN, GTDB_STEPS),
        "scaling": "strong", "n_gpus": world, "job_reads_per_s": job, "ms_per_step": ms / GTDB_STEPS,
        "algorithmic_bytes_per_read": int(total_bytes / max(1, n * GTDB_STEPS)),
        "aggregate_probe_GBps_over_wall": (total_bytes / 1e9) / (ms / 1e3), "roofline_reads_per_s_at_peak": peak * 1e9 * world / max(1.0, total_bytes / max(1, n * GTDB_STEPS)),
        "frac_of_roofline": job / (peak * 1e9 * world / max(1.0, total_bytes / max(1, n * GTDB_STEPS))),
        "hits_per_step": int(sum(c for c, _d in digests) / max(1, len(digests))),
        "hit_list_digest": "%016x" % (sum(d for _c, d in digests) & (2**64 - 1)),
        "digest_note": "sum over the timed steps of kmcpg_hits_digest of the merged (query, target, count) list in (query, target) order: must be equal at every N",
        "per_rank": per_rank,
    }


def c5_block(torch, dist, api, multigpu, ctx, stream, dev, rank, world, run_id, barrier, max_over_ranks, gather_objects):
    """BASELINE.json configs[4]: HiFi-shaped reads (≈ 9,300 k-mers each → sort + unique, U:874-908) against the GTDB-scale index that
    gtdb_block left sharded over the ranks; the same reads at every N (strong scaling)"""
    info = ctx.db_info()
    steps = C5_WARMUP + C5_STEPS
    batches, offs = [], []
    if rank == 0:
        for s in range(steps):
            off, seq = hifi_batch(C5_SEED, s, C5_READS, GTDB_SEED, GTDB_GENOMES, GTDB_GL)
            offs.append(off)
            batches.append(np.concatenate([off.view(np.uint8), seq]))
        sizes = [int(b.size) for b in batches]
    else:
        sizes = None
    if world > 1:
        box = [sizes]
        dist.broadcast_object_list(box, src=0)
        sizes = box[0]
    bb = max(sizes)
    d_batches = None
    if rank == 0:
        d_batches = torch.zeros(steps * bb, dtype=torch.uint8, device="cuda")
        for s, b in enumerate(batches):
            d_batches[s * bb:s * bb + b.size].copy_(torch.from_numpy(b))
        torch.cuda.synchronize()
    sh = multigpu.ShardedSearch(ctx, rank, world, C5_READS, bb, hit_cap=max(1 << 20, 64 * C5_READS), name=run_id + "h", device=dev, dist=dist, params=ctx.default_params())
    digests = []

    def feed(first):
        return lambda s: (d_batches[(first + s) * bb:(first + s + 1) * bb], False)

    def consume(s, lists, meta):
        merged = multigpu.merge_lists(lists, 0, C5_READS)
        digests.append((len(merged), multigpu.hits_digest(merged)))

    with torch.cuda.stream(stream):
        sh.run(C5_WARMUP, feed(0), consume)              # offsets differ from step to step: the library fetches them from the device
        digests.clear()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        outs = sh.run(C5_STEPS, feed(C5_WARMUP), consume)
        ev1.record(stream)
        barrier()
        ms = max_over_ranks(ev0.elapsed_time(ev1))
    sh.close()
    mine = {"rank": rank, "probe_ms": sum(o.ms_probe for o in outs), "probe_bytes": sum(o.probe_row_bytes for o in outs), "prep_ms": sum(o.ms_hash for o in outs),
            "probe_launches": sum(o.probe_launches for o in outs), "sum_row_bytes": int(info.sum_row_bytes)}
    per_rank = gather_objects(mine)
    if rank != 0:
        return None
    peak, _src = measured_peak()
    for r in per_rank:
        r["probe_GBps"] = (r["probe_bytes"] / 1e9) / (r["probe_ms"] / 1e3) if r["probe_ms"] > 0 else 0.0
        r["probe_frac"] = r["probe_GBps"] / peak
    n_total = C5_READS * C5_STEPS
    bases = int(sum(int(o[-1]) for o in offs[C5_WARMUP:]))
    total_bytes = sum(r["probe_bytes"] for r in per_rank)
    return {
        "workload": "HiFi-shaped reads (log-normal, median 8.3 kb, 45 .. 45,000 bp, 0.5 %% substitutions; mean %.0f bp here) against the GTDB-scale index above, "
                    "%d reads per step, %d timed steps" % (bases / n_total, C5_READS, C5_STEPS),
        "scaling": "strong", "n_gpus": world, "job_reads_per_s": n_total / (ms / 1e3), "job_Mbases_per_s": bases / (ms / 1e3) / 1e6, "ms_per_step": ms / C5_STEPS,
        "algorithmic_bytes_per_read": int(total_bytes / n_total), "aggregate_probe_GBps_over_wall": (total_bytes / 1e9) / (ms / 1e3),
        "frac_of_roofline": (total_bytes / 1e9) / (ms / 1e3) / (peak * world),
        "hits_per_step": int(sum(c for c, _d in digests) / max(1, len(digests))), "hit_list_digest": "%016x" % (sum(d for _c, d in digests) & (2**64 - 1)),
        "per_rank": per_rank,
    }


def c3_block(torch, api, ctx):
    """BASELINE.json configs[2]: FracMinHash genome-vs-genome sketch search on one GPU — 1,000 seeded 4 Mb assemblies (scale 1000, k=21, h=3,
    fpr 0.001: the tutorial's setting), whole genomes as -g queries.  The index (tens of MB) lives in L2: the path is bound by kernel 1
    (hash every base once), so the figure is bases/s against "genome bytes read once"."""
    t0 = time.perf_counter()
    ctx.build_synth_db(GENOME_SEED, C3_GENOMES, C3_GL, k=K, n_chunks=1, overlap=0, num_hashes=3, fpr=0.001, block_size=0, scale=1000)
    build_s = time.perf_counter() - t0
    info = ctx.db_info()
    nq = C3_QUERIES
    d = torch.empty(nq * C3_GL, dtype=torch.uint8, device="cuda")
    ctx.synth_genomes(GENOME_SEED, 0, nq, C3_GL, d.data_ptr())
    off = np.arange(nq + 1, dtype=np.uint64) * np.uint64(C3_GL)
    d_off = torch.from_numpy(off.view(np.int64)).cuda()
    torch.cuda.synchronize()
    p = ctx.default_params(min_query_cov=0.5)
    outs = []
    for rep in range(4):
        t0 = time.perf_counter()
        o = ctx.wait(ctx.submit(d.data_ptr(), d_off.data_ptr(), nq, p, device=True, host_off_ptr=off.ctypes.data))
        if rep:
            outs.append((time.perf_counter() - t0, o))
    dt = min(t for t, _o in outs)
    o = outs[-1][1]
    # every query genome is in the database: it must find itself with all of its k-mers
    self_hits = int(np.sum(o.hits["count"] == o.n_kmers[o.hits["query"]]))
    peak, _src = measured_peak()
    return {
        "workload": "FracMinHash sketch search: %d seeded %.1f Mb assemblies (scale 1000, k=%d, h=3, fpr 0.001, %d blocks, index %.1f MB), %d whole genomes as queries, device resident"
                    % (C3_GENOMES, C3_GL / 1e6, K, info.n_blocks, info.resident_bytes / 1e6, nq),
        "genomes_per_s": nq / dt, "Gbases_per_s": nq * C3_GL / dt / 1e9, "call_ms": dt * 1e3, "db_build_s": round(build_s, 2),
        "sketch_kmers_per_query": int(np.mean(o.n_kmers)), "prep_ms": o.ms_hash, "probe_ms": o.ms_probe, "hits": int(len(o.hits)), "queries_matching_themselves_fully": self_hits,
        "roofline": {"bound": "hbm (genome bytes read once; the probe works out of L2)", "achieved_GBps": nq * C3_GL / dt / 1e9, "peak_GBps": peak,
                     "frac": nq * C3_GL / dt / 1e9 / peak, "note": "kernel 1 is ALU / shared-memory bound, not bandwidth bound: one base per k-mer position and ~40 integer operations for it"},
    }


if __name__ == "__main__":
    main()

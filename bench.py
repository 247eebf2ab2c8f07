#!/usr/bin/env python
"""bench.py — reads/s of the `kmcp search` hot path on B200 (BASELINE.json metric), one JSON line on rank 0.

Workload (N=1) = BASELINE.json configs[1]: synthetic 10k-chunk COBS index (1,000 seeded random genomes x
4 Mb, 10 chunks, k=21, h=1, fpr 0.3, ONE block of 10,000 targets ≈ 1.4 GB ≫ L2) resident in HBM, 150 bp
synthetic reads (80 % sampled from the genomes with 1 % substitutions, 20 % random).  A step = one batch of
READS_PER_STEP reads through the whole hot path (hash → locs → probe → hit sort → D2H).  Every step uses
different reads; the index is far larger than L2, so no explicit L2 flush is needed (stated in `config`).

  value   : reads/s, inputs resident in HBM when the timed region starts (kmcpg_search_batch_device).
  e2e     : reads/s through the host-facing engine call with PINNED HOST buffers: H2D of the reads,
            kernels, D2H of hits, host post-filter (tCov/FPR/sort) — the number to compare with the CPU arm.
  roofline: probe kernel only; achieved = algorithmic row bytes (n_kmers·h·Σ numRowBytes per read) of a
            launch ÷ its CUDA-event duration on the launching stream; peak = MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline: the oracle's restatement of the reference algorithm (64-row buffer, byte transpose,
            positional popcount) on all host threads, bounded sample — kind "port" (the Go reference cannot
            be built here: no Go toolchain).

N>1 (torchrun, one rank per GPU): index blocks shard across ranks (rank r holds block r, 10,000 targets;
the DB grows with N = weak scaling), the read batch is broadcast with NCCL, per-rank hit lists are
concatenated on the host of rank 0.  No data-path collective besides the broadcast.

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
else:  # quick functional check of the harness
    N_GENOMES, GENOME_LEN, READS_PER_STEP, CPU_SAMPLE0 = 100, 200_000, 100_000, 5_000
N_CHUNKS, OVERLAP, K, H, FPR, READ_LEN = 10, 150, 21, 1, 0.3, 150
BLOCK_SIZE = N_GENOMES * N_CHUNKS
GENOME_SEED, READ_SEED = 1, 2
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
    O.write_db_yml(os.path.join(r001, "__db.yml"), {
        "version": 4, "unikiVersion": 4, "alias": "bench", "k": K, "ks": [K], "hashed": True, "canonical": True, "scaled": False,
        "scale": 0, "minimizer": False, "minimizer-w": 0, "syncmer": False, "syncmer-s": 0, "split-seq": True, "split-size": 0,
        "split-num": N_CHUNKS, "split-overlap": OVERLAP, "compact-size": False, "hashes": H, "fpr": FPR,
        "numNameGroups": int(info.n_targets), "blocksize": BLOCK_SIZE, "totalKmers": 0, "files": files})
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


def stage_reference_dbs(tmp, world, n_reads_total):
    """the index of every shard of this arm's config (world blocks of 10,000 targets, seeds GENOME_SEED + r) as .uniki files the
    CPU port can open, plus the seeded reads: built by the GPU index builder (byte-identical to the oracle's builder,
    tests/test_gpu_parity.py) because the 1.4 GB blocks take minutes on the CPU.  Returns ([R001 dirs], reads u8)."""
    from kmcp_b200 import api
    ctx = api.Context(0)
    try:
        dirs = []
        for r in range(world):
            ctx.build_synth_db(GENOME_SEED + r, N_GENOMES, GENOME_LEN, k=K, n_chunks=N_CHUNKS, overlap=OVERLAP, num_hashes=H, fpr=FPR, block_size=BLOCK_SIZE)
            dirs.append(dump_db_for_cpu(ctx, os.path.join(tmp, "shard%d" % r)))
        d = ctx.device_alloc(n_reads_total * READ_LEN)
        ctx.synth_reads(READ_SEED, 0, n_reads_total, READ_LEN, GENOME_SEED, N_GENOMES, GENOME_LEN, d)
        reads = ctx.d2h(d, n_reads_total * READ_LEN)
        ctx.device_free(d)
    finally:
        ctx.close()
    return dirs, reads


def run_reference(args, rank, world, stage=stage_reference_dbs):
    """--impl reference: the CPU port of the reference algorithm, rank 0 only, on this arm's config: at N ranks the index is
    N blocks of 10,000 targets (one per GPU in the b200 arm), every read is searched against all of them, and `value`
    counts read x shard probes exactly as the b200 arm does."""
    if rank != 0:
        return
    from oracle import oracle as O
    O.build()
    tmp = "/dev/shm/kmcp_bench_ref" if os.path.isdir("/dev/shm") else "/tmp/kmcp_bench_ref"
    shutil.rmtree(tmp, ignore_errors=True)
    os.makedirs(tmp)
    built_by = "gpu index builder (byte-identical to the oracle builder, tests/test_gpu_parity.py)"
    step_reads = max(1000, (50_000 if SCALE == "full" else 5_000) // world)      # the CPU work per step stays the same at every N
    n_total = step_reads * (args.steps + args.warmup)
    try:
        dirs, reads = stage(tmp, world, n_total)
    except Exception as e:  # no usable GPU: nothing to build the 1.4 GB index with in reasonable time
        shutil.rmtree(tmp, ignore_errors=True)
        print(json.dumps({"impl": "reference", "unavailable": "cannot stage the synthetic index without the GPU builder: %s" % str(e)[:120]}))
        return
    odbs = [O.DB(d) for d in dirs]
    cores = os.cpu_count() or 1
    off = np.arange(step_reads + 1, dtype=np.uint64) * np.uint64(READ_LEN)
    n_hits = [0]

    def step(i):
        batch = reads[i * step_reads * READ_LEN:(i + 1) * step_reads * READ_LEN]
        for odb in odbs:          # explicit thread count: torchrun sets OMP_NUM_THREADS=1
            n_hits[0] += len(odb.search(packed=(batch, off), threads=cores, algo=1).hits)

    for i in range(args.warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(args.warmup + i)
    dt = time.perf_counter() - t0
    v = world * step_reads * args.steps / dt
    for odb in odbs:
        odb.close()
    shutil.rmtree(tmp, ignore_errors=True)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "reads/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64/u8 bitset",
        "data": "synthetic", "config": {"workload": workload_name(), "reads_per_step_cpu_sample": step_reads, "db_built_by": built_by,
                                        "blocks_searched": world,
                                        "multi_gpu_units": "value counts read×shard probes (every read against each of the %d 10k-target blocks)" % world if world > 1 else "reads"},
        "job_reads_per_s": step_reads * args.steps / dt,
        "cpu_baseline": {"value": v, "unit": "reads/s", "cores": cores, "kind": "port",
                         "sample": "%d reads per step against %d block(s), restated reference algorithm (oracle algo=1), OpenMP all threads" % (step_reads, world)},
        "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            del os.environ["NCCL_DEBUG"]             # keeps NCCL's version banner out of stdout: one JSON line only
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    stream = torch.cuda.Stream()
    ctx = api.Context(local_rank)
    ctx.set_stream(stream.cuda_stream)

    t_build = time.perf_counter()
    ctx.build_synth_db(GENOME_SEED + rank, N_GENOMES, GENOME_LEN, k=K, n_chunks=N_CHUNKS, overlap=OVERLAP, num_hashes=H, fpr=FPR, block_size=BLOCK_SIZE)
    t_build = time.perf_counter() - t_build
    info = ctx.db_info()

    n_steps_total = args.warmup + args.steps
    step_bytes = READS_PER_STEP * READ_LEN
    # reads of every step resident in HBM before any timing; same seeds on every rank = the broadcast batch
    d_reads = torch.empty(n_steps_total * step_bytes, dtype=torch.uint8, device="cuda")
    for s in range(n_steps_total):
        ctx.synth_reads(READ_SEED, s * READS_PER_STEP, READS_PER_STEP, READ_LEN, GENOME_SEED, N_GENOMES, GENOME_LEN, d_reads.data_ptr() + s * step_bytes)
    off_np = np.arange(READS_PER_STEP + 1, dtype=np.uint64) * np.uint64(READ_LEN)
    d_off = torch.from_numpy(off_np.view(np.int64)).cuda()
    params = ctx.default_params()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- value: device-resident inputs ----------------
    def dev_step(s):
        return ctx.search_batch_ptr(d_reads.data_ptr() + s * step_bytes, d_off.data_ptr(), READS_PER_STEP, params, device=True, seq_bytes=step_bytes, copy=False)

    with torch.cuda.stream(stream):
        for s in range(args.warmup):
            dev_step(s)
        barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        outs = [dev_step(args.warmup + s) for s in range(args.steps)]
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

    # ---------------- e2e: pinned host buffers through the engine ----------------
    h_reads, h_ptr = api.pinned_array(args.steps * step_bytes)
    h_reads[:] = d_reads[args.warmup * step_bytes:].cpu().numpy()
    h_off, h_off_ptr = api.pinned_array(off_np.nbytes)
    h_off[:] = off_np.view(np.uint8)
    eopts = ctx.default_engine_opts()
    d2h_bytes = 0
    e2e_matches = 0
    e2e_break = {"search_call_ms": 0.0, "post_filter_ms": 0.0, "engine_call_ms": 0.0}
    with torch.cuda.stream(stream):
        if world == 1:
            ctx.engine_search_ptr(h_ptr, h_off_ptr, READS_PER_STEP, eopts)          # warm the host-side caches once
            barrier()
            t0 = time.perf_counter()
            for s in range(args.steps):
                r = ctx.engine_search_ptr(h_ptr + s * step_bytes, h_off_ptr, READS_PER_STEP, eopts, copy=False)
                e2e_matches += r.n_matches
                e2e_break["search_call_ms"] += r.ms_gpu_total / args.steps; e2e_break["post_filter_ms"] += r.ms_post / args.steps
                e2e_break["engine_call_ms"] += r.ms_total / args.steps
            torch.cuda.synchronize()
            e2e_s = time.perf_counter() - t0
            d2h_bytes = int(12 * n_hits / args.steps + 8 * READS_PER_STEP)
        else:
            # two staging buffers: step s+1 is copied to rank 0's GPU and broadcast (side stream) while step s is searched
            stage = [torch.empty(step_bytes, dtype=torch.uint8, device="cuda") for _ in range(2)]
            side = torch.cuda.Stream()
            staged = [torch.cuda.Event(), torch.cuda.Event()]
            dev = torch.device("cuda", local_rank)

            def prefetch(s):
                with torch.cuda.stream(side):
                    if rank == 0:
                        stage[s & 1].copy_(torch.from_numpy(h_reads[s * step_bytes:(s + 1) * step_bytes]), non_blocking=True)
                    dist.broadcast(stage[s & 1], src=0)                                # NCCL over NVLink: the only data-path collective
                    staged[s & 1].record(side)

            # untimed: the first gather builds NCCL's point-to-point channels
            multigpu.gather_hits_padded(np.zeros(1024, dtype=api.HIT_DTYPE), rank, world, dev)
            multigpu.gather_hits_padded(np.zeros(900_000, dtype=api.HIT_DTYPE), rank, world, dev)
            barrier()
            t0 = time.perf_counter()
            prefetch(0)
            for s in range(args.steps):
                staged[s & 1].synchronize()
                if s + 1 < args.steps:
                    prefetch(s + 1)
                o = ctx.search_batch_ptr(stage[s & 1].data_ptr(), d_off.data_ptr(), READS_PER_STEP, params, device=True, seq_bytes=step_bytes)
                # hit lists (disjoint by target) → rank 0: one padded NCCL gather, global target numbering added on the way
                merged = multigpu.gather_hits_padded(o.hits, rank, world, dev, target_base=rank * BLOCK_SIZE)
                if rank == 0:
                    e2e_matches += len(merged)
            barrier()
            e2e_s = time.perf_counter() - t0
            t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
            d2h_bytes = int(12 * n_hits / args.steps + 8 * READS_PER_STEP)
    e2e_value = world * READS_PER_STEP * args.steps / e2e_s

    # ---------------- CPU baseline beside it (rank 0, N=1 only) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        tmp = "/dev/shm/kmcp_bench_cpu" if os.path.isdir("/dev/shm") else "/tmp/kmcp_bench_cpu"
        shutil.rmtree(tmp, ignore_errors=True)
        try:
            r001 = dump_db_for_cpu(ctx, tmp)
            v, cores, n1 = time_cpu_port(r001, h_reads, args.steps * READS_PER_STEP)
            cpu = {"value": v, "unit": "reads/s", "cores": cores, "kind": "port",
                   "sample": "%d reads of the timed set, restated reference algorithm (oracle algo=1: 64-row buffer, byte transpose, "
                             "positional popcount), OpenMP on all %d host threads, same index in RAM" % (n1, cores)}
        finally:
            shutil.rmtree(tmp, ignore_errors=True)

    api.host_free(h_ptr); api.host_free(h_off_ptr)
    if rank == 0:
        peak, peak_src = measured_peak()
        achieved = (probe_bytes / 1e9) / (probe_ms / 1e3) if probe_ms > 0 else 0.0
        traffic = ncu_traffic()
        line = {
            "metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64 hash / u32 bit-sliced counters", "data": "synthetic",
            "config": {"workload": workload_name(), "reads_per_step": READS_PER_STEP, "index_bytes_per_gpu": int(info.resident_bytes),
                       "index_disk_bytes_per_gpu": int(info.disk_bytes), "targets_per_gpu": int(info.n_targets),
                       "algorithmic_bytes_per_read": int(probe_bytes / max(1, READS_PER_STEP * args.steps)),
                       "l2": "index (%.2f GB) ≫ 126 MB L2 and every step probes different random rows: no explicit flush" % (info.resident_bytes / 1e9),
                       "parallelism": "blocks sharded over %d GPU(s), read batch broadcast" % world, "db_build_s": round(t_build, 2),
                       "hits_per_step": int(n_hits / args.steps),
                       "multi_gpu_units": "value counts read×shard probes (each rank probes every read against its own 10k-target block)" if world > 1 else "reads"},
            "job_reads_per_s": READS_PER_STEP * args.steps / (ms_total / 1e3),
            "roofline": {"bound": "hbm", "kernel": "probe_kernel<1,8>", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_src, "launches": probe_launches, "avg_launch_ms": probe_ms / max(1, probe_launches),
                         "algorithmic_bytes_per_launch": probe_bytes / max(1, probe_launches),
                         "traffic": (traffic["dram_over_algorithmic"] * probe_bytes / max(1, probe_launches)) if traffic else None,
                         "traffic_note": ("average launch of this run x the DRAM/algorithmic ratio of the committed ncu --set full capture: " + traffic["source"])
                         if traffic else "no ncu --set full capture committed yet"},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "reads/s", "h2d_bytes_per_step": step_bytes + off_np.nbytes, "d2h_bytes_per_step": d2h_bytes,
                    "matches_per_step": int(e2e_matches / args.steps), "ms_per_step": e2e_s / args.steps * 1e3, "breakdown_ms_per_step": e2e_break,
                    "path": "kmcpg_engine_search (pinned host reads → H2D → kernels → D2H hits → host tCov/FPR/sort)" if world == 1 else
                            "rank0 pinned reads → H2D → ncclBroadcast (prefetched one step ahead) → kmcpg_search_batch_device on every rank → padded NCCL gather of the hit lists → rank 0 host"},
            "gpu_launches": int(launches), "clocks": clocks,
            "stage_ms_per_step": {"hash": sum(o.ms_hash for o in outs) / args.steps, "locs": sum(o.ms_locs for o in outs) / args.steps,
                                  "probe": probe_ms / args.steps, "call_wall": sum(o.ms_total for o in outs) / args.steps},
        }
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Host-side plumbing of the multi-GPU search with ONE PROCESS PER GPU (torchrun / torch.distributed).

The path shards by index BLOCK (the reference's block fan-out + gather, util-db-search.go:939-964, and across
processes `kmcp merge`, merge.go:190-256): every rank holds the blocks kmcpg_open_db(shard_rank, shard_world) gives it, the read
batch is broadcast — NCCL over NVLink, the only collective on the data path —, every rank probes the whole batch against its
blocks, and the per-rank hit lists — disjoint by target — are concatenated on the host of rank 0.

The hit lists do not travel through a collective: every rank's library copies its list device→host straight into a shared-memory
segment of its own (`HitExchange`, kmcpg_shm_open + kmcpg_batch.hits_dst; each GPU uses its own PCIe link) and rank 0 reads the
segments in place.  `ShardedSearch` runs the per-step pipeline: the broadcast of step s+1 and the hit return of step s-1 both
happen while the GPUs probe step s (two jobs are kept submitted, kmcpg_search_submit).
"""
from __future__ import annotations

import ctypes as C
import threading
import time
from typing import Callable, List, Optional, Sequence

import numpy as np

from . import api

_HDR = 64          # bytes per header (one cache line)


class HitExchange:
    """Per-rank hit mailboxes in shared host memory.

    Segment of rank r:  [ctrl 64 B | slot 0: header 64 B + cap hits | slot 1: ...].  A rank writes the hits of step s into slot
    s % slots (the library's device→host copy lands there), then the header {n_hits, step + 1}; rank 0 polls the headers of all
    ranks, reads the lists in place and finally stores the number of consumed steps in the ctrl line of ITS segment, which the
    other ranks poll before they reuse a slot.  x86 keeps the store order; the hit bytes are complete before the header is
    written because kmcpg_search_wait returns only after the copy's CUDA event."""

    def __init__(self, name: str, rank: int, world: int, cap_hits: int, barrier: Callable[[], None], slots: int = 2, cuda_register: bool = True):
        self.name, self.rank, self.world, self.cap, self.slots, self.reg = name, rank, world, int(cap_hits), slots, bool(cuda_register)
        self.slot_bytes = _HDR + self.cap * 12
        self.bytes = _HDR + slots * self.slot_bytes
        self._L = api.load()
        self._ptr = {}
        self._open(rank, create=True)
        self._view(rank)[:] = 0
        barrier()
        for r in (range(world) if rank == 0 else (0,)):
            if r != rank:
                self._open(r, create=False)
        barrier()
        self.step = 0              # steps handed out so far (every rank counts in lockstep)

    def _seg(self, r):
        return ("%s.%d" % (self.name, r)).encode()

    def _open(self, r, create):
        p = C.c_void_p()
        rc = self._L.kmcpg_shm_open(self._seg(r), self.bytes, 1 if create else 0, 1 if (create and self.reg) else 0, C.byref(p))
        if rc:
            raise api.KmcpGpuError(rc, (self._L.kmcpg_last_error(None) or b"").decode())
        self._ptr[r] = p.value

    def _view(self, r, off=0, n=None, dtype=np.uint8):
        n = self.bytes - off if n is None else n
        return np.frombuffer((C.c_uint8 * n).from_address(self._ptr[r] + off), dtype=dtype)

    def _hdr(self, r, step):
        return self._view(r, _HDR + (step % self.slots) * self.slot_bytes, 16, np.uint64)        # [n_hits, step + 1]

    # ---- every rank
    def hits_ptr(self, step: int) -> int:
        return self._ptr[self.rank] + _HDR + (step % self.slots) * self.slot_bytes + _HDR

    def wait_free(self, step: int, timeout: float = 120.0):
        """the slot of `step` was last used by step - slots: rank 0 must have consumed that one"""
        need = step - self.slots + 1
        if need <= 0:
            return
        done = self._view(0, 0, 8, np.uint64)
        t0 = time.perf_counter()
        while int(done[0]) < need:
            if time.perf_counter() - t0 > timeout:
                raise TimeoutError("rank %d: hit slot of step %d never freed" % (self.rank, step))
            time.sleep(20e-6)

    def publish(self, step: int, n_hits: int):
        h = self._hdr(self.rank, step)
        h[0] = n_hits
        h[1] = step + 1

    # ---- rank 0
    def collect(self, step: int, timeout: float = 120.0) -> List[np.ndarray]:
        """views (api.HIT_DTYPE) of every rank's hit list of `step`, in rank order; valid until release(step)"""
        out = []
        t0 = time.perf_counter()
        for r in range(self.world):
            h = self._hdr(r, step)
            while int(h[1]) != step + 1:
                if time.perf_counter() - t0 > timeout:
                    raise TimeoutError("rank 0: no hit list of rank %d for step %d" % (r, step))
                time.sleep(20e-6)
            n = int(h[0])
            base = _HDR + (step % self.slots) * self.slot_bytes + _HDR
            out.append(self._view(r, base, n * 12).view(api.HIT_DTYPE) if n else np.zeros(0, api.HIT_DTYPE))
        return out

    def release(self, step: int):
        self._view(0, 0, 8, np.uint64)[0] = step + 1

    def close(self):
        for r, p in list(self._ptr.items()):
            self._L.kmcpg_shm_close(self._seg(r), p, self.bytes, 1 if (r == self.rank and self.reg) else 0, 1 if r == self.rank else 0)
        self._ptr = {}


def merge_lists(lists: Sequence[np.ndarray], first_query: int, n_queries: int, threads: int = 0, out: Optional[np.ndarray] = None) -> np.ndarray:
    """kmcpg_merge_hits: the union of per-rank hit lists (disjoint by target, each sorted) in (query, target) order"""
    L = api.load()
    total = sum(len(a) for a in lists)
    if out is None or len(out) < total:
        out = np.empty(max(total, 1), dtype=api.HIT_DTYPE)
    ptrs = (C.c_void_p * max(1, len(lists)))(*[a.ctypes.data if len(a) else None for a in lists])
    ns = (C.c_uint64 * max(1, len(lists)))(*[len(a) for a in lists])
    rc = L.kmcpg_merge_hits(ptrs, ns, len(lists), first_query, n_queries, threads, out.ctypes.data)
    if rc:
        raise api.KmcpGpuError(rc, "kmcpg_merge_hits")
    return out[:total]


def hits_digest(hits: np.ndarray, first_index: int = 0) -> int:
    """kmcpg_hits_digest: order-sensitive 64-bit digest of a hit list"""
    if not len(hits):
        return 0
    return int(api.load().kmcpg_hits_digest(hits.ctypes.data, len(hits), first_index))


def postfilter(opts, n_kmers: np.ndarray, query_len: np.ndarray, hits: np.ndarray, target_sizes: np.ndarray, fpr: float, k: int, copy: bool = True):
    """kmcpg_engine_postfilter: the engine's tCov / FPR / sort / top-N over a merged hit list (host only)"""
    L = api.load()
    r = api.Results()
    nk = np.ascontiguousarray(n_kmers, dtype=np.int32)
    ql = np.ascontiguousarray(query_len, dtype=np.int32)
    ts = np.ascontiguousarray(target_sizes, dtype=np.float64)
    rc = L.kmcpg_engine_postfilter(C.byref(opts), len(nk), nk.ctypes.data, ql.ctypes.data, hits.ctypes.data if len(hits) else None, len(hits),
                                   ts.ctypes.data, len(ts), fpr, k, 0, C.byref(r))
    if rc:
        raise api.KmcpGpuError(rc, "kmcpg_engine_postfilter")
    nq = r.n_queries
    if copy:
        out = api.EngineResults(api._np_from(r.query_len, nq, 4, np.int32), api._np_from(r.n_kmers, nq, 4, np.int32), api._np_from(r.k_used, nq, 4, np.int32),
                                api._np_from(r.match_off, nq + 1, 8, np.uint64), api._np_from(r.matches, r.n_matches, C.sizeof(api.Match), api.MATCH_DTYPE),
                                0.0, 0, 0)
    else:
        out = api.EngineResults(np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.uint64), np.zeros(0, api.MATCH_DTYPE), 0.0, 0, 0)
    out.n_matches = int(r.n_matches); out.ms_post = r.ms_post; out.ms_total = r.ms_total
    L.kmcpg_free_results(C.byref(r))
    return out


class ShardedSearch:
    """The per-step pipeline of one rank.  Rank 0 owns the batches; all ranks call run() with the same number of steps.

        feed(s)      rank 0 only: returns (tensor, is_host) — the packed batch of step s, `batch_bytes` uint8: the (n_seqs + 1) u64
                     offsets followed by the sequence bytes — in pinned host memory (copied to the GPU first) or already on rank 0's GPU
        consume(s, lists, meta)   rank 0 only, on a helper thread, while the GPUs work on the next steps: lists = per-rank hit lists
                     (views into the shared segments, valid during the call), meta = rank 0's BatchHits of step s (n_kmers, query_len)
    """

    def __init__(self, ctx: api.Context, rank: int, world: int, n_seqs: int, batch_bytes: int, hit_cap: int, name: str, device, dist, params=None):
        import torch
        self.torch, self.dist = torch, dist
        self.ctx, self.rank, self.world, self.n_seqs, self.batch_bytes = ctx, rank, world, n_seqs, batch_bytes
        self.params = params or ctx.default_params()
        self.off_bytes = (n_seqs + 1) * 8
        self.stage = [torch.empty(batch_bytes, dtype=torch.uint8, device=device) for _ in range(2)]
        self.comm = torch.cuda.Stream(device=device, priority=-1)      # ahead of the queued probe CTAs
        self.staged = [torch.cuda.Event(), torch.cuda.Event()]
        self.hx = HitExchange(name, rank, world, hit_cap, barrier=(lambda: dist.barrier()) if world > 1 else (lambda: None), cuda_register=True)
        self.base = 0          # exchange steps used by earlier run() calls

    def _prefetch(self, s, feed):
        torch = self.torch
        buf = self.stage[s & 1]
        with torch.cuda.stream(self.comm):
            if self.rank == 0:
                src, is_host = feed(s)
                buf.copy_(src, non_blocking=True)                      # pinned host → GPU (e2e), or a device slice (inputs resident)
            if self.world > 1:
                self.dist.broadcast(buf, src=0)                        # NCCL over NVLink: the only collective on the data path
            self.staged[s & 1].record(self.comm)

    def run(self, steps: int, feed, consume=None, host_off: Optional[np.ndarray] = None):
        """returns the list of per-step BatchHits summaries of this rank"""
        ctx, hx = self.ctx, self.hx
        outs = [None] * steps
        cv = threading.Condition()
        err: List[BaseException] = []
        stop = [False]
        metas = {}

        def consumer():
            try:
                for s in range(steps):
                    with cv:
                        while s not in metas and not err and not stop[0]:
                            cv.wait(0.05)
                        if err or s not in metas:
                            return
                        meta = metas.pop(s)
                    lists = hx.collect(self.base + s)
                    if consume is not None:
                        consume(s, lists, meta)
                    hx.release(self.base + s)
            except BaseException as e:      # noqa: BLE001 - re-raised on the main thread
                err.append(e)

        th = None
        if self.rank == 0:
            th = threading.Thread(target=consumer, daemon=True)
            th.start()
        jobs = {}
        hoff_ptr = host_off.ctypes.data if host_off is not None else 0

        def finish(s):
            r = ctx.wait(jobs.pop(s), copy="meta" if self.rank == 0 and consume is not None else False)
            hx.publish(self.base + s, r.n_hits)
            outs[s] = r
            if self.rank == 0:
                with cv:
                    metas[s] = r
                    cv.notify_all()

        try:
            self._prefetch(0, feed)
            for s in range(steps):
                if err:
                    raise err[0]
                hx.wait_free(self.base + s)
                buf = self.stage[s & 1]
                jobs[s] = ctx.submit(buf.data_ptr() + self.off_bytes, buf.data_ptr(), self.n_seqs, self.params, device=True, host_off_ptr=hoff_ptr,
                                     hits_dst=hx.hits_ptr(self.base + s), hits_cap=hx.cap, ready_event=self.staged[s & 1].cuda_event)
                if s >= 1:
                    finish(s - 1)
                if s + 1 < steps:
                    self._prefetch(s + 1, feed)
            if steps:
                finish(steps - 1)
        except BaseException:
            stop[0] = True
            raise
        finally:
            for j in list(jobs.values()):          # an error above: do not leave jobs behind
                try:
                    ctx.wait(j, copy=False)
                except Exception:
                    pass
            if th is not None:
                th.join()
        if err:
            raise err[0]
        self.base += steps
        return outs

    def close(self):
        self.hx.close()

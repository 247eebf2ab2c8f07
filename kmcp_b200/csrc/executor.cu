// executor.cu — the batch executor of libkmcp_gpu: query preparation, the probe launches and the result return of
// kmcpg_search_batch* / kmcpg_search_submit (reference: UnikIndexDB.handleQuery U:763-941 + every UnikIndex worker U:6613-7741).
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cub/cub.cuh>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "ctx_internal.h"
#include "nvtx_ranges.h"

using namespace kmcpg;

namespace kmcpg {

int planes_for(uint64_t max_n) {
    if (max_n <= 255) return 8;
    if (max_n <= 65535) return 16;
    if (max_n < (1ull << 24)) return 24;
    return 32;
}

struct Timing { float ms_hash = 0, ms_locs = 0, ms_probe = 0; uint64_t probe_bytes = 0; uint32_t probe_launches = 0; };

static int ensure_events(kmcpg_ctx *ctx, WorkSet &w) {
    for (cudaEvent_t *e : {&w.ev_in, &w.ev_a0, &w.ev_hash, &w.ev_a, &w.ev_cnt, &w.ev_sorted, &w.ev_b})
        if (!*e) CU(cudaEventCreate(e));
    while (w.probe_ev.size() < ctx->blocks.size() * 3 + 3) {
        cudaEvent_t e;
        CU(cudaEventCreate(&e));
        w.probe_ev.push_back(e);
    }
    CU(w.h_cnt.ensure(64));
    CU(w.counters.ensure(64 + ctx->blocks.size() * 8));       // [0] hit count, [1] Σ n_kmers, [2 + b] task counter of block b
    return KMCPG_OK;
}

// widens a u32 count to u64 while scanning
struct U32ToU64 { __host__ __device__ uint64_t operator()(uint32_t v) const { return (uint64_t)v; } };

// hashing of one HashArgs job: warp per query for short sequences, warp per 4096-position tile (+ gather) when a query is long
static int hash_any(kmcpg_ctx *ctx, WorkSet &w, const HashArgs &ha, uint32_t n_seqs, uint64_t total_slots, uint64_t max_query_slots, cudaStream_t st) {
    if (!ha.raw && !ha.scaled && max_query_slots <= HASH_GROUP_MAX_KMERS) {        // short reads, every k-mer kept: eight lanes per query
        CU(launch_hash_groups(ha, st)); ctx->launches++;
        return KMCPG_OK;
    }
    if (max_query_slots <= 2ull * HASH_TILE_POS) {
        CU(launch_hash(ha, st)); ctx->launches++;
        return KMCPG_OK;
    }
    const uint64_t max_tiles = total_slots / HASH_TILE_POS + n_seqs;
    CU(w.tile_n.ensure((n_seqs + 1) * 8ull)); CU(w.tile_off.ensure((n_seqs + 1) * 8ull));
    CU(launch_tiles_per_seq(ha.seq_off, n_seqs, ha.k, w.tile_n.as<uint64_t>(), st));
    size_t t1 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, t1, w.tile_n.as<uint64_t>(), w.tile_off.as<uint64_t>(), (int)(n_seqs + 1), st);
    CU(w.tmp.ensure(t1));
    CU(cub::DeviceScan::ExclusiveSum(w.tmp.p, t1, w.tile_n.as<uint64_t>(), w.tile_off.as<uint64_t>(), (int)(n_seqs + 1), st));
    ctx->launches += 3;
    if (ha.raw) {                                    // position-indexed output: the tiles write straight to their place
        CU(launch_hash_tiles(ha, n_seqs, w.tile_off.as<uint64_t>(), max_tiles, nullptr, nullptr, st)); ctx->launches++;
        return KMCPG_OK;
    }
    if (max_tiles >= (1ull << 31)) return fail(ctx, KMCPG_EUNSUPPORTED, "too many tiles in one part");
    CU(w.tile_cnt.ensure((max_tiles + 1) * 4)); CU(w.tile_pre.ensure((max_tiles + 1) * 8));
    CU(w.codes2.ensure(std::max<uint64_t>(total_slots, 1) * 8));
    CU(cudaMemsetAsync(w.tile_cnt.p, 0, (max_tiles + 1) * 4, st));
    CU(launch_hash_tiles(ha, n_seqs, w.tile_off.as<uint64_t>(), max_tiles, w.codes2.as<uint64_t>(), w.tile_cnt.as<uint32_t>(), st));
    cub::TransformInputIterator<uint64_t, U32ToU64, const uint32_t *> it(w.tile_cnt.as<uint32_t>(), U32ToU64());
    size_t t2 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, t2, it, w.tile_pre.as<uint64_t>(), (int)(max_tiles + 1), st);
    CU(w.tmp.ensure(t2));
    CU(cub::DeviceScan::ExclusiveSum(w.tmp.p, t2, it, w.tile_pre.as<uint64_t>(), (int)(max_tiles + 1), st));
    CU(launch_gather_tiles(ha, n_seqs, w.tile_off.as<uint64_t>(), w.tile_pre.as<uint64_t>(), w.tile_cnt.as<uint32_t>(), w.codes2.as<uint64_t>(), max_tiles, st));
    ctx->launches += 4;
    return KMCPG_OK;
}

// slot scan → hash → (sort+unique) → verdict, all on stream `st` (nullptr: the compute stream); sb.d_seq/d_off must already be valid there
int run_hash_stage(kmcpg_ctx *ctx, WorkSet &w, const kmcpg_search_params &p, int k, const SubBatch &sb, uint32_t nq, uint64_t **codes_out, cudaStream_t st) {
    const DbMeta &m = ctx->meta;
    if (!st) st = ctx->st;
    int rc = ensure_events(ctx, w);
    if (rc) return rc;
    CU(w.slot_cnt.ensure((sb.n_seqs + 1) * 8ull));
    CU(w.slot_off.ensure((sb.n_seqs + 1) * 8ull));
    CU(launch_slot_bounds(sb.d_off, sb.n_seqs, k, w.slot_cnt.as<uint64_t>(), st)); ctx->launches++;
    size_t tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, w.slot_cnt.as<uint64_t>(), w.slot_off.as<uint64_t>(), (int)(sb.n_seqs + 1), st);
    CU(w.tmp.ensure(tmp));
    CU(cub::DeviceScan::ExclusiveSum(w.tmp.p, tmp, w.slot_cnt.as<uint64_t>(), w.slot_off.as<uint64_t>(), (int)(sb.n_seqs + 1), st));
    ctx->launches += 2;
    CU(w.codes.ensure(std::max<uint64_t>(sb.total_slots, 1) * 8));
    for (DevBuf *b : {&w.ncodes, &w.qlen, &w.nk, &w.neff, &w.thresh}) CU(b->ensure(std::max<uint32_t>(nq, 1) * 4ull));

    HashArgs ha;
    memset(&ha, 0, sizeof(ha));
    ha.seq = sb.d_seq; ha.seq_off = sb.d_off; ha.slot_off = w.slot_off.as<uint64_t>();
    ha.codes = w.codes.as<uint64_t>(); ha.n_codes = w.ncodes.as<uint32_t>(); ha.query_len = w.qlen.as<int32_t>();
    ha.n_queries = nq; ha.paired = p.paired; ha.mate_select = p.mate_select; ha.k = k; ha.canonical = m.canonical;
    ha.scaled = m.scaled;
    ha.max_hash = ~0ull;
    if (m.scaled) {                                   // U:1040-1043: uint64(float64(^uint64(0)) / float64(scale))
        double v = 18446744073709551616.0 / (double)m.scale;
        ha.max_hash = v >= 18446744073709551616.0 ? ~0ull : (uint64_t)v;
    }
    ha.minimizer = m.minimizer; ha.minimizer_w = m.minimizer_w; ha.syncmer = m.syncmer; ha.syncmer_s = m.syncmer_s;
    ha.min_query_len = p.min_query_len;
    if (!m.minimizer && !m.syncmer) {
        rc = hash_any(ctx, w, ha, sb.n_seqs, sb.total_slots, sb.max_query_slots, st);
        if (rc) return rc;
    } else {
        // sketch databases: hash every position (k-mers, and s-mers for syncmers), then select per window
        CU(w.ck.ensure(std::max<uint64_t>(sb.total_slots, 1) * 8));
        HashArgs hr = ha;
        hr.raw = 1; hr.n_queries = sb.n_seqs; hr.codes = w.ck.as<uint64_t>();
        rc = hash_any(ctx, w, hr, sb.n_seqs, sb.total_slots, sb.max_query_slots, st);
        if (rc) return rc;
        SelectArgs sa;
        memset(&sa, 0, sizeof(sa));
        sa.seq_off = sb.d_off; sa.ck = w.ck.as<uint64_t>(); sa.slot_off = w.slot_off.as<uint64_t>();
        if (m.syncmer) {
            const int s = (int)m.syncmer_s;
            if (s < 1 || s >= k) return fail(ctx, KMCPG_EFORMAT, "syncmer-s must be in 1..k-1");
            CU(w.cs_cnt.ensure((sb.n_seqs + 1) * 8ull)); CU(w.cs_off.ensure((sb.n_seqs + 1) * 8ull));
            CU(launch_slot_bounds(sb.d_off, sb.n_seqs, s, w.cs_cnt.as<uint64_t>(), st));
            size_t t1 = 0;
            cub::DeviceScan::ExclusiveSum(nullptr, t1, w.cs_cnt.as<uint64_t>(), w.cs_off.as<uint64_t>(), (int)(sb.n_seqs + 1), st);
            CU(w.tmp.ensure(t1));
            CU(cub::DeviceScan::ExclusiveSum(w.tmp.p, t1, w.cs_cnt.as<uint64_t>(), w.cs_off.as<uint64_t>(), (int)(sb.n_seqs + 1), st));
            // every sequence has at most k-s more s-mers than k-mers (plus the ones shorter than k)
            CU(w.cs.ensure((sb.total_slots + (uint64_t)sb.n_seqs * (uint64_t)(k - s + 1) + 1) * 8));
            HashArgs hs = hr;
            hs.k = s; hs.slot_off = w.cs_off.as<uint64_t>(); hs.codes = w.cs.as<uint64_t>();
            rc = hash_any(ctx, w, hs, sb.n_seqs, sb.total_slots + (uint64_t)sb.n_seqs * (uint64_t)(k - s + 1), sb.max_query_slots + 2ull * (k - s), st);
            if (rc) return rc;
            ctx->launches += 3;
            sa.cs = w.cs.as<uint64_t>(); sa.cs_off = w.cs_off.as<uint64_t>(); sa.syncmer_s = s;
        }
        sa.codes = w.codes.as<uint64_t>(); sa.n_codes = w.ncodes.as<uint32_t>(); sa.query_len = w.qlen.as<int32_t>();
        sa.n_queries = nq; sa.paired = p.paired; sa.mate_select = p.mate_select; sa.k = k; sa.minimizer_w = m.minimizer_w;
        sa.scaled = m.scaled; sa.max_hash = ha.max_hash; sa.min_query_len = p.min_query_len;
        CU(launch_select(sa, st)); ctx->launches++;
    }

    uint64_t *codes = w.codes.as<uint64_t>();
    int do_unique = 0;
    if (sb.max_query_slots > (uint64_t)p.dedup_threshold && sb.total_slots > 0) {
        // U:874-908: sort + unique of queries with more than dedup_threshold k-mers.
        // up to SMALL_DEDUP_MAX k-mers: inside one warp, in place; longer queries: CUB segmented sort
        CU(launch_small_dedup(w.codes.as<uint64_t>(), w.slot_off.as<uint64_t>(), w.ncodes.as<uint32_t>(), nq, p.paired, p.dedup_threshold,
                              p.min_matched, sb.max_query_slots, st));
        ctx->launches++;
        if (sb.max_query_slots > (uint64_t)SMALL_DEDUP_MAX) {
            if (sb.total_slots >= (1ull << 31)) return fail(ctx, KMCPG_EUNSUPPORTED, "sub-batch too large for the dedup sort");
            CU(w.segb.ensure(nq * 4ull)); CU(w.sege.ensure(nq * 4ull));
            CU(w.codes2.ensure(sb.total_slots * 8));
            CU(launch_sort_segments(w.slot_off.as<uint64_t>(), w.ncodes.as<uint32_t>(), nq, p.paired, std::max(p.dedup_threshold, SMALL_DEDUP_MAX),
                                    w.segb.as<int>(), w.sege.as<int>(), st));
            CU(cudaMemcpyAsync(w.codes2.p, w.codes.p, sb.total_slots * 8, cudaMemcpyDeviceToDevice, st));
            size_t t2 = 0;
            cub::DeviceSegmentedSort::SortKeys(nullptr, t2, w.codes.as<uint64_t>(), w.codes2.as<uint64_t>(), (int)sb.total_slots, (int)nq,
                                               w.segb.as<int>(), w.sege.as<int>(), st);
            CU(w.tmp.ensure(t2));
            CU(cub::DeviceSegmentedSort::SortKeys(w.tmp.p, t2, w.codes.as<uint64_t>(), w.codes2.as<uint64_t>(), (int)sb.total_slots, (int)nq,
                                                  w.segb.as<int>(), w.sege.as<int>(), st));
            ctx->launches += 4;
            codes = w.codes2.as<uint64_t>();
        }
        do_unique = 1;
    }
    CU(cudaMemsetAsync(w.counters.p, 0, 16, st));     // [0] hit count, [1] Σ n_kmers
    FinalizeArgs fa;
    fa.codes = codes; fa.slot_off = w.slot_off.as<uint64_t>(); fa.n_codes = w.ncodes.as<uint32_t>();
    fa.n_kmers_out = w.nk.as<int32_t>(); fa.n_eff = w.neff.as<uint32_t>(); fa.thresh = w.thresh.as<uint32_t>();
    fa.n_sum = w.counters.as<unsigned long long>() + 1;
    fa.n_queries = nq; fa.paired = p.paired; fa.dedup_threshold = p.dedup_threshold; fa.do_unique = do_unique;
    fa.min_matched = p.min_matched; fa.min_query_cov = p.min_query_cov;
    CU(launch_finalize(fa, st)); ctx->launches++;
    *codes_out = codes;
    w.order_valid = false;
    if (sb.max_query_slots > 255 && nq > 1) {
        // long queries differ a lot in length (HiFi reads: 45 bp .. 45 kb): the probe's warps draw their tasks from a counter, longest
        // queries first, so that the draws left for the end of a launch are short ones
        CU(w.order.ensure(nq * 4ull)); CU(w.order_in.ensure(nq * 4ull)); CU(w.neff_sorted.ensure(nq * 4ull));
        CU(launch_iota(w.order_in.as<uint32_t>(), nq, st));
        size_t t4 = 0;
        cub::DeviceRadixSort::SortPairsDescending(nullptr, t4, w.neff.as<uint32_t>(), w.neff_sorted.as<uint32_t>(), w.order_in.as<uint32_t>(), w.order.as<uint32_t>(), (int)nq, 0, 32, st);
        CU(w.tmp.ensure(t4));
        CU(cub::DeviceRadixSort::SortPairsDescending(w.tmp.p, t4, w.neff.as<uint32_t>(), w.neff_sorted.as<uint32_t>(), w.order_in.as<uint32_t>(), w.order.as<uint32_t>(), (int)nq, 0, 32, st));
        ctx->launches += 3;
        w.order_valid = true;
    }
    return KMCPG_OK;
}

// bits of the largest global target index (hit keys are query << bits | target)
static int target_bits_of(const kmcpg_ctx *ctx) {
    int b = 1;
    while (b < 32 && (1ll << b) < std::max<int64_t>(ctx->meta.n_targets, 1)) b++;
    return b;
}

// Per resident block: the row indices (locs kernel, on the query-preparation stream: the indices of block b+1 are computed while
// block b is probed; two buffers alternate) and ONE probe launch on the compute stream; then the counters travel to the host on
// their own stream.  Blocks with numSigs >= 2^32-1 need no locs kernel: their probe derives 64-bit indices itself.
static int enqueue_probes(kmcpg_ctx *ctx, WorkSet &w, const kmcpg_search_params &p, cudaStream_t hs) {
    cudaStream_t st = ctx->st;
    const int H = ctx->meta.num_hashes;
    CU(w.hkeys.ensure(w.cap * 8)); CU(w.hvals.ensure(w.cap * 4));
    bool in_kernel = false;
#ifdef KMCPG_DEV
    static const bool dev_in_kernel = getenv("KMCPG_PROBE_LOCS") && !strcmp(getenv("KMCPG_PROBE_LOCS"), "kernel");
    in_kernel = dev_in_kernel;
#endif
    const bool by_query = ctx->meta.scaled || ctx->meta.minimizer || ctx->meta.syncmer;
    bool any_locs = false;
    for (auto &b : ctx->blocks) any_locs = any_locs || (!in_kernel && ctx->meta.blocks[b.meta_idx].num_sigs < 0xFFFFFFFFull);
    const size_t locs_bytes = std::max<uint64_t>(w.sb.total_slots, 1) * 4ull * H;
    if (any_locs) {
        CU(w.locs[0].ensure(locs_bytes));
        if (ctx->blocks.size() > 1) CU(w.locs[1].ensure(locs_bytes));
    }
    CU(cudaMemsetAsync(w.counters.p, 0, 8, st));
    if (w.planes > 8) CU(cudaMemsetAsync(w.counters.as<unsigned long long>() + 2, 0, ctx->blocks.size() * 8, st));
    size_t bi = 0;
    for (auto &b : ctx->blocks) {
        const BlockMeta &bm = ctx->meta.blocks[b.meta_idx];
        const bool use_locs = !in_kernel && bm.num_sigs < 0xFFFFFFFFull;
        uint32_t *locs = use_locs ? w.locs[bi & 1].as<uint32_t>() : nullptr;
        if (use_locs) {
            if (bi >= 2 && hs != st) CU(cudaStreamWaitEvent(hs, w.probe_ev[(bi - 2) * 3 + 2], 0));   // the probe that read this buffer last
            if (by_query) CU(launch_locs_by_query(w.codes_ptr, w.slot_off.as<uint64_t>(), w.neff.as<uint32_t>(), w.nq, p.paired, H, b.fm, locs, hs));
            else CU(launch_locs(w.codes_ptr, w.sb.total_slots, H, b.fm, locs, hs));
            ctx->launches++;
            if (hs != st) {
                CU(cudaEventRecord(w.probe_ev[bi * 3], hs));
                CU(cudaStreamWaitEvent(st, w.probe_ev[bi * 3], 0));
            }
        }
        CU(cudaEventRecord(w.probe_ev[bi * 3 + 1], st));
        ProbeArgs pa;
        memset(&pa, 0, sizeof(pa));
        pa.rows = b.d_rows; pa.pitch = b.pitch; pa.row_bytes = b.row_bytes;
        pa.n_names = b.n_cols; pa.target_base = (uint32_t)(bm.target_base + b.col0); pa.num_hashes = H;
        pa.codes = w.codes_ptr; pa.fm = b.fm; pa.locs = locs; pa.slot_off = w.slot_off.as<uint64_t>();
        pa.n_eff = w.neff.as<uint32_t>(); pa.thresh = w.thresh.as<uint32_t>(); pa.n_queries = w.nq; pa.paired = p.paired;
        pa.hit_keys = w.hkeys.as<uint64_t>(); pa.hit_vals = w.hvals.as<uint32_t>(); pa.target_bits = target_bits_of(ctx);
        pa.hit_count = w.counters.as<unsigned long long>(); pa.hit_cap = w.cap; pa.dense_counts = nullptr; pa.planes = w.planes;
        pa.task_counter = w.counters.as<unsigned long long>() + 2 + bi;        // one counter per block of this part, zeroed above
        pa.order = w.order_valid ? w.order.as<uint32_t>() : nullptr;
        CU(launch_probe(pa, ctx->sm_count, st)); ctx->launches++;
        CU(cudaEventRecord(w.probe_ev[bi * 3 + 2], st));
        bi++;
    }
    CU(cudaEventRecord(w.ev_a, st));
    CU(cudaStreamWaitEvent(ctx->cnt_st, w.ev_a, 0));
    CU(cudaMemcpyAsync(w.h_cnt.p, w.counters.p, 16, cudaMemcpyDeviceToHost, ctx->cnt_st));
    CU(cudaEventRecord(w.ev_cnt, ctx->cnt_st));
    return KMCPG_OK;
}

// stage A of a part: input staging, query preparation (on the hash stream, so that it runs beside the probes of the part
// before it) and the probe launches (compute stream).  host_seq != nullptr → the inputs are staged through the input stream.
static int enqueue_part(kmcpg_ctx *ctx, WorkSet &w, const kmcpg_search_params &p, int k, const SubBatch &sb_in, const uint8_t *host_seq,
                        const uint64_t *host_off, uint64_t host_bytes, cudaEvent_t ready) {
    NvtxRange nvtx("kmcpg:part enqueue (H2D, query preparation, locs, probe)");
    cudaStream_t st = ctx->st, hs = ctx->hash_st;
#ifdef KMCPG_DEV
    static const bool one_stream = getenv("KMCPG_HASH_STREAM") && atoi(getenv("KMCPG_HASH_STREAM")) == 0;
    if (one_stream) hs = st;
#endif
    int rc = ensure_events(ctx, w);
    if (rc) return rc;
    w.sb = sb_in;
    w.nq = p.paired ? sb_in.n_seqs / 2 : sb_in.n_seqs;
    // the part that used this work set two parts ago must have moved its results out before the buffers are rewritten: stream
    // dependencies, not a host wait, so the host keeps enqueueing ahead of the GPU
    if (w.busy) {
        if (hs != st) CU(cudaStreamWaitEvent(hs, w.ev_b, 0));
        CU(cudaStreamWaitEvent(st, w.ev_b, 0));
    }
    if (host_seq) {
        CU(w.h_off.ensure((sb_in.n_seqs + 1) * 8ull));
        CU(w.off.ensure((sb_in.n_seqs + 1) * 8ull));
        CU(w.seq.ensure(std::max<uint64_t>(host_bytes, 1) + 64));
        if (w.busy) CU(cudaStreamWaitEvent(ctx->in_st, w.ev_b, 0));
        uint64_t *ho = w.h_off.as<uint64_t>();
        const uint64_t base = host_off[0];
        for (uint32_t i = 0; i <= sb_in.n_seqs; i++) ho[i] = host_off[i] - base;
        CU(cudaMemcpyAsync(w.off.p, ho, (sb_in.n_seqs + 1) * 8ull, cudaMemcpyHostToDevice, ctx->in_st));
        if (host_bytes) CU(cudaMemcpyAsync(w.seq.p, host_seq + base, host_bytes, cudaMemcpyHostToDevice, ctx->in_st));
        CU(cudaEventRecord(w.ev_in, ctx->in_st));
        CU(cudaStreamWaitEvent(hs, w.ev_in, 0));
        w.sb.d_seq = w.seq.as<uint8_t>();
        w.sb.d_off = w.off.as<uint64_t>();
    } else if (ready) {
        CU(cudaStreamWaitEvent(hs, ready, 0));                          // device input that is still being produced (kmcpg_batch.ready_event)
    }
    CU(cudaEventRecord(w.ev_a0, hs));
    rc = run_hash_stage(ctx, w, p, k, w.sb, w.nq, &w.codes_ptr, hs);
    if (rc) return rc;
    CU(cudaEventRecord(w.ev_hash, hs));
    if (hs != st) CU(cudaStreamWaitEvent(st, w.ev_hash, 0));
    w.planes = planes_for(w.sb.max_query_slots);
    w.cap = std::max<uint64_t>(1u << 20, 4ull * w.nq);
    if (w.hkeys.cap / 8 > w.cap) w.cap = w.hkeys.cap / 8;
    rc = enqueue_probes(ctx, w, p, hs);
    if (rc) return rc;
    w.busy = true;
    return KMCPG_OK;
}

static int grow_hits(kmcpg_ctx *ctx, HitsPriv &res, uint64_t need) {
    if (res.hits.cap >= need * sizeof(kmcpg_hit)) return KMCPG_OK;
    if (res.ext_hits) return fail(ctx, KMCPG_ENOMEM, "the hit list does not fit the caller's buffer (kmcpg_batch.hits_cap)");
    // results already copied (or in flight on the copy stream) must land before they are moved
    CU(cudaStreamSynchronize(ctx->copy_st));
    PinBuf nb;
    int rc = pin_acquire(ctx, std::max<uint64_t>(need * 2, 1u << 16) * sizeof(kmcpg_hit), nb);
    if (rc) return rc;
    if (res.nh) memcpy(nb.p, res.hits.p, res.nh * sizeof(kmcpg_hit));
    pin_release(ctx, res.hits);
    res.hits = nb;
    return KMCPG_OK;
}

// stage B: hit count known → sort, pack, results to the host (asynchronously, on the copy stream)
static int finish_probes(kmcpg_ctx *ctx, WorkSet &w, const kmcpg_search_params &p, HitsPriv &res, Timing &tm) {
    NvtxRange nvtx("kmcpg:part finish (hit count, hit sort, D2H)");
    cudaStream_t st = ctx->st;
    for (int attempt = 0;; attempt++) {
        CU(cudaEventSynchronize(w.ev_cnt));
        w.n_hits = w.h_cnt.as<uint64_t>()[0];
        for (size_t i = 0; i < ctx->blocks.size(); i++) {
            float b = 0;
            cudaEventElapsedTime(&b, w.probe_ev[i * 3 + 1], w.probe_ev[i * 3 + 2]);
            tm.ms_probe += b; tm.probe_launches++;
        }
        if (w.n_hits <= w.cap) break;
        if (attempt == 2) return fail(ctx, KMCPG_ENOMEM, "hit list keeps overflowing");
        // rare: the hit list overflowed.  Drain, grow, redo the probe phase of this part.
        CU(cudaStreamSynchronize(st));
        CU(cudaStreamSynchronize(ctx->copy_st));
        w.cap = w.n_hits + w.n_hits / 4 + 1024;
        int rc = enqueue_probes(ctx, w, p, st);
        if (rc) return rc;
    }
    float a = 0;
    cudaEventElapsedTime(&a, w.ev_a0, w.ev_hash);
    tm.ms_hash += a;
    tm.probe_bytes += w.h_cnt.as<uint64_t>()[1] * (uint64_t)ctx->meta.num_hashes * (uint64_t)ctx->sum_row_bytes;

    const uint64_t n_hits = w.n_hits;
    w.hit_dst = res.nh;
    // the part's probes are done (the host has just read their counters): sort + pack on the post stream, so the results
    // reach the host while the compute stream is already busy with the next part's probe kernel
    cudaStream_t ps = ctx->post_st;
    CU(cudaStreamWaitEvent(ps, w.ev_a, 0));
    if (n_hits) {
        int rc = grow_hits(ctx, res, res.nh + n_hits);
        if (rc) return rc;
        CU(w.hkeys2.ensure(n_hits * 8)); CU(w.hvals2.ensure(n_hits * 4)); CU(w.hits.ensure(n_hits * sizeof(kmcpg_hit)));
        int qbits = 1; while ((1ull << qbits) < w.nq) qbits++;
        const int tbits = target_bits_of(ctx);                   // keys are query << tbits | target: the sort looks at tbits + qbits bits only
        size_t t3 = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, t3, w.hkeys.as<uint64_t>(), w.hkeys2.as<uint64_t>(), w.hvals.as<uint32_t>(), w.hvals2.as<uint32_t>(),
                                        (int64_t)n_hits, 0, tbits + qbits, ps);
        CU(w.tmp2.ensure(t3));
        CU(cub::DeviceRadixSort::SortPairs(w.tmp2.p, t3, w.hkeys.as<uint64_t>(), w.hkeys2.as<uint64_t>(), w.hvals.as<uint32_t>(), w.hvals2.as<uint32_t>(),
                                           (int64_t)n_hits, 0, tbits + qbits, ps));
        CU(launch_pack_hits(w.hkeys2.as<uint64_t>(), w.hvals2.as<uint32_t>(), n_hits, w.sb.query_base + res.first_query, tbits, w.hits.as<kmcpg_hit>(), ps));
        ctx->launches += 4;
    }
    CU(cudaEventRecord(w.ev_sorted, ps));
    CU(cudaStreamWaitEvent(ctx->copy_st, w.ev_sorted, 0));
    if (n_hits) CU(cudaMemcpyAsync((kmcpg_hit *)res.hits.p + res.nh, w.hits.p, n_hits * sizeof(kmcpg_hit), cudaMemcpyDeviceToHost, ctx->copy_st));
    CU(cudaMemcpyAsync((int32_t *)res.nk.p + w.sb.query_base, w.nk.p, w.nq * 4ull, cudaMemcpyDeviceToHost, ctx->copy_st));
    CU(cudaMemcpyAsync((int32_t *)res.ql.p + w.sb.query_base, w.qlen.p, w.nq * 4ull, cudaMemcpyDeviceToHost, ctx->copy_st));
    CU(cudaEventRecord(w.ev_b, ctx->copy_st));
    res.nh += n_hits;
    return KMCPG_OK;
}

static void fill_out(kmcpg_hits *out, HitsPriv *priv, const Timing &tm, float ms_total, uint32_t launches) {
    out->n_queries = priv->nq;
    out->n_hits = priv->nh;
    out->n_kmers = (int32_t *)priv->nk.p;
    out->query_len = (int32_t *)priv->ql.p;
    out->hits = (kmcpg_hit *)priv->hits.p;
    out->ms_hash = tm.ms_hash; out->ms_locs = tm.ms_locs; out->ms_probe = tm.ms_probe; out->ms_total = ms_total;
    out->probe_launches = tm.probe_launches;
    out->probe_row_bytes = tm.probe_bytes;
    out->kernel_launches = launches;
    out->_priv = priv;
}

static void drop_priv(HitsPriv *priv) {
    if (!priv) return;
    if (priv->ctx) {
        pin_release(priv->ctx, priv->nk); pin_release(priv->ctx, priv->ql);
        if (!priv->ext_hits) pin_release(priv->ctx, priv->hits);
    }
    delete priv;
}

static int check_search_args(kmcpg_ctx *ctx, const kmcpg_search_params *p, const void *seq, const void *off, uint32_t n_seqs, int *k) {
    if (!ctx) return KMCPG_EINVAL;
    if (!p || (n_seqs && (!seq || !off))) return fail(ctx, KMCPG_EINVAL, "null argument");
    if (!ctx->has_db) return fail(ctx, KMCPG_EINVAL, "no database open");
    if (p->paired && (n_seqs & 1)) return fail(ctx, KMCPG_EINVAL, "paired batch needs an even number of sequences");
    *k = p->k > 0 ? p->k : ctx->meta.ks.front();
    if (std::find(ctx->meta.ks.begin(), ctx->meta.ks.end(), *k) == ctx->meta.ks.end()) return fail(ctx, KMCPG_EINVAL, "k is not one of the database's k values");
    if (*k > 64 || *k < 1) return fail(ctx, KMCPG_EUNSUPPORTED, "k must be in 1..64");
    if (p->min_matched < 1) return fail(ctx, KMCPG_EINVAL, "min_matched must be >= 1");
    if (!(p->min_query_cov >= 0 && p->min_query_cov <= 1)) return fail(ctx, KMCPG_EINVAL, "min_query_cov must be in [0,1]");
    return KMCPG_OK;
}

struct Part { uint32_t a, b; uint64_t slots, maxq; };

static const uint64_t PART_SLOTS = 32ull << 20;   // k-mer slots per part (≈ 250 k reads of 150 bp)
static const uint32_t PART_SEQS = 2u << 20;

// greedy parts [a, b) from host-visible offsets.  Sequences are taken in blocks of 4096 whose slot sum / max are
// computed by a branch-free (vectorisable) loop; only blocks that are large by themselves are walked one by one.
static int cut_parts(kmcpg_ctx *ctx, const uint64_t *off, uint32_t n_seqs, uint32_t step, int k, std::vector<Part> &parts) {
    const uint64_t kk = (uint64_t)k;
    const uint32_t BLK = 4096;
    // A part should keep the GPU busy for a few milliseconds, or the per-part host round trips (hit count, sort launch, delivery) show:
    // a context that holds only a narrow share of the rows (a column-range shard of one block, a small sketch database) takes
    // proportionally more k-mers per part — up to 4x, the work-set buffers grow with it.
    uint64_t PART_SLOTS = kmcpg::PART_SLOTS;
    {
        // (the widest shard of the plan decides, so that all shards of one database cut a batch into the same parts)
        const uint64_t row_work = (uint64_t)std::max<int64_t>(1, ctx->part_row_bytes > 0 ? ctx->part_row_bytes : ctx->sum_row_bytes) * (uint64_t)std::max(1, ctx->meta.num_hashes);
        if (row_work < 1250) PART_SLOTS = std::min<uint64_t>(4 * PART_SLOTS, PART_SLOTS * 1250 / row_work);
    }
    uint32_t a = 0, b = 0;
    uint64_t slots = 0, maxq = 0;
    auto limit = [&]() { return a == 0 ? PART_SLOTS / 4 : PART_SLOTS; };   // a short first part fills the pipeline quickly
    auto close = [&]() { parts.push_back({a, b, slots, maxq}); a = b; slots = 0; maxq = 0; };
    while (b < n_seqs) {
        const uint32_t e = std::min<uint32_t>(n_seqs, b + BLK);
        uint64_t bsum = 0, bmax = 0, bad = 0;
        for (uint32_t i = b; i < e; i++) {
            const uint64_t lo = off[i], hi = off[i + 1];
            bad |= (uint64_t)(hi < lo);
            const uint64_t len = hi - lo;
            const uint64_t qs = len >= kk ? len - kk + 1 : 0;
            bsum += qs;
            bmax = bmax > qs ? bmax : qs;
        }
        if (bad) return fail(ctx, KMCPG_EINVAL, "offsets must be non-decreasing");
        if (step == 2) bmax *= 2;                                   // upper bound of a query's two mates
        if (bsum <= PART_SLOTS / 16) {                              // a small block moves as one unit
            if (b > a && (slots + bsum > limit() || (b - a) + (e - b) > PART_SEQS)) close();
            slots += bsum; maxq = std::max(maxq, bmax); b = e;
        } else {                                                    // long sequences: query by query
            for (uint32_t i = b; i < e; i += step) {
                uint64_t qs = 0;
                for (uint32_t m = 0; m < step; m++) { const uint64_t len = off[i + m + 1] - off[i + m]; qs += len >= kk ? len - kk + 1 : 0; }
                if (i > a && (slots + qs > limit() || (i - a) >= PART_SEQS)) { b = i; close(); }
                slots += qs; maxq = std::max(maxq, qs);
            }
            b = e;
        }
    }
    if (b > a) close();
    // A short LAST part drains the pipeline quickly: what follows it — hit sort, result copy, the caller's post-filter of that part — is not
    // hidden behind another probe unless a second batch is queued.  The tail of a big last part becomes a part of its own (~ 1/8 of a part).
    if (!parts.empty() && parts.back().slots > PART_SLOTS / 4) {
        Part &last = parts.back();
        uint64_t tail = 0, tmax = 0;
        uint32_t c = last.b;
        while (c > last.a + step && tail < PART_SLOTS / 8) {
            uint64_t qs = 0;
            for (uint32_t m = 0; m < step; m++) { const uint64_t len = off[c - step + m + 1] - off[c - step + m]; qs += len >= kk ? len - kk + 1 : 0; }
            if (tail + qs > PART_SLOTS / 4) break;
            tail += qs; tmax = std::max(tmax, qs); c -= step;
        }
        if (c > last.a && c < last.b && tail > 0) {
            const Part t{c, last.b, tail, tmax};
            last.b = c; last.slots -= tail;                     // its maxq stays as an upper bound
            parts.push_back(t);
        }
    }
    return KMCPG_OK;
}

struct PartDone { uint32_t first_query, nq; uint64_t hit_dst, n_hits; cudaEvent_t ev; int ws; };

}  // namespace kmcpg

// One submitted batch (kmcpg_search_submit): its parts travel through the context's executor thread.
struct kmcpg_job {
    kmcpg_ctx *ctx = nullptr;
    kmcpg_search_params p;
    int k = 0;
    const uint8_t *host_seq = nullptr; const uint64_t *host_off = nullptr;       // host input: parts are staged by the executor
    const uint8_t *d_seq = nullptr; const uint64_t *d_off = nullptr;             // device input
    PinBuf hoff_own;                                                              // offsets fetched from the device when the caller had no host copy
    cudaEvent_t ready = nullptr;
    uint32_t n_seqs = 0;
    std::vector<Part> parts;
    kmcpg_part_cb cb = nullptr; void *user = nullptr;
    HitsPriv *priv = nullptr;
    Timing tm;
    std::vector<PartDone> done;
    size_t enqueued = 0, finished = 0, delivered = 0;
    uint32_t launches = 0;
    std::chrono::steady_clock::time_point t0;
    int rc = KMCPG_OK;
    std::string err;
    bool complete = false;
};

namespace kmcpg {

// The executor: one host thread per context that feeds the GPU.  Parts of all submitted jobs form ONE stream of work: part n
// uses work set n & 1, its kernels are enqueued BEFORE the host waits for the hit count of part n-1 — also when part n-1 belongs
// to the job before — so neither the count round trip, nor the hit sort + result copies, nor the boundary between two batches
// leaves the GPU idle as long as the caller keeps a second batch submitted.
struct Executor {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::deque<kmcpg_job *> queue;       // submitted, not yet completely enqueued
    bool stop = false, idle = true;
    uint64_t seq = 0;                    // parts enqueued so far
};

static void sync_all_streams(kmcpg_ctx *ctx) {
    cudaStreamSynchronize(ctx->hash_st);
    cudaStreamSynchronize(ctx->st);
    cudaStreamSynchronize(ctx->copy_st);
    cudaStreamSynchronize(ctx->cnt_st);
    cudaStreamSynchronize(ctx->in_st);
    cudaStreamSynchronize(ctx->post_st);
    for (auto &w : ctx->ws) w.busy = false;
}

static const bool g_trace = getenv("KMCPG_TRACE") != nullptr;
static const auto g_t0 = std::chrono::steady_clock::now();
static float now_ms() { return std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - g_t0).count(); }

// hands every finished part of the job whose results have reached the host to the caller, in order
static int deliver_job(kmcpg_ctx *ctx, kmcpg_job *job) {
    for (; job->delivered < job->done.size(); job->delivered++) {
        const PartDone &d = job->done[job->delivered];
        const float ta = now_ms();
        {
            NvtxRange nvtx("kmcpg:part wait for D2H");
            CU(cudaEventSynchronize(d.ev));
        }
        ctx->ws[d.ws].busy = false;                  // the newest user of that work set is this part: everything it queued is done
        const float tb = now_ms();
        if (job->cb) {
            HitsPriv &res = *job->priv;
            kmcpg_part pt;
            pt.first_query = d.first_query + res.first_query; pt.n_queries = d.nq;
            pt.n_kmers = (const int32_t *)res.nk.p + d.first_query; pt.query_len = (const int32_t *)res.ql.p + d.first_query;
            pt.hits = (const kmcpg_hit *)res.hits.p + d.hit_dst; pt.n_hits = d.n_hits;
            NvtxRange nvtx("kmcpg:part callback (host post-filter)");
            job->cb(job->user, &pt);
        }
        if (g_trace) fprintf(stderr, "[trace] job %p part %zu: d2h wait %.2f..%.2f cb ..%.2f (%u q, %llu hits)\n", (void *)job, job->delivered, ta, tb, now_ms(), d.nq, (unsigned long long)d.n_hits);
    }
    return KMCPG_OK;
}

static void complete_job(kmcpg_ctx *ctx, kmcpg_job *job, int rc) {
    Executor &ex = *ctx->exec;
    std::lock_guard<std::mutex> lk(ex.mu);
    if (rc && !job->rc) { job->rc = rc; job->err = ctx->err; }
    job->complete = true;
    ex.cv_done.notify_all();
}

static void executor_main(kmcpg_ctx *ctx) {
    cudaSetDevice(ctx->device);
    Executor &ex = *ctx->exec;
    struct Pending { kmcpg_job *job; int ws; };
    std::deque<Pending> pending;                 // enqueued on the GPU, hit count not read yet (at most two)
    // A failed part ends its job.  A CUDA failure leaves the pipeline in an unknown state: drain it and fail every job that has parts
    // in flight; any other failure (a caller's hit buffer that is too small, an allocation) concerns that job alone — its parts
    // are taken out of the pipeline, the parts of the other jobs stay where they are.
    auto fail_inflight = [&](kmcpg_job *culprit, int rc) {
        sync_all_streams(ctx);
        (void)cudaGetLastError();
        std::vector<kmcpg_job *> hit{culprit};
        if (rc == KMCPG_ECUDA) {
            for (auto &pd : pending) if (std::find(hit.begin(), hit.end(), pd.job) == hit.end()) hit.push_back(pd.job);
            pending.clear();
        } else {
            for (auto it = pending.begin(); it != pending.end();) it = it->job == culprit ? pending.erase(it) : it + 1;
        }
        {
            std::lock_guard<std::mutex> lk(ex.mu);
            for (kmcpg_job *j : hit) {
                auto it = std::find(ex.queue.begin(), ex.queue.end(), j);
                if (it != ex.queue.end()) ex.queue.erase(it);
            }
        }
        for (kmcpg_job *j : hit) complete_job(ctx, j, rc);
    };
    for (;;) {
        kmcpg_job *job = nullptr;
        {
            std::unique_lock<std::mutex> lk(ex.mu);
            if (pending.empty() && ex.queue.empty()) {
                ex.idle = true;
                ex.cv_done.notify_all();
                ex.cv_work.wait(lk, [&] { return ex.stop || !ex.queue.empty(); });
                if (ex.queue.empty()) return;                        // stop
                ex.idle = false;
            }
            if (!ex.queue.empty()) job = ex.queue.front();
        }
        bool enq = false;
        if (job) {
            WorkSet &w = ctx->ws[ex.seq & 1];
            const Part &pt = job->parts[job->enqueued];
            const uint32_t step = job->p.paired ? 2 : 1;
            SubBatch sb{job->d_seq, job->d_off ? job->d_off + pt.a : nullptr, pt.b - pt.a, pt.slots, pt.maxq, pt.a / step};
            const uint32_t l0 = ctx->launches;
            int rc;
            if (job->host_seq) rc = enqueue_part(ctx, w, job->p, job->k, sb, job->host_seq, job->host_off + pt.a, job->host_off[pt.b] - job->host_off[pt.a], nullptr);
            else rc = enqueue_part(ctx, w, job->p, job->k, sb, nullptr, nullptr, 0, job->ready);
            job->launches += ctx->launches - l0;
            if (rc) { fail_inflight(job, rc); continue; }
            if (g_trace) fprintf(stderr, "[trace] job %p part %zu enqueued at %.2f\n", (void *)job, job->enqueued, now_ms());
            pending.push_back({job, (int)(ex.seq & 1)});
            ex.seq++;
            enq = true;
            if (++job->enqueued == job->parts.size()) {
                std::lock_guard<std::mutex> lk(ex.mu);
                ex.queue.pop_front();
            }
        }
        // read the hit count of the older part once a newer one is queued behind it on the GPU — or at once when there is nothing
        // more to enqueue right now
        while (pending.size() > (enq ? 1u : 0u)) {
            const Pending pd = pending.front();
            pending.pop_front();
            kmcpg_job *j = pd.job;
            WorkSet &w = ctx->ws[pd.ws];
            const float tf = now_ms();
            const uint32_t l0 = ctx->launches;
            int rc = finish_probes(ctx, w, j->p, *j->priv, j->tm);
            j->launches += ctx->launches - l0;
            if (rc) { fail_inflight(j, rc); break; }
            if (g_trace) fprintf(stderr, "[trace] job %p part %zu probes finished: wait %.2f..%.2f\n", (void *)j, j->finished, tf, now_ms());
            j->done.push_back({w.sb.query_base, w.nq, w.hit_dst, w.n_hits, w.ev_b, pd.ws});
            j->finished++;
            // its sort + copies run beside the probes of the next part: hand it to the caller as soon as it has landed
            rc = deliver_job(ctx, j);
            if (rc) { fail_inflight(j, rc); break; }
            if (j->finished == j->parts.size()) complete_job(ctx, j, KMCPG_OK);
        }
    }
}

static void start_executor(kmcpg_ctx *ctx) {
    if (ctx->exec) return;
    ctx->exec = new Executor();
    ctx->exec->th = std::thread(executor_main, ctx);
}

// blocks until the executor has nothing in flight (callers that are about to change the database or borrow the work sets
// hold ctx->mu, so nothing new can be submitted meanwhile)
void executor_drain(kmcpg_ctx *ctx) {
    if (!ctx->exec) return;
    Executor &ex = *ctx->exec;
    std::unique_lock<std::mutex> lk(ex.mu);
    ex.cv_done.wait(lk, [&] { return ex.idle && ex.queue.empty(); });
}

void executor_stop(kmcpg_ctx *ctx) {
    if (!ctx->exec) return;
    executor_drain(ctx);
    {
        std::lock_guard<std::mutex> lk(ctx->exec->mu);
        ctx->exec->stop = true;
    }
    ctx->exec->cv_work.notify_all();
    ctx->exec->th.join();
    delete ctx->exec;
    ctx->exec = nullptr;
}

}  // namespace kmcpg

// ========================================================================================================
// C ABI of the search path
// ========================================================================================================
extern "C" {

void kmcpg_default_params(kmcpg_search_params *p) {
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->min_query_len = 30; p->min_matched = 10; p->dedup_threshold = 256; p->min_query_cov = 0.55;   // S:1055-1069
}

int kmcpg_search_submit(kmcpg_ctx *ctx, const kmcpg_search_params *p, const kmcpg_batch *b, kmcpg_job **out) {
    if (!ctx) return KMCPG_EINVAL;
    if (!b || !out) return fail(ctx, KMCPG_EINVAL, "null argument");
    *out = nullptr;
    int k = 0;
    int rc = check_search_args(ctx, p, b->seq, b->off, b->n_seqs, &k);
    if (rc) return rc;
    if (b->hits_dst && b->hits_cap == 0) return fail(ctx, KMCPG_EINVAL, "hits_dst needs hits_cap");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    kmcpg_job *job = new kmcpg_job();
    auto bail = [&](int code) { if (job->priv) drop_priv(job->priv); pin_release(ctx, job->hoff_own); delete job; return code; };
    job->ctx = ctx; job->p = *p; job->k = k; job->n_seqs = b->n_seqs; job->cb = b->cb; job->user = b->user;
    const uint32_t first_query = b->first_query;
    job->t0 = std::chrono::steady_clock::now();
    const uint64_t *cut_off = b->off;
    static const uint64_t zero_off[1] = {0};
    if (b->on_device) {
        job->d_seq = b->seq; job->d_off = b->off; job->ready = (cudaEvent_t)b->ready_event;
        cut_off = b->host_off;
        if (!cut_off && b->n_seqs) {
            // the lengths live on the device only: fetch the offsets (8 B per sequence) to cut the batch into parts
            rc = pin_acquire(ctx, ((size_t)b->n_seqs + 1) * 8, job->hoff_own);
            if (rc) return bail(rc);
            cudaError_t e = cudaSuccess;
            if (job->ready) e = cudaStreamWaitEvent(ctx->cnt_st, job->ready, 0);
            if (e == cudaSuccess) e = cudaMemcpyAsync(job->hoff_own.p, b->off, ((size_t)b->n_seqs + 1) * 8, cudaMemcpyDeviceToHost, ctx->cnt_st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->cnt_st);
            if (e != cudaSuccess) { (void)cudaGetLastError(); fail(ctx, KMCPG_ECUDA, std::string("CUDA error: ") + cudaGetErrorString(e)); return bail(KMCPG_ECUDA); }
            cut_off = (const uint64_t *)job->hoff_own.p;
        }
    } else {
        job->host_seq = b->seq ? b->seq : (const uint8_t *)"";
        job->host_off = b->off;
    }
    if (!b->n_seqs) cut_off = zero_off;
    const uint32_t step = p->paired ? 2 : 1;
    rc = cut_parts(ctx, cut_off, b->n_seqs, step, k, job->parts);
    if (rc) return bail(rc);
    HitsPriv *priv = new HitsPriv();
    job->priv = priv;
    priv->ctx = ctx; priv->nq = b->n_seqs / step; priv->first_query = first_query;
    rc = pin_acquire(ctx, std::max<uint32_t>(priv->nq, 1) * 4ull, priv->nk);
    if (!rc) rc = pin_acquire(ctx, std::max<uint32_t>(priv->nq, 1) * 4ull, priv->ql);
    if (!rc) {
        if (b->hits_dst) { priv->hits.p = b->hits_dst; priv->hits.cap = b->hits_cap * sizeof(kmcpg_hit); priv->ext_hits = true; }
        else rc = pin_acquire(ctx, std::max<uint64_t>(1u << 16, 2ull * priv->nq) * sizeof(kmcpg_hit), priv->hits);
    }
    if (rc) return bail(rc);
    start_executor(ctx);
    if (job->parts.empty()) job->complete = true;               // nothing to search
    else {
        std::lock_guard<std::mutex> lk2(ctx->exec->mu);
        ctx->exec->queue.push_back(job);
        ctx->exec->idle = false;
    }
    ctx->exec->cv_work.notify_all();
    *out = job;
    return KMCPG_OK;
}

int kmcpg_search_wait(kmcpg_job *job, kmcpg_hits *out) {
    if (!job) return KMCPG_EINVAL;
    kmcpg_ctx *ctx = job->ctx;
    {
        std::unique_lock<std::mutex> lk(ctx->exec->mu);
        ctx->exec->cv_done.wait(lk, [&] { return job->complete; });
    }
    const int rc = job->rc;
    pin_release(ctx, job->hoff_own);
    if (rc || !out) {
        if (rc) { std::lock_guard<std::mutex> lk(ctx->mu); ctx->err = job->err; }
        drop_priv(job->priv);
    } else {
        memset(out, 0, sizeof(*out));
        const float ms_total = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - job->t0).count();
        fill_out(out, job->priv, job->tm, ms_total, job->launches);
    }
    delete job;
    return rc;
}

static int search_sync(kmcpg_ctx *ctx, const kmcpg_search_params *p, const kmcpg_batch &b, kmcpg_hits *out) {
    if (!ctx) return KMCPG_EINVAL;
    if (!out) return fail(ctx, KMCPG_EINVAL, "null argument");
    kmcpg_job *job = nullptr;
    int rc = kmcpg_search_submit(ctx, p, &b, &job);
    if (rc) return rc;
    return kmcpg_search_wait(job, out);
}

int kmcpg_search_batch(kmcpg_ctx *ctx, const kmcpg_search_params *p, const uint8_t *seq, const uint64_t *off, uint32_t n_seqs, kmcpg_hits *out) {
    kmcpg_batch b;
    memset(&b, 0, sizeof(b));
    b.seq = seq; b.off = off; b.n_seqs = n_seqs;
    return search_sync(ctx, p, b, out);
}

int kmcpg_search_batch_cb(kmcpg_ctx *ctx, const kmcpg_search_params *p, const uint8_t *seq, const uint64_t *off, uint32_t n_seqs, kmcpg_part_cb cb, void *user,
                          kmcpg_hits *out) {
    if (ctx && !cb) return fail(ctx, KMCPG_EINVAL, "callback is NULL");
    kmcpg_batch b;
    memset(&b, 0, sizeof(b));
    b.seq = seq; b.off = off; b.n_seqs = n_seqs; b.cb = cb; b.user = user;
    return search_sync(ctx, p, b, out);
}

int kmcpg_search_batch_device(kmcpg_ctx *ctx, const kmcpg_search_params *p, const uint8_t *d_seq, const uint64_t *d_off, uint32_t n_seqs,
                              uint64_t seq_bytes, kmcpg_hits *out) {
    (void)seq_bytes;
    kmcpg_batch b;
    memset(&b, 0, sizeof(b));
    b.seq = d_seq; b.off = d_off; b.n_seqs = n_seqs; b.on_device = 1;
    return search_sync(ctx, p, b, out);
}

void kmcpg_free_hits(kmcpg_hits *h) {
    if (!h) return;
    drop_priv((HitsPriv *)h->_priv);
    memset(h, 0, sizeof(*h));
}

}  // extern "C"

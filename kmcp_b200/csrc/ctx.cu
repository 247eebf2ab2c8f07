// ctx.cu — context, database residency in HBM, the batch executor, and the C ABI of include/kmcp_gpu.h.
//
// Data layout in HBM (per resident block): the on-disk row-major bit matrix numSigs × numRowBytes
// (X:307-349; row r byte i bit 7-j ⇔ Bloom bit r of target 8i+j) is re-pitched on upload so that every
// row starts 16-byte aligned (128-byte aligned for rows wider than 256 B) and is zero padded; the disk
// format is untouched.
//
// Batch executor: a call is cut into parts (≤ 32 M k-mer slots each).  ALL kernels run in order on one
// compute stream, so every launch can be timed exactly with events; transfers run on a copy stream.
// Part i goes  [H2D] → slot scan → hash → (sort+unique of long queries) → verdict → per block {locs, probe}
// and, once the host has read its hit count, → radix sort of the hit list → pack → [D2H into pinned results].
// Two work sets alternate, and the kernels of part i+1 are enqueued BEFORE the host waits for part i's hit
// count, so the round trip and the result copies hide behind the next part's probe kernel.
#include <cuda_runtime.h>
#include <fcntl.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cub/cub.cuh>
#include <mutex>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include "ctx_internal.h"

using namespace kmcpg;

thread_local std::string g_create_error;

namespace kmcpg {

int fail(kmcpg_ctx *c, int code, const std::string &msg) {
    if (c) c->err = msg; else g_create_error = msg;
    return code;
}

uint32_t pitch_for(uint32_t row_bytes) {
    if (row_bytes > 256) return (row_bytes + 127) / 128 * 128;
    return (row_bytes + 15) / 16 * 16;
}

void layout_block(DeviceBlock &b, const BlockMeta &m, uint32_t col0, uint32_t n_cols) {
    if (n_cols == 0) { col0 = 0; n_cols = (uint32_t)m.n_names; }      // the whole block
    b.col0 = col0; b.n_cols = n_cols;
    b.whole = col0 == 0 && n_cols == (uint32_t)m.n_names;
    b.row_bytes = b.whole ? (uint32_t)m.row_bytes : (n_cols + 7) / 8;
    b.pitch = pitch_for(b.row_bytes);
    b.row16 = (b.row_bytes + 15) / 16;
    uint32_t g = 1;
    while (g < b.row16 && g < 8) g <<= 1;       // informational: the probe kernel derives its task geometry from row_bytes
    b.G = g;
    b.chunks = (b.row16 + g - 1) / g;
    b.fm = make_fastmod(m.num_sigs);
}

void free_db(kmcpg_ctx *ctx) {
    for (auto &b : ctx->blocks) if (b.d_rows) cudaFree(b.d_rows);
    ctx->blocks.clear();
    ctx->resident_of.clear();
    ctx->target_sizes.clear();
    ctx->has_db = false;
    ctx->sum_row_bytes = ctx->resident_bytes = ctx->disk_bytes = 0;
}

void WorkSet::release() {
    for (DevBuf *b : {&seq, &off, &slot_cnt, &slot_off, &codes, &codes2, &locs, &ncodes, &qlen, &nk, &neff, &thresh, &hkeys, &hvals, &hkeys2, &hvals2,
                      &hits, &counters, &tmp, &tmp2, &segb, &sege, &ck, &cs, &cs_cnt, &cs_off, &tile_n, &tile_off, &tile_cnt, &tile_pre})
        b->release();
    h_off.release(); h_cnt.release();
    for (cudaEvent_t *e : {&ev_in, &ev_a0, &ev_hash, &ev_a, &ev_cnt, &ev_sorted, &ev_b}) if (*e) { cudaEventDestroy(*e); *e = nullptr; }
    for (auto e : probe_ev) cudaEventDestroy(e);
    probe_ev.clear();
}

int pin_acquire(kmcpg_ctx *ctx, size_t bytes, PinBuf &out) {
    if (bytes == 0) bytes = 8;
    {
        std::lock_guard<std::mutex> lk(ctx->pin_mu);
        int best = -1;
        for (size_t i = 0; i < ctx->pin_pool.size(); i++)
            if (ctx->pin_pool[i].cap >= bytes && (best < 0 || ctx->pin_pool[i].cap < ctx->pin_pool[best].cap)) best = (int)i;
        if (best >= 0 && ctx->pin_pool[best].cap <= bytes * 4 + (1u << 20)) {
            out = ctx->pin_pool[best];
            ctx->pin_pool.erase(ctx->pin_pool.begin() + best);
            return KMCPG_OK;
        }
    }
    size_t want = bytes + bytes / 4 + 4096;
    void *p = nullptr;
    cudaError_t e = cudaMallocHost(&p, want);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return fail(ctx, KMCPG_ENOMEM, std::string("pinned allocation failed: ") + cudaGetErrorString(e)); }
    out.p = p; out.cap = want;
    return KMCPG_OK;
}

void pin_release(kmcpg_ctx *ctx, PinBuf &b) {
    if (!b.p) return;
    std::lock_guard<std::mutex> lk(ctx->pin_mu);
    ctx->pin_pool.push_back(b);
    b = PinBuf();
    while (ctx->pin_pool.size() > 12) {      // keep the pool small: drop the smallest buffer
        size_t m = 0;
        for (size_t i = 1; i < ctx->pin_pool.size(); i++) if (ctx->pin_pool[i].cap < ctx->pin_pool[m].cap) m = i;
        cudaFreeHost(ctx->pin_pool[m].p);
        ctx->pin_pool.erase(ctx->pin_pool.begin() + m);
    }
}

// blocks → shards: largest first onto the least loaded shard (deterministic in every rank)
static void plan_shards(const DbMeta &m, int world, std::vector<int> &owner, std::vector<uint64_t> &load) {
    std::vector<int> order(m.blocks.size());
    std::iota(order.begin(), order.end(), 0);
    auto bytes_of = [&](int i) { return (uint64_t)m.blocks[i].num_sigs * pitch_for((uint32_t)m.blocks[i].row_bytes); };
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return bytes_of(a) > bytes_of(b); });
    load.assign(world, 0);
    owner.assign(m.blocks.size(), 0);
    for (int i : order) {
        int best = (int)(std::min_element(load.begin(), load.end()) - load.begin());
        owner[i] = best;
        load[best] += bytes_of(i);
    }
}

// The pieces every shard keeps.  A DB with at least as many blocks as shards is split by whole blocks
// (plan_shards).  With FEWER blocks than shards, blocks are cut by column range: the columns of all blocks, in units of
// 128 targets (16 row bytes) weighted by the block's numSigs, are laid end to end and shard s takes the s-th of `world`
// equal-cost stretches — so a shard holds a contiguous run of columns that may span a block boundary.
void plan_pieces(const DbMeta &m, int world, std::vector<ShardPiece> &pieces, std::vector<uint64_t> &load) {
    pieces.clear();
    load.assign(world, 0);
    if (world <= 1 || m.blocks.size() >= (size_t)world) {
        std::vector<int> owner;
        plan_shards(m, world < 1 ? 1 : world, owner, load);
        for (size_t i = 0; i < m.blocks.size(); i++) pieces.push_back({(int)i, owner[i], 0u, (uint32_t)m.blocks[i].n_names});
        return;
    }
    const uint32_t UNIT = 128;
    std::vector<uint64_t> units(m.blocks.size());
    long double total = 0;
    for (size_t i = 0; i < m.blocks.size(); i++) {
        units[i] = ((uint64_t)m.blocks[i].n_names + UNIT - 1) / UNIT;
        total += (long double)units[i] * (long double)m.blocks[i].num_sigs;
    }
    long double done = 0;             // cost of the units already handed out
    int s = 0;
    for (size_t i = 0; i < m.blocks.size(); i++) {
        const long double w = (long double)m.blocks[i].num_sigs;
        uint64_t u = 0;
        while (u < units[i]) {
            // units of this block that still fit below the end of shard s's stretch (at least one, so every piece is non-empty)
            const long double end_s = s + 1 >= world ? total : total * (long double)(s + 1) / (long double)world;
            uint64_t take = w > 0 ? (uint64_t)((end_s - done) / w + 0.5L) : units[i] - u;
            if (s + 1 >= world) take = units[i] - u;
            take = std::max<uint64_t>(1, std::min<uint64_t>(take, units[i] - u));
            const uint32_t c0 = (uint32_t)(u * UNIT);
            const uint32_t c1 = (uint32_t)std::min<uint64_t>((u + take) * UNIT, (uint64_t)m.blocks[i].n_names);
            pieces.push_back({(int)i, s, c0, c1 - c0});
            load[s] += (uint64_t)m.blocks[i].num_sigs * pitch_for((c1 - c0 + 7) / 8);
            u += take;
            done += (long double)take * w;
            if (s + 1 < world && done + 0.5L * w >= total * (long double)(s + 1) / (long double)world) s++;
        }
    }
}

static int planes_for(uint64_t max_n) {
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
    CU(w.counters.ensure(64));
    return KMCPG_OK;
}

// widens a u32 count to u64 while scanning
struct U32ToU64 { __host__ __device__ uint64_t operator()(uint32_t v) const { return (uint64_t)v; } };

// hashing of one HashArgs job: warp per query for short sequences, warp per 4096-position tile (+ gather) when a query is long
static int hash_any(kmcpg_ctx *ctx, WorkSet &w, const HashArgs &ha, uint32_t n_seqs, uint64_t total_slots, uint64_t max_query_slots) {
    cudaStream_t st = ctx->st;
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

// slot scan → hash → (sort+unique) → verdict, all on the compute stream; sb.d_seq/d_off must already be valid there
int run_hash_stage(kmcpg_ctx *ctx, WorkSet &w, const kmcpg_search_params &p, int k, const SubBatch &sb, uint32_t nq, uint64_t **codes_out) {
    const DbMeta &m = ctx->meta;
    cudaStream_t st = ctx->st;
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
        rc = hash_any(ctx, w, ha, sb.n_seqs, sb.total_slots, sb.max_query_slots);
        if (rc) return rc;
    } else {
        // sketch databases: hash every position (k-mers, and s-mers for syncmers), then select per window
        CU(w.ck.ensure(std::max<uint64_t>(sb.total_slots, 1) * 8));
        HashArgs hr = ha;
        hr.raw = 1; hr.n_queries = sb.n_seqs; hr.codes = w.ck.as<uint64_t>();
        rc = hash_any(ctx, w, hr, sb.n_seqs, sb.total_slots, sb.max_query_slots);
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
            rc = hash_any(ctx, w, hs, sb.n_seqs, sb.total_slots + (uint64_t)sb.n_seqs * (uint64_t)(k - s + 1), sb.max_query_slots + 2ull * (k - s));
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
    return KMCPG_OK;
}

// per block {locs, probe} on the compute stream, then the counters travel to the host on the copy stream
static int enqueue_probes(kmcpg_ctx *ctx, WorkSet &w, const kmcpg_search_params &p) {
    cudaStream_t st = ctx->st;
    const int H = ctx->meta.num_hashes;
    CU(w.locs.ensure(std::max<uint64_t>(w.sb.total_slots, 1) * 4ull * H));
    CU(w.hkeys.ensure(w.cap * 8)); CU(w.hvals.ensure(w.cap * 4));
    CU(cudaMemsetAsync(w.counters.p, 0, 8, st));
    size_t bi = 0;
    for (auto &b : ctx->blocks) {
        const BlockMeta &bm = ctx->meta.blocks[b.meta_idx];
        CU(cudaEventRecord(w.probe_ev[bi * 3], st));
        if (ctx->meta.scaled || ctx->meta.minimizer || ctx->meta.syncmer)
            CU(launch_locs_by_query(w.codes_ptr, w.slot_off.as<uint64_t>(), w.neff.as<uint32_t>(), w.nq, p.paired, H, b.fm, w.locs.as<uint32_t>(), st));
        else
            CU(launch_locs(w.codes_ptr, w.sb.total_slots, H, b.fm, w.locs.as<uint32_t>(), st));
        ctx->launches++;
        CU(cudaEventRecord(w.probe_ev[bi * 3 + 1], st));
        ProbeArgs pa;
        memset(&pa, 0, sizeof(pa));
        pa.rows = b.d_rows; pa.pitch = b.pitch; pa.row_bytes = b.row_bytes;
        pa.n_names = b.n_cols; pa.target_base = (uint32_t)(bm.target_base + b.col0); pa.num_hashes = H;
        pa.locs = w.locs.as<uint32_t>(); pa.slot_off = w.slot_off.as<uint64_t>();
        pa.n_eff = w.neff.as<uint32_t>(); pa.thresh = w.thresh.as<uint32_t>(); pa.n_queries = w.nq; pa.paired = p.paired;
        pa.hit_keys = w.hkeys.as<uint64_t>(); pa.hit_vals = w.hvals.as<uint32_t>();
        pa.hit_count = w.counters.as<unsigned long long>(); pa.hit_cap = w.cap; pa.dense_counts = nullptr; pa.planes = w.planes;
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

// stage A of a part: everything up to the probes.  host_seq != nullptr → stage the inputs through the copy stream.
static int enqueue_part(kmcpg_ctx *ctx, WorkSet &w, const kmcpg_search_params &p, int k, const SubBatch &sb_in, const uint8_t *host_seq,
                        const uint64_t *host_off, uint64_t host_bytes) {
    cudaStream_t st = ctx->st;
    int rc = ensure_events(ctx, w);
    if (rc) return rc;
    w.sb = sb_in;
    w.nq = p.paired ? sb_in.n_seqs / 2 : sb_in.n_seqs;
    if (host_seq) {
        CU(w.h_off.ensure((sb_in.n_seqs + 1) * 8ull));
        CU(w.off.ensure((sb_in.n_seqs + 1) * 8ull));
        CU(w.seq.ensure(std::max<uint64_t>(host_bytes, 1) + 64));
        if (w.busy) CU(cudaStreamWaitEvent(ctx->in_st, w.ev_b, 0));     // the previous part of this work set is completely done
        uint64_t *ho = w.h_off.as<uint64_t>();
        const uint64_t base = host_off[0];
        for (uint32_t i = 0; i <= sb_in.n_seqs; i++) ho[i] = host_off[i] - base;
        CU(cudaMemcpyAsync(w.off.p, ho, (sb_in.n_seqs + 1) * 8ull, cudaMemcpyHostToDevice, ctx->in_st));
        if (host_bytes) CU(cudaMemcpyAsync(w.seq.p, host_seq + base, host_bytes, cudaMemcpyHostToDevice, ctx->in_st));
        CU(cudaEventRecord(w.ev_in, ctx->in_st));
        CU(cudaStreamWaitEvent(st, w.ev_in, 0));
        w.sb.d_seq = w.seq.as<uint8_t>();
        w.sb.d_off = w.off.as<uint64_t>();
    }
    CU(cudaEventRecord(w.ev_a0, st));
    rc = run_hash_stage(ctx, w, p, k, w.sb, w.nq, &w.codes_ptr);
    if (rc) return rc;
    CU(cudaEventRecord(w.ev_hash, st));
    w.planes = planes_for(w.sb.max_query_slots);
    w.cap = std::max<uint64_t>(1u << 20, 4ull * w.nq);
    if (w.hkeys.cap / 8 > w.cap) w.cap = w.hkeys.cap / 8;
    rc = enqueue_probes(ctx, w, p);
    if (rc) return rc;
    w.busy = true;
    return KMCPG_OK;
}

static int grow_hits(kmcpg_ctx *ctx, HitsPriv &res, uint64_t need) {
    if (res.hits.cap >= need * sizeof(kmcpg_hit)) return KMCPG_OK;
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
    cudaStream_t st = ctx->st;
    for (int attempt = 0;; attempt++) {
        CU(cudaEventSynchronize(w.ev_cnt));
        w.n_hits = w.h_cnt.as<uint64_t>()[0];
        for (size_t i = 0; i < ctx->blocks.size(); i++) {
            float a = 0, b = 0;
            cudaEventElapsedTime(&a, w.probe_ev[i * 3], w.probe_ev[i * 3 + 1]);
            cudaEventElapsedTime(&b, w.probe_ev[i * 3 + 1], w.probe_ev[i * 3 + 2]);
            tm.ms_locs += a; tm.ms_probe += b; tm.probe_launches++;
        }
        if (w.n_hits <= w.cap) break;
        if (attempt == 2) return fail(ctx, KMCPG_ENOMEM, "hit list keeps overflowing");
        // rare: the hit list overflowed.  Drain, grow, redo the probe phase of this part.
        CU(cudaStreamSynchronize(st));
        CU(cudaStreamSynchronize(ctx->copy_st));
        w.cap = w.n_hits + w.n_hits / 4 + 1024;
        int rc = enqueue_probes(ctx, w, p);
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
        size_t t3 = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, t3, w.hkeys.as<uint64_t>(), w.hkeys2.as<uint64_t>(), w.hvals.as<uint32_t>(), w.hvals2.as<uint32_t>(),
                                        (int64_t)n_hits, 0, 32 + qbits, ps);
        CU(w.tmp2.ensure(t3));
        CU(cub::DeviceRadixSort::SortPairs(w.tmp2.p, t3, w.hkeys.as<uint64_t>(), w.hkeys2.as<uint64_t>(), w.hvals.as<uint32_t>(), w.hvals2.as<uint32_t>(),
                                           (int64_t)n_hits, 0, 32 + qbits, ps));
        CU(launch_pack_hits(w.hkeys2.as<uint64_t>(), w.hvals2.as<uint32_t>(), n_hits, w.sb.query_base, w.hits.as<kmcpg_hit>(), ps));
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

static int wait_part(kmcpg_ctx *ctx, WorkSet &w) {
    if (!w.busy) return KMCPG_OK;
    CU(cudaEventSynchronize(w.ev_b));
    w.busy = false;
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
    if (priv->ctx) { pin_release(priv->ctx, priv->nk); pin_release(priv->ctx, priv->ql); pin_release(priv->ctx, priv->hits); }
    delete priv;
}

static int check_search_args(kmcpg_ctx *ctx, const kmcpg_search_params *p, const void *seq, const void *off, uint32_t n_seqs, kmcpg_hits *out, int *k) {
    if (!ctx) return KMCPG_EINVAL;
    if (!p || !out || (n_seqs && (!seq || !off))) return fail(ctx, KMCPG_EINVAL, "null argument");
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
    return KMCPG_OK;
}

// runs the two-deep pipeline over the parts
struct PartDone { uint32_t first_query, nq; uint64_t hit_dst, n_hits; cudaEvent_t ev; };

static int run_parts(kmcpg_ctx *ctx, const kmcpg_search_params &p, int k, const std::vector<Part> &parts, const uint8_t *host_seq, const uint64_t *host_off,
                     const uint8_t *d_seq, const uint64_t *d_off, HitsPriv &res, Timing &tm, kmcpg_part_cb cb, void *user) {
    const uint32_t step = p.paired ? 2 : 1;
    static const bool trace = getenv("KMCPG_TRACE") != nullptr;
    const auto T0 = std::chrono::steady_clock::now();
    auto now = [&]() { return std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - T0).count(); };
    int rc = KMCPG_OK;
    std::vector<PartDone> done;
    size_t delivered = 0;
    // hands every part whose results have reached the host to the caller, in order (up to and including `upto`)
    auto deliver = [&](size_t upto) -> int {
        for (; cb && delivered < done.size() && delivered <= upto; delivered++) {
            const PartDone &d = done[delivered];
            const float ta = now();
            CU(cudaEventSynchronize(d.ev));
            const float tb = now();
            kmcpg_part pt;
            pt.first_query = d.first_query; pt.n_queries = d.nq;
            pt.n_kmers = (const int32_t *)res.nk.p + d.first_query; pt.query_len = (const int32_t *)res.ql.p + d.first_query;
            pt.hits = (const kmcpg_hit *)res.hits.p + d.hit_dst; pt.n_hits = d.n_hits;
            cb(user, &pt);
            if (trace) fprintf(stderr, "[trace] part %zu: d2h wait %.2f..%.2f cb ..%.2f (%u q, %llu hits)\n", delivered, ta, tb, now(), d.nq, (unsigned long long)d.n_hits);
        }
        return KMCPG_OK;
    };
    for (size_t i = 0; i <= parts.size(); i++) {
        if (i < parts.size()) {
            WorkSet &w = ctx->ws[i & 1];
            // part i-2 used this work set: its result copies (copy stream) must finish before these buffers are
            // rewritten — a stream dependency, not a host wait, so the host keeps enqueueing ahead of the GPU
            if (w.busy) CU(cudaStreamWaitEvent(ctx->st, w.ev_b, 0));
            const Part &pt = parts[i];
            SubBatch sb{d_seq, d_off ? d_off + pt.a : nullptr, pt.b - pt.a, pt.slots, pt.maxq, pt.a / step};
            if (host_seq) rc = enqueue_part(ctx, w, p, k, sb, host_seq, host_off + pt.a, host_off[pt.b] - host_off[pt.a]);
            else rc = enqueue_part(ctx, w, p, k, sb, nullptr, nullptr, 0);
            if (rc) return rc;
            if (trace) fprintf(stderr, "[trace] part %zu enqueued at %.2f\n", i, now());
        }
        if (i >= 1) {
            WorkSet &w = ctx->ws[(i - 1) & 1];
            const float tf = now();
            rc = finish_probes(ctx, w, p, res, tm);
            if (rc) return rc;
            if (trace) fprintf(stderr, "[trace] part %zu probes finished: wait %.2f..%.2f\n", i - 1, tf, now());
            done.push_back({w.sb.query_base, w.nq, w.hit_dst, w.n_hits, w.ev_b});
            // its sort + copies run beside the probe of part i: hand part i-1 to the caller as soon as it has landed, while
            // the GPU keeps working
            rc = deliver(i - 1);
            if (rc) return rc;
        }
    }
    for (auto &w : ctx->ws) { rc = wait_part(ctx, w); if (rc) return rc; }
    return deliver(parts.size());
}

static void abort_parts(kmcpg_ctx *ctx) {
    cudaStreamSynchronize(ctx->st);
    cudaStreamSynchronize(ctx->copy_st);
    cudaStreamSynchronize(ctx->cnt_st);
    cudaStreamSynchronize(ctx->in_st);
    cudaStreamSynchronize(ctx->post_st);
    for (auto &w : ctx->ws) w.busy = false;
}

static int search_common(kmcpg_ctx *ctx, const kmcpg_search_params *p, int k, const uint64_t *host_off, uint32_t n_seqs, const uint8_t *host_seq,
                         const uint8_t *d_seq, const uint64_t *d_off, kmcpg_hits *out, kmcpg_part_cb cb = nullptr, void *user = nullptr) {
    auto t0 = std::chrono::steady_clock::now();
    const uint32_t launches0 = ctx->launches;
    const uint32_t step = p->paired ? 2 : 1;
    std::vector<Part> parts;
    int rc = cut_parts(ctx, host_off, n_seqs, step, k, parts);
    if (rc) return rc;
    HitsPriv *priv = new HitsPriv();
    priv->ctx = ctx; priv->nq = n_seqs / step;
    rc = pin_acquire(ctx, std::max<uint32_t>(priv->nq, 1) * 4ull, priv->nk);
    if (!rc) rc = pin_acquire(ctx, std::max<uint32_t>(priv->nq, 1) * 4ull, priv->ql);
    if (!rc) rc = pin_acquire(ctx, std::max<uint64_t>(1u << 16, 2ull * priv->nq) * sizeof(kmcpg_hit), priv->hits);
    Timing tm;
    if (!rc) rc = run_parts(ctx, *p, k, parts, host_seq, host_off, d_seq, d_off, *priv, tm, cb, user);
    if (rc) { abort_parts(ctx); drop_priv(priv); return rc; }
    float ms_total = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    fill_out(out, priv, tm, ms_total, ctx->launches - launches0);
    return KMCPG_OK;
}

}  // namespace kmcpg

// ========================================================================================================
// C ABI
// ========================================================================================================
extern "C" {

int kmcpg_abi_version(void) { return KMCPG_ABI_VERSION; }

const char *kmcpg_last_error(const kmcpg_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int kmcpg_create(int device, kmcpg_ctx **out) {
    if (!out) return fail(nullptr, KMCPG_EINVAL, "out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        (void)cudaGetLastError();
        return fail(nullptr, KMCPG_ECUDA, std::string("no CUDA device available (") + (e == cudaSuccess ? "device count 0" : cudaGetErrorString(e)) +
                                              "); libkmcp_gpu has no CPU fallback");
    }
    if (device < 0 || device >= n) return fail(nullptr, KMCPG_EINVAL, "device ordinal out of range");
    kmcpg_ctx *ctx = new kmcpg_ctx();
    ctx->device = device;
    auto bail = [&](cudaError_t err, const char *what) {
        std::string m = std::string(what) + ": " + cudaGetErrorString(err);
        delete ctx;
        return fail(nullptr, KMCPG_ECUDA, m);
    };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bail(e, "cudaSetDevice");
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return bail(e, "cudaGetDeviceProperties");
    ctx->sm_count = prop.multiProcessorCount;
    if ((e = cudaStreamCreateWithFlags(&ctx->own_st, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    if ((e = cudaStreamCreateWithFlags(&ctx->copy_st, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    if ((e = cudaStreamCreateWithFlags(&ctx->cnt_st, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    if ((e = cudaStreamCreateWithFlags(&ctx->in_st, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    {   // the post stream gets the highest priority: its tiny sort/pack kernels slip in as soon as probe CTAs retire
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        if ((e = cudaStreamCreateWithPriority(&ctx->post_st, cudaStreamNonBlocking, hi)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    }
    ctx->st = ctx->own_st;
    if ((e = ctx->h_small.ensure(256)) != cudaSuccess) return bail(e, "cudaMallocHost");
    *out = ctx;
    return KMCPG_OK;
}

int kmcpg_set_stream(kmcpg_ctx *ctx, void *stream) {
    if (!ctx) return KMCPG_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->st);
    ctx->st = stream ? (cudaStream_t)stream : ctx->own_st;
    return KMCPG_OK;
}

int kmcpg_close(kmcpg_ctx *ctx) {
    if (!ctx) return KMCPG_EINVAL;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->st);
    cudaStreamSynchronize(ctx->copy_st);
    cudaStreamSynchronize(ctx->cnt_st);
    cudaStreamSynchronize(ctx->in_st);
    cudaStreamSynchronize(ctx->post_st);
    free_db(ctx);
    for (auto &w : ctx->ws) w.release();
    for (DevBuf *b : {&ctx->d_tmp, &ctx->d_dense, &ctx->d_scal, &ctx->d_genome}) b->release();
    ctx->h_stage.release(); ctx->h_small.release();
    for (auto &b : ctx->pin_pool) cudaFreeHost(b.p);
    ctx->pin_pool.clear();
    if (ctx->own_st) cudaStreamDestroy(ctx->own_st);
    if (ctx->copy_st) cudaStreamDestroy(ctx->copy_st);
    if (ctx->cnt_st) cudaStreamDestroy(ctx->cnt_st);
    if (ctx->in_st) cudaStreamDestroy(ctx->in_st);
    if (ctx->post_st) cudaStreamDestroy(ctx->post_st);
    delete ctx;
    return KMCPG_OK;
}

int kmcpg_open_db(kmcpg_ctx *ctx, const char *dir, const kmcpg_db_opts *opts) {
    if (!ctx || !dir) return KMCPG_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    free_db(ctx);
    std::string err;
    int rc = read_db_meta(dir, ctx->meta, err);
    if (rc) return fail(ctx, rc, err);
    const DbMeta &m = ctx->meta;
    if (m.num_hashes < 1 || m.num_hashes > 4) return fail(ctx, KMCPG_EFORMAT, "number of hashes must be 1..4");   // I:196
    int world = opts && opts->shard_world > 1 ? opts->shard_world : 1;
    int rank = opts && world > 1 ? opts->shard_rank : 0;
    if (rank < 0 || rank >= world) return fail(ctx, KMCPG_EINVAL, "shard_rank out of range");
    std::vector<ShardPiece> pieces;
    std::vector<uint64_t> load;
    plan_pieces(m, world, pieces, load);
    if (opts && opts->max_resident_bytes > 0 && (int64_t)load[rank] > opts->max_resident_bytes)
        return fail(ctx, KMCPG_ENOMEM, "resident blocks exceed max_resident_bytes");
    ctx->resident_of.assign(m.blocks.size(), -1);
    const size_t CHUNK = 256ull << 20;
    CU(ctx->h_stage.ensure(CHUNK));
    for (const ShardPiece &pc : pieces) {
        if (pc.shard != rank) continue;
        const size_t i = (size_t)pc.block;
        const BlockMeta &bm = m.blocks[i];
        if (bm.num_sigs >= (1ull << 32) - 1) return fail(ctx, KMCPG_EUNSUPPORTED, "blocks with >= 2^32-1 signatures are not supported");
        DeviceBlock b;
        b.meta_idx = (int)i;
        layout_block(b, bm, pc.col0, pc.n_cols);
        b.bytes = (size_t)bm.num_sigs * b.pitch;
        CU(cudaMalloc((void **)&b.d_rows, std::max<size_t>(b.bytes, 16)));
        ctx->blocks.push_back(b);
        if (ctx->resident_of[i] < 0) ctx->resident_of[i] = (int)ctx->blocks.size() - 1;
        const int fd = ::open(bm.path.c_str(), O_RDONLY);
        if (fd < 0) return fail(ctx, KMCPG_EIO, "cannot open " + bm.path);
        struct FdGuard { int fd; ~FdGuard() { ::close(fd); } } guard{fd};
        const uint64_t rows_per_chunk = std::max<uint64_t>(1, CHUNK / (uint64_t)bm.row_bytes);
        if ((uint64_t)bm.row_bytes > CHUNK) return fail(ctx, KMCPG_EUNSUPPORTED, "row wider than the staging buffer");
        // a chunk is read by several streams at once (one fread stream does ~3 GB/s from the page cache, far below the H2D copy)
        const int io_threads = std::max(1, std::min(8, (int)std::thread::hardware_concurrency() / 4));
        for (uint64_t r0 = 0; r0 < bm.num_sigs; r0 += rows_per_chunk) {
            uint64_t nr = std::min<uint64_t>(rows_per_chunk, bm.num_sigs - r0);
            size_t bytes = (size_t)nr * bm.row_bytes;
            if (!pread_parallel(fd, ctx->h_stage.p, bytes, bm.data_offset + r0 * (uint64_t)bm.row_bytes, io_threads))
                return fail(ctx, KMCPG_EIO, "kmcp: truncated index file: " + bm.path);
            cudaError_t e = ctx->d_tmp.ensure(bytes);
            if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->d_tmp.p, ctx->h_stage.p, bytes, cudaMemcpyHostToDevice, ctx->st);
            // whole rows travel; the re-pitch kernel keeps the bytes [col0/8, col0/8 + row_bytes) of every row
            if (e == cudaSuccess) e = launch_repitch_cols(ctx->d_tmp.as<uint8_t>(), b.d_rows + r0 * b.pitch, nr, (uint32_t)bm.row_bytes, b.col0 / 8, b.row_bytes, b.pitch, ctx->st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->st);
            CU(e);
        }
        ctx->sum_row_bytes += b.row_bytes;
        ctx->resident_bytes += (int64_t)b.bytes;
        ctx->disk_bytes += (int64_t)(bm.num_sigs * (uint64_t)b.row_bytes);
    }
    ctx->target_sizes.resize((size_t)m.n_targets);
    for (auto &bm : m.blocks)
        for (int c = 0; c < bm.n_names; c++) ctx->target_sizes[(size_t)bm.target_base + c] = (double)bm.sizes[c];
    ctx->has_db = true;
    return KMCPG_OK;
}

int kmcpg_shard_plan(const char *dir, int shard_world, int32_t *owner_out, int32_t n_owner) {
    if (!dir || !owner_out || shard_world < 1) return KMCPG_EINVAL;
    DbMeta m;
    std::string err;
    int rc = read_db_meta(dir, m, err);
    if (rc) return fail(nullptr, rc, err);
    if ((size_t)n_owner < m.blocks.size()) return fail(nullptr, KMCPG_EINVAL, "owner array too small");
    std::vector<ShardPiece> pieces;
    std::vector<uint64_t> load;
    plan_pieces(m, shard_world, pieces, load);
    for (size_t i = 0; i < m.blocks.size(); i++) owner_out[i] = -1;
    for (const ShardPiece &pc : pieces) if (owner_out[pc.block] < 0) owner_out[pc.block] = pc.shard;   // a split block: the shard of its first columns
    return (int)m.blocks.size();
}

int kmcpg_shard_pieces(const char *dir, int shard_world, kmcpg_shard_piece *out, int32_t cap) {
    if (!dir || shard_world < 1 || (cap > 0 && !out)) return KMCPG_EINVAL;
    DbMeta m;
    std::string err;
    int rc = read_db_meta(dir, m, err);
    if (rc) return fail(nullptr, rc, err);
    std::vector<ShardPiece> pieces;
    std::vector<uint64_t> load;
    plan_pieces(m, shard_world, pieces, load);
    if ((size_t)cap < pieces.size()) return fail(nullptr, KMCPG_EINVAL, "piece array too small");
    for (size_t i = 0; i < pieces.size(); i++) {
        out[i].block = pieces[i].block; out[i].shard = pieces[i].shard; out[i].col0 = pieces[i].col0; out[i].n_cols = pieces[i].n_cols;
        out[i].resident_bytes = m.blocks[pieces[i].block].num_sigs * (uint64_t)pitch_for((pieces[i].n_cols + 7) / 8);
    }
    return (int)pieces.size();
}

int kmcpg_db_info(const kmcpg_ctx *ctx, kmcpg_db_info_t *o) {
    if (!ctx || !o || !ctx->has_db) return KMCPG_EINVAL;
    const DbMeta &m = ctx->meta;
    memset(o, 0, sizeof(*o));
    o->n_ks = (int32_t)std::min<size_t>(m.ks.size(), 8);
    for (int i = 0; i < o->n_ks; i++) o->ks[i] = m.ks[i];
    o->canonical = m.canonical; o->num_hashes = m.num_hashes; o->scaled = m.scaled; o->scale = m.scale;
    o->minimizer = m.minimizer; o->minimizer_w = m.minimizer_w; o->syncmer = m.syncmer; o->syncmer_s = m.syncmer_s;
    o->fpr = m.fpr; o->n_blocks = (int32_t)m.blocks.size(); o->n_resident_blocks = (int32_t)ctx->blocks.size();
    o->n_targets = m.n_targets; o->sum_row_bytes = ctx->sum_row_bytes; o->resident_bytes = ctx->resident_bytes; o->disk_bytes = ctx->disk_bytes;
    return KMCPG_OK;
}

int kmcpg_target(const kmcpg_ctx *ctx, int64_t g, kmcpg_target_t *o) {
    if (!ctx || !o || !ctx->has_db) return KMCPG_EINVAL;
    const auto &bl = ctx->meta.blocks;
    size_t lo = 0, hi = bl.size();
    while (lo + 1 < hi) { size_t mid = (lo + hi) / 2; if (bl[mid].target_base <= g) lo = mid; else hi = mid; }
    if (bl.empty() || g < bl[lo].target_base || g >= bl[lo].target_base + bl[lo].n_names) return KMCPG_EINVAL;
    int c = (int)(g - bl[lo].target_base);
    o->name = bl[lo].names[c].c_str(); o->index = bl[lo].indices[c]; o->genome_size = bl[lo].gsizes[c]; o->n_kmers = bl[lo].sizes[c];
    o->block = (int32_t)lo; o->col = c; o->resident = 0;
    for (const DeviceBlock &b : ctx->blocks)
        if (b.meta_idx == (int)lo && (uint32_t)c >= b.col0 && (uint32_t)c < b.col0 + b.n_cols) { o->resident = 1; break; }
    return KMCPG_OK;
}

// internal (engine.cpp): 1 when every block of the DB is resident here with all of its columns (not a shard)
int kmcpg_internal_holds_whole_db(const kmcpg_ctx *ctx) {
    if (!ctx || !ctx->has_db) return 0;
    size_t whole = 0;
    for (const DeviceBlock &b : ctx->blocks) whole += b.whole;
    return whole == ctx->meta.blocks.size() && ctx->blocks.size() == ctx->meta.blocks.size();
}

// internal (engine.cpp): Sizes[t] of every target as float64, valid while the DB is open
const double *kmcpg_internal_target_sizes(const kmcpg_ctx *ctx) { return ctx && ctx->has_db ? ctx->target_sizes.data() : nullptr; }

void kmcpg_default_params(kmcpg_search_params *p) {
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->min_query_len = 30; p->min_matched = 10; p->dedup_threshold = 256; p->min_query_cov = 0.55;   // S:1055-1069
}

int kmcpg_search_batch(kmcpg_ctx *ctx, const kmcpg_search_params *p, const uint8_t *seq, const uint64_t *off, uint32_t n_seqs, kmcpg_hits *out) {
    int k = 0;
    int rc = check_search_args(ctx, p, seq, off, n_seqs, out, &k);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    memset(out, 0, sizeof(*out));
    return search_common(ctx, p, k, off, n_seqs, seq ? seq : (const uint8_t *)"", nullptr, nullptr, out);
}

int kmcpg_search_batch_cb(kmcpg_ctx *ctx, const kmcpg_search_params *p, const uint8_t *seq, const uint64_t *off, uint32_t n_seqs, kmcpg_part_cb cb, void *user,
                          kmcpg_hits *out) {
    int k = 0;
    int rc = check_search_args(ctx, p, seq, off, n_seqs, out, &k);
    if (rc) return rc;
    if (!cb) return fail(ctx, KMCPG_EINVAL, "callback is NULL");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    memset(out, 0, sizeof(*out));
    return search_common(ctx, p, k, off, n_seqs, seq ? seq : (const uint8_t *)"", nullptr, nullptr, out, cb, user);
}

int kmcpg_search_batch_device(kmcpg_ctx *ctx, const kmcpg_search_params *p, const uint8_t *d_seq, const uint64_t *d_off, uint32_t n_seqs,
                              uint64_t seq_bytes, kmcpg_hits *out) {
    int k = 0;
    int rc = check_search_args(ctx, p, d_seq, d_off, n_seqs, out, &k);
    if (rc) return rc;
    (void)seq_bytes;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    memset(out, 0, sizeof(*out));
    // the lengths live on the device: fetch the offsets once (8 B per sequence) to cut the batch into parts
    PinBuf hb;
    rc = pin_acquire(ctx, ((size_t)n_seqs + 1) * 8, hb);
    if (rc) return rc;
    uint64_t *hoff = (uint64_t *)hb.p;
    hoff[0] = 0;
    if (n_seqs) {
        cudaError_t e = cudaMemcpyAsync(hoff, d_off, (n_seqs + 1) * 8ull, cudaMemcpyDeviceToHost, ctx->st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->st);
        if (e != cudaSuccess) { pin_release(ctx, hb); CU(e); }
    }
    rc = search_common(ctx, p, k, hoff, n_seqs, nullptr, d_seq, d_off, out);
    pin_release(ctx, hb);
    return rc;
}

void kmcpg_free_hits(kmcpg_hits *h) {
    if (!h) return;
    drop_priv((HitsPriv *)h->_priv);
    memset(h, 0, sizeof(*h));
}

int kmcpg_host_alloc(void **p, size_t bytes) {
    if (!p) return KMCPG_EINVAL;
    cudaError_t e = cudaMallocHost(p, bytes ? bytes : 1);
    if (e != cudaSuccess) { (void)cudaGetLastError(); g_create_error = cudaGetErrorString(e); return KMCPG_ENOMEM; }
    return KMCPG_OK;
}
int kmcpg_host_free(void *p) { return cudaFreeHost(p) == cudaSuccess ? KMCPG_OK : KMCPG_ECUDA; }

int kmcpg_device_memory(kmcpg_ctx *ctx, size_t *free_bytes, size_t *total_bytes) {
    if (!ctx || !free_bytes || !total_bytes) return KMCPG_EINVAL;
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemGetInfo(free_bytes, total_bytes));
    return KMCPG_OK;
}

int kmcpg_device_alloc(kmcpg_ctx *ctx, void **p, size_t bytes) {
    if (!ctx || !p) return KMCPG_EINVAL;
    CU(cudaSetDevice(ctx->device));
    CU(cudaMalloc(p, bytes ? bytes : 1));
    return KMCPG_OK;
}
int kmcpg_device_free(kmcpg_ctx *ctx, void *p) {
    if (!ctx) return KMCPG_EINVAL;
    CU(cudaSetDevice(ctx->device));
    CU(cudaFree(p));
    return KMCPG_OK;
}
int kmcpg_memcpy_h2d(kmcpg_ctx *ctx, void *d, const void *h, size_t bytes) {
    if (!ctx) return KMCPG_EINVAL;
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    return KMCPG_OK;
}
int kmcpg_memcpy_d2h(kmcpg_ctx *ctx, void *h, const void *d, size_t bytes) {
    if (!ctx) return KMCPG_EINVAL;
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    return KMCPG_OK;
}

void kmcpg_free(void *p) { free(p); }

int kmcpg_generate_kmers(kmcpg_ctx *ctx, const kmcpg_sketch_params *sp, const uint8_t *seq, const uint64_t *off, uint32_t n_seqs,
                         uint64_t **out_codes, uint64_t **out_off) {
    if (!ctx || !sp || !out_codes || !out_off || (n_seqs && (!seq || !off))) return KMCPG_EINVAL;
    if (sp->k < 1 || sp->k > 64) return fail(ctx, KMCPG_EUNSUPPORTED, "k must be in 1..64");
    if (sp->syncmer && (sp->syncmer_s < 1 || (int)sp->syncmer_s >= sp->k)) return fail(ctx, KMCPG_EINVAL, "syncmer_s must be in 1..k-1");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    // borrow the pipeline's hash stage with a throw-away DbMeta view
    DbMeta saved = ctx->meta;
    DbMeta &m = ctx->meta;
    m.canonical = sp->canonical; m.scaled = sp->scaled; m.scale = sp->scale; m.minimizer = sp->minimizer; m.minimizer_w = sp->minimizer_w;
    m.syncmer = sp->syncmer; m.syncmer_s = sp->syncmer_s;
    kmcpg_search_params p;
    kmcpg_default_params(&p);
    p.min_query_len = 0; p.min_matched = 1; p.dedup_threshold = 0x7fffffff; p.min_query_cov = 0;
    uint64_t total = 0, mx = 0;
    for (uint32_t i = 0; i < n_seqs; i++) {
        uint64_t len = off[i + 1] - off[i];
        uint64_t s = len >= (uint64_t)sp->k ? len - sp->k + 1 : 0;
        total += s; mx = std::max(mx, s);
    }
    const uint64_t nbytes = n_seqs ? off[n_seqs] - off[0] : 0;
    int rc = KMCPG_OK;
    uint64_t *codes = nullptr;
    std::vector<uint64_t> ho(n_seqs + 1);
    for (uint32_t i = 0; i <= n_seqs; i++) ho[i] = n_seqs ? off[i] - off[0] : 0;
    WorkSet &w = ctx->ws[0];
    auto body = [&]() -> int {
        CU(w.off.ensure((n_seqs + 1) * 8ull));
        CU(w.seq.ensure(nbytes + 64));
        CU(cudaMemcpyAsync(w.off.p, ho.data(), (n_seqs + 1) * 8ull, cudaMemcpyHostToDevice, ctx->st));
        if (nbytes) CU(cudaMemcpyAsync(w.seq.p, seq + off[0], nbytes, cudaMemcpyHostToDevice, ctx->st));
        SubBatch sb{w.seq.as<uint8_t>(), w.off.as<uint64_t>(), n_seqs, total, mx, 0};
        int r = run_hash_stage(ctx, w, p, sp->k, sb, n_seqs, &codes);
        if (r) return r;
        std::vector<uint32_t> nc(n_seqs);
        std::vector<uint64_t> so(n_seqs + 1);
        CU(cudaMemcpyAsync(nc.data(), w.ncodes.p, n_seqs * 4ull, cudaMemcpyDeviceToHost, ctx->st));
        CU(cudaMemcpyAsync(so.data(), w.slot_off.p, (n_seqs + 1) * 8ull, cudaMemcpyDeviceToHost, ctx->st));
        CU(cudaStreamSynchronize(ctx->st));
        std::vector<uint64_t> all(total ? total : 1);
        if (total) {
            CU(cudaMemcpyAsync(all.data(), codes, total * 8, cudaMemcpyDeviceToHost, ctx->st));
            CU(cudaStreamSynchronize(ctx->st));
        }
        uint64_t *oo = (uint64_t *)malloc((n_seqs + 1) * 8ull);
        uint64_t sum = 0;
        for (uint32_t i = 0; i < n_seqs; i++) { oo[i] = sum; sum += nc[i] == 0xFFFFFFFFu ? 0 : nc[i]; }
        oo[n_seqs] = sum;
        uint64_t *oc = (uint64_t *)malloc(std::max<uint64_t>(sum, 1) * 8);
        for (uint32_t i = 0; i < n_seqs; i++) memcpy(oc + oo[i], all.data() + so[i], (oo[i + 1] - oo[i]) * 8);
        *out_codes = oc; *out_off = oo;
        return KMCPG_OK;
    };
    rc = n_seqs ? body() : KMCPG_OK;
    if (!n_seqs) { *out_codes = (uint64_t *)malloc(8); *out_off = (uint64_t *)calloc(1, 8); }
    std::string keep_err = ctx->err;
    ctx->meta = saved;
    ctx->err = keep_err;
    return rc;
}

int kmcpg_count_codes(kmcpg_ctx *ctx, const uint64_t *codes, uint64_t n, uint32_t *counts) {
    if (!ctx || !counts || (n && !codes)) return KMCPG_EINVAL;
    if (!ctx->has_db) return fail(ctx, KMCPG_EINVAL, "no database open");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    const int H = ctx->meta.num_hashes;
    const uint64_t nt = (uint64_t)ctx->meta.n_targets;
    cudaStream_t st = ctx->st;
    WorkSet &w = ctx->ws[0];
    CU(ctx->d_dense.ensure(std::max<uint64_t>(nt, 1) * 4));
    CU(cudaMemsetAsync(ctx->d_dense.p, 0, std::max<uint64_t>(nt, 1) * 4, st));
    if (n > 0) {
        if (n >= (1ull << 32) - 1) return fail(ctx, KMCPG_EUNSUPPORTED, "too many codes");
        CU(w.codes.ensure(n * 8)); CU(w.locs.ensure(n * 4 * H)); CU(w.slot_off.ensure(16));
        CU(w.neff.ensure(4)); CU(w.thresh.ensure(4)); CU(w.counters.ensure(64));
        CU(w.hkeys.ensure(8)); CU(w.hvals.ensure(4));
        uint64_t so[2] = {0, n};
        uint32_t neff = (uint32_t)n, th = 0xFFFFFFFFu;
        CU(cudaMemcpyAsync(w.codes.p, codes, n * 8, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(w.slot_off.p, so, 16, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(w.neff.p, &neff, 4, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(w.thresh.p, &th, 4, cudaMemcpyHostToDevice, st));
        CU(cudaMemsetAsync(w.counters.p, 0, 8, st));
        CU(cudaStreamSynchronize(st));
        for (auto &b : ctx->blocks) {
            const BlockMeta &bm = ctx->meta.blocks[b.meta_idx];
            CU(launch_locs(w.codes.as<uint64_t>(), n, H, b.fm, w.locs.as<uint32_t>(), st));
            ProbeArgs pa;
            memset(&pa, 0, sizeof(pa));
            pa.rows = b.d_rows; pa.pitch = b.pitch; pa.row_bytes = b.row_bytes;
            pa.n_names = b.n_cols; pa.target_base = (uint32_t)(bm.target_base + b.col0); pa.num_hashes = H;
            pa.locs = w.locs.as<uint32_t>(); pa.slot_off = w.slot_off.as<uint64_t>();
            pa.n_eff = w.neff.as<uint32_t>(); pa.thresh = w.thresh.as<uint32_t>(); pa.n_queries = 1; pa.paired = 0;
            pa.hit_keys = w.hkeys.as<uint64_t>(); pa.hit_vals = w.hvals.as<uint32_t>();
            pa.hit_count = w.counters.as<unsigned long long>(); pa.hit_cap = 0; pa.dense_counts = ctx->d_dense.as<uint32_t>();
            pa.planes = planes_for(n);
            CU(launch_probe(pa, ctx->sm_count, st));
        }
    }
    CU(cudaMemcpyAsync(counts, ctx->d_dense.p, nt * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return KMCPG_OK;
}

}  // extern "C"

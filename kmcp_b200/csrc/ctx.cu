// ctx.cu — context, database residency in HBM, the batch executor, and the C ABI of include/kmcp_gpu.h.
//
// Data layout in HBM (per resident block): the on-disk row-major bit matrix numSigs × numRowBytes
// (X:307-349; row r byte i bit 7-j ⇔ Bloom bit r of target 8i+j) is re-pitched on upload so that every
// row starts 16-byte aligned (128-byte aligned for rows wider than 256 B) and is zero padded; the disk
// format is untouched.
//
// Batch executor: a call is cut into parts (≤ 32 M k-mer slots each).  ALL kernels run in order on one
// compute stream, so every launch can be timed exactly with events; transfers run on a copy stream.
// Part i goes  [H2D] → slot scan → hash → (sort+unique of long queries) → verdict → per block probe (row indices derived in the kernel)
// and, once the host has read its hit count, → radix sort of the hit list → pack → [D2H into pinned results].
// Two work sets alternate, and the kernels of part i+1 are enqueued BEFORE the host waits for part i's hit
// count, so the round trip and the result copies hide behind the next part's probe kernel.
#include <cuda_runtime.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
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
#include "nvtx_ranges.h"

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

int64_t widest_shard_row_bytes(const std::vector<ShardPiece> &pieces, int world) {
    std::vector<int64_t> rb((size_t)std::max(1, world), 0);
    for (const ShardPiece &pc : pieces) if (pc.shard >= 0 && pc.shard < (int)rb.size()) rb[pc.shard] += (pc.n_cols + 7) / 8;
    return *std::max_element(rb.begin(), rb.end());
}

void free_db(kmcpg_ctx *ctx) {
    for (auto &b : ctx->blocks) if (b.d_rows) cudaFree(b.d_rows);
    ctx->blocks.clear();
    ctx->resident_of.clear();
    ctx->target_sizes.clear();
    ctx->has_db = false;
    ctx->sum_row_bytes = ctx->resident_bytes = ctx->disk_bytes = ctx->part_row_bytes = 0;
}

void WorkSet::release() {
    for (DevBuf *b : {&seq, &off, &slot_cnt, &slot_off, &codes, &codes2, &locs[0], &locs[1], &ncodes, &qlen, &nk, &neff, &thresh, &hkeys, &hvals, &hkeys2, &hvals2,
                      &hits, &counters, &tmp, &tmp2, &segb, &sege, &ck, &cs, &cs_cnt, &cs_off, &tile_n, &tile_off, &tile_cnt, &tile_pre, &order, &order_in, &neff_sorted})
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
    {   // query preparation of the next part runs beside the probes of the current one, at the probe stream's own (lowest) priority:
        // its CTAs take the SM slots the probe's last wave leaves idle; ahead of the queued probe CTAs they only displace them (measured:
        // 27.1 ms per C2 step at the same priority, 28.0 ms at a higher one, 27.6 ms on one stream; profiles/README.md round 2)
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        int prio = lo;
#ifdef KMCPG_DEV
        if (const char *v = getenv("KMCPG_HASH_PRIO")) prio = !strcmp(v, "low") ? lo : (!strcmp(v, "high") ? hi : (lo + hi) / 2);
#endif
        if ((e = cudaStreamCreateWithPriority(&ctx->hash_st, cudaStreamNonBlocking, prio)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    }
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
    executor_drain(ctx);
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->st);
    ctx->st = stream ? (cudaStream_t)stream : ctx->own_st;
    return KMCPG_OK;
}

int kmcpg_close(kmcpg_ctx *ctx) {
    if (!ctx) return KMCPG_EINVAL;
    executor_stop(ctx);
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->st);
    cudaStreamSynchronize(ctx->hash_st);
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
    for (auto &b : ctx->stage_pool) b.release();
    ctx->stage_pool.clear();
    if (ctx->own_st) cudaStreamDestroy(ctx->own_st);
    if (ctx->copy_st) cudaStreamDestroy(ctx->copy_st);
    if (ctx->cnt_st) cudaStreamDestroy(ctx->cnt_st);
    if (ctx->in_st) cudaStreamDestroy(ctx->in_st);
    if (ctx->hash_st) cudaStreamDestroy(ctx->hash_st);
    if (ctx->post_st) cudaStreamDestroy(ctx->post_st);
    delete ctx;
    return KMCPG_OK;
}

int kmcpg_open_db(kmcpg_ctx *ctx, const char *dir, const kmcpg_db_opts *opts) {
    if (!ctx || !dir) return KMCPG_EINVAL;
    NvtxRange nvtx("kmcpg:open_db (pread, H2D, re-pitch)");
    std::lock_guard<std::mutex> lk(ctx->mu);
    executor_drain(ctx);
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
    ctx->part_row_bytes = widest_shard_row_bytes(pieces, world);
    if (opts && opts->max_resident_bytes > 0 && (int64_t)load[rank] > opts->max_resident_bytes)
        return fail(ctx, KMCPG_ENOMEM, "resident blocks exceed max_resident_bytes");
    ctx->resident_of.assign(m.blocks.size(), -1);
    const size_t CHUNK = 256ull << 20;
    CU(ctx->h_stage.ensure(CHUNK));
    for (const ShardPiece &pc : pieces) {
        if (pc.shard != rank) continue;
        const size_t i = (size_t)pc.block;
        const BlockMeta &bm = m.blocks[i];
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

int kmcpg_host_alloc(void **p, size_t bytes) {
    if (!p) return KMCPG_EINVAL;
    cudaError_t e = cudaMallocHost(p, bytes ? bytes : 1);
    if (e != cudaSuccess) { (void)cudaGetLastError(); g_create_error = cudaGetErrorString(e); return KMCPG_ENOMEM; }
    return KMCPG_OK;
}
int kmcpg_host_free(void *p) { return cudaFreeHost(p) == cudaSuccess ? KMCPG_OK : KMCPG_ECUDA; }

// ---- host memory shared between the processes of a one-process-per-GPU run ------------------------------------------------
static std::string shm_path(const char *name) {
    std::string n = name;
    for (char &c : n) if (c == '/') c = '_';
    struct stat st;
    const char *dir = (stat("/dev/shm", &st) == 0 && S_ISDIR(st.st_mode)) ? "/dev/shm/" : "/tmp/";
    return std::string(dir) + n;
}

int kmcpg_shm_open(const char *name, size_t bytes, int create, int cuda_register, void **ptr) {
    if (!name || !*name || !ptr || bytes == 0) return fail(nullptr, KMCPG_EINVAL, "bad shared-memory arguments");
    *ptr = nullptr;
    const std::string path = shm_path(name);
    const int fd = ::open(path.c_str(), create ? (O_RDWR | O_CREAT) : O_RDWR, 0600);
    if (fd < 0) return fail(nullptr, KMCPG_EIO, "cannot open shared-memory segment " + path);
    if (create && ftruncate(fd, (off_t)bytes) != 0) { ::close(fd); return fail(nullptr, KMCPG_EIO, "cannot size shared-memory segment " + path); }
    struct stat st;
    if (fstat(fd, &st) != 0 || (size_t)st.st_size < bytes) { ::close(fd); return fail(nullptr, KMCPG_EIO, "shared-memory segment is smaller than asked for: " + path); }
    void *p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED | (create ? MAP_POPULATE : 0), fd, 0);
    ::close(fd);
    if (p == MAP_FAILED) return fail(nullptr, KMCPG_ENOMEM, "mmap of shared-memory segment failed: " + path);
    if (cuda_register) {
        cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            munmap(p, bytes);
            return fail(nullptr, KMCPG_ECUDA, std::string("cudaHostRegister of a shared-memory segment failed: ") + cudaGetErrorString(e));
        }
    }
    *ptr = p;
    return KMCPG_OK;
}

int kmcpg_shm_close(const char *name, void *ptr, size_t bytes, int cuda_registered, int unlink_it) {
    if (ptr) {
        if (cuda_registered) { cudaHostUnregister(ptr); (void)cudaGetLastError(); }
        munmap(ptr, bytes);
    }
    if (unlink_it && name && *name) ::unlink(shm_path(name).c_str());
    return KMCPG_OK;
}

// ---- one batch staged for several contexts of this process (kmcpg_engine_search_sharded) ----------------------------------
// The batch crosses PCIe ONCE, into the first context's device; the other contexts get it by peer-to-peer copies (NVLink), piece by
// piece, so every context can start on piece j while piece j+1 is still travelling.  ready(i, j) is the event after which piece j
// is complete on context i (kmcpg_batch.ready_event of the job that searches it).
struct kmcpg_stage {
    std::vector<kmcpg_ctx *> ctxs;
    std::vector<DevBuf> seq, off;
    std::vector<std::vector<cudaEvent_t>> ready;      // [ctx][piece]
    std::vector<uint64_t> rebased;                    // offsets made relative to the first byte staged
};

static void stage_release(kmcpg_stage *st) {
    for (size_t i = 0; i < st->ctxs.size(); i++) {
        cudaSetDevice(st->ctxs[i]->device);
        for (cudaEvent_t e : st->ready[i]) cudaEventDestroy(e);
        std::lock_guard<std::mutex> lk(st->ctxs[i]->pin_mu);
        if (st->seq[i].p) st->ctxs[i]->stage_pool.push_back(st->seq[i]);
        if (st->off[i].p) st->ctxs[i]->stage_pool.push_back(st->off[i]);
        while (st->ctxs[i]->stage_pool.size() > 8) { st->ctxs[i]->stage_pool.front().release(); st->ctxs[i]->stage_pool.erase(st->ctxs[i]->stage_pool.begin()); }
    }
    delete st;
}

static cudaError_t stage_buffer(kmcpg_ctx *c, size_t bytes, DevBuf &out) {
    {
        std::lock_guard<std::mutex> lk(c->pin_mu);
        int best = -1;
        for (size_t i = 0; i < c->stage_pool.size(); i++)
            if (c->stage_pool[i].cap >= bytes && (best < 0 || c->stage_pool[i].cap < c->stage_pool[best].cap)) best = (int)i;
        if (best >= 0) { out = c->stage_pool[best]; c->stage_pool.erase(c->stage_pool.begin() + best); return cudaSuccess; }
    }
    return out.ensure(bytes);
}

int kmcpg_internal_stage_begin(kmcpg_ctx *const *ctxs, int n_ctx, const uint8_t *seq, const uint64_t *off, uint32_t n_seqs, const uint32_t *cuts, int n_pieces,
                               kmcpg_stage **out) {
    if (!ctxs || n_ctx < 1 || !seq || !off || !cuts || n_pieces < 1 || !out || cuts[0] != 0 || cuts[n_pieces] != n_seqs) return KMCPG_EINVAL;
    kmcpg_ctx *ctx = ctxs[0];                         // errors are reported on the first context
    kmcpg_stage *st = new kmcpg_stage();
    st->ctxs.assign(ctxs, ctxs + n_ctx);
    st->seq.resize(n_ctx); st->off.resize(n_ctx); st->ready.resize(n_ctx);
    const uint64_t base = off[0], bytes = off[n_seqs] - base;
    const uint64_t *hoff = off;
    if (base) {
        st->rebased.resize((size_t)n_seqs + 1);
        for (uint32_t i = 0; i <= n_seqs; i++) st->rebased[i] = off[i] - base;
        hoff = st->rebased.data();
    }
    auto body = [&]() -> int {
        for (int i = 0; i < n_ctx; i++) {
            kmcpg_ctx *c = ctxs[i];
            CU(cudaSetDevice(c->device));
            CU(stage_buffer(c, bytes + 64, st->seq[i]));
            CU(stage_buffer(c, ((size_t)n_seqs + 1) * 8, st->off[i]));
            st->ready[i].resize(n_pieces);
            for (int j = 0; j < n_pieces; j++) CU(cudaEventCreateWithFlags(&st->ready[i][j], cudaEventDisableTiming));
            if (i > 0 && c->device != ctxs[0]->device) {
                int can = 0;
                cudaDeviceCanAccessPeer(&can, c->device, ctxs[0]->device);
                if (can) { cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[0]->device, 0); if (e != cudaSuccess) (void)cudaGetLastError(); }   // "already enabled" is fine
            }
        }
        kmcpg_ctx *r = ctxs[0];
        for (int j = 0; j < n_pieces; j++) {
            const uint64_t o0 = hoff[cuts[j]], o1 = hoff[cuts[j + 1]];
            CU(cudaSetDevice(r->device));
            if (j == 0) CU(cudaMemcpyAsync(st->off[0].p, hoff, ((size_t)n_seqs + 1) * 8, cudaMemcpyHostToDevice, r->in_st));
            if (o1 > o0) CU(cudaMemcpyAsync((uint8_t *)st->seq[0].p + o0, seq + base + o0, o1 - o0, cudaMemcpyHostToDevice, r->in_st));
            CU(cudaEventRecord(st->ready[0][j], r->in_st));
            for (int i = 1; i < n_ctx; i++) {
                kmcpg_ctx *c = ctxs[i];
                CU(cudaSetDevice(c->device));
                CU(cudaStreamWaitEvent(c->in_st, st->ready[0][j], 0));
                if (j == 0) CU(cudaMemcpyPeerAsync(st->off[i].p, c->device, st->off[0].p, r->device, ((size_t)n_seqs + 1) * 8, c->in_st));
                if (o1 > o0) CU(cudaMemcpyPeerAsync((uint8_t *)st->seq[i].p + o0, c->device, (uint8_t *)st->seq[0].p + o0, r->device, o1 - o0, c->in_st));
                CU(cudaEventRecord(st->ready[i][j], c->in_st));
            }
        }
        return KMCPG_OK;
    };
    const int rc = body();
    if (rc) {
        for (int i = 0; i < n_ctx; i++) { cudaSetDevice(ctxs[i]->device); cudaStreamSynchronize(ctxs[i]->in_st); }
        stage_release(st);
        return rc;
    }
    *out = st;
    return KMCPG_OK;
}

int kmcpg_internal_stage_get(kmcpg_stage *st, int i, int piece, const uint8_t **d_seq, const uint64_t **d_off, void **ready) {
    if (!st || i < 0 || i >= (int)st->ctxs.size() || piece < 0 || piece >= (int)st->ready[i].size()) return KMCPG_EINVAL;
    *d_seq = (const uint8_t *)st->seq[i].p; *d_off = (const uint64_t *)st->off[i].p; *ready = (void *)st->ready[i][piece];
    return KMCPG_OK;
}

// after every job that reads the staged batch has been waited for
void kmcpg_internal_stage_end(kmcpg_stage *st) {
    if (!st) return;
    for (kmcpg_ctx *c : st->ctxs) { cudaSetDevice(c->device); cudaStreamSynchronize(c->in_st); }
    stage_release(st);
}

int kmcpg_target_sizes(const kmcpg_ctx *ctx, double *out, int64_t n) {
    if (!ctx || !out || !ctx->has_db || n < (int64_t)ctx->target_sizes.size()) return KMCPG_EINVAL;
    memcpy(out, ctx->target_sizes.data(), ctx->target_sizes.size() * sizeof(double));
    return (int)KMCPG_OK;
}

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

// test hook (tests/test_gpu_parity.py): the device's row-index arithmetic alone — out[i*num_hashes + j] = row of hash j of codes[i] in a
// block of num_sigs signatures (hashValues H:125-141 + fastdiv.Mod U:6811), through the 32-bit kernel form when num_sigs < 2^32-1
int kmcpg_internal_row_indices(kmcpg_ctx *ctx, const uint64_t *codes, uint64_t n, int num_hashes, uint64_t num_sigs, uint64_t *out) {
    if (!ctx || !out || (n && !codes) || num_hashes < 1 || num_hashes > 4 || num_sigs < 1) return KMCPG_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    if (!n) return KMCPG_OK;
    const FastMod fm = make_fastmod(num_sigs);
    const bool wide = num_sigs >= 0xFFFFFFFFull;
    DevBuf dc, dl;
    auto body = [&]() -> int {
        CU(dc.ensure(n * 8)); CU(dl.ensure(n * (uint64_t)num_hashes * 8));
        CU(cudaMemcpyAsync(dc.p, codes, n * 8, cudaMemcpyHostToDevice, ctx->st));
        if (wide) CU(launch_locs64(dc.as<uint64_t>(), n, num_hashes, fm, dl.as<uint64_t>(), ctx->st));
        else CU(launch_locs(dc.as<uint64_t>(), n, num_hashes, fm, dl.as<uint32_t>(), ctx->st));
        if (wide) {
            CU(cudaMemcpyAsync(out, dl.p, n * (uint64_t)num_hashes * 8, cudaMemcpyDeviceToHost, ctx->st));
            CU(cudaStreamSynchronize(ctx->st));
        } else {
            std::vector<uint32_t> tmp(n * (uint64_t)num_hashes);
            CU(cudaMemcpyAsync(tmp.data(), dl.p, tmp.size() * 4, cudaMemcpyDeviceToHost, ctx->st));
            CU(cudaStreamSynchronize(ctx->st));
            for (size_t i = 0; i < tmp.size(); i++) out[i] = tmp[i];
        }
        return KMCPG_OK;
    };
    const int rc = body();
    dc.release(); dl.release();
    return rc;
}

int kmcpg_generate_kmers(kmcpg_ctx *ctx, const kmcpg_sketch_params *sp, const uint8_t *seq, const uint64_t *off, uint32_t n_seqs,
                         uint64_t **out_codes, uint64_t **out_off) {
    if (!ctx || !sp || !out_codes || !out_off || (n_seqs && (!seq || !off))) return KMCPG_EINVAL;
    if (sp->k < 1 || sp->k > 64) return fail(ctx, KMCPG_EUNSUPPORTED, "k must be in 1..64");
    if (sp->syncmer && (sp->syncmer_s < 1 || (int)sp->syncmer_s >= sp->k)) return fail(ctx, KMCPG_EINVAL, "syncmer_s must be in 1..k-1");
    std::lock_guard<std::mutex> lk(ctx->mu);
    executor_drain(ctx);
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
    executor_drain(ctx);
    CU(cudaSetDevice(ctx->device));
    const int H = ctx->meta.num_hashes;
    const uint64_t nt = (uint64_t)ctx->meta.n_targets;
    cudaStream_t st = ctx->st;
    WorkSet &w = ctx->ws[0];
    CU(ctx->d_dense.ensure(std::max<uint64_t>(nt, 1) * 4));
    CU(cudaMemsetAsync(ctx->d_dense.p, 0, std::max<uint64_t>(nt, 1) * 4, st));
    if (n > 0) {
        if (n >= (1ull << 32) - 1) return fail(ctx, KMCPG_EUNSUPPORTED, "too many codes");
        CU(w.codes.ensure(n * 8)); CU(w.locs[0].ensure(n * 4 * H)); CU(w.slot_off.ensure(16));
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
            ProbeArgs pa;
            memset(&pa, 0, sizeof(pa));
            pa.rows = b.d_rows; pa.pitch = b.pitch; pa.row_bytes = b.row_bytes;
            pa.n_names = b.n_cols; pa.target_base = (uint32_t)(bm.target_base + b.col0); pa.num_hashes = H;
            pa.codes = w.codes.as<uint64_t>(); pa.fm = b.fm; pa.slot_off = w.slot_off.as<uint64_t>();
            if (bm.num_sigs < 0xFFFFFFFFull) {
                CU(launch_locs(w.codes.as<uint64_t>(), n, H, b.fm, w.locs[0].as<uint32_t>(), st));
                pa.locs = w.locs[0].as<uint32_t>();
            }
            pa.n_eff = w.neff.as<uint32_t>(); pa.thresh = w.thresh.as<uint32_t>(); pa.n_queries = 1; pa.paired = 0;
            pa.hit_keys = w.hkeys.as<uint64_t>(); pa.hit_vals = w.hvals.as<uint32_t>(); pa.target_bits = 32;
            pa.hit_count = w.counters.as<unsigned long long>(); pa.hit_cap = 0; pa.dense_counts = ctx->d_dense.as<uint32_t>();
            pa.task_counter = w.counters.as<unsigned long long>() + 2;
            CU(cudaMemsetAsync(w.counters.as<unsigned long long>() + 2, 0, 8, st));
            pa.planes = planes_for(n);
            CU(launch_probe(pa, ctx->sm_count, st));
        }
    }
    CU(cudaMemcpyAsync(counts, ctx->d_dense.p, nt * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return KMCPG_OK;
}

}  // extern "C"

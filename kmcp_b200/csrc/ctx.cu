// ctx.cu — context, database residency in HBM, the batch pipeline, and the C ABI of include/kmcp_gpu.h.
//
// Data layout in HBM (per resident block): the on-disk row-major bit matrix numSigs × numRowBytes
// (X:307-349; row r byte i bit 7-j ⇔ Bloom bit r of target 8i+j) is re-pitched on upload so that every
// row starts 16-byte aligned (128-byte aligned for rows wider than 256 B) and is zero padded; the disk
// format is untouched.  One batch flows  H2D → hash → (sort+unique of long queries) → per block
// {locs, probe} → radix sort of the hit list → D2H,  all on the context's stream.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <cub/cub.cuh>
#include <mutex>
#include <numeric>
#include <string>
#include <vector>

#include <chrono>

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

void layout_block(DeviceBlock &b, const BlockMeta &m) {
    b.pitch = pitch_for((uint32_t)m.row_bytes);
    b.row16 = ((uint32_t)m.row_bytes + 15) / 16;
    uint32_t g = 1;
    while (g < b.row16 && g < 8) g <<= 1;       // a task covers up to 128 B of a row
    b.G = g;
    b.chunks = (b.row16 + g - 1) / g;
    b.fm = make_fastmod(m.num_sigs);
}

void free_db(kmcpg_ctx *ctx) {
    for (auto &b : ctx->blocks) if (b.d_rows) cudaFree(b.d_rows);
    ctx->blocks.clear();
    ctx->resident_of.clear();
    ctx->has_db = false;
    ctx->sum_row_bytes = ctx->resident_bytes = ctx->disk_bytes = 0;
}

int planes_for(uint64_t max_n) {
    if (max_n <= 255) return 8;
    if (max_n <= 65535) return 16;
    if (max_n < (1ull << 24)) return 24;
    return 32;
}

struct Timing { float ms_hash = 0, ms_locs = 0, ms_probe = 0; uint64_t probe_bytes = 0; uint32_t probe_launches = 0; };

int run_hash_stage(kmcpg_ctx *ctx, const kmcpg_search_params &p, int k, const SubBatch &sb, uint32_t nq, uint64_t **codes_out) {
    const DbMeta &m = ctx->meta;
    cudaStream_t st = ctx->st;
    CU(ctx->d_slot_cnt.ensure((sb.n_seqs + 1) * 8ull));
    CU(ctx->d_slot_off.ensure((sb.n_seqs + 1) * 8ull));
    CU(launch_slot_bounds(sb.d_off, sb.n_seqs, k, ctx->d_slot_cnt.as<uint64_t>(), st)); ctx->launches++;
    size_t tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, ctx->d_slot_cnt.as<uint64_t>(), ctx->d_slot_off.as<uint64_t>(), (int)(sb.n_seqs + 1), st);
    CU(ctx->d_tmp.ensure(tmp));
    CU(cub::DeviceScan::ExclusiveSum(ctx->d_tmp.p, tmp, ctx->d_slot_cnt.as<uint64_t>(), ctx->d_slot_off.as<uint64_t>(), (int)(sb.n_seqs + 1), st));
    ctx->launches += 2;
    CU(ctx->d_codes.ensure(std::max<uint64_t>(sb.total_slots, 1) * 8));
    CU(ctx->d_ncodes.ensure(std::max<uint32_t>(nq, 1) * 4ull));
    CU(ctx->d_qlen.ensure(std::max<uint32_t>(nq, 1) * 4ull));
    CU(ctx->d_nk.ensure(std::max<uint32_t>(nq, 1) * 4ull));
    CU(ctx->d_neff.ensure(std::max<uint32_t>(nq, 1) * 4ull));
    CU(ctx->d_thresh.ensure(std::max<uint32_t>(nq, 1) * 4ull));

    HashArgs ha;
    memset(&ha, 0, sizeof(ha));
    ha.seq = sb.d_seq; ha.seq_off = sb.d_off; ha.slot_off = ctx->d_slot_off.as<uint64_t>();
    ha.codes = ctx->d_codes.as<uint64_t>(); ha.n_codes = ctx->d_ncodes.as<uint32_t>(); ha.query_len = ctx->d_qlen.as<int32_t>();
    ha.n_queries = nq; ha.paired = p.paired; ha.mate_select = p.mate_select; ha.k = k; ha.canonical = m.canonical;
    ha.scaled = m.scaled;
    ha.max_hash = ~0ull;
    if (m.scaled) {                                   // U:1040-1043: uint64(float64(^uint64(0)) / float64(scale))
        double v = 18446744073709551616.0 / (double)m.scale;
        ha.max_hash = v >= 18446744073709551616.0 ? ~0ull : (uint64_t)v;
    }
    ha.minimizer = m.minimizer; ha.minimizer_w = m.minimizer_w; ha.syncmer = m.syncmer; ha.syncmer_s = m.syncmer_s;
    ha.min_query_len = p.min_query_len;
    CU(launch_hash(ha, st)); ctx->launches++;

    uint64_t *codes = ctx->d_codes.as<uint64_t>();
    int do_unique = 0;
    if (sb.max_query_slots > (uint64_t)p.dedup_threshold && sb.total_slots > 0) {
        // U:874-908: sort + unique of queries with more than dedup_threshold k-mers
        if (sb.total_slots >= (1ull << 31)) return fail(ctx, KMCPG_EUNSUPPORTED, "sub-batch too large for the dedup sort");
        CU(ctx->d_segb.ensure(nq * 4ull)); CU(ctx->d_sege.ensure(nq * 4ull));
        CU(ctx->d_codes2.ensure(sb.total_slots * 8));
        CU(launch_sort_segments(ctx->d_slot_off.as<uint64_t>(), ctx->d_ncodes.as<uint32_t>(), nq, p.paired, p.dedup_threshold,
                                ctx->d_segb.as<int>(), ctx->d_sege.as<int>(), st));
        CU(cudaMemcpyAsync(ctx->d_codes2.p, ctx->d_codes.p, sb.total_slots * 8, cudaMemcpyDeviceToDevice, st));
        size_t t2 = 0;
        cub::DeviceSegmentedSort::SortKeys(nullptr, t2, ctx->d_codes.as<uint64_t>(), ctx->d_codes2.as<uint64_t>(), (int)sb.total_slots,
                                           (int)nq, ctx->d_segb.as<int>(), ctx->d_sege.as<int>(), st);
        CU(ctx->d_tmp.ensure(t2));
        CU(cub::DeviceSegmentedSort::SortKeys(ctx->d_tmp.p, t2, ctx->d_codes.as<uint64_t>(), ctx->d_codes2.as<uint64_t>(),
                                              (int)sb.total_slots, (int)nq, ctx->d_segb.as<int>(), ctx->d_sege.as<int>(), st));
        ctx->launches += 4;
        codes = ctx->d_codes2.as<uint64_t>();
        do_unique = 1;
    }
    FinalizeArgs fa;
    fa.codes = codes; fa.slot_off = ctx->d_slot_off.as<uint64_t>(); fa.n_codes = ctx->d_ncodes.as<uint32_t>();
    fa.n_kmers_out = ctx->d_nk.as<int32_t>(); fa.n_eff = ctx->d_neff.as<uint32_t>(); fa.thresh = ctx->d_thresh.as<uint32_t>();
    fa.n_queries = nq; fa.paired = p.paired; fa.dedup_threshold = p.dedup_threshold; fa.do_unique = do_unique;
    fa.min_matched = p.min_matched; fa.min_query_cov = p.min_query_cov;
    CU(launch_finalize(fa, st)); ctx->launches++;
    *codes_out = codes;
    return KMCPG_OK;
}

int run_subbatch(kmcpg_ctx *ctx, const kmcpg_search_params &p, int k, const SubBatch &sb, HitsPriv &res, Timing &tm) {
    cudaStream_t st = ctx->st;
    const uint32_t nq = p.paired ? sb.n_seqs / 2 : sb.n_seqs;
    if (nq == 0) return KMCPG_OK;
    CU(cudaEventRecord(ctx->ev[0], st));
    uint64_t *codes = nullptr;
    int rc = run_hash_stage(ctx, p, k, sb, nq, &codes);
    if (rc) return rc;
    CU(cudaEventRecord(ctx->ev[1], st));

    const int H = ctx->meta.num_hashes;
    CU(ctx->d_locs.ensure(std::max<uint64_t>(sb.total_slots, 1) * 4ull * H));
    CU(ctx->d_hitcount.ensure(8));
    uint64_t cap = std::max<uint64_t>(1u << 20, 4ull * nq);
    if (ctx->d_hkeys.cap / 8 > cap) cap = ctx->d_hkeys.cap / 8;
    uint64_t n_hits = 0;
    const int planes = planes_for(sb.max_query_slots);
    for (int attempt = 0; attempt < 3; attempt++) {
        CU(ctx->d_hkeys.ensure(cap * 8)); CU(ctx->d_hvals.ensure(cap * 4));
        CU(cudaMemsetAsync(ctx->d_hitcount.p, 0, 8, st));
        while (ctx->probe_ev.size() < ctx->blocks.size() * 3) {
            cudaEvent_t e;
            CU(cudaEventCreate(&e));
            ctx->probe_ev.push_back(e);
        }
        size_t bi = 0;
        for (auto &b : ctx->blocks) {
            const BlockMeta &bm = ctx->meta.blocks[b.meta_idx];
            CU(cudaEventRecord(ctx->probe_ev[bi * 3], st));
            CU(launch_locs(codes, sb.total_slots, H, b.fm, ctx->d_locs.as<uint32_t>(), st)); ctx->launches++;
            CU(cudaEventRecord(ctx->probe_ev[bi * 3 + 1], st));
            ProbeArgs pa;
            memset(&pa, 0, sizeof(pa));
            pa.rows = b.d_rows; pa.pitch = b.pitch; pa.row16 = b.row16; pa.lanes_per_task = b.G; pa.chunks = b.chunks;
            pa.n_names = (uint32_t)bm.n_names; pa.target_base = (uint32_t)bm.target_base; pa.num_hashes = H;
            pa.locs = ctx->d_locs.as<uint32_t>(); pa.slot_off = ctx->d_slot_off.as<uint64_t>();
            pa.n_eff = ctx->d_neff.as<uint32_t>(); pa.thresh = ctx->d_thresh.as<uint32_t>(); pa.n_queries = nq; pa.paired = p.paired;
            pa.hit_keys = ctx->d_hkeys.as<uint64_t>(); pa.hit_vals = ctx->d_hvals.as<uint32_t>();
            pa.hit_count = ctx->d_hitcount.as<unsigned long long>(); pa.hit_cap = cap; pa.dense_counts = nullptr; pa.planes = planes;
            CU(launch_probe(pa, ctx->sm_count, st)); ctx->launches++;
            CU(cudaEventRecord(ctx->probe_ev[bi * 3 + 2], st));
            bi++;
        }
        CU(cudaMemcpyAsync(ctx->h_small.p, ctx->d_hitcount.p, 8, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        n_hits = *ctx->h_small.as<uint64_t>();
        for (size_t i = 0; i < ctx->blocks.size(); i++) {
            float a = 0, b = 0;
            cudaEventElapsedTime(&a, ctx->probe_ev[i * 3], ctx->probe_ev[i * 3 + 1]);
            cudaEventElapsedTime(&b, ctx->probe_ev[i * 3 + 1], ctx->probe_ev[i * 3 + 2]);
            tm.ms_locs += a; tm.ms_probe += b; tm.probe_launches++;
        }
        if (n_hits <= cap) break;
        cap = n_hits + n_hits / 4 + 1024;            // hit list overflowed: grow and redo the probe phase
        if (attempt == 2) return fail(ctx, KMCPG_ENOMEM, "hit list keeps overflowing");
    }
    CU(cudaEventRecord(ctx->ev[2], st));

    // canonical order: sort by (query, target)
    const size_t old_hits = res.hits.size(), old_q = res.n_kmers.size();
    res.n_kmers.resize(old_q + nq); res.query_len.resize(old_q + nq);
    if (n_hits) {
        CU(ctx->d_hkeys2.ensure(n_hits * 8)); CU(ctx->d_hvals2.ensure(n_hits * 4)); CU(ctx->d_hits.ensure(n_hits * sizeof(kmcpg_hit)));
        int qbits = 1; while ((1ull << qbits) < nq) qbits++;
        size_t t3 = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, t3, ctx->d_hkeys.as<uint64_t>(), ctx->d_hkeys2.as<uint64_t>(), ctx->d_hvals.as<uint32_t>(),
                                        ctx->d_hvals2.as<uint32_t>(), (int64_t)n_hits, 0, 32 + qbits, st);
        CU(ctx->d_tmp.ensure(t3));
        CU(cub::DeviceRadixSort::SortPairs(ctx->d_tmp.p, t3, ctx->d_hkeys.as<uint64_t>(), ctx->d_hkeys2.as<uint64_t>(),
                                           ctx->d_hvals.as<uint32_t>(), ctx->d_hvals2.as<uint32_t>(), (int64_t)n_hits, 0, 32 + qbits, st));
        CU(launch_pack_hits(ctx->d_hkeys2.as<uint64_t>(), ctx->d_hvals2.as<uint32_t>(), n_hits, sb.query_base, ctx->d_hits.as<kmcpg_hit>(), st));
        ctx->launches += 4;
        res.hits.resize(old_hits + n_hits);
        CU(cudaMemcpyAsync(res.hits.data() + old_hits, ctx->d_hits.p, n_hits * sizeof(kmcpg_hit), cudaMemcpyDeviceToHost, st));
    }
    CU(cudaMemcpyAsync(res.n_kmers.data() + old_q, ctx->d_nk.p, nq * 4ull, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(res.query_len.data() + old_q, ctx->d_qlen.p, nq * 4ull, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    float a = 0;
    cudaEventElapsedTime(&a, ctx->ev[0], ctx->ev[1]);
    tm.ms_hash += a;
    uint64_t nsum = 0;
    for (uint32_t i = 0; i < nq; i++) nsum += (uint64_t)res.n_kmers[old_q + i];
    tm.probe_bytes += nsum * (uint64_t)H * (uint64_t)ctx->sum_row_bytes;
    return KMCPG_OK;
}

void fill_out(kmcpg_hits *out, HitsPriv *priv, const Timing &tm, float ms_total, uint32_t launches) {
    out->n_queries = (uint32_t)priv->n_kmers.size();
    out->n_hits = priv->hits.size();
    out->n_kmers = priv->n_kmers.data();
    out->query_len = priv->query_len.data();
    out->hits = priv->hits.data();
    out->ms_hash = tm.ms_hash; out->ms_locs = tm.ms_locs; out->ms_probe = tm.ms_probe; out->ms_total = ms_total;
    out->probe_launches = tm.probe_launches;
    out->probe_row_bytes = tm.probe_bytes;
    out->kernel_launches = launches;
    out->_priv = priv;
}

int check_search_args(kmcpg_ctx *ctx, const kmcpg_search_params *p, const void *seq, const void *off, uint32_t n_seqs, kmcpg_hits *out, int *k) {
    if (!ctx) return KMCPG_EINVAL;
    if (!p || !out || (n_seqs && (!seq || !off))) return fail(ctx, KMCPG_EINVAL, "null argument");
    if (!ctx->has_db) return fail(ctx, KMCPG_EINVAL, "no database open");
    if (p->paired && (n_seqs & 1)) return fail(ctx, KMCPG_EINVAL, "paired batch needs an even number of sequences");
    *k = p->k > 0 ? p->k : ctx->meta.ks.front();
    if (std::find(ctx->meta.ks.begin(), ctx->meta.ks.end(), *k) == ctx->meta.ks.end()) return fail(ctx, KMCPG_EINVAL, "k is not one of the database's k values");
    if (*k > 64 || *k < 1) return fail(ctx, KMCPG_EUNSUPPORTED, "k must be in 1..64");
    if (ctx->meta.minimizer || ctx->meta.syncmer) return fail(ctx, KMCPG_EUNSUPPORTED, "minimizer/syncmer sketches are not on the device path yet");
    if (p->min_matched < 1) return fail(ctx, KMCPG_EINVAL, "min_matched must be >= 1");
    if (!(p->min_query_cov >= 0 && p->min_query_cov <= 1)) return fail(ctx, KMCPG_EINVAL, "min_query_cov must be in [0,1]");
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
    ctx->st = ctx->own_st;
    for (auto &ev : ctx->ev) if ((e = cudaEventCreate(&ev)) != cudaSuccess) return bail(e, "cudaEventCreate");
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
    free_db(ctx);
    for (DevBuf *b : {&ctx->d_seq, &ctx->d_off, &ctx->d_slot_cnt, &ctx->d_slot_off, &ctx->d_codes, &ctx->d_codes2, &ctx->d_locs, &ctx->d_ncodes,
                      &ctx->d_qlen, &ctx->d_nk, &ctx->d_neff, &ctx->d_thresh, &ctx->d_hkeys, &ctx->d_hvals, &ctx->d_hkeys2, &ctx->d_hvals2,
                      &ctx->d_hits, &ctx->d_hitcount, &ctx->d_tmp, &ctx->d_segb, &ctx->d_sege, &ctx->d_dense, &ctx->d_scal})
        b->release();
    ctx->h_stage.release(); ctx->h_off.release(); ctx->h_small.release();
    for (auto &ev : ctx->ev) if (ev) cudaEventDestroy(ev);
    for (auto &ev : ctx->probe_ev) cudaEventDestroy(ev);
    if (ctx->own_st) cudaStreamDestroy(ctx->own_st);
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
    // blocks → shards: largest first onto the least loaded shard (deterministic in every rank)
    std::vector<int> order(m.blocks.size());
    std::iota(order.begin(), order.end(), 0);
    auto bytes_of = [&](int i) { return (uint64_t)m.blocks[i].num_sigs * pitch_for((uint32_t)m.blocks[i].row_bytes); };
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return bytes_of(a) > bytes_of(b); });
    std::vector<uint64_t> load(world, 0);
    std::vector<int> owner(m.blocks.size(), 0);
    for (int i : order) {
        int best = (int)(std::min_element(load.begin(), load.end()) - load.begin());
        owner[i] = best;
        load[best] += bytes_of(i);
    }
    if (opts && opts->max_resident_bytes > 0 && (int64_t)load[rank] > opts->max_resident_bytes)
        return fail(ctx, KMCPG_ENOMEM, "resident blocks exceed max_resident_bytes");
    ctx->resident_of.assign(m.blocks.size(), -1);
    const size_t CHUNK = 256ull << 20;
    CU(ctx->h_stage.ensure(CHUNK));
    for (size_t i = 0; i < m.blocks.size(); i++) {
        if (owner[i] != rank) continue;
        const BlockMeta &bm = m.blocks[i];
        if (bm.num_sigs >= (1ull << 32)) return fail(ctx, KMCPG_EUNSUPPORTED, "blocks with >= 2^32 signatures are not supported");
        DeviceBlock b;
        b.meta_idx = (int)i;
        layout_block(b, bm);
        b.bytes = (size_t)bm.num_sigs * b.pitch;
        CU(cudaMalloc((void **)&b.d_rows, std::max<size_t>(b.bytes, 16)));
        ctx->blocks.push_back(b);
        ctx->resident_of[i] = (int)ctx->blocks.size() - 1;
        FILE *f = fopen(bm.path.c_str(), "rb");
        if (!f) return fail(ctx, KMCPG_EIO, "cannot open " + bm.path);
        fseek(f, (long)bm.data_offset, SEEK_SET);
        const uint64_t rows_per_chunk = std::max<uint64_t>(1, CHUNK / (uint64_t)bm.row_bytes);
        if ((uint64_t)bm.row_bytes > CHUNK) { fclose(f); return fail(ctx, KMCPG_EUNSUPPORTED, "row wider than the staging buffer"); }
        for (uint64_t r0 = 0; r0 < bm.num_sigs; r0 += rows_per_chunk) {
            uint64_t nr = std::min<uint64_t>(rows_per_chunk, bm.num_sigs - r0);
            size_t bytes = (size_t)nr * bm.row_bytes;
            if (fread(ctx->h_stage.p, 1, bytes, f) != bytes) { fclose(f); return fail(ctx, KMCPG_EIO, "kmcp: truncated index file: " + bm.path); }
            cudaError_t e = ctx->d_tmp.ensure(bytes);
            if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->d_tmp.p, ctx->h_stage.p, bytes, cudaMemcpyHostToDevice, ctx->st);
            if (e == cudaSuccess) e = launch_repitch(ctx->d_tmp.as<uint8_t>(), b.d_rows + r0 * b.pitch, nr, (uint32_t)bm.row_bytes, b.pitch, ctx->st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->st);
            if (e != cudaSuccess) { fclose(f); CU(e); }
        }
        fclose(f);
        ctx->sum_row_bytes += bm.row_bytes;
        ctx->resident_bytes += (int64_t)b.bytes;
        ctx->disk_bytes += (int64_t)(bm.num_sigs * (uint64_t)bm.row_bytes);
    }
    ctx->has_db = true;
    return KMCPG_OK;
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
    // blocks are few: binary search over target_base
    size_t lo = 0, hi = bl.size();
    while (lo + 1 < hi) { size_t mid = (lo + hi) / 2; if (bl[mid].target_base <= g) lo = mid; else hi = mid; }
    if (bl.empty() || g < bl[lo].target_base || g >= bl[lo].target_base + bl[lo].n_names) return KMCPG_EINVAL;
    int c = (int)(g - bl[lo].target_base);
    o->name = bl[lo].names[c].c_str(); o->index = bl[lo].indices[c]; o->genome_size = bl[lo].gsizes[c]; o->n_kmers = bl[lo].sizes[c];
    o->block = (int32_t)lo; o->col = c; o->resident = ctx->resident_of[lo] >= 0;
    return KMCPG_OK;
}

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
    HitsPriv *priv = new HitsPriv();
    Timing tm;
    const uint32_t launches0 = ctx->launches;
    auto t0 = std::chrono::steady_clock::now();
    const uint64_t MAX_SLOTS = 64ull << 20;
    const uint32_t MAX_SEQS = 4u << 20;
    const uint32_t step = p->paired ? 2 : 1;
    uint32_t a = 0;
    while (a < n_seqs) {
        // greedy sub-batch [a, b)
        uint64_t slots = 0, maxq = 0;
        uint32_t b = a;
        while (b < n_seqs && (b - a) < MAX_SEQS) {
            uint64_t qs = 0;
            for (uint32_t m = 0; m < step; m++) {
                uint64_t len = off[b + m + 1] - off[b + m];
                if (off[b + m + 1] < off[b + m]) { delete priv; return fail(ctx, KMCPG_EINVAL, "offsets must be non-decreasing"); }
                qs += len >= (uint64_t)k ? len - k + 1 : 0;
            }
            if (b > a && slots + qs > MAX_SLOTS) break;
            slots += qs; maxq = std::max(maxq, qs);
            b += step;
        }
        const uint32_t ns = b - a;
        const uint64_t nbytes = off[b] - off[a];
        // stage offsets (rebased) + bytes, H2D
        cudaError_t e = ctx->h_off.ensure((ns + 1) * 8ull);
        if (e == cudaSuccess) e = ctx->d_off.ensure((ns + 1) * 8ull);
        if (e == cudaSuccess) e = ctx->d_seq.ensure(std::max<uint64_t>(nbytes, 1) + 64);
        if (e != cudaSuccess) { delete priv; CU(e); }
        uint64_t *ho = ctx->h_off.as<uint64_t>();
        for (uint32_t i = 0; i <= ns; i++) ho[i] = off[a + i] - off[a];
        e = cudaMemcpyAsync(ctx->d_off.p, ho, (ns + 1) * 8ull, cudaMemcpyHostToDevice, ctx->st);
        if (e == cudaSuccess && nbytes) e = cudaMemcpyAsync(ctx->d_seq.p, seq + off[a], nbytes, cudaMemcpyHostToDevice, ctx->st);
        if (e != cudaSuccess) { delete priv; CU(e); }
        SubBatch sb{ctx->d_seq.as<uint8_t>(), ctx->d_off.as<uint64_t>(), ns, slots, maxq, a / step};
        rc = run_subbatch(ctx, *p, k, sb, *priv, tm);
        if (rc) { delete priv; return rc; }
        a = b;
    }
    float ms_total = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    fill_out(out, priv, tm, ms_total, ctx->launches - launches0);
    return KMCPG_OK;
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
    auto t0 = std::chrono::steady_clock::now();
    const uint32_t launches0 = ctx->launches;
    // slot totals are needed on the host to size buffers: Σ and max of the per-sequence bounds
    CU(ctx->d_slot_cnt.ensure((n_seqs + 1) * 8ull));
    CU(ctx->d_scal.ensure(16));
    CU(launch_slot_bounds(d_off, n_seqs, k, ctx->d_slot_cnt.as<uint64_t>(), ctx->st));
    size_t t1 = 0, t2 = 0;
    cub::DeviceReduce::Sum(nullptr, t1, ctx->d_slot_cnt.as<uint64_t>(), ctx->d_scal.as<uint64_t>(), (int)n_seqs, ctx->st);
    cub::DeviceReduce::Max(nullptr, t2, ctx->d_slot_cnt.as<uint64_t>(), ctx->d_scal.as<uint64_t>() + 1, (int)n_seqs, ctx->st);
    CU(ctx->d_tmp.ensure(std::max(t1, t2)));
    CU(cub::DeviceReduce::Sum(ctx->d_tmp.p, t1, ctx->d_slot_cnt.as<uint64_t>(), ctx->d_scal.as<uint64_t>(), (int)n_seqs, ctx->st));
    CU(cub::DeviceReduce::Max(ctx->d_tmp.p, t2, ctx->d_slot_cnt.as<uint64_t>(), ctx->d_scal.as<uint64_t>() + 1, (int)n_seqs, ctx->st));
    ctx->launches += 3;
    CU(cudaMemcpyAsync(ctx->h_small.p, ctx->d_scal.p, 16, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    uint64_t total = ctx->h_small.as<uint64_t>()[0], mx = ctx->h_small.as<uint64_t>()[1];
    if (p->paired) mx *= 2;
    HitsPriv *priv = new HitsPriv();
    Timing tm;
    SubBatch sb{d_seq, d_off, n_seqs, total, mx, 0};
    rc = run_subbatch(ctx, *p, k, sb, *priv, tm);
    if (rc) { delete priv; return rc; }
    float ms_total = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    fill_out(out, priv, tm, ms_total, ctx->launches - launches0);
    return KMCPG_OK;
}

void kmcpg_free_hits(kmcpg_hits *h) {
    if (!h) return;
    delete (HitsPriv *)h->_priv;
    memset(h, 0, sizeof(*h));
}

int kmcpg_host_alloc(void **p, size_t bytes) {
    if (!p) return KMCPG_EINVAL;
    cudaError_t e = cudaMallocHost(p, bytes ? bytes : 1);
    if (e != cudaSuccess) { (void)cudaGetLastError(); g_create_error = cudaGetErrorString(e); return KMCPG_ENOMEM; }
    return KMCPG_OK;
}
int kmcpg_host_free(void *p) { return cudaFreeHost(p) == cudaSuccess ? KMCPG_OK : KMCPG_ECUDA; }

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
    if (sp->minimizer || sp->syncmer) return fail(ctx, KMCPG_EUNSUPPORTED, "minimizer/syncmer sketches are not on the device path yet");
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
    auto body = [&]() -> int {
        CU(ctx->d_off.ensure((n_seqs + 1) * 8ull));
        CU(ctx->d_seq.ensure(nbytes + 64));
        CU(cudaMemcpyAsync(ctx->d_off.p, ho.data(), (n_seqs + 1) * 8ull, cudaMemcpyHostToDevice, ctx->st));
        if (nbytes) CU(cudaMemcpyAsync(ctx->d_seq.p, seq + off[0], nbytes, cudaMemcpyHostToDevice, ctx->st));
        SubBatch sb{ctx->d_seq.as<uint8_t>(), ctx->d_off.as<uint64_t>(), n_seqs, total, mx, 0};
        int r = run_hash_stage(ctx, p, sp->k, sb, n_seqs, &codes);
        if (r) return r;
        std::vector<uint32_t> nc(n_seqs);
        std::vector<uint64_t> so(n_seqs + 1);
        CU(cudaMemcpyAsync(nc.data(), ctx->d_ncodes.p, n_seqs * 4ull, cudaMemcpyDeviceToHost, ctx->st));
        CU(cudaMemcpyAsync(so.data(), ctx->d_slot_off.p, (n_seqs + 1) * 8ull, cudaMemcpyDeviceToHost, ctx->st));
        CU(cudaStreamSynchronize(ctx->st));
        std::vector<uint64_t> all(total ? total : 1);
        if (total) CU(cudaMemcpy(all.data(), codes, total * 8, cudaMemcpyDeviceToHost));
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
    CU(ctx->d_dense.ensure(std::max<uint64_t>(nt, 1) * 4));
    CU(cudaMemsetAsync(ctx->d_dense.p, 0, std::max<uint64_t>(nt, 1) * 4, st));
    if (n > 0) {
        if (n >= (1ull << 32) - 1) return fail(ctx, KMCPG_EUNSUPPORTED, "too many codes");
        CU(ctx->d_codes.ensure(n * 8)); CU(ctx->d_locs.ensure(n * 4 * H)); CU(ctx->d_slot_off.ensure(16));
        CU(ctx->d_neff.ensure(4)); CU(ctx->d_thresh.ensure(4)); CU(ctx->d_hitcount.ensure(8));
        CU(ctx->d_hkeys.ensure(8)); CU(ctx->d_hvals.ensure(4));
        uint64_t so[2] = {0, n};
        uint32_t neff = (uint32_t)n, th = 0xFFFFFFFFu;
        CU(cudaMemcpyAsync(ctx->d_codes.p, codes, n * 8, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(ctx->d_slot_off.p, so, 16, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(ctx->d_neff.p, &neff, 4, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(ctx->d_thresh.p, &th, 4, cudaMemcpyHostToDevice, st));
        CU(cudaMemsetAsync(ctx->d_hitcount.p, 0, 8, st));
        CU(cudaStreamSynchronize(st));
        for (auto &b : ctx->blocks) {
            const BlockMeta &bm = ctx->meta.blocks[b.meta_idx];
            CU(launch_locs(ctx->d_codes.as<uint64_t>(), n, H, b.fm, ctx->d_locs.as<uint32_t>(), st));
            ProbeArgs pa;
            memset(&pa, 0, sizeof(pa));
            pa.rows = b.d_rows; pa.pitch = b.pitch; pa.row16 = b.row16; pa.lanes_per_task = b.G; pa.chunks = b.chunks;
            pa.n_names = (uint32_t)bm.n_names; pa.target_base = (uint32_t)bm.target_base; pa.num_hashes = H;
            pa.locs = ctx->d_locs.as<uint32_t>(); pa.slot_off = ctx->d_slot_off.as<uint64_t>();
            pa.n_eff = ctx->d_neff.as<uint32_t>(); pa.thresh = ctx->d_thresh.as<uint32_t>(); pa.n_queries = 1; pa.paired = 0;
            pa.hit_keys = ctx->d_hkeys.as<uint64_t>(); pa.hit_vals = ctx->d_hvals.as<uint32_t>();
            pa.hit_count = ctx->d_hitcount.as<unsigned long long>(); pa.hit_cap = 0; pa.dense_counts = ctx->d_dense.as<uint32_t>();
            pa.planes = planes_for(n);
            CU(launch_probe(pa, ctx->sm_count, st));
        }
    }
    CU(cudaMemcpyAsync(counts, ctx->d_dense.p, nt * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return KMCPG_OK;
}

}  // extern "C"

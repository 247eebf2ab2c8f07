// synth.cu — seeded synthetic workloads made ON the device, and the device index builder.
//
// The generators are pure functions of (seed, index) and are mirrored byte for byte by
// oracle/oracle.py (synth_genome / synth_read) so the CPU oracle can rebuild any input.
//
// kmcpg_build_synth_db is the GPU form of `kmcp compute` + `kmcp index` for single-record genomes in
// split mode (reference kmcp/cmd/compute.go C:685-745 chunk windows, C:746-826 code sets;
// kmcp/cmd/index.go I:667 sort by k-mer count, I:670-682 block size, I:936-948 + 1023 numSigs,
// I:1157 / I:1188 bit set): hash every chunk with the search path's own hash kernel, sort+unique,
// then scatter bits into the block matrices in HBM.  SURVEY.md §8(f1).
#include <algorithm>
#include <cstring>
#include <cub/cub.cuh>
#include <numeric>

#include "ctx_internal.h"

namespace kmcpg {

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

__device__ __forceinline__ uint32_t genome_base(uint64_t gkey, uint64_t pos) {
    uint64_t w = splitmix64(gkey + (pos >> 5));
    return (uint32_t)(w >> (2 * (pos & 31))) & 3u;
}

__global__ void synth_genome_kernel(uint64_t gseed, uint32_t genome, uint64_t start, uint64_t len, uint8_t *__restrict__ out) {
    const uint64_t gkey = splitmix64(gseed * 0x100000001B3ULL + genome);
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < len; i += stride) out[i] = "ACGT"[genome_base(gkey, start + i)];
}

cudaError_t launch_synth_genome(uint64_t gseed, uint32_t genome, uint64_t start, uint64_t len, uint8_t *out, cudaStream_t st) {
    if (!len) return cudaSuccess;
    uint64_t blocks = (len + 255) / 256;
    synth_genome_kernel<<<(unsigned)std::min<uint64_t>(blocks, 148 * 32), 256, 0, st>>>(gseed, genome, start, len, out);
    return cudaGetLastError();
}

// read r: 80 % sampled from a genome (random strand, 1 % substitutions), 20 % uniform random
__global__ void synth_reads_kernel(uint64_t seed, uint64_t first, uint32_t n_reads, uint32_t L, uint64_t gseed, uint32_t n_genomes,
                                   uint32_t glen, uint8_t *__restrict__ out) {
    const uint64_t total = (uint64_t)n_reads * L;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        const uint64_t r = i / L;
        uint32_t j = (uint32_t)(i - r * L);
        const uint64_t u = splitmix64(seed * 0x100000001B3ULL + (first + r));
        uint32_t b;
        if ((u & 0xFF) < 204 && glen >= L) {
            const uint32_t g = (uint32_t)((u >> 8) & 0x7FFFFF) % n_genomes;
            const uint64_t pos = (u >> 32) % (uint64_t)(glen - L + 1);
            const bool rc = ((u >> 31) & 1) != 0;
            const uint32_t jj = rc ? (L - 1 - j) : j;               // position inside the forward-strand fragment
            const uint64_t gkey = splitmix64(gseed * 0x100000001B3ULL + g);
            b = genome_base(gkey, pos + jj);
            const uint64_t v = splitmix64(u ^ ((uint64_t)jj * 0xD1342543DE82EF95ULL));
            if ((v & 0xFFFF) < 655) b = (b + 1 + (uint32_t)((v >> 16) % 3)) & 3;
            if (rc) b = 3 - b;
        } else {
            uint64_t w = splitmix64(u + (j >> 5));
            b = (uint32_t)(w >> (2 * (j & 31))) & 3u;
        }
        out[i] = "ACGT"[b];
    }
}

cudaError_t launch_synth_reads(uint64_t seed, uint64_t first, uint32_t n_reads, uint32_t read_len, uint64_t gseed, uint32_t n_genomes,
                               uint32_t genome_len, uint8_t *out, cudaStream_t st) {
    if (!n_reads || !read_len) return cudaSuccess;
    synth_reads_kernel<<<148 * 32, 256, 0, st>>>(seed, first, n_reads, read_len, gseed, n_genomes, genome_len, out);
    return cudaGetLastError();
}

__device__ __forceinline__ uint64_t fastmod_dev2(uint64_t a, uint64_t m_hi, uint64_t m_lo, uint64_t d) {
    uint64_t lo = m_lo * a;
    uint64_t hi = __umul64hi(m_lo, a) + m_hi * a;
    uint64_t bottom = __umul64hi(lo, d);
    uint64_t top_lo = hi * d;
    uint64_t top_hi = __umul64hi(hi, d);
    uint64_t sum = bottom + top_lo;
    return top_hi + (sum < bottom ? 1 : 0);
}

// sigs[loc] |= 1 << (7-j)  (I:1157); h>1: every hashValues location (I:1188)
__global__ void set_bits_kernel(const uint64_t *__restrict__ codes, uint64_t n, int H, FastMod fm, uint8_t *__restrict__ rows, uint32_t pitch, uint32_t col) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint32_t byte = col >> 3;
    const uint32_t shift = (byte & 3) * 8 + (7 - (col & 7));       // bit inside the aligned 32-bit word (little endian)
    for (; i < n; i += stride) {
        const uint64_t code = codes[i];
        const uint32_t x = (uint32_t)(code >> 32), y = (uint32_t)code;
        for (int j = 0; j < H; j++) {
            const uint64_t v = H == 1 ? code : (uint64_t)(uint32_t)(x + y * (uint32_t)j);
            const uint64_t loc = fastmod_dev2(v, fm.m_hi, fm.m_lo, fm.d);
            uint32_t *w = (uint32_t *)(rows + loc * pitch + (byte & ~3u));
            atomicOr(w, 1u << shift);
        }
    }
}

cudaError_t launch_set_bits(const uint64_t *codes, uint64_t n, int H, FastMod fm, uint8_t *rows, uint32_t pitch, uint32_t col, cudaStream_t st) {
    if (!n) return cudaSuccess;
    uint64_t blocks = (n + 255) / 256;
    set_bits_kernel<<<(unsigned)std::min<uint64_t>(blocks, 148 * 16), 256, 0, st>>>(codes, n, H, fm, rows, pitch, col);
    return cudaGetLastError();
}

namespace {

struct Window { uint64_t start, len; };

// C:685-745 split windows of a single-record genome (non circular, greedy)
std::vector<Window> split_windows(uint64_t L, int n_chunks, int overlap, int k, int split_min_ref = 1000) {
    std::vector<Window> w;
    uint64_t size = L, step = L;
    if (n_chunks > 1 && L >= (uint64_t)split_min_ref) {
        size = (L + (uint64_t)(n_chunks - 1) * overlap + n_chunks - 1) / n_chunks;
        step = size - overlap;
    }
    if (step == 0) return w;
    for (uint64_t i = 0; i < L; i += step) {
        uint64_t len = std::min<uint64_t>(size, L - i);
        if (n_chunks > 1 && (int64_t)len - 1 <= overlap) continue;      // C:713, 742
        if (len < (uint64_t)k) continue;
        w.push_back({i, len});
    }
    return w;
}

}  // namespace
}  // namespace kmcpg

using namespace kmcpg;

extern "C" {

int kmcpg_synth_reads(kmcpg_ctx *ctx, uint64_t seed, uint64_t first, uint32_t n_reads, uint32_t read_len, uint64_t genome_seed,
                      uint32_t n_genomes, uint32_t genome_len, uint8_t *d_out) {
    if (!ctx || !d_out || !n_genomes) return KMCPG_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    CU(launch_synth_reads(seed, first, n_reads, read_len, genome_seed, n_genomes, genome_len, d_out, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    return KMCPG_OK;
}

int kmcpg_synth_genomes(kmcpg_ctx *ctx, uint64_t genome_seed, uint32_t first, uint32_t n_genomes, uint32_t genome_len, uint8_t *d_out) {
    if (!ctx || !d_out) return KMCPG_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    for (uint32_t g = 0; g < n_genomes; g++)
        CU(launch_synth_genome(genome_seed, first + g, 0, genome_len, d_out + (uint64_t)g * genome_len, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    return KMCPG_OK;
}

int kmcpg_build_synth_db(kmcpg_ctx *ctx, const kmcpg_synth_db *spec) {
    if (!ctx || !spec) return KMCPG_EINVAL;
    if (spec->k < 1 || spec->k > 64 || spec->n_genomes < 1 || spec->genome_len < (uint32_t)spec->k || spec->num_hashes < 1 || spec->num_hashes > 4 ||
        spec->n_chunks < 1 || spec->overlap < 0 || !(spec->fpr > 0 && spec->fpr < 1))
        return fail(ctx, KMCPG_EINVAL, "bad synthetic DB spec");
    std::lock_guard<std::mutex> lk(ctx->mu);
    executor_drain(ctx);
    CU(cudaSetDevice(ctx->device));
    free_db(ctx);
    cudaStream_t st = ctx->st;

    DbMeta &m = ctx->meta;
    m = DbMeta();
    m.dir = "<synthetic>";
    m.version = 4; m.index_version = 4; m.ks = {spec->k}; m.canonical = true; m.num_hashes = spec->num_hashes; m.fpr = spec->fpr;
    m.scaled = spec->scale > 1; m.scale = m.scaled ? spec->scale : 0;

    const uint64_t L = spec->genome_len;
    const std::vector<Window> wins = split_windows(L, spec->n_chunks, spec->overlap, spec->k);
    if (wins.empty()) return fail(ctx, KMCPG_EINVAL, "genome too short to split");
    const uint32_t nw = (uint32_t)wins.size();
    const uint64_t n_targets = (uint64_t)spec->n_genomes * nw;

    // genomes per pass: keep the code arrays of one pass around 1.5 GB
    uint64_t slots_per_genome = 0;
    for (auto &w : wins) slots_per_genome += w.len - spec->k + 1;
    uint32_t gpp = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(spec->n_genomes, (96ull << 20) / std::max<uint64_t>(slots_per_genome, 1)));

    kmcpg_search_params hp;
    kmcpg_default_params(&hp);
    hp.min_query_len = 0; hp.min_matched = 1; hp.dedup_threshold = 0; hp.min_query_cov = 0;   // sort+unique every window (C:815-823)

    const int world = spec->shard_world > 1 ? spec->shard_world : 1;
    const int rank = world > 1 ? spec->shard_rank : 0;
    if (rank < 0 || rank >= world) return fail(ctx, KMCPG_EINVAL, "shard_rank out of range");

    std::vector<uint64_t> sizes(n_targets, 0);
    std::vector<uint64_t> h_off;
    std::vector<uint32_t> h_nc;
    std::vector<uint64_t> h_slot;
    std::vector<uint32_t> sel;            // windows of the running pass: target ids (genome*nw + window)
    WorkSet &w = ctx->ws[0];

    // target order / block composition are known only after pass 0
    std::vector<uint32_t> order;          // sorted position → target id (genome*nw + window)
    std::vector<uint32_t> pos_of;         // target id → sorted position
    std::vector<int> owner;               // block → shard (kmcpg_open_db's plan)
    int block_size = 0;

    for (int pass = 0; pass < 2; pass++) {
        if (pass == 1) {
            // I:667: ascending by k-mer count (stable), I:670-682 block size
            order.resize(n_targets);
            std::iota(order.begin(), order.end(), 0u);
            std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return sizes[a] < sizes[b]; });
            pos_of.resize(n_targets);
            for (uint64_t i = 0; i < n_targets; i++) pos_of[order[i]] = (uint32_t)i;
            block_size = spec->block_size > 0 ? spec->block_size : (int)std::max<uint64_t>(8, std::min<uint64_t>(n_targets, ((n_targets / 16) + 7) / 8 * 8));
            const uint64_t nb = (n_targets + block_size - 1) / block_size;
            m.blocks.resize(nb);
            for (uint64_t bi = 0; bi < nb; bi++) {
                BlockMeta &bm = m.blocks[bi];
                const uint64_t t0 = bi * block_size, t1 = std::min<uint64_t>(n_targets, t0 + block_size);
                bm.path = "<synthetic>"; bm.k = spec->k; bm.canonical = true; bm.num_hashes = spec->num_hashes;
                bm.n_names = (int)(t1 - t0); bm.row_bytes = (bm.n_names + 7) / 8; bm.target_base = (int64_t)t0;
                uint64_t mx = 0;
                for (uint64_t t = t0; t < t1; t++) {
                    const uint32_t id = order[t];
                    char nm[64];
                    snprintf(nm, sizeof(nm), "synth_%06u", id / nw);
                    bm.names.push_back(nm);
                    bm.indices.push_back((id % nw) | (nw << 16));
                    bm.gsizes.push_back(L);
                    bm.sizes.push_back(sizes[id]);
                    mx = std::max(mx, sizes[id]);
                }
                bm.num_sigs = calc_signature_size(mx, spec->num_hashes, spec->fpr);     // I:936-948, I:1023
                if (bm.num_sigs == 0) return fail(ctx, KMCPG_EUNSUPPORTED, "synthetic block has an unsupported number of signatures");
            }
            m.n_targets = (int64_t)n_targets;
            // the shard's blocks: exactly what kmcpg_open_db(shard_rank, shard_world) keeps of this database
            std::vector<ShardPiece> pieces;
            std::vector<uint64_t> load;
            plan_pieces(m, world, pieces, load);
            ctx->part_row_bytes = widest_shard_row_bytes(pieces, world);
            owner.assign(nb, -1);
            for (const ShardPiece &pc : pieces) {
                if (pc.col0 != 0 || pc.n_cols != (uint32_t)m.blocks[pc.block].n_names)
                    return fail(ctx, KMCPG_EUNSUPPORTED, "the synthetic builder shards by whole blocks only (fewer blocks than shards)");
                owner[pc.block] = pc.shard;
            }
            ctx->resident_of.assign(nb, -1);
            for (uint64_t bi = 0; bi < nb; bi++) {
                if (owner[bi] != rank) continue;
                const BlockMeta &bm = m.blocks[bi];
                DeviceBlock db;
                db.meta_idx = (int)bi;
                layout_block(db, bm);
                db.bytes = (size_t)bm.num_sigs * db.pitch;
                CU(cudaMalloc((void **)&db.d_rows, std::max<size_t>(db.bytes, 16)));
                CU(cudaMemsetAsync(db.d_rows, 0, db.bytes, st));
                ctx->blocks.push_back(db);
                ctx->resident_of[bi] = (int)ctx->blocks.size() - 1;
                ctx->sum_row_bytes += bm.row_bytes;
                ctx->resident_bytes += (int64_t)db.bytes;
                ctx->disk_bytes += (int64_t)(bm.num_sigs * (uint64_t)bm.row_bytes);
            }
            ctx->target_sizes.resize(n_targets);
            for (auto &bm : m.blocks)
                for (int c = 0; c < bm.n_names; c++) ctx->target_sizes[(size_t)bm.target_base + c] = (double)bm.sizes[c];
            // bits are set idempotently: the second pass needs no sort + unique, and only this shard's windows
            hp.dedup_threshold = 0x7fffffff;
        }
        const uint32_t gstep = pass == 0 ? gpp : (uint32_t)std::min<uint64_t>(spec->n_genomes, (uint64_t)gpp * (uint64_t)world);
        for (uint32_t g0 = 0; g0 < spec->n_genomes; g0 += gstep) {
            const uint32_t ng = std::min<uint32_t>(gstep, spec->n_genomes - g0);
            // windows of this round as independent sequences (overlapping windows need their own byte ranges, so each window is
            // generated separately); the second pass takes only the windows whose block lives in this shard
            sel.clear();
            for (uint32_t g = g0; g < g0 + ng; g++)
                for (uint32_t wi = 0; wi < nw; wi++) {
                    const uint32_t id = g * nw + wi;
                    if (pass == 1 && owner[pos_of[id] / (uint32_t)block_size] != rank) continue;
                    sel.push_back(id);
                }
            const uint32_t ns = (uint32_t)sel.size();
            if (!ns) continue;
            h_off.resize((size_t)ns + 1); h_nc.resize(ns); h_slot.resize((size_t)ns + 1);
            uint64_t bytes = 0, total = 0, mx = 0;
            for (uint32_t s = 0; s < ns; s++) {
                const Window &wn = wins[sel[s] % nw];
                h_off[s] = bytes; bytes += wn.len;
                const uint64_t c = wn.len - spec->k + 1;
                total += c; mx = std::max(mx, c);
            }
            h_off[ns] = bytes;
            CU(w.off.ensure(((uint64_t)ns + 1) * 8));
            CU(w.seq.ensure(bytes + 64));
            for (uint32_t s = 0; s < ns; s++)
                CU(launch_synth_genome(spec->genome_seed, sel[s] / nw, wins[sel[s] % nw].start, wins[sel[s] % nw].len, w.seq.as<uint8_t>() + h_off[s], st));
            CU(cudaMemcpyAsync(w.off.p, h_off.data(), (ns + 1) * 8ull, cudaMemcpyHostToDevice, st));
            SubBatch sb{w.seq.as<uint8_t>(), w.off.as<uint64_t>(), ns, total, mx, 0};
            uint64_t *codes = nullptr;
            int rc = run_hash_stage(ctx, w, hp, spec->k, sb, ns, &codes);
            if (rc) return rc;
            CU(cudaMemcpyAsync(h_nc.data(), w.ncodes.p, ns * 4ull, cudaMemcpyDeviceToHost, st));
            CU(cudaMemcpyAsync(h_slot.data(), w.slot_off.p, (ns + 1) * 8ull, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            for (uint32_t s = 0; s < ns; s++) {
                const uint32_t id = sel[s];
                const uint32_t nc = h_nc[s] == 0xFFFFFFFFu ? 0 : h_nc[s];
                if (pass == 0) { sizes[id] = nc; continue; }
                const uint32_t pos = pos_of[id];
                const uint32_t bi = pos / block_size, col = pos % block_size;
                DeviceBlock &db = ctx->blocks[ctx->resident_of[bi]];
                CU(launch_set_bits(codes + h_slot[s], nc, spec->num_hashes, db.fm, db.d_rows, db.pitch, col, st));
            }
            CU(cudaStreamSynchronize(st));
        }
    }
    ctx->has_db = true;
    return KMCPG_OK;
}

int kmcpg_write_block(kmcpg_ctx *ctx, int rb, const char *path) {
    if (!ctx || !path) return KMCPG_EINVAL;
    if (!ctx->has_db || rb < 0 || rb >= (int)ctx->blocks.size()) return fail(ctx, KMCPG_EINVAL, "no such resident block");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    const DeviceBlock &b = ctx->blocks[rb];
    const BlockMeta &bm = ctx->meta.blocks[b.meta_idx];
    if (!b.whole) return fail(ctx, KMCPG_EUNSUPPORTED, "this context holds only a column range of the block");
    const size_t bytes = (size_t)bm.num_sigs * bm.row_bytes;
    CU(ctx->d_tmp.ensure(std::max<size_t>(bytes, 16)));
    CU(launch_unpitch(b.d_rows, ctx->d_tmp.as<uint8_t>(), bm.num_sigs, (uint32_t)bm.row_bytes, b.pitch, ctx->st));
    std::vector<uint8_t> host(bytes ? bytes : 1);
    CU(cudaMemcpyAsync(host.data(), ctx->d_tmp.p, bytes, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    std::string err;
    int rc = write_block_file(path, bm, host.data(), err);
    if (rc) return fail(ctx, rc, err);
    return KMCPG_OK;
}

}  // extern "C"

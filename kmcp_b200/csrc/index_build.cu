// index_build.cu — `kmcp compute` + `kmcp index` fused on the GPU: FASTA/Q files → `.uniki` blocks,
// `__db.yml`, `__name_mapping.tsv`, with no `.unik` intermediates, and the database left resident in HBM.
//
// Reference semantics restated on the host (chunking, naming, block assembly) and executed on the device (hashing with
// the search path's own kernels, sort + unique per target, bit scatter):
//   compute: record filter -B (C:587-600), N-join of the kept records in split mode (C:612-626), split windows
//            (C:685-745), non-split mode = every record hashed on its own and pooled (C:676-681, 805-807, 905-914),
//            target name by -N regexp on the file name (C:826-838, 919-929), sort+unique code sets (C:815-823);
//   index:   ascending by k-mer count (I:667), block size rule (I:670-682), numSigs from the largest set of the block
//            (I:936-948, 1023), bit 7-(j&7) of byte j>>3 (I:1157 / 1188), Indices = chunkIdx | nChunks<<16 (I:1096),
//            R001/_blockNNN.uniki + __db.yml + __name_mapping.tsv (I:1283-1285, 1353-1399).
// Not restated: the special block sizes for huge genomes (-x/-X/-8/-1, I:213-259, 787-880), --by-seq, --circular,
// --split-size, multiple k.  Two passes over the input (sizes, then bits) keep host and device memory bounded.
#include <sys/stat.h>
#include <zlib.h>

#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <thread>
#include <cub/cub.cuh>
#include <numeric>
#include <regex>
#include <set>

#include "ctx_internal.h"

namespace kmcpg {
namespace {

// the zlib line reader the builder started with: kept as the yardstick of the loader self-test only
struct LegacyFastxReader {
    gzFile f = nullptr;
    std::string pending;
    bool have_pending = false;
    bool open(const std::string &p) { f = gzopen(p.c_str(), "rb"); if (f) gzbuffer(f, 1 << 20); return f != nullptr; }
    void close() { if (f) gzclose(f); f = nullptr; }
    bool getline(std::string &out) {
        if (have_pending) { out.swap(pending); have_pending = false; return true; }
        out.clear();
        char buf[1 << 16];
        for (;;) {
            if (!gzgets(f, buf, sizeof(buf))) return !out.empty();
            size_t n = strlen(buf);
            out.append(buf, n);
            if (n && buf[n - 1] == '\n') break;
        }
        while (!out.empty() && (out.back() == '\n' || out.back() == '\r')) out.pop_back();
        return true;
    }
    bool next(std::string &header, std::string &seq) {
        std::string l;
        do { if (!getline(l)) return false; } while (l.empty());
        if (l[0] != '>' && l[0] != '@') return false;
        const bool fq = l[0] == '@';
        header.assign(l, 1, std::string::npos);
        seq.clear();
        if (fq) {
            if (!getline(l)) return true;
            seq = l;
            std::string plus, qual;
            if (!getline(plus)) return true;
            while (plus.empty() || plus[0] != '+') { seq += plus; if (!getline(plus)) return true; }
            size_t got = 0;
            while (got < seq.size() && getline(qual)) got += qual.size();
        } else {
            while (getline(l)) {
                if (!l.empty() && l[0] == '>') { pending.swap(l); have_pending = true; break; }
                seq += l;
            }
        }
        return true;
    }
};

std::string base_name(const std::string &p) { size_t s = p.find_last_of('/'); return s == std::string::npos ? p : p.substr(s + 1); }

std::string trim_ext(const std::string &file) {
    std::string b = base_name(file);
    for (const char *z : {".gz", ".xz", ".zst", ".bz2"}) {
        size_t n = strlen(z);
        if (b.size() > n && b.compare(b.size() - n, n, z) == 0) { b.resize(b.size() - n); break; }
    }
    size_t dot = b.find_last_of('.');
    if (dot != std::string::npos && dot > 0) b.resize(dot);
    return b;
}

std::regex make_regex(std::string pat) {          // Go RE2 `(?i)` prefix → icase flag
    if (pat.rfind("(?i)", 0) == 0) pat = pat.substr(4);
    return std::regex(pat, std::regex::ECMAScript | std::regex::icase | std::regex::optimize);
}

struct Target {
    std::string name;
    uint32_t chunk_idx = 0, n_chunks = 1;
    uint64_t gsize = 0, size = 0;
};

// one input file → the sequences to hash and how they group into targets
struct Genome {
    std::string name;
    uint64_t gsize = 0;
    std::vector<std::string> seqs;        // windows (split mode) or kept records (non-split)
    bool split = false;                   // split mode: one target per sequence; else one target for all
};

// records through the reader stage's parser and gzip decoder (reader.cpp: fastx_reader.h / fastgz.h, about three times zlib)
struct FastxReader {
    FastxFile *f = nullptr;
    std::string err;
    bool failed = false;
    bool open(const std::string &p) { f = fastx_open(p); return f != nullptr; }
    void close() { if (f) fastx_close(f); f = nullptr; }
    bool next(std::string &header, std::string &seq) {
        const int rc = fastx_next(f, header, seq, err);
        if (rc < 0) failed = true;
        return rc == 1;
    }
};

inline bool reader_failed(const FastxReader &r) { return r.failed; }
inline std::string reader_error(const FastxReader &r) { return r.err; }
inline bool reader_failed(const LegacyFastxReader &) { return false; }
inline std::string reader_error(const LegacyFastxReader &) { return std::string(); }

template <class R>
int load_genome_with(const std::string &file, const kmcpg_index_params &p, const std::vector<std::regex> &filters, const std::regex *name_re, Genome &g,
                     std::string &err);

int load_genome(const std::string &file, const kmcpg_index_params &p, const std::vector<std::regex> &filters, const std::regex *name_re, Genome &g,
                std::string &err) {
    return load_genome_with<FastxReader>(file, p, filters, name_re, g, err);
}

template <class R>
int load_genome_with(const std::string &file, const kmcpg_index_params &p, const std::vector<std::regex> &filters, const std::regex *name_re, Genome &g,
                     std::string &err) {
    R r;
    if (!r.open(file)) { err = "cannot open " + file; return KMCPG_EIO; }
    g = Genome();
    const std::string base = base_name(file);
    std::smatch m;
    g.name = trim_ext(file);
    if (name_re && std::regex_search(base, m, *name_re) && m.size() > 1) g.name = m[1].str();
    std::vector<std::string> recs;
    std::string header, seq;
    while (r.next(header, seq)) {
        bool drop = false;
        for (auto &f : filters) if (std::regex_search(header, f)) { drop = true; break; }       // C:587-600
        if (!drop && !seq.empty()) recs.push_back(seq);
    }
    if (reader_failed(r)) { err = reader_error(r); r.close(); return KMCPG_EIO; }
    r.close();
    const int k = p.k;
    g.split = p.split_number > 1;
    if (!g.split) {
        for (auto &s : recs) g.gsize += s.size();                                               // C:668
        g.seqs.swap(recs);
        return KMCPG_OK;
    }
    std::string big;
    for (size_t i = 0; i < recs.size(); i++) { if (i) big.append((size_t)(k - 1), 'N'); big += recs[i]; }   // C:612-626
    g.gsize = big.size();
    const uint64_t L = big.size();
    if (L == 0) return KMCPG_OK;
    uint64_t size = L, step = L;
    if (L >= (uint64_t)p.split_min_ref) {                                                       // C:676
        size = (L + (uint64_t)(p.split_number - 1) * p.split_overlap + p.split_number - 1) / p.split_number;   // C:691
        step = size - p.split_overlap;
        if (size <= (uint64_t)p.split_overlap) { err = "split overlap too large for " + file; return KMCPG_EINVAL; }
    }
    for (uint64_t i = 0; i < L; i += step) {                                                    // seq.Slider(size, step, false, greedy)
        const uint64_t len = std::min<uint64_t>(size, L - i);
        if ((int64_t)len - 1 <= p.split_overlap || len < (uint64_t)k) continue;                 // C:713, 742
        g.seqs.push_back(big.substr(i, len));
    }
    return KMCPG_OK;
}

// The input files loaded a few ahead of the thread that feeds the GPU, by several threads (inflate + parse + windows are pure
// host work, about 0.25 GB/s per thread; the kernels need a tiny fraction of that time).  get(i) returns file i, in order.
class GenomeLoader {
  public:
    GenomeLoader(const char *const *files, int n_files, const kmcpg_index_params &p, const std::vector<std::regex> &filters, const std::regex *name_re,
                 int threads)
        : files_(files), n_(n_files), p_(p), filters_(filters), name_re_(name_re), slots_((size_t)n_files) {
        threads = std::max(1, std::min(threads, n_files));
        window_ = 2 * threads + 2;
        for (int t = 0; t < threads; t++) th_.emplace_back([this] { work(); });
    }
    ~GenomeLoader() {
        { std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
        cv_.notify_all();
        for (auto &t : th_) t.join();
    }
    int get(int i, Genome &g, std::string &err) {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return slots_[(size_t)i].ready; });
        Slot &s = slots_[(size_t)i];
        g = std::move(s.g);
        err = s.err;
        const int rc = s.rc;
        s.g = Genome();
        consumed_ = i + 1;
        cv_.notify_all();
        return rc;
    }

  private:
    struct Slot { Genome g; int rc = KMCPG_OK; std::string err; bool ready = false; };
    void work() {
        for (;;) {
            int i;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return stop_ || (next_ < n_ && next_ < consumed_ + window_); });
                if (stop_) return;
                i = next_++;
            }
            Genome g;
            std::string err;
            const int rc = load_genome(files_[i], p_, filters_, name_re_, g, err);
            std::lock_guard<std::mutex> lk(mu_);
            Slot &s = slots_[(size_t)i];
            s.g = std::move(g); s.rc = rc; s.err = err; s.ready = true;
            cv_.notify_all();
        }
    }
    const char *const *files_;
    int n_;
    const kmcpg_index_params &p_;
    const std::vector<std::regex> &filters_;
    const std::regex *name_re_;
    std::vector<Slot> slots_;
    std::vector<std::thread> th_;
    std::mutex mu_;
    std::condition_variable cv_;
    int next_ = 0, consumed_ = 0, window_ = 4;
    bool stop_ = false;
};

// unique count of sorted segments: one warp per segment
__global__ void count_unique_kernel(const uint64_t *__restrict__ codes, const int *__restrict__ seg_begin, const int *__restrict__ seg_end, uint32_t n_seg,
                                    uint64_t *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t s = warp; s < n_seg; s += n_warps) {
        const int b = seg_begin[s], e = seg_end[s];
        uint32_t cnt = 0;
        for (int i = b + lane; i < e; i += 32) cnt += (i == b || codes[i] != codes[i - 1]) ? 1u : 0u;
        for (int d = 16; d > 0; d >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, d);
        if (lane == 0) out[s] = cnt;
    }
}

}  // namespace
}  // namespace kmcpg

using namespace kmcpg;

extern "C" {

void kmcpg_default_index_params(kmcpg_index_params *p) {
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->k = 21; p->num_hashes = 1; p->fpr = 0.3; p->split_number = 1; p->split_overlap = -1; p->split_min_ref = 1000;   // compute.go:1040-1068, index.go:1444-1450
    p->scale = 1; p->threads = 16;
}

int kmcpg_index_fasta(kmcpg_ctx *ctx, const kmcpg_index_params *pin, const char *const *files, int n_files, const char *out_dir) {
    if (!ctx || !pin || !files || n_files < 1) return KMCPG_EINVAL;
    kmcpg_index_params p = *pin;
    if (p.k < 1 || p.k > 64) return fail(ctx, KMCPG_EINVAL, "k must be in 1..64");
    if (p.num_hashes < 1 || p.num_hashes > 4) return fail(ctx, KMCPG_EINVAL, "value of -n/--num-hash too big");           // I:196
    if (!(p.fpr > 0 && p.fpr < 1)) return fail(ctx, KMCPG_EINVAL, "false positive rate must be in (0,1)");
    if (p.split_number < 1) p.split_number = 1;
    if (p.split_number > 65535) return fail(ctx, KMCPG_EINVAL, "value of flag -s/--split-number should not be greater than 65535");
    if (p.split_overlap < 0) p.split_overlap = p.k - 1;                                                                  // C:268-270
    if (p.syncmer_s > 0 && p.minimizer_w > 0) return fail(ctx, KMCPG_EINVAL, "flag -W/--minimizer-w and -S/--syncmer-s are incompatible");
    if (p.syncmer_s > 0 && (int)p.syncmer_s >= p.k) return fail(ctx, KMCPG_EINVAL, "syncmer-s must be smaller than k");
    std::lock_guard<std::mutex> lk(ctx->mu);
    executor_drain(ctx);
    CU(cudaSetDevice(ctx->device));
    free_db(ctx);
    cudaStream_t st = ctx->st;
    WorkSet &w = ctx->ws[0];

    std::vector<std::regex> filters;
    std::regex name_re_obj;
    const std::regex *name_re = nullptr;
    try {
        if (p.seq_name_filters)
            for (int i = 0; i < p.n_seq_name_filters; i++) filters.push_back(make_regex(p.seq_name_filters[i]));
        if (p.ref_name_regexp && *p.ref_name_regexp) { name_re_obj = make_regex(p.ref_name_regexp); name_re = &name_re_obj; }
    } catch (const std::regex_error &e) {
        return fail(ctx, KMCPG_EINVAL, std::string("failed to parse regular expression: ") + e.what());
    }

    DbMeta &m = ctx->meta;
    m = DbMeta();
    m.dir = out_dir ? out_dir : "<memory>";
    m.version = 4; m.index_version = 4; m.ks = {p.k}; m.canonical = true; m.num_hashes = p.num_hashes; m.fpr = p.fpr;
    m.scaled = p.scale > 1; m.scale = m.scaled ? p.scale : 0;
    m.minimizer = p.minimizer_w > 0; m.minimizer_w = p.minimizer_w; m.syncmer = p.syncmer_s > 0; m.syncmer_s = p.syncmer_s;

    kmcpg_search_params hp;
    kmcpg_default_params(&hp);
    hp.min_query_len = 0; hp.min_matched = 1; hp.dedup_threshold = 0x7fffffff; hp.min_query_cov = 0;   // plain code lists per sequence

    std::vector<Target> targets;                  // in file order, chunks in order
    std::vector<uint32_t> first_target(n_files, 0);
    std::vector<uint32_t> order, pos_of;
    int block_size = 0;
    const uint64_t BATCH_BYTES = 192ull << 20;
    std::string err;

    for (int pass = 0; pass < 2; pass++) {
        if (pass == 1) {
            // ---- plan: I:667 ascending by k-mer count (stable), I:670-682 block size, I:936-1023 numSigs ----
            std::vector<uint32_t> keep;
            for (uint32_t t = 0; t < targets.size(); t++) if (targets[t].size > 0) keep.push_back(t);      // empty sets are skipped (I:798-800)
            if (keep.empty()) return fail(ctx, KMCPG_EINVAL, "no valid sequences / k-mers in the input files");
            order = keep;
            std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return targets[a].size < targets[b].size; });
            pos_of.assign(targets.size(), 0xFFFFFFFFu);
            for (uint32_t i = 0; i < order.size(); i++) pos_of[order[i]] = i;
            const int nf = (int)order.size();
            block_size = p.block_size > 0 ? p.block_size : ((int)((double)nf / (double)std::max(1, p.threads)) + 7) / 8 * 8;
            if (p.block_size <= 0) { if (block_size > nf) block_size = nf; if (block_size < 8) block_size = 8; }
            const uint64_t nb = ((uint64_t)nf + block_size - 1) / block_size;
            m.blocks.resize(nb);
            for (uint64_t bi = 0; bi < nb; bi++) {
                BlockMeta &bm = m.blocks[bi];
                const uint64_t t0 = bi * block_size, t1 = std::min<uint64_t>((uint64_t)nf, t0 + block_size);
                char fn[64];
                snprintf(fn, sizeof(fn), "_block%03llu.uniki", (unsigned long long)(bi + 1));       // I:637
                bm.path = fn; bm.k = p.k; bm.canonical = true; bm.num_hashes = p.num_hashes;
                bm.n_names = (int)(t1 - t0); bm.row_bytes = (bm.n_names + 7) / 8; bm.target_base = (int64_t)t0;
                uint64_t mx = 0;
                for (uint64_t t = t0; t < t1; t++) {
                    const Target &tg = targets[order[t]];
                    bm.names.push_back(tg.name); bm.indices.push_back(tg.chunk_idx | (tg.n_chunks << 16));
                    bm.gsizes.push_back(tg.gsize); bm.sizes.push_back(tg.size);
                    mx = std::max(mx, tg.size);
                }
                bm.num_sigs = calc_signature_size(mx, p.num_hashes, p.fpr);
                if (bm.num_sigs == 0) return fail(ctx, KMCPG_EUNSUPPORTED, "block has an unsupported number of signatures");
                DeviceBlock db;
                db.meta_idx = (int)bi;
                layout_block(db, bm);
                db.bytes = (size_t)bm.num_sigs * db.pitch;
                CU(cudaMalloc((void **)&db.d_rows, std::max<size_t>(db.bytes, 16)));
                CU(cudaMemsetAsync(db.d_rows, 0, db.bytes, st));
                ctx->blocks.push_back(db);
                ctx->sum_row_bytes += bm.row_bytes;
                ctx->resident_bytes += (int64_t)db.bytes;
                ctx->disk_bytes += (int64_t)(bm.num_sigs * (uint64_t)bm.row_bytes);
                m.files.push_back(fn);
            }
            m.n_targets = nf;
            ctx->resident_of.resize(nb);
            std::iota(ctx->resident_of.begin(), ctx->resident_of.end(), 0);
            ctx->target_sizes.resize((size_t)nf);
            for (auto &bm : m.blocks)
                for (int c = 0; c < bm.n_names; c++) ctx->target_sizes[(size_t)bm.target_base + c] = (double)bm.sizes[c];
        }

        // ---- stream the files in batches of sequences ----
        std::vector<uint8_t> bytes;
        std::vector<uint64_t> off(1, 0);
        std::vector<uint32_t> seq_target;                    // target of every batched sequence
        auto flush = [&]() -> int {
            const uint32_t ns = (uint32_t)seq_target.size();
            if (!ns) return KMCPG_OK;
            uint64_t total = 0, mx = 0;
            for (uint32_t s = 0; s < ns; s++) { uint64_t len = off[s + 1] - off[s]; uint64_t c = len >= (uint64_t)p.k ? len - p.k + 1 : 0; total += c; mx = std::max(mx, c); }
            if (total >= (1ull << 31)) return fail(ctx, KMCPG_EUNSUPPORTED, "batch too large");
            CU(w.seq.ensure(bytes.size() + 64)); CU(w.off.ensure((ns + 1) * 8ull));
            CU(cudaMemcpyAsync(w.seq.p, bytes.data(), bytes.size(), cudaMemcpyHostToDevice, st));
            CU(cudaMemcpyAsync(w.off.p, off.data(), (ns + 1) * 8ull, cudaMemcpyHostToDevice, st));
            SubBatch sb{w.seq.as<uint8_t>(), w.off.as<uint64_t>(), ns, total, mx, 0};
            uint64_t *codes = nullptr;
            int rc = run_hash_stage(ctx, w, hp, p.k, sb, ns, &codes);
            if (rc) return rc;
            std::vector<uint32_t> nc(ns);
            std::vector<uint64_t> so(ns + 1);
            CU(cudaMemcpyAsync(nc.data(), w.ncodes.p, ns * 4ull, cudaMemcpyDeviceToHost, st));
            CU(cudaMemcpyAsync(so.data(), w.slot_off.p, (ns + 1) * 8ull, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            for (auto &x : nc) if (x == 0xFFFFFFFFu) x = 0;
            if (pass == 1) {
                for (uint32_t s = 0; s < ns; s++) {
                    const uint32_t pos = pos_of[seq_target[s]];
                    if (pos == 0xFFFFFFFFu || !nc[s]) continue;
                    DeviceBlock &db = ctx->blocks[pos / block_size];
                    CU(launch_set_bits(codes + so[s], nc[s], p.num_hashes, db.fm, db.d_rows, db.pitch, pos % block_size, st));
                }
                CU(cudaStreamSynchronize(st));
            } else {
                // gather the code lists of every target contiguously, sort per target, count distinct (C:815-823 / 905-914)
                std::vector<int> segb, sege;
                std::vector<uint32_t> seg_target;
                CU(w.codes2.ensure(std::max<uint64_t>(total, 1) * 8));
                uint64_t dst = 0;
                for (uint32_t s = 0; s < ns; s++) {
                    if (seg_target.empty() || seg_target.back() != seq_target[s]) { if (!seg_target.empty()) sege.push_back((int)dst); seg_target.push_back(seq_target[s]); segb.push_back((int)dst); }
                    if (nc[s]) CU(cudaMemcpyAsync(w.codes2.as<uint64_t>() + dst, codes + so[s], nc[s] * 8ull, cudaMemcpyDeviceToDevice, st));
                    dst += nc[s];
                }
                sege.push_back((int)dst);
                const uint32_t nseg = (uint32_t)seg_target.size();
                CU(w.segb.ensure(nseg * 4ull)); CU(w.sege.ensure(nseg * 4ull)); CU(w.codes.ensure(std::max<uint64_t>(dst, 1) * 8)); CU(ctx->d_scal.ensure(nseg * 8ull + 16));
                CU(cudaMemcpyAsync(w.segb.p, segb.data(), nseg * 4ull, cudaMemcpyHostToDevice, st));
                CU(cudaMemcpyAsync(w.sege.p, sege.data(), nseg * 4ull, cudaMemcpyHostToDevice, st));
                if (dst) {
                    size_t t2 = 0;
                    cub::DeviceSegmentedSort::SortKeys(nullptr, t2, w.codes2.as<uint64_t>(), w.codes.as<uint64_t>(), (int)dst, (int)nseg, w.segb.as<int>(), w.sege.as<int>(), st);
                    CU(w.tmp.ensure(t2));
                    CU(cub::DeviceSegmentedSort::SortKeys(w.tmp.p, t2, w.codes2.as<uint64_t>(), w.codes.as<uint64_t>(), (int)dst, (int)nseg, w.segb.as<int>(),
                                                          w.sege.as<int>(), st));
                }
                count_unique_kernel<<<std::min<uint32_t>((nseg + 7) / 8, 148 * 8), 256, 0, st>>>(w.codes.as<uint64_t>(), w.segb.as<int>(), w.sege.as<int>(), nseg,
                                                                                                  ctx->d_scal.as<uint64_t>());
                CU(cudaGetLastError());
                std::vector<uint64_t> uq(nseg);
                CU(cudaMemcpyAsync(uq.data(), ctx->d_scal.p, nseg * 8ull, cudaMemcpyDeviceToHost, st));
                CU(cudaStreamSynchronize(st));
                for (uint32_t i = 0; i < nseg; i++) targets[seg_target[i]].size += uq[i];
            }
            bytes.clear(); off.assign(1, 0); seq_target.clear();
            return KMCPG_OK;
        };

        GenomeLoader loader(files, n_files, p, filters, name_re, std::max(1, std::min(16, (int)std::thread::hardware_concurrency() / 4)));
        for (int fi = 0; fi < n_files; fi++) {
            Genome g;
            int rc = loader.get(fi, g, err);
            if (rc) return fail(ctx, rc, err);
            if (pass == 0) {
                first_target[fi] = (uint32_t)targets.size();
                if (g.split) {
                    for (uint32_t c = 0; c < g.seqs.size(); c++) { Target t; t.name = g.name; t.chunk_idx = c; t.n_chunks = (uint32_t)g.seqs.size(); t.gsize = g.gsize; targets.push_back(t); }
                } else if (!g.seqs.empty()) {
                    Target t; t.name = g.name; t.chunk_idx = 0; t.n_chunks = 1; t.gsize = g.gsize; targets.push_back(t);
                }
            }
            // a target's sequences must stay in one batch (its code set is sorted as one segment)
            uint64_t gbytes = 0;
            for (auto &s : g.seqs) gbytes += s.size();
            if (!g.split && !bytes.empty() && bytes.size() + gbytes > BATCH_BYTES) { rc = flush(); if (rc) return rc; }
            for (uint32_t s = 0; s < g.seqs.size(); s++) {
                if (g.split && !bytes.empty() && bytes.size() + g.seqs[s].size() > BATCH_BYTES) { rc = flush(); if (rc) return rc; }
                bytes.insert(bytes.end(), g.seqs[s].begin(), g.seqs[s].end());
                off.push_back(bytes.size());
                seq_target.push_back(first_target[fi] + (g.split ? s : 0));
            }
        }
        int rc = flush();
        if (rc) return rc;
    }
    ctx->has_db = true;

    // ---- files (I:1283-1285, 1353-1399) ----
    if (out_dir && *out_dir) {
        const std::string r001 = std::string(out_dir) + "/R001";
        mkdir(out_dir, 0755);
        mkdir(r001.c_str(), 0755);
        std::vector<uint8_t> host;
        uint64_t total_kmers = 0;
        for (size_t bi = 0; bi < ctx->blocks.size(); bi++) {
            const DeviceBlock &b = ctx->blocks[bi];
            BlockMeta &bm = m.blocks[bi];
            const size_t nbytes = (size_t)bm.num_sigs * bm.row_bytes;
            CU(ctx->d_tmp.ensure(std::max<size_t>(nbytes, 16)));
            CU(launch_unpitch(b.d_rows, ctx->d_tmp.as<uint8_t>(), bm.num_sigs, (uint32_t)bm.row_bytes, b.pitch, st));
            host.resize(nbytes ? nbytes : 1);
            CU(cudaMemcpyAsync(host.data(), ctx->d_tmp.p, nbytes, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            const std::string path = r001 + "/" + bm.path;
            int rc = write_block_file(path, bm, host.data(), err);
            if (rc) return fail(ctx, rc, err);
            bm.path = path;
            for (auto s : bm.sizes) total_kmers += s;
        }
        FILE *f = fopen((r001 + "/__db.yml").c_str(), "w");                          // util-db-info.go:46-79 key order
        if (!f) return fail(ctx, KMCPG_EIO, "fail to write kmcp database info file: " + r001 + "/__db.yml");
        fprintf(f, "version: 4\nunikiVersion: 4\nalias: %s\nk: %d\nks:\n- %d\nhashed: true\ncanonical: true\n", base_name(out_dir).c_str(), p.k, p.k);
        fprintf(f, "scaled: %s\nscale: %u\nminimizer: %s\nminimizer-w: %u\nsyncmer: %s\nsyncmer-s: %u\n", m.scaled ? "true" : "false", m.scale,
                m.minimizer ? "true" : "false", m.minimizer_w, m.syncmer ? "true" : "false", m.syncmer_s);
        fprintf(f, "split-seq: %s\nsplit-size: 0\nsplit-num: %d\nsplit-overlap: %d\ncompact-size: true\n", p.split_number > 1 ? "true" : "false",
                p.split_number > 1 ? p.split_number : 0, p.split_number > 1 ? p.split_overlap : 0);
        char fprs[64];
        for (int prec = 1; prec <= 17; prec++) { snprintf(fprs, sizeof(fprs), "%.*g", prec, p.fpr); if (strtod(fprs, nullptr) == p.fpr) break; }   // shortest round trip
        fprintf(f, "hashes: %d\nfpr: %s\nnumNameGroups: %lld\nblocksize: %d\ntotalKmers: %llu\nfiles:\n", p.num_hashes, fprs, (long long)m.n_targets, block_size,
                (unsigned long long)total_kmers);
        for (auto &fn : m.files) fprintf(f, "- %s\n", fn.c_str());
        fclose(f);
        std::set<std::string> names;
        for (auto &bm : m.blocks) for (auto &n : bm.names) names.insert(n);
        f = fopen((r001 + "/__name_mapping.tsv").c_str(), "w");
        if (f) { for (auto &n : names) fprintf(f, "%s\t%s\n", n.c_str(), n.c_str()); fclose(f); }
        m.dir = r001;
    }
    return KMCPG_OK;
}

}  // extern "C"

// test hook (host only, tests/test_abi.py): the files through GenomeLoader on `threads` threads and one by one through load_genome —
// a running FNV-1a over (name, genome size, split flag, every sequence) must agree.  Returns 0, or a KMCPG_E* code / -100 on a difference.
extern "C" int kmcpg_internal_genome_loader_selftest(const char *const *files, int n_files, int k, int split_number, int split_overlap, int threads,
                                                       uint64_t *digest_out) {
    if (!files || n_files < 1) return KMCPG_EINVAL;
    kmcpg_index_params p;
    kmcpg_default_index_params(&p);
    p.k = k; p.split_number = split_number < 1 ? 1 : split_number; p.split_overlap = split_overlap < 0 ? k - 1 : split_overlap;
    std::vector<std::regex> filters;
    auto fold = [](uint64_t h, const void *d, size_t n) { const uint8_t *b = (const uint8_t *)d; for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; } return h; };
    auto digest = [&](uint64_t h, const Genome &g) {
        h = fold(h, g.name.data(), g.name.size()); h = fold(h, &g.gsize, 8);
        const uint8_t sp = g.split; h = fold(h, &sp, 1);
        for (auto &s : g.seqs) { const uint64_t n = s.size(); h = fold(h, &n, 8); h = fold(h, s.data(), s.size()); }
        return h;
    };
    uint64_t a = 1469598103934665603ull, b = a;
    std::string err;
    {
        GenomeLoader loader(files, n_files, p, filters, nullptr, threads);
        for (int i = 0; i < n_files; i++) { Genome g; const int rc = loader.get(i, g, err); if (rc) return rc; a = digest(a, g); }
    }
    for (int i = 0; i < n_files; i++) {        // one by one, through the zlib line reader
        Genome g;
        const int rc = load_genome_with<LegacyFastxReader>(files[i], p, filters, nullptr, g, err);
        if (rc) return rc;
        b = digest(b, g);
    }
    if (digest_out) *digest_out = a;
    return a == b ? KMCPG_OK : -100;
}

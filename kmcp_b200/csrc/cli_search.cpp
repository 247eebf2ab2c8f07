// kmcp-gpu — drop-in for `kmcp search` on B200: same flags, same input formats, same 15-column TSV.
// Mirrors kmcp/cmd/search.go of the reference: flag set S:1031-1107 (names, shorthands, defaults), DB discovery
// S:299-324, query construction S:793-1000 (single-end, paired-end -1/-2, whole-file -g), ordered output and
// the 15 columns S:437 + S:460-575, trailer S:1023-1025, log lines S:1011-1017.
// All searching goes through libkmcp_gpu (include/kmcp_gpu.h); there is no CPU search path in this program.
#include <dirent.h>
#include <sys/stat.h>
#include <zlib.h>
#include <fcntl.h>
#include <unistd.h>
#include <errno.h>

#include <algorithm>
#include <atomic>
#include <cmath>

#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <deque>
#include <future>
#include <map>
#include <mutex>
#include <regex>
#include <string>
#include <thread>
#include <vector>

#include "nvtx_ranges.h"
#include "../../include/kmcp_gpu.h"
#include "fastx_reader.h"
#include "tsv_format.h"
#if defined(__x86_64__)
#include <immintrin.h>
#endif

extern "C" void kmcpg_internal_reader_stats(uint64_t *pieces, uint64_t *fallbacks);      // test hook of reader.cpp
extern "C" int kmcpg_internal_unmatched_results(const uint64_t *off, uint32_t n_seqs, int paired, int k, kmcpg_results *out);   // test hook of engine.cpp

namespace {

using fastx::Reader;

bool g_quiet = false;
FILE *g_log = nullptr;

void logf(const char *level, const char *fmt, ...) {
    char buf[4096];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    auto now = std::chrono::system_clock::now();
    time_t t = std::chrono::system_clock::to_time_t(now);
    int ms = (int)(std::chrono::duration_cast<std::chrono::milliseconds>(now.time_since_epoch()).count() % 1000);
    struct tm tmv;
    localtime_r(&t, &tmv);
    char ts[32];
    strftime(ts, sizeof(ts), "%H:%M:%S", &tmv);
    bool err = !strcmp(level, "ERRO");
    if (!g_quiet || err) fprintf(stderr, "%s.%03d [%s] %s\n", ts, ms, level, buf);
    if (g_log) fprintf(g_log, "%s.%03d [%s] %s\n", ts, ms, level, buf);
}

[[noreturn]] void die(const char *fmt, ...) {     // checkError → log + os.Exit(-1) (util-cli.go:35-40)
    char buf[4096];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    logf("ERRO", "%s", buf);
    fflush(stderr);
    if (g_log) fflush(g_log);
    _exit(255);          // like os.Exit: at once, without unwinding under the reader / writer / engine threads that are still running
}

bool is_dir(const std::string &p) { struct stat st; return stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode); }
bool is_file(const std::string &p) { struct stat st; return stat(p.c_str(), &st) == 0 && S_ISREG(st.st_mode); }

struct Less {                       // the engine's match order (U:105-145), used again when several databases are merged
    int sort_by;
    int score(const kmcpg_match &a, const kmcpg_match &b) const {      // <0: a first
        if (sort_by == 0) { if (a.qcov != b.qcov) return a.qcov > b.qcov ? -1 : 1; if (a.tcov != b.tcov) return a.tcov > b.tcov ? -1 : 1; }
        else if (sort_by == 1) { if (a.tcov != b.tcov) return a.tcov > b.tcov ? -1 : 1; if (a.count != b.count) return a.count > b.count ? -1 : 1; }
        else { if (a.jacc != b.jacc) return a.jacc > b.jacc ? -1 : 1; if (a.count != b.count) return a.count > b.count ? -1 : 1; }
        return 0;
    }
};

struct Opts {
    std::string out_file = "-", read1, read2, sort_by = "qcov", query_id, log_file, ref_counts_file;
    double rc_min_qcov = 0.55, rc_max_fpr = 0.01;   // `kmcp profile` -t / -f for --ref-counts
    std::vector<std::string> db_dirs, files, name_maps;   // several -d: every database is searched, results merged as `kmcp merge` does
    int dedup = 256, min_kmers = 10, min_qlen = 30, top_scores = 0, threads = 0;
    std::vector<int> devices{0};     // --gpu N, or --gpus LIST|all: every database is sharded over these devices
    bool all_devices = false;
    std::string gpu_mode = "auto";   // --gpu-mode: shard (index split over the devices), replicate (reads split), auto (replicate when the index fits)
    double qcov = 0.55, tcov = 0, max_fpr = 0.01;
    bool dry_run = false;            // --dry-run: no device, no database: every query comes out unmatched (host-only tests of the plumbing)
    bool try_se = false, whole_file = false, use_filename = false, default_name_map = false, keep_unmatched = false, no_header = false,
         do_not_sort = false;
    size_t batch_reads = 1u << 18, batch_bytes = 256u << 20;
};

void usage() {
    fputs(
        "Search sequences against a kmcp database on a B200 GPU (drop-in for `kmcp search`)\n\n"
        "Usage:\n  kmcp-gpu search [flags] [-w] -d <kmcp db> [-t <min-query-cov>] [read1.fq.gz] [read2.fq.gz] [unpaired.fq.gz] [-o read.tsv.gz]\n\n"
        "Flags (same names, shorthands and defaults as kmcp search):\n"
        "  -d, --db-dir string              database directory created by \"kmcp index\"; may be given several times: the hits of a\n"
        "                                   query in all of them are united and re-sorted, as \"kmcp merge\" does with the outputs of\n"
        "                                   separate searches (-n/--keep-top-scores applies per database, as it would there;\n"
        "                                   databases with several sub-databases R001, R002, ... are not supported)\n"
        "  -1, --read1 string / -2, --read2 string   paired-end read files\n"
        "      --try-se                     if paired-end reads have no hits, re-search with read1, then read2\n"
        "  -u, --kmer-dedup-threshold int   remove duplicated kmers for a query with >= X k-mers (default 256)\n"
        "  -g, --query-whole-file           use the whole file as a query\n"
        "  -G, --use-filename               use file name as query ID with -g\n"
        "      --query-id string            custom query ID with -g\n"
        "  -c, --min-kmers int              minimum number of matched k-mers (default 10)\n"
        "  -m, --min-query-len int          minimum query length (default 30)\n"
        "  -t, --min-query-cov float        minimum query coverage (default 0.55)\n"
        "  -T, --min-target-cov float       minimum target coverage (default 0)\n"
        "  -f, --max-fpr float              maximum false positive rate of a query (default 0.01)\n"
        "  -o, --out-file string            out file, \".gz\" suffix supported (default \"-\")\n"
        "  -N, --name-map strings           two-column file(s) mapping reference IDs to user-defined values\n"
        "  -D, --default-name-map           load ${db}/__name_mapping.tsv first\n"
        "  -K, --keep-unmatched             keep unmatched query sequence information\n"
        "  -n, --keep-top-scores int        keep matches with the top N scores, 0 for all\n"
        "  -H, --no-header-row              do not print header row\n"
        "  -s, --sort-by string             qcov, tcov or jacc (default \"qcov\")\n"
        "  -S, --do-not-sort                do not sort matches of a query\n"
        "  -w, --load-whole-db / --low-mem  accepted for compatibility (the index always lives in HBM)\n"
        "  -j, --threads int                host threads for the post-filter (default all)\n"
        "  -q, --quiet / --log string       logging\n"
        "  -i, --infile-list string         file of input files list (one file per line)\n"
        "      --gpu int                    CUDA device ordinal (default 0)\n"
        "      --gpus list|all              several devices, e.g. 0,1,2,3: every database is sharded over them by index block\n"
        "                                   (by column range when it has fewer blocks than devices); every device sees every read\n"
        "      --gpu-mode string            with several devices: \"shard\" splits the index (every device searches every read),\n"
        "                                   \"replicate\" loads the whole index on every device and splits the reads,\n"
        "                                   \"auto\" (default) replicates when the index fits into every device's free memory\n"
        "      --compression-level int      level of the .gz output, 1-9 (default 4)\n"
        "      --inflate-threads int        threads that decompress ONE .gz input side by side (chunk-parallel inflate; default: by itself\n"
        "                                   on machines with >= 32 hardware threads for files >= 32 MB; 1 = sequential decoder)\n"
        "      --parse-threads int          threads that parse ONE FASTQ input side by side (default 1: one parser thread per input)\n"
        "      --ref-counts file            also write the per-reference, per-chunk read counters of `kmcp profile` stage 1/4\n"
        "                                   (match, uniqMatch, uniqMatchHic), computed from the result stream\n"
        "      --ref-counts-min-qcov float  profile -t/--min-query-cov for --ref-counts (default 0.55)\n"
        "      --ref-counts-max-fpr float   profile -f/--max-fpr for --ref-counts (default 0.01)\n",
        stderr);
}

// one complete gzip member for a block of text (concatenated members are a valid .gz stream, as pgzip writes them)
std::string gz_member(const char *data, size_t n, int level) {
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (deflateInit2(&zs, level, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) die("zlib init failed");
    std::string out(deflateBound(&zs, (uLong)n) + 64, '\0');
    zs.next_in = (Bytef *)data; zs.avail_in = (uInt)n;
    zs.next_out = (Bytef *)&out[0]; zs.avail_out = (uInt)out.size();
    if (deflate(&zs, Z_FINISH) != Z_STREAM_END) die("zlib deflate failed");
    out.resize(zs.total_out);
    deflateEnd(&zs);
    return out;
}

fastx::Tuning g_tune;             // --inflate-threads, --parse-threads and the test knobs behind them
int g_compression_level = 4;      // --compression-level (the reference's pgzip default is a fast level as well)

// text → file.  ".gz" output is compressed in 1 MB blocks by several threads (independent gzip members, written in order) and
// write() only queues the text: the next batch is formatted while this one is still being compressed and written.
struct Writer {
    FILE *fp = nullptr;
    bool gz = false;
    size_t max_inflight = 64;                       // blocks being compressed or waiting to be written
    std::thread flusher;
    std::mutex mu;
    std::condition_variable cv;
    std::deque<std::future<std::string>> q;
    bool closing = false;
    void open(const std::string &p) {
        if (p == "-") fp = stdout;
        else { fp = fopen(p.c_str(), "wb"); if (!fp) die("fail to write %s", p.c_str()); }
        gz = p.size() > 3 && p.compare(p.size() - 3, 3, ".gz") == 0;
        if (gz) {
            max_inflight = std::max<size_t>(8, std::min<size_t>(64, std::thread::hardware_concurrency()));
            flusher = std::thread([this] {
                for (;;) {
                    std::future<std::string> f;
                    {
                        std::unique_lock<std::mutex> lk(mu);
                        cv.wait(lk, [&] { return !q.empty() || closing; });
                        if (q.empty()) return;
                        f = std::move(q.front());
                    }
                    const std::string z = f.get();                 // members leave in the order they were queued
                    if (fwrite(z.data(), 1, z.size(), fp) != z.size()) die("write error");
                    std::lock_guard<std::mutex> lk(mu);
                    q.pop_front();                                 // only now: the block counts as in flight until it is on disk
                    cv.notify_all();
                }
            });
        }
    }
    void write(std::string &&t) {
        if (t.empty()) return;
        if (!gz) { if (fwrite(t.data(), 1, t.size(), fp) != t.size()) die("write error"); return; }
        const size_t BLK = 1u << 20;
        auto keep = std::make_shared<const std::string>(std::move(t));
        const int level = g_compression_level;
        for (size_t o = 0; o < keep->size(); o += BLK) {
            const size_t len = std::min(BLK, keep->size() - o);
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return q.size() < max_inflight; });
            q.push_back(std::async(std::launch::async, [keep, o, len, level] { return gz_member(keep->data() + o, len, level); }));
            cv.notify_all();
        }
    }
    void write(const char *s, size_t n) { write(std::string(s, n)); }
    void write(const std::string &t) { write(std::string(t)); }
    void close() {
        if (flusher.joinable()) {
            { std::lock_guard<std::mutex> lk(mu); closing = true; }
            cv.notify_all();
            flusher.join();
        }
        if (fp && fp != stdout) fclose(fp); else if (fp) fflush(fp);
        fp = nullptr;
    }
};


// a batch from the library's reader stage (kmcpg_reader_*), released with the object
struct Batch {
    kmcpg_read_batch b;
    Batch() { memset(&b, 0, sizeof(b)); }
    Batch(const Batch &) = delete;
    Batch &operator=(const Batch &) = delete;
    ~Batch() { kmcpg_reader_free_batch(&b); }
    size_t n_ids() const { return b.n_queries; }
    const char *id(size_t q) const { return b.ids + b.id_off[q]; }
    size_t id_len(size_t q) const { return (size_t)(b.id_off[q + 1] - b.id_off[q]); }
};

struct ReaderSetup {
    bool paired = false, whole_file = false, use_filename = false;
    std::string read1, read2, query_id;
    std::vector<std::string> files;
    size_t batch_reads = 0, batch_bytes = 0;
    int kmax = 21;
};
// S:793-1000 through the library: every batch of the input, in order, to `take` (which owns it)
void read_batches(const ReaderSetup &c, const std::function<void(Batch *)> &take) {
    kmcpg_reader_opts ro;
    kmcpg_default_reader_opts(&ro);
    std::vector<const char *> fl;
    for (auto &f : c.files) fl.push_back(f.c_str());
    if (c.paired) { ro.read1 = c.read1.c_str(); ro.read2 = c.read2.c_str(); }
    else { ro.files = fl.data(); ro.n_files = (int32_t)fl.size(); }
    ro.whole_file = c.whole_file; ro.use_filename = c.use_filename; ro.query_id = c.query_id.empty() ? nullptr : c.query_id.c_str();
    ro.k = c.kmax; ro.batch_reads = (uint32_t)c.batch_reads; ro.batch_bytes = c.batch_bytes;
    ro.inflate_threads = g_tune.inflate_threads; ro.parse_threads = g_tune.parse_threads;
    ro.inflate_chunk = g_tune.inflate_chunk; ro.inflate_cap = g_tune.inflate_cap; ro.parse_piece = g_tune.parse_piece;
    ro.log = [](void *, const char *level, const char *msg) { logf(level, "%s", msg); };
    kmcpg_reader *rd = nullptr;
    if (kmcpg_reader_open(&ro, &rd) != KMCPG_OK) die("cannot start the reader: invalid input files");
    for (;;) {
        Batch *bt = new Batch();
        const int rc = kmcpg_reader_next(rd, &bt->b);
        if (rc == 1) { take(bt); continue; }
        delete bt;
        if (rc < 0) die("%s", kmcpg_reader_error(rd));
        break;
    }
    kmcpg_reader_close(rd);
}

void load_kv(const std::string &path, std::map<std::string, std::string> &m) {
    Reader r;
    if (!r.open(path)) die("fail to read name mapping file: %s", path.c_str());
    std::string l;
    while (r.getline(l)) {
        size_t t = l.find('\t');
        if (l.empty() || t == std::string::npos) continue;
        m[l.substr(0, t)] = l.substr(t + 1);
    }
    r.close();
}

}  // namespace

// kmcp-gpu index: `kmcp compute` + `kmcp index` in one step on the GPU (flags keep the reference's names; -n is
// compute's --split-number, index's -n/--num-hash is spelled --num-hash here)
int index_main(int argc, char **argv) {
    kmcpg_index_params p;
    kmcpg_default_index_params(&p);
    std::string in_dir, out_dir, file_re = "\\.(f[aq](st[aq])?|fna)(.gz)?$", name_re = "(?i)(.+)\\.(f[aq](st[aq])?|fna)(.gz)?$";
    std::vector<std::string> files, filters;
    bool force = false;
    int device = 0;
    auto need = [&](int &i) -> const char * { if (i + 1 >= argc) die("flag needs an argument: %s", argv[i]); return argv[++i]; };
    for (int i = 2; i < argc; i++) {
        std::string a = argv[i];
        if (a == "-I" || a == "--in-dir") in_dir = need(i);
        else if (a == "-r" || a == "--file-regexp") file_re = need(i);
        else if (a == "-O" || a == "--out-dir") out_dir = need(i);
        else if (a == "-k" || a == "--kmer") p.k = atoi(need(i));
        else if (a == "-n" || a == "--split-number") p.split_number = atoi(need(i));
        else if (a == "-l" || a == "--split-overlap") p.split_overlap = atoi(need(i));
        else if (a == "-m" || a == "--split-min-ref") p.split_min_ref = atoi(need(i));
        else if (a == "-B" || a == "--seq-name-filter") filters.push_back(need(i));
        else if (a == "-N" || a == "--ref-name-regexp") name_re = need(i);
        else if (a == "-D" || a == "--scale") p.scale = (uint32_t)atoi(need(i));
        else if (a == "-W" || a == "--minimizer-w") p.minimizer_w = (uint32_t)atoi(need(i));
        else if (a == "-S" || a == "--syncmer-s") p.syncmer_s = (uint32_t)atoi(need(i));
        else if (a == "-f" || a == "--false-positive-rate") p.fpr = atof(need(i));
        else if (a == "--num-hash") p.num_hashes = atoi(need(i));
        else if (a == "-b" || a == "--block-size") p.block_size = atoi(need(i));
        else if (a == "-j" || a == "--threads") p.threads = atoi(need(i));
        else if (a == "--force") force = true;
        else if (a == "-q" || a == "--quiet") g_quiet = true;
        else if (a == "--gpu") device = atoi(need(i));
        else if (a == "-h" || a == "--help") {
            fputs("kmcp-gpu index [-I <dir> | files...] -O <out.kmcp> [-k 21] [-n split-number] [-l split-overlap] [-B regexp]... [-N regexp]\n"
                  "               [-D scale] [-W minimizer-w] [-S syncmer-s] [-f fpr] [--num-hash n] [-b block-size] [-j threads] [--force]\n", stderr);
            return 0;
        } else if (a.size() > 1 && a[0] == '-') die("unknown flag: %s", a.c_str());
        else files.push_back(a);
    }
    if (out_dir.empty()) die("flag -O/--out-dir needed");
    if (is_dir(out_dir) && !force) die("out-dir not empty: %s, use --force to overwrite", out_dir.c_str());
    if (!in_dir.empty()) {
        std::regex re(file_re, std::regex::ECMAScript | std::regex::icase);
        DIR *d = opendir(in_dir.c_str());
        if (!d) die("fail to read directory: %s", in_dir.c_str());
        while (dirent *e = readdir(d)) { std::string n = e->d_name; if (std::regex_search(n, re) && is_file(in_dir + "/" + n)) files.push_back(in_dir + "/" + n); }
        closedir(d);
    }
    std::sort(files.begin(), files.end());
    if (files.empty()) die("no files given");
    logf("INFO", "kmcp-gpu index: %zu input file(s)", files.size());
    kmcpg_ctx *ctx = nullptr;
    if (kmcpg_create(device, &ctx)) die("%s", kmcpg_last_error(nullptr));
    std::vector<const char *> fp, flt;
    for (auto &f : files) fp.push_back(f.c_str());
    for (auto &f : filters) flt.push_back(f.c_str());
    p.ref_name_regexp = name_re.c_str();
    p.seq_name_filters = flt.empty() ? nullptr : flt.data();
    p.n_seq_name_filters = (int)flt.size();
    auto t0 = std::chrono::steady_clock::now();
    if (kmcpg_index_fasta(ctx, &p, fp.data(), (int)fp.size(), out_dir.c_str())) die("%s", kmcpg_last_error(ctx));
    kmcpg_db_info_t info;
    kmcpg_db_info(ctx, &info);
    {   // `kmcp index` moves references with more than -x 10M / -8 20M / -1 200M k-mers into blocks of -X 256 / 8 / 1 (I:213-259, 787-880) so that one
        // huge genome does not set the Bloom-filter size of a whole -b wide block; this builder keeps one block width
        uint64_t big = 0, mx = 0;
        for (int64_t t = 0; t < info.n_targets; t++) {
            kmcpg_target_t tg;
            if (kmcpg_target(ctx, t, &tg) == 0) { big += tg.n_kmers > 10000000ull; mx = std::max<uint64_t>(mx, tg.n_kmers); }
        }
        if (big)
            logf("WARN", "%llu target(s) hold more than 10M k-mers (largest: %llu): `kmcp index` would give them narrower blocks (-x/-8/-1); "
                         "this build keeps -b wide blocks, so the index is larger than the reference's", (unsigned long long)big, (unsigned long long)mx);
    }
    logf("INFO", "kmcp database with %lld targets in %d block(s) saved to: %s (%.1f s)", (long long)info.n_targets, info.n_blocks, out_dir.c_str(),
         std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    kmcpg_close(ctx);
    return 0;
}

// kmcp-gpu parse [--ahead] [-1 a -2 b | files...]: the FASTA/Q reader alone (no GPU): one line per record,
// "id<TAB>length<TAB>crc32 of the sequence bytes"; with -1/-2 the mates alternate.  Used by the host-only tests.
int parse_main(int argc, char **argv) {
    std::vector<std::string> files;
    std::string r1, r2;
    bool ahead = false, count_only = false, batches = false, whole = false;
    size_t batch_reads = 1u << 18;
    for (int i = 2; i < argc; i++) {
        std::string a = argv[i];
        if (a == "--ahead") ahead = true;
        else if (a == "--batches") batches = true;           // through the search command's batch builder (parser threads per file)
        else if (a == "--batch-reads" && i + 1 < argc) batch_reads = (size_t)atol(argv[++i]);
        else if (a == "-g") whole = true;
        else if (a == "--count") count_only = true;          // the reader's rate alone: records and bases to stderr, no per-record output
        else if (a == "--inflate-threads" && i + 1 < argc) g_tune.inflate_threads = atoi(argv[++i]);
        else if (a == "--parse-threads" && i + 1 < argc) g_tune.parse_threads = atoi(argv[++i]);
        else if (a == "--parse-piece" && i + 1 < argc) g_tune.parse_piece = (size_t)atol(argv[++i]);
        else if (a == "--inflate-chunk" && i + 1 < argc) g_tune.inflate_chunk = (size_t)atol(argv[++i]);
        else if (a == "--inflate-cap" && i + 1 < argc) g_tune.inflate_cap = (size_t)atol(argv[++i]);
        else if (a == "-1" && i + 1 < argc) r1 = argv[++i];
        else if (a == "-2" && i + 1 < argc) r2 = argv[++i];
        else files.push_back(a);
    }
    std::string id, out;
    std::vector<uint8_t> seq;
    char line[64];
    auto emit = [&]() {
        int n = snprintf(line, sizeof(line), "\t%zu\t%08lx\n", seq.size(), (unsigned long)crc32(0L, seq.data(), (uInt)seq.size()));
        out += id; out.append(line, (size_t)n);
        seq.clear();
        if (out.size() > (1u << 20)) { fwrite(out.data(), 1, out.size(), stdout); out.clear(); }
    };
    if (batches) {
        // one line per query of every batch: id, then length and CRC-32 of each of its sequences; "# batch" lines in between
        ReaderSetup rc;
        rc.paired = !r1.empty() && !r2.empty(); rc.read1 = r1; rc.read2 = r2; rc.files = files; rc.whole_file = whole; rc.batch_reads = batch_reads;
        const auto t0 = std::chrono::steady_clock::now();
        uint64_t nq = 0;
        read_batches(rc, [&](Batch *bt) {
            const size_t step = rc.paired ? 2 : 1;
            if (!count_only) {
                out += "# batch of " + std::to_string(bt->n_ids()) + "\n";
                for (size_t q = 0; q < bt->n_ids(); q++) {
                    out.append(bt->id(q), bt->id_len(q));
                    for (size_t m = 0; m < step; m++) {
                        const uint64_t a = bt->b.off[q * step + m], b = bt->b.off[q * step + m + 1];
                        int n = snprintf(line, sizeof(line), "\t%llu\t%08lx", (unsigned long long)(b - a), (unsigned long)crc32(0L, bt->b.seq + a, (uInt)(b - a)));
                        out.append(line, (size_t)n);
                    }
                    out += '\n';
                }
                fwrite(out.data(), 1, out.size(), stdout); out.clear();
            }
            nq += bt->n_ids();
            delete bt;
        });
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        uint64_t st[2] = {0, 0};
        kmcpg_internal_reader_stats(&st[0], &st[1]);
        if (count_only)
            fprintf(stderr, "%llu queries, %.3f s, %.2f M queries/s (pieces parsed in parallel: %llu, files handed back to the general reader: %llu)\n",
                    (unsigned long long)nq, dt, nq / dt / 1e6, (unsigned long long)st[0], (unsigned long long)st[1]);
        return 0;
    }
    if (count_only) {
        const auto t0 = std::chrono::steady_clock::now();
        uint64_t n = 0, bases = 0;
        std::vector<char> ids;                               // what the search pipeline keeps of a record: ID and sequence, back to back
        std::vector<uint64_t> id_off, off;
        for (auto &f : files) {
            Reader r;
            if (!r.open(f, ahead, g_tune)) die("%s: no such file", f.c_str());
            while (r.next(id, seq)) {
                n++; ids.insert(ids.end(), id.begin(), id.end()); id_off.push_back(ids.size()); off.push_back(seq.size());
                if (off.size() >= (1u << 18)) { bases += seq.size(); seq.clear(); ids.clear(); id_off.clear(); off.clear(); }
            }
            r.close();
        }
        bases += seq.size();
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        fprintf(stderr, "%llu records, %llu bases, %.3f s, %.2f M records/s\n", (unsigned long long)n, (unsigned long long)bases, dt, n / dt / 1e6);
        return 0;
    }
    if (!r1.empty() && !r2.empty()) {
        Reader a, b;
        if (!a.open(r1, ahead, g_tune) || !b.open(r2, ahead, g_tune)) die("no such file");
        for (;;) {
            if (!a.next(id, seq)) break;
            emit();
            if (!b.next(id, seq)) break;
            emit();
        }
        a.close(); b.close();
    } else {
        for (auto &f : files) {
            Reader r;
            if (!r.open(f, ahead, g_tune)) die("%s: no such file", f.c_str());
            while (r.next(id, seq)) emit();
            r.close();
        }
    }
    fwrite(out.data(), 1, out.size(), stdout);
    return 0;
}

// kmcp-gpu gunzip [--read-size N] [--chunk N] file|- : the input decoder alone (fastgz.h), decoded bytes to stdout.
// Exit code 1 and a message on a malformed stream.  Used by the host-only tests.
int gunzip_main(int argc, char **argv) {
    size_t read_size = 1u << 20, chunk = 4u << 20, par_chunk = 2u << 20;
    int threads = 0;
    bool stats = false;
    std::string file;
    for (int i = 2; i < argc; i++) {
        std::string a = argv[i];
        if (a == "--read-size" && i + 1 < argc) read_size = (size_t)atol(argv[++i]);
        else if (a == "--chunk" && i + 1 < argc) chunk = (size_t)atol(argv[++i]);
        else if (a == "--threads" && i + 1 < argc) threads = atoi(argv[++i]);          // > 0: the chunk-parallel decoder (pargz.h)
        else if (a == "--par-chunk" && i + 1 < argc) par_chunk = (size_t)atol(argv[++i]);
        else if (a == "--par-cap" && i + 1 < argc) g_tune.inflate_cap = (size_t)atol(argv[++i]);
        else if (a == "--stats") stats = true;
        else file = a;
    }
    if (file.empty() || !read_size || !chunk) { fputs("usage: kmcp-gpu gunzip [--read-size N] [--chunk N] [--threads T [--par-chunk B] [--stats]] file|-\n", stderr); return 2; }
    const int fd = file == "-" ? 0 : ::open(file.c_str(), O_RDONLY);
    if (fd < 0) die("%s: no such file", file.c_str());
    if (threads > 0) {
        if (!fastgz::ParallelInflater::usable(fd)) { fprintf(stderr, "kmcp-gpu gunzip: %s: not a seekable gzip file\n", file.c_str()); return 2; }
        fastgz::ParallelInflater inf(fd, threads, par_chunk, g_tune.inflate_cap);
        std::vector<char> buf(chunk);
        for (;;) {
            const ssize_t r = inf.read(buf.data(), buf.size());
            if (r < 0) { fprintf(stderr, "kmcp-gpu gunzip: %s: %s\n", file.c_str(), inf.error()); return inf.too_big() ? 4 : 1; }
            if (r == 0) break;
            if (fwrite(buf.data(), 1, (size_t)r, stdout) != (size_t)r) return 3;
        }
        if (stats) fprintf(stderr, "chunks used %llu, stretches decoded again in order %llu, symbols resolved %llu%s\n", (unsigned long long)inf.chunks_used(),
                           (unsigned long long)inf.chunks_redone(), (unsigned long long)inf.symbols_resolved(), inf.gave_up() ? ", gave up looking for block starts" : "");
        return 0;
    }
    fastgz::Inflater inf([fd, read_size](void *dst, size_t n) -> ssize_t { return ::read(fd, dst, std::min(n, read_size)); });
    std::vector<char> buf(chunk);
    for (;;) {
        const ssize_t r = inf.read(buf.data(), buf.size());
        if (r < 0) { fprintf(stderr, "kmcp-gpu gunzip: %s: %s\n", file.c_str(), inf.error()); return 1; }
        if (r == 0) break;
        if (fwrite(buf.data(), 1, (size_t)r, stdout) != (size_t)r) return 3;
    }
    return 0;
}

// kmcp-gpu gzip-write <text file> <out[.gz]> [piece bytes]: the result writer alone (no GPU), fed in pieces the size of a
// batch's TSV text; prints its rate.  Used by the host-only tests and to size the writer.
int gzip_write_main(int argc, char **argv) {
    if (argc < 4) { fputs("usage: kmcp-gpu gzip-write <text file> <out[.gz]> [piece bytes]\n", stderr); return 2; }
    FILE *f = fopen(argv[2], "rb");
    if (!f) die("%s: no such file", argv[2]);
    std::string text;
    std::vector<char> buf(1u << 24);
    for (size_t n; (n = fread(buf.data(), 1, buf.size(), f)) > 0;) text.append(buf.data(), n);
    fclose(f);
    const size_t piece = argc > 4 ? (size_t)atol(argv[4]) : (size_t)19 << 20;
    Writer w;
    w.open(argv[3]);
    const auto t0 = std::chrono::steady_clock::now();
    for (size_t o = 0; o < text.size(); o += piece) w.write(text.data() + o, std::min(piece, text.size() - o));
    w.close();
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    fprintf(stderr, "%zu bytes in %.3f s: %.1f MB/s\n", text.size(), dt, text.size() / dt / 1e6);
    return 0;
}

int run(int argc, char **argv) {
    if (argc > 1 && !strcmp(argv[1], "index")) return index_main(argc, argv);
    if (argc > 1 && !strcmp(argv[1], "gzip-write")) return gzip_write_main(argc, argv);
    if (argc > 1 && !strcmp(argv[1], "gunzip")) return gunzip_main(argc, argv);
    if (argc > 1 && !strcmp(argv[1], "fmt-selftest")) {        // kmcp-gpu fmt-selftest [n [seed]]: tsv_format.h against printf (host-only test)
        const uint64_t n = argc > 2 ? strtoull(argv[2], nullptr, 10) : 2000000, seed = argc > 3 ? strtoull(argv[3], nullptr, 10) : 1;
        const uint64_t bad = tsvfmt::selftest(n, seed);
        printf("%llu mismatches in %llu values\n", (unsigned long long)bad, (unsigned long long)n);
        return bad ? 1 : 0;
    }
    if (argc > 1 && !strcmp(argv[1], "parse")) return parse_main(argc, argv);
    Opts o;
    int ai = 1;
    if (ai < argc && !strcmp(argv[ai], "search")) ai++;
    else if (ai < argc && (!strcmp(argv[ai], "-h") || !strcmp(argv[ai], "--help"))) { usage(); return 0; }
    auto need = [&](int &i) -> const char * { if (i + 1 >= argc) die("flag needs an argument: %s", argv[i]); return argv[++i]; };
    for (int i = ai; i < argc; i++) {
        std::string a = argv[i], val;
        bool has_eq = false;
        if (a.rfind("--", 0) == 0) { size_t eq = a.find('='); if (eq != std::string::npos) { val = a.substr(eq + 1); a = a.substr(0, eq); has_eq = true; } }
        auto sval = [&]() -> std::string { return has_eq ? val : std::string(need(i)); };
        if (a == "-h" || a == "--help") { usage(); return 0; }
        else if (a == "-d" || a == "--db-dir") o.db_dirs.push_back(sval());
        else if (a == "--ref-counts") o.ref_counts_file = sval();
        else if (a == "--ref-counts-min-qcov") o.rc_min_qcov = atof(sval().c_str());
        else if (a == "--ref-counts-max-fpr") o.rc_max_fpr = atof(sval().c_str());
        else if (a == "-o" || a == "--out-file") o.out_file = sval();
        else if (a == "-1" || a == "--read1") o.read1 = sval();
        else if (a == "-2" || a == "--read2") o.read2 = sval();
        else if (a == "--try-se") o.try_se = true;
        else if (a == "-u" || a == "--kmer-dedup-threshold") o.dedup = atoi(sval().c_str());
        else if (a == "-g" || a == "--query-whole-file") o.whole_file = true;
        else if (a == "-G" || a == "--use-filename") o.use_filename = true;
        else if (a == "--query-id") o.query_id = sval();
        else if (a == "-c" || a == "--min-kmers") o.min_kmers = atoi(sval().c_str());
        else if (a == "-m" || a == "--min-query-len") o.min_qlen = atoi(sval().c_str());
        else if (a == "-t" || a == "--min-query-cov") o.qcov = atof(sval().c_str());
        else if (a == "-T" || a == "--min-target-cov") o.tcov = atof(sval().c_str());
        else if (a == "-f" || a == "--max-fpr") o.max_fpr = atof(sval().c_str());
        else if (a == "-N" || a == "--name-map") o.name_maps.push_back(sval());
        else if (a == "-D" || a == "--default-name-map") o.default_name_map = true;
        else if (a == "-K" || a == "--keep-unmatched") o.keep_unmatched = true;
        else if (a == "-n" || a == "--keep-top-scores") o.top_scores = atoi(sval().c_str());
        else if (a == "-H" || a == "--no-header-row") o.no_header = true;
        else if (a == "-s" || a == "--sort-by") o.sort_by = sval();
        else if (a == "-S" || a == "--do-not-sort") o.do_not_sort = true;
        else if (a == "-w" || a == "--load-whole-db" || a == "--low-mem") {}
        else if (a == "-j" || a == "--threads") o.threads = atoi(sval().c_str());
        else if (a == "-q" || a == "--quiet") g_quiet = true;
        else if (a == "--log") o.log_file = sval();
        else if (a == "--gpu") o.devices.assign(1, atoi(sval().c_str()));
        else if (a == "--gpu-mode") o.gpu_mode = sval();
        else if (a == "--dry-run") o.dry_run = true;
        else if (a == "--batch-reads") o.batch_reads = (size_t)std::max(1L, atol(sval().c_str()));     // queries per engine call (default 262144)
        else if (a == "--compression-level") { g_compression_level = atoi(sval().c_str()); if (g_compression_level < 1 || g_compression_level > 9) die("--compression-level should be in range [1, 9]"); }
        else if (a == "--inflate-threads") g_tune.inflate_threads = atoi(sval().c_str());
        else if (a == "--parse-threads") g_tune.parse_threads = atoi(sval().c_str());
        else if (a == "--parse-piece") g_tune.parse_piece = (size_t)atol(sval().c_str());
        else if (a == "--inflate-chunk") g_tune.inflate_chunk = (size_t)atol(sval().c_str());
        else if (a == "--inflate-cap") g_tune.inflate_cap = (size_t)atol(sval().c_str());
        else if (a == "--gpus") {
            const std::string v = sval();
            o.devices.clear();
            if (v == "all") o.all_devices = true;
            else
                for (size_t b = 0; b <= v.size();) {
                    size_t e = v.find(',', b);
                    if (e == std::string::npos) e = v.size();
                    if (e > b) o.devices.push_back(atoi(v.substr(b, e - b).c_str()));
                    b = e + 1;
                }
            if (!o.all_devices && o.devices.empty()) die("invalid value for flag --gpus: %s", v.c_str());
        }
        else if (a == "-i" || a == "--infile-list") { Reader r; std::string f = sval(), l; if (!r.open(f)) die("fail to read %s", f.c_str()); while (r.getline(l)) if (!l.empty()) o.files.push_back(l); r.close(); }
        else if (a.size() > 1 && a[0] == '-' && a != "-") die("unknown flag: %s", a.c_str());
        else o.files.push_back(a);
    }
    if (!o.log_file.empty()) g_log = fopen(o.log_file.c_str(), "w");
    auto t_start = std::chrono::steady_clock::now();

    // ---- flag checks (S:157-205) ----
    if (o.db_dirs.empty() && !o.dry_run) die("flag -d/--db-dir needed");
    if (o.dry_run) { o.db_dirs.clear(); o.ref_counts_file.clear(); }
    if (o.min_qlen < 0) die("value of flag --min-query-len should be greater than or equal to 0");
    if (o.min_kmers <= 0) die("value of flag --min-kmers should be greater than 0");
    if (!(o.max_fpr > 0)) die("value of flag --max-fpr should be greater than 0");
    if (o.dedup <= 0) die("value of flag --kmer-dedup-threshold should be greater than 0");
    if (o.top_scores < 0) die("value of flag --keep-top-scores should be greater than or equal to 0");
    int sort_by = o.sort_by == "qcov" ? 0 : o.sort_by == "tcov" ? 1 : o.sort_by == "jacc" ? 2 : -1;
    if (sort_by < 0) die("invalid value for flag -s/--sort-by: %s. Available: qcov/tsov/jacc", o.sort_by.c_str());
    if (o.qcov < 0 || o.qcov > 1) die("value of -t/--min-query-cov should be in range [0, 1]");
    if (o.tcov < 0 || o.tcov > 1) die("value of -T/-target-cov should be in range [0, 1]");
    if (o.do_not_sort && o.top_scores > 0) logf("WARN", "flag -n/--keep-top-scores ignored when -S/--do-not-sort given");
    if (o.gpu_mode != "auto" && o.gpu_mode != "shard" && o.gpu_mode != "replicate") die("invalid value for flag --gpu-mode: %s. Available: auto/shard/replicate", o.gpu_mode.c_str());

    logf("INFO", "kmcp-gpu (B200 search path of kmcp v0.9.5)");
    logf("INFO", "checking input files ...");
    bool paired = false;
    std::vector<std::string> files;
    if (o.read1.empty()) { if (!o.read2.empty()) files.push_back(o.read2); }
    else if (o.read2.empty()) files.push_back(o.read1);
    else { paired = true; logf("INFO", "paired end files given: %s, %s", o.read1.c_str(), o.read2.c_str()); }
    if (o.try_se && !paired) { logf("WARN", "flag --try-se ignored for single-end input(s)"); o.try_se = false; }
    if (!paired) {
        for (auto &f : o.files) files.push_back(f);
        if (files.empty()) files.push_back("-");
        logf("INFO", "  %zu input file(s) given", files.size());
        for (auto &f : files) if (f != "-" && f == o.out_file) die("out file should not be one of the input file");
    }

    // ---- three-stage pipeline: reader thread (inflate + parse + pack) → this thread (GPU engine, every database) →
    //      writer thread (merge across databases, TSV formatting on several threads, parallel gzip), all in input order ----
    struct Job { Batch *batch = nullptr; std::vector<kmcpg_results> res; };
    std::mutex mu;
    std::condition_variable cv;
    std::deque<Batch *> in_q;
    std::deque<Job *> out_q;
    bool in_done = false, out_done = false;
    uint64_t total = 0, matched = 0;
    std::thread reader;
    auto start_reader = [&](int kmax) {
        ReaderSetup rc;
        rc.paired = paired; rc.read1 = o.read1; rc.read2 = o.read2; rc.files = files;
        rc.whole_file = o.whole_file; rc.use_filename = o.use_filename; rc.query_id = o.query_id;
        rc.batch_reads = o.batch_reads; rc.batch_bytes = o.batch_bytes; rc.kmax = kmax;
        reader = std::thread([&, rc] {
            read_batches(rc, [&](Batch *bt) {
                kmcpg::NvtxRange nvtx("kmcp-gpu:reader hands over a batch");
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return in_q.size() < 2; });
                in_q.push_back(bt);
                cv.notify_all();
            });
            std::lock_guard<std::mutex> lk(mu);
            in_done = true;
            cv.notify_all();
        });
    };
    // the first batches are read while the index travels to the GPU — unless the reader needs something of the database first
    // (-g joins records with k-1 'N'; with several devices the batch size depends on how the index is laid out)
    const bool early_reader = !o.whole_file && o.devices.size() == 1 && !o.all_devices;
    if (early_reader) start_reader(21);

    // ---- databases (S:299-324): every child directory of a -d holding __db.yml; several -d = several databases ----
    struct Db {
        std::string dir;
        std::vector<kmcpg_ctx *> ctxs;      // one per device: shard i of ctxs.size() (a single context holds the whole database)
        bool replicas = false;              // every context holds the whole database and the reads are split instead
        kmcpg_db_info_t info;
        std::vector<kmcpg_target_t> targets;
        std::vector<const std::string *> mapped;
        std::map<std::string, std::string> name_map;
    };
    std::vector<Db> dbs;
    for (auto &dd : o.db_dirs) {
        logf("INFO", "checking the database: %s", dd.c_str());
        std::vector<std::string> subs;
        DIR *d = opendir(dd.c_str());
        if (!d) die("read database error: open %s: no such file or directory", dd.c_str());
        while (dirent *e = readdir(d)) {
            if (e->d_name[0] == '.') continue;
            std::string p = dd + "/" + e->d_name;
            if (is_dir(p) && is_file(p + "/__db.yml")) subs.push_back(p);
        }
        closedir(d);
        std::sort(subs.begin(), subs.end());
        if (subs.empty()) die("invalid kmcp database: %s", dd.c_str());
        if (subs.size() > 1) die("databases with several repeats (R001, R002, ...) are not supported by kmcp-gpu yet: %s", dd.c_str());
        dbs.emplace_back();
        dbs.back().dir = subs[0];
    }
    if (o.dry_run) {                      // a database of nothing: k = 21, no targets, no device
        dbs.emplace_back();
        memset(&dbs[0].info, 0, sizeof(dbs[0].info));
        dbs[0].info.n_ks = 1; dbs[0].info.ks[0] = 21;
        logf("WARN", "--dry-run: no device and no database are used, every query is reported unmatched");
    }
    if (o.all_devices && !o.dry_run) {    // every visible device: probe the ordinals until one is refused
        for (int d = 0; d < 64; d++) {
            kmcpg_ctx *probe = nullptr;
            if (kmcpg_create(d, &probe)) break;
            kmcpg_close(probe);
            o.devices.push_back(d);
        }
        if (o.devices.empty()) die("%s", kmcpg_last_error(nullptr));
    }
    for (auto &db : dbs) {
        if (o.dry_run) break;
        const int world = (int)o.devices.size();
        auto t_db = std::chrono::steady_clock::now();
        std::vector<kmcpg_ctx *> shard((size_t)world, nullptr);
        std::vector<std::string> errs((size_t)world);
        size_t min_free = ~(size_t)0;
        for (int r = 0; r < world; r++) {
            if (kmcpg_create(o.devices[(size_t)r], &shard[(size_t)r])) die("%s", kmcpg_last_error(nullptr));
            size_t fr = 0, tot = 0;
            if (world > 1 && kmcpg_device_memory(shard[(size_t)r], &fr, &tot) == KMCPG_OK) min_free = std::min(min_free, fr);
        }
        if (world > 1) {
            // replicate when the whole index (plus 8 GB of batch work space) fits into every device; otherwise split it
            uint64_t db_bytes = 0;
            std::vector<kmcpg_shard_piece> pcs(1 << 16);
            const int np = kmcpg_shard_pieces(db.dir.c_str(), 1, pcs.data(), (int32_t)pcs.size());
            if (np < 0) die("open kmcp db: %s: %s", db.dir.c_str(), kmcpg_last_error(nullptr));
            for (int i = 0; i < np; i++) db_bytes += pcs[(size_t)i].resident_bytes;
            const bool fits = min_free != ~(size_t)0 && db_bytes + (8ull << 30) <= (uint64_t)min_free;
            db.replicas = o.gpu_mode == "replicate" || (o.gpu_mode == "auto" && fits);
            if (o.gpu_mode == "replicate" && !fits) logf("WARN", "--gpu-mode replicate: the index (%.1f GB) may not fit into every device", db_bytes / 1e9);
        }
        logf("INFO", "loading database into HBM: %s%s", db.dir.c_str(), world > 1 ? (db.replicas ? " (a replica per device, reads are split)" : " (sharded over the devices)") : "");
        auto load = [&](int r) {          // devices load side by side: each reads only the blocks (or column ranges) it keeps
            kmcpg_db_opts dopt;
            memset(&dopt, 0, sizeof(dopt));
            dopt.shard_rank = r; dopt.shard_world = world;
            if (kmcpg_open_db(shard[(size_t)r], db.dir.c_str(), world > 1 && !db.replicas ? &dopt : nullptr)) errs[(size_t)r] = kmcpg_last_error(shard[(size_t)r]);
        };
        {
            std::vector<std::thread> lt;
            for (int r = 1; r < world; r++) lt.emplace_back(load, r);
            load(0);
            for (auto &t : lt) t.join();
        }
        for (int r = 0; r < world; r++) if (!errs[(size_t)r].empty()) die("open kmcp db: %s: %s", db.dir.c_str(), errs[(size_t)r].c_str());
        double gb = 0;
        for (int r = 0; r < world; r++) {
            kmcpg_db_info_t si;
            kmcpg_db_info(shard[(size_t)r], &si);
            if (r == 0) db.info = si;
            if (world > 1 && si.n_resident_blocks == 0) {
                // fewer 128-target column units than devices: this device would hash every read and probe nothing
                logf("WARN", "device %d gets no part of %s and stays idle", o.devices[(size_t)r], db.dir.c_str());
                kmcpg_close(shard[(size_t)r]);
                continue;
            }
            gb += si.resident_bytes / 1e9;
            if (world > 1) logf("INFO", "  device %d: %d block piece(s), %.2f GB", o.devices[(size_t)r], si.n_resident_blocks, si.resident_bytes / 1e9);
            db.ctxs.push_back(shard[(size_t)r]);
        }
        if (db.ctxs.empty()) die("invalid kmcp database (no index blocks): %s", db.dir.c_str());
        logf("INFO", "database loaded: %s (%d blocks, %lld targets, %.2f GB in HBM on %zu device(s), %.1f s)", db.dir.c_str(), db.info.n_blocks,
             (long long)db.info.n_targets, gb, db.ctxs.size(), std::chrono::duration<double>(std::chrono::steady_clock::now() - t_db).count());
        if (o.qcov <= db.info.fpr)      // S:405-409
            logf("WARN", "the value of -t/--min-query-cov (%f) is <= FPR (%f) of the database, you may get many false positives", o.qcov, db.info.fpr);
        if (db.info.ks[0] != dbs[0].info.ks[0]) die("databases with different k cannot be searched together");
        if (o.default_name_map && is_file(db.dir + "/__name_mapping.tsv")) load_kv(db.dir + "/__name_mapping.tsv", db.name_map);
        for (auto &f : o.name_maps) load_kv(f, db.name_map);       // user maps override the default one (U:322-330)
        db.targets.resize((size_t)db.info.n_targets);
        db.mapped.assign((size_t)db.info.n_targets, nullptr);
        for (int64_t t = 0; t < db.info.n_targets; t++) {
            kmcpg_target(db.ctxs[0], t, &db.targets[(size_t)t]);
            auto it = db.name_map.find(db.targets[(size_t)t].name);
            if (it != db.name_map.end()) db.mapped[(size_t)t] = &it->second;
        }
    }
    kmcpg_refcounts *refcounts = nullptr;
    if (!o.ref_counts_file.empty()) {
        if (dbs.size() != 1) die("--ref-counts needs exactly one -d database");
        if (sort_by != 0 || o.do_not_sort) die("--ref-counts needs matches sorted by qcov (the default), as `kmcp profile` does");
        kmcpg_refcount_params rp;
        kmcpg_default_refcount_params(&rp);
        rp.min_query_cov = o.rc_min_qcov; rp.max_fpr = o.rc_max_fpr;
        if (kmcpg_refcounts_create(dbs[0].ctxs[0], nullptr, &rp, &refcounts)) die("%s", kmcpg_last_error(dbs[0].ctxs[0]));
    }
    logf("INFO", "-------------------- [main parameters] --------------------");
    logf("INFO", "  minimum    query length: %d", o.min_qlen);
    logf("INFO", "  minimum  matched k-mers: %d", o.min_kmers);
    logf("INFO", "  minimum  query coverage: %f", o.qcov);
    logf("INFO", "  minimum target coverage: %f", o.tcov);
    logf("INFO", "-------------------- [main parameters] --------------------");

    Writer w;
    w.open(o.out_file);
    if (!o.no_header) {
        const char *h = "#query\tqLen\tqKmers\tFPR\thits\ttarget\tchunkIdx\tchunks\ttLen\tkSize\tmKmers\tqCov\ttCov\tjacc\tqueryIdx\n";   // S:437
        w.write(h, strlen(h));
    }

    {   // replicas split every batch between them: keep each device's share at the usual batch size
        size_t nrep = 1;
        for (auto &db : dbs) if (db.replicas) nrep = std::max(nrep, db.ctxs.size());
        o.batch_reads = std::min<size_t>(o.batch_reads * nrep, (size_t)1 << 22);
        o.batch_bytes = std::min<size_t>(o.batch_bytes * nrep, (size_t)2 << 30);
    }
    kmcpg_engine_opts eo;
    kmcpg_default_engine_opts(&eo);
    eo.min_query_len = o.min_qlen; eo.min_matched = o.min_kmers; eo.dedup_threshold = o.dedup; eo.min_query_cov = o.qcov; eo.min_target_cov = o.tcov;
    eo.max_fpr = o.max_fpr; eo.sort_by = sort_by; eo.do_not_sort = o.do_not_sort; eo.top_n_scores = o.top_scores; eo.try_se = o.try_se;
    eo.paired = paired; eo.threads = o.threads;

    if (!early_reader) start_reader(dbs[0].info.ks[0]);

    const Less less{sort_by};
    std::thread writer([&] {
        const int FT = 16;                                            // formatting threads per batch
        for (;;) {
            Job *job = nullptr;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return !out_q.empty() || out_done; });
                if (out_q.empty()) return;
                job = out_q.front(); out_q.pop_front();
                cv.notify_all();
            }
            kmcpg::NvtxRange nvtx("kmcp-gpu:writer (TSV formatting, gzip)");
            const Batch &bt = *job->batch;
            const uint32_t nq = (uint32_t)bt.n_ids();
            std::vector<std::string> text(FT);
            std::vector<uint64_t> nmatched(FT, 0);
            auto fmt = [&](int t) {
                char line[512];
                static thread_local tsvfmt::E4Cache e4;                // the %.4e strings of the FPR values seen lately
                std::string &out = text[t];
                const uint32_t lo = (uint32_t)((uint64_t)nq * t / FT), hi = (uint32_t)((uint64_t)nq * (t + 1) / FT);
                out.reserve((size_t)(hi - lo) * 96);
                std::vector<std::pair<kmcpg_match, int>> merged;      // (match, database) of one query when several databases are searched
                for (uint32_t q = lo; q < hi; q++) {
                    const char *const idp = bt.id(q);
                    const size_t idn = bt.id_len(q);
                    const kmcpg_results &r0 = job->res[0];
                    uint64_t hits = 0;
                    for (auto &r : job->res) hits += r.match_off[q + 1] - r.match_off[q];
                    if (hits == 0) {
                        if (!o.keep_unmatched) continue;
                        char *p = line;                                   // S:460-511: "\t%d\t%d\t0\t0\t\t-1\t0\t0\t%d\t0\t0\t0\t0\t%d\n"
                        *p++ = '\t'; p = tsvfmt::put_int(p, r0.query_len[q]);
                        *p++ = '\t'; p = tsvfmt::put_int(p, r0.n_kmers[q]);
                        memcpy(p, "\t0\t0\t\t-1\t0\t0\t", 13); p += 13;
                        p = tsvfmt::put_int(p, r0.k_used[q]);
                        memcpy(p, "\t0\t0\t0\t0\t", 9); p += 9;
                        p = tsvfmt::put_uint(p, bt.b.first_query + q);
                        *p++ = '\n';
                        const int n = (int)(p - line);
                        out.append(idp, idn); out.append(line, (size_t)n);
                        continue;
                    }
                    nmatched[t]++;
                    auto put = [&](const kmcpg_results &r, const Db &db, const kmcpg_match &m) {
                        // S:517-575: "%s\t%d\t%d\t%.4e\t%d\t%s\t%d\t%d\t%d\t%d\t%d\t%.4f\t%.4f\t%.4f\t%d\n", by hand (tsv_format.h: byte for byte what printf prints)
                        const kmcpg_target_t &tg = db.targets[m.target];
                        const std::string *mp = db.mapped[m.target];
                        char *p = line;
                        *p++ = '\t'; p = tsvfmt::put_int(p, r.query_len[q]);
                        *p++ = '\t'; p = tsvfmt::put_int(p, r.n_kmers[q]);
                        *p++ = '\t'; p = e4.put(p, m.fpr);
                        *p++ = '\t'; p = tsvfmt::put_uint(p, hits);
                        *p++ = '\t';
                        out.append(idp, idn); out.append(line, (size_t)(p - line));
                        if (mp) out.append(*mp); else out.append(tg.name);
                        p = line;
                        *p++ = '\t'; p = tsvfmt::put_uint(p, tg.index & 0xFFFFu);                               // S:532-539
                        *p++ = '\t'; p = tsvfmt::put_uint(p, tg.index >> 16);
                        *p++ = '\t'; p = tsvfmt::put_uint(p, tg.genome_size);
                        *p++ = '\t'; p = tsvfmt::put_int(p, r.k_used[q]);
                        *p++ = '\t'; p = tsvfmt::put_uint(p, m.count);
                        *p++ = '\t'; p = tsvfmt::put_f4(p, m.qcov);
                        *p++ = '\t'; p = tsvfmt::put_f4(p, m.tcov);
                        *p++ = '\t'; p = tsvfmt::put_f4(p, m.jacc);
                        *p++ = '\t'; p = tsvfmt::put_uint(p, bt.b.first_query + q);
                        *p++ = '\n';
                        out.append(line, (size_t)(p - line));
                    };
                    if (job->res.size() == 1) {
                        for (uint64_t i = r0.match_off[q]; i < r0.match_off[q + 1]; i++) put(r0, dbs[0], r0.matches[i]);
                    } else {                                          // union of the databases' hits, re-sorted (merge.go:190-256)
                        merged.clear();
                        for (size_t d = 0; d < job->res.size(); d++)
                            for (uint64_t i = job->res[d].match_off[q]; i < job->res[d].match_off[q + 1]; i++) merged.push_back({job->res[d].matches[i], (int)d});
                        if (!o.do_not_sort)
                            std::stable_sort(merged.begin(), merged.end(), [&](const std::pair<kmcpg_match, int> &a, const std::pair<kmcpg_match, int> &b) {
                                const int c = less.score(a.first, b.first);           // ties: database order, then target index
                                if (c) return c < 0;
                                if (a.second != b.second) return a.second < b.second;
                                return a.first.target < b.first.target;
                            });
                        for (auto &mm : merged) put(job->res[mm.second], dbs[mm.second], mm.first);
                    }
                }
            };
            std::vector<std::future<void>> fs;
            for (int t = 1; t < FT; t++) fs.push_back(std::async(std::launch::async, fmt, t));
            fmt(0);
            for (auto &f : fs) f.get();
            for (auto &t : text) w.write(std::move(t));               // in range order: no copy into one string first
            for (auto v : nmatched) matched += v;
            total += nq;
            if (refcounts && kmcpg_refcounts_add(refcounts, &job->res[0])) die("--ref-counts: inconsistent chunk numbering in the database");
            for (auto &r : job->res) kmcpg_free_results(&r);
            delete job->batch;
            delete job;
            if (!g_quiet) fprintf(stderr, "processed queries: %llu\r", (unsigned long long)total);
        }
    });

    logf("INFO", "searching ...");
    auto t_search = std::chrono::steady_clock::now();
    for (;;) {
        Batch *bt = nullptr;
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return !in_q.empty() || in_done; });
            if (in_q.empty()) break;
            bt = in_q.front(); in_q.pop_front();
            cv.notify_all();
        }
        Job *job = new Job();
        job->batch = bt;
        job->res.resize(dbs.size());
        for (size_t d = 0; d < dbs.size(); d++) {
            const uint32_t ns = bt->b.n_seqs;
            const int nc = (int)dbs[d].ctxs.size();
            const int rc = o.dry_run         ? kmcpg_internal_unmatched_results(bt->b.off, ns, paired, dbs[d].info.ks[0], &job->res[d])
                           : nc == 1         ? kmcpg_engine_search(dbs[d].ctxs[0], &eo, bt->b.seq, bt->b.off, ns, &job->res[d])
                           : dbs[d].replicas ? kmcpg_engine_search_replicas(dbs[d].ctxs.data(), nc, &eo, bt->b.seq, bt->b.off, ns, &job->res[d])
                                             : kmcpg_engine_search_sharded(dbs[d].ctxs.data(), nc, &eo, bt->b.seq, bt->b.off, ns, &job->res[d]);
            if (rc) {
                std::string msg;
                for (auto *c : dbs[d].ctxs) { const char *m = kmcpg_last_error(c); if (m && *m) { msg = m; break; } }
                die("search failed (%d): %s", rc, msg.c_str());
            }
        }
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return out_q.size() < 2; });
        out_q.push_back(job);
        cv.notify_all();
    }
    reader.join();
    { std::lock_guard<std::mutex> lk(mu); out_done = true; cv.notify_all(); }
    writer.join();
    char line[512];

    double minutes = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_search).count() / 60.0;
    if (!g_quiet) fprintf(stderr, "\n");
    logf("INFO", "");
    logf("INFO", "processed queries: %llu, speed: %.3f million queries per minute", (unsigned long long)total, total / 1e6 / (minutes > 0 ? minutes : 1e-9));
    logf("INFO", "%.4f%% (%llu/%llu) queries matched", total ? (double)matched / (double)total * 100 : 0.0, (unsigned long long)matched, (unsigned long long)total);
    logf("INFO", "done searching");
    if (o.out_file != "-") logf("INFO", "search results saved to: %s", o.out_file.c_str());
    int n = snprintf(line, sizeof(line), "# input queries: %llu\n# matched queries: %llu\n# matched percentage: %.4f%%\n", (unsigned long long)total,    // S:1023-1025
                     (unsigned long long)matched, total ? (double)matched / (double)total * 100 : NAN);
    w.write(line, (size_t)n);
    w.close();
    if (refcounts) {
        kmcpg_refcount_table tb;
        kmcpg_refcounts_get(refcounts, &tb);
        Writer rw;
        rw.open(o.ref_counts_file);
        std::string t = "#ref\tchunkIdx\tchunks\tgenomeSize\tmatch\tuniqMatch\tuniqMatchHic\n";
        for (uint32_t i = 0; i < tb.n_refs; i++)
            for (uint32_t c = 0; c < tb.rows[i].n_chunks; c++) {
                int k = snprintf(line, sizeof(line), "\t%u\t%u\t%llu\t%.17g\t%.17g\t%.17g\n", c, tb.rows[i].n_chunks, (unsigned long long)tb.rows[i].genome_size,
                                 tb.rows[i].match[c], tb.rows[i].uniq_match[c], tb.rows[i].uniq_match_hic[c]);
                t += tb.rows[i].name; t.append(line, (size_t)k);
            }
        int k = snprintf(line, sizeof(line), "# reads: %llu\n# references: %u\n", (unsigned long long)tb.n_reads, tb.n_refs);
        t.append(line, (size_t)k);
        rw.write(t);
        rw.close();
        logf("INFO", "reference counters of %u references saved to: %s", tb.n_refs, o.ref_counts_file.c_str());
        kmcpg_refcounts_free(refcounts);
    }
    for (auto &db : dbs) for (auto *c : db.ctxs) kmcpg_close(c);
    logf("INFO", "");
    logf("INFO", "elapsed time: %.3fs", std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count());
    if (g_log) fclose(g_log);
    return 0;
}

int main(int argc, char **argv) {
    try {
        return run(argc, argv);
    } catch (const fastx::ReaderError &e) {       // a reader used directly (name maps, file lists, the parse subcommand)
        die("%s", e.what());
    }
}

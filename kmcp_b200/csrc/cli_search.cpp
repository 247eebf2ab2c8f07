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

#include "../../include/kmcp_gpu.h"
#include "fastgz.h"
#include "pargz.h"
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace {

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
    exit(255);
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
    bool try_se = false, whole_file = false, use_filename = false, default_name_map = false, keep_unmatched = false, no_header = false,
         do_not_sort = false;
    size_t batch_reads = 1u << 18, batch_bytes = 256u << 20;
};

void usage() {
    fputs(
        "Search sequences against a kmcp database on a B200 GPU (drop-in for `kmcp search`)\n\n"
        "Usage:\n  kmcp-gpu search [flags] [-w] -d <kmcp db> [-t <min-query-cov>] [read1.fq.gz] [read2.fq.gz] [unpaired.fq.gz] [-o read.tsv.gz]\n\n"
        "Flags (same names, shorthands and defaults as kmcp search):\n"
        "  -d, --db-dir string              database directory created by \"kmcp index\"\n"
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

// inflate on a thread of its own: 4 MB chunks travel through a short queue to the parsing thread, so the two mates of a
// paired-end run (and the next file of a list) are decompressed side by side with the parsing (the reference reads through
// pgzip/xopen readers that also decompress ahead of the parser).  The decoder is fastgz.h (about three times zlib's rate).
struct InflateAhead {
    static constexpr size_t CHUNK = 4u << 20, DEPTH = 4;
    struct Chunk { std::vector<char> data; int n = 0; };
    std::function<ssize_t(void *, size_t)> f;
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::deque<Chunk *> ready, spare;
    Chunk *cur = nullptr;
    size_t cur_pos = 0;
    bool stop = false, done = false;
    void start(std::function<ssize_t(void *, size_t)> source) {
        f = std::move(source);
        th = std::thread([this] {
            for (;;) {
                Chunk *c = nullptr;
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&] { return stop || ready.size() < DEPTH; });
                    if (stop) return;
                    if (!spare.empty()) { c = spare.front(); spare.pop_front(); }
                }
                if (!c) { c = new Chunk(); c->data.resize(CHUNK); }
                c->n = (int)f(c->data.data(), CHUNK);
                const bool last = c->n <= 0;           // 0: end of file, < 0: error (reported by the consumer)
                {
                    std::lock_guard<std::mutex> lk(mu);
                    ready.push_back(c);
                    if (last) done = true;
                }
                cv.notify_all();
                if (last) return;
            }
        });
    }
    // like gzread: bytes copied (> 0), 0 at end of file, < 0 on a read error
    int read(char *dst, size_t cap) {
        if (!cur || cur_pos == (size_t)cur->n) {
            std::unique_lock<std::mutex> lk(mu);
            if (cur) { spare.push_back(cur); cur = nullptr; cv.notify_all(); }
            cv.wait(lk, [&] { return !ready.empty(); });
            cur = ready.front(); ready.pop_front(); cur_pos = 0;
            cv.notify_all();
            if (cur->n <= 0) { const int r = cur->n; ready.push_front(cur); cur = nullptr; return r; }   // stays at the head: every later read sees it too
        }
        const size_t n = std::min(cap, (size_t)cur->n - cur_pos);
        memcpy(dst, cur->data.data() + cur_pos, n);
        cur_pos += n;
        return (int)n;
    }
    void finish() {
        { std::lock_guard<std::mutex> lk(mu); stop = true; }
        cv.notify_all();
        if (th.joinable()) th.join();
        for (Chunk *c : ready) delete c;
        for (Chunk *c : spare) delete c;
        delete cur;
        ready.clear(); spare.clear(); cur = nullptr;
    }
};

int g_inflate_threads = 0;              // --inflate-threads: 0 = decide per file, 1 = always the sequential decoder
size_t g_inflate_chunk = 2u << 20;      // --inflate-chunk: compressed bytes per task of the chunk-parallel decoder
size_t g_inflate_cap = (size_t)256 << 20;   // --inflate-cap: most bytes a chunk may decode to before the sequential decoder takes over

// offsets (base + i) of every '\n' in p[0, n): 64 bytes per step with AVX2 where the CPU has it
#if defined(__x86_64__)
__attribute__((target("avx2"))) void scan_newlines_avx2(const char *p, size_t n, uint32_t base, std::vector<uint32_t> &out) {
    const __m256i nl = _mm256_set1_epi8('\n');
    size_t i = 0;
    for (; i + 64 <= n; i += 64) {
        const uint32_t m0 = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_loadu_si256((const __m256i *)(p + i)), nl));
        const uint32_t m1 = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_loadu_si256((const __m256i *)(p + i + 32)), nl));
        uint64_t m = (uint64_t)m0 | ((uint64_t)m1 << 32);
        while (m) { out.push_back(base + (uint32_t)i + (uint32_t)__builtin_ctzll(m)); m &= m - 1; }
    }
    for (; i < n; i++) if (p[i] == '\n') out.push_back(base + (uint32_t)i);
}
#endif
void scan_newlines(const char *p, size_t n, uint32_t base, std::vector<uint32_t> &out) {
#if defined(__x86_64__)
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2) { scan_newlines_avx2(p, n, base, out); return; }
#endif
    for (const char *q = p, *e = p + n; q < e;) {
        const char *h = (const char *)memchr(q, '\n', (size_t)(e - q));
        if (!h) break;
        out.push_back(base + (uint32_t)(h - p));
        q = h + 1;
    }
}

struct Reader {          // FASTA/Q, plain or gzip (bio/seqio/fastx default reader: ID = header up to first blank)
    int fd = -1;
    fastgz::Inflater *f = nullptr;       // gzip members are inflated, anything else passes through (as gzread does)
    fastgz::ParallelInflater *pf = nullptr;   // big gzip files on machines with cores to spare: one stream decoded by several threads
    std::string path;
    std::vector<char> buf;   // block buffer: lines are found with memchr, no per-line allocation
    size_t pos = 0, end = 0;
    bool eof = false;
    InflateAhead *ahead = nullptr;
    // line ends of buf[0, end) (offsets of '\n'), kept for the four-line FASTQ fast path: found 64 bytes at a time when the
    // buffer is filled instead of one memchr call per (short) line
    std::vector<uint32_t> nl;
    size_t nl_i = 0;
    bool open(const std::string &p, bool inflate_ahead = false) {
        path = p;
        fd = p == "-" ? 0 : ::open(p.c_str(), O_RDONLY);
        if (fd >= 0) {
            const int h = fd, T = inflate_threads(h);
            if (T >= 2) pf = new fastgz::ParallelInflater(h, T, g_inflate_chunk, g_inflate_cap);
            else f = sequential(h);
        }
        buf.resize(16u << 20);
        pos = end = 0; eof = false; raw_total = 0;
        nl.clear(); nl_i = 0;
        if (fd >= 0 && inflate_ahead) { ahead = new InflateAhead(); ahead->start([this](void *dst, size_t n) { return raw_read(dst, n); }); }
        return fd >= 0;
    }
    void close() {
        if (ahead) { ahead->finish(); delete ahead; ahead = nullptr; }
        delete f; delete pf;
        f = nullptr; pf = nullptr;
        if (fd > 0) ::close(fd);
        fd = -1;
    }
    static fastgz::Inflater *sequential(int h) {
        return new fastgz::Inflater([h](void *dst, size_t n) -> ssize_t {
            for (;;) { const ssize_t r = ::read(h, dst, n); if (r >= 0 || errno != EINTR) return r; }
        });
    }
    uint64_t raw_total = 0;              // decoded bytes handed out so far
    ssize_t raw_read(void *dst, size_t n) {
        if (pf) {
            const ssize_t r = pf->read(dst, n);
            if (r >= 0 || !pf->too_big()) { if (r > 0) raw_total += (uint64_t)r; return r; }
            // a stretch that expands beyond what the chunk-parallel decoder keeps in memory (compression ratios in the hundreds):
            // the sequential decoder streams; it starts over and drops what was handed out already
            delete pf; pf = nullptr;
            if (lseek(fd, 0, SEEK_SET) != 0) return -1;
            f = sequential(fd);
            std::vector<char> scratch(4u << 20);
            for (uint64_t left = raw_total; left;) {
                const ssize_t k = f->read(scratch.data(), (size_t)std::min<uint64_t>(left, scratch.size()));
                if (k <= 0) return -1;
                left -= (uint64_t)k;
            }
        }
        const ssize_t r = f->read(dst, n);
        if (r > 0) raw_total += (uint64_t)r;
        return r;
    }
    // the decoded text itself (for callers that cut it into records themselves); like read(2)
    int read_text(char *dst, size_t n) { n = std::min<size_t>(n, 1u << 30); return ahead ? ahead->read(dst, n) : (int)raw_read(dst, n); }
    const char *error() const { return pf ? pf->error() : (f ? f->error() : ""); }
    // worker threads for one input: --inflate-threads N, or by itself on big machines for big seekable gzip files
    static int inflate_threads(int h) {
        if (g_inflate_threads == 1 || !fastgz::ParallelInflater::usable(h)) return 1;
        if (g_inflate_threads > 1) return g_inflate_threads;
        struct stat st;
        if (fstat(h, &st) != 0 || st.st_size < (32 << 20)) return 1;
        const int hw = (int)std::thread::hardware_concurrency();
        const int T = std::min(8, hw / 8);              // the chunk-parallel decoder does ~1.7x the work: it pays from 4 threads on
        return T >= 4 ? T : 1;
    }
    bool fill() {            // keeps [pos, end), reads more behind it; false at end of file
        if (eof) return false;
        if (pos) {
            memmove(buf.data(), buf.data() + pos, end - pos);
            size_t w = 0;                                  // line ends behind pos move with the bytes
            for (size_t i = nl_i; i < nl.size(); i++) if (nl[i] >= pos) nl[w++] = nl[i] - (uint32_t)pos;
            nl.resize(w); nl_i = 0;
            end -= pos; pos = 0;
        }
        if (end == buf.size()) buf.resize(buf.size() * 2);
        const size_t room = std::min<size_t>(buf.size() - end, 1u << 30);
        const int r = ahead ? ahead->read(buf.data() + end, room) : (int)raw_read(buf.data() + end, room);
        if (r < 0) die("read error in %s: %s", path.c_str(), pf ? pf->error() : f->error());
        if (r == 0) { eof = true; return false; }
        if (end + (size_t)r < ((size_t)1 << 32)) scan_newlines(buf.data() + end, (size_t)r, (uint32_t)end, nl);
        else { nl.clear(); nl_i = 0; fast_ok = false; }          // a single line of gigabytes: offsets no longer fit
        end += (size_t)r;
        return true;
    }
    bool fast_ok = true;
    // A FASTQ record written as exactly four lines (header, sequence, '+', quality of the same length) taken from the line-end
    // table: the ID is returned as a range of the buffer (valid until the next call), the sequence is appended to dst.
    // false = not such a record here (FASTA, wrapped FASTQ, blank lines, the unterminated tail of a file): the general
    // reader below takes it from the same position.
    template <class V>
    bool next_four_line(const char *&idp, size_t &idn, V &dst) {
        if (!fast_ok) return false;
        if (pos == end && !fill()) return false;
        if (buf[pos] != '@') return false;                     // FASTA (or a blank line): do not wait for four lines of a genome
        for (;;) {
            while (nl_i < nl.size() && nl[nl_i] < pos) nl_i++;
            if (nl.size() - nl_i >= 4) break;
            if (!fill()) return false;
        }
        const char *b = buf.data();
        const size_t h0 = pos, s0 = (size_t)nl[nl_i] + 1, p0 = (size_t)nl[nl_i + 1] + 1, q0 = (size_t)nl[nl_i + 2] + 1;
        size_t h1 = nl[nl_i], s1 = nl[nl_i + 1], q1 = nl[nl_i + 3];
        if (b[h0] != '@' || b[p0] != '+') return false;          // an empty line holds its own '\n' here, so both tests also refuse blank lines
        while (h1 > h0 && b[h1 - 1] == '\r') h1--;
        while (s1 > s0 && b[s1 - 1] == '\r') s1--;
        while (q1 > q0 && b[q1 - 1] == '\r') q1--;
        if (q1 - q0 != s1 - s0) return false;
        size_t e = h0 + 1;
        while (e < h1 && b[e] != ' ' && b[e] != '\t') e++;
        idp = b + h0 + 1; idn = e - h0 - 1;
        dst.insert(dst.end(), b + s0, b + s1);
        pos = (size_t)nl[nl_i + 3] + 1;
        nl_i += 4;
        return true;
    }
    // next line without its end-of-line bytes; the pointer is valid until the next call
    bool line(const char *&sp, size_t &n) {
        size_t scanned = pos;
        for (;;) {
            const char *eol = (const char *)memchr(buf.data() + scanned, '\n', end - scanned);
            if (eol) { sp = buf.data() + pos; n = (size_t)(eol - sp); pos = (size_t)(eol - buf.data()) + 1; break; }
            const size_t had = end - pos;
            if (!fill()) { if (pos == end) return false; sp = buf.data() + pos; n = end - pos; pos = end; break; }
            scanned = pos + had;
        }
        while (n && (sp[n - 1] == '\r' || sp[n - 1] == '\n')) n--;
        return true;
    }
    int peek() {             // first byte of the next line, -1 at end of file
        if (pos == end && !fill()) return -1;
        return (unsigned char)buf[pos];
    }
    bool getline(std::string &out) {
        const char *sp; size_t n;
        if (!line(sp, n)) return false;
        out.assign(sp, n);
        return true;
    }
    // one record: ID into `id`, sequence bytes APPENDED to `dst`; false at end of file
    template <class V>
    bool next(std::string &id, V &dst) {
        const char *l; size_t n;
        if (next_four_line(l, n, dst)) { id.assign(l, n); return true; }
        do { if (!line(l, n)) return false; } while (n == 0);
        if (l[0] != '>' && l[0] != '@') die("invalid FASTA/Q record in %s", path.c_str());
        const bool fq = l[0] == '@';
        size_t e = 1;
        while (e < n && l[e] != ' ' && l[e] != '\t') e++;
        id.assign(l + 1, e - 1);
        if (fq) {
            if (!line(l, n)) return true;                // first line after the header is sequence
            dst.insert(dst.end(), l, l + n);
            size_t slen = n;
            for (;;) {                                   // more sequence lines up to the '+' line (multi-line FASTQ)
                if (!line(l, n)) return true;
                if (n && l[0] == '+') break;
                dst.insert(dst.end(), l, l + n); slen += n;
            }
            size_t got = 0;
            while (got < slen && line(l, n)) got += n;
        } else {
            for (;;) {
                const int c = peek();
                if (c < 0 || c == '>') break;
                if (!line(l, n)) break;
                dst.insert(dst.end(), l, l + n);
            }
        }
        return true;
    }
};

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


std::string trim_ext(const std::string &file) {     // filepathTrimExtension: strip dir, .gz/.xz/.zst/.bz2, then one extension
    std::string b = file.substr(file.find_last_of('/') == std::string::npos ? 0 : file.find_last_of('/') + 1);
    for (const char *z : {".gz", ".xz", ".zst", ".bz2"}) {
        size_t n = strlen(z);
        if (b.size() > n && b.compare(b.size() - n, n, z) == 0) { b.resize(b.size() - n); break; }
    }
    size_t dot = b.find_last_of('.');
    if (dot != std::string::npos && dot > 0) b.resize(dot);
    return b;
}

// ---- reader stage: files → packed batches ---------------------------------------------------------------------------------
// what the engine call takes: the sequences of the batch back to back + offsets (two per query for paired-end input), the IDs in
// one arena (no allocation per read)
struct Batch {
    std::vector<char> id_buf;
    std::vector<uint64_t> id_off{0};
    std::vector<uint8_t> seq;
    std::vector<uint64_t> off{0};
    uint64_t base = 0;
    size_t n_ids() const { return id_off.size() - 1; }
    void add_id(const char *p, size_t n) { id_buf.insert(id_buf.end(), p, p + n); id_off.push_back(id_buf.size()); }
    void add_id(const std::string &id) { add_id(id.data(), id.size()); }
};

int g_parse_threads = 0;                // --parse-threads: 0 = decide per file, 1 = one parser thread per input
size_t g_parse_piece = 8u << 20;        // --parse-piece: bytes of text per task of the parallel parser
std::atomic<uint64_t> g_stat_pieces{0}, g_stat_fallbacks{0};      // pieces parsed by the workers / files handed back to the general reader

// One input file parsed into blocks of records (IDs and sequences back to back) ahead of the thread that builds the batches:
// with paired-end input the two mates are inflated AND parsed side by side, and the batch builder only copies.
//
// Two ways to get the blocks.  The plain one is a thread that calls the reader record by record.  For regular files on machines
// with cores to spare the text itself is cut into pieces of ~8 MB at record boundaries (a line that starts with '@' whose second
// next line starts with '+' — a quality line that starts with '@' is followed by a header and a sequence, never by a '+' line) and
// the pieces are parsed by several workers.  A worker accepts a piece only if EVERY record in it is written as exactly four lines
// ('@' header, sequence, '+' line, quality of the same length); such a piece starts and ends on record boundaries and parses to
// what the general reader returns for it.  The first piece that is anything else (FASTA, wrapped FASTQ, blank lines, a truncated
// last record) ends the parallel mode: the file is opened again, the records already handed out are skipped, and the plain
// thread carries on — the result is the general reader's in every case.
struct RecordStream {
    static constexpr size_t BLOCK_RECS = 1u << 15, BLOCK_BYTES = 64u << 20, DEPTH = 4;
    static constexpr size_t PIECE_MAX = 256u << 20;
    struct Block {
        std::vector<char> ids;
        std::vector<uint8_t> seq;
        std::vector<uint32_t> id_end, seq_end;       // ends of record i inside ids / seq
        bool bad = false;                            // parallel mode: the piece was not made of four-line records
        size_t n() const { return id_end.size(); }
        void clear() { ids.clear(); seq.clear(); id_end.clear(); seq_end.clear(); bad = false; }
    };
    struct Text { std::vector<char> d; size_t n = 0; };
    Reader r;
    std::string path;
    std::thread th;                                  // the plain parser, or the cutter of the parallel mode
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv;
    std::deque<Block *> ready, spare;
    bool stop = false, done = false;
    Block *cur = nullptr;
    size_t i = 0;
    uint64_t delivered = 0;                          // records handed out so far
    // parallel mode
    bool par = false;
    std::deque<std::pair<uint64_t, Text *>> work_q;  // pieces waiting for a worker
    std::deque<Text *> text_pool;
    std::map<uint64_t, Block *> parsed;              // finished pieces by number
    uint64_t cut_n = 0, next_n = 0;                  // pieces cut / pieces handed out
    size_t max_inflight = 8;

    static int parse_threads(const std::string &p) {
        if (g_parse_threads == 1 || p == "-") return 1;
        struct stat st;
        if (stat(p.c_str(), &st) != 0 || !S_ISREG(st.st_mode)) return 1;
        // opt-in until it has been timed on the GPU box: on the 8 vCPUs it was written on the extra copies eat the gain
        return g_parse_threads > 1 ? g_parse_threads : 1;
    }
    bool open(const std::string &p) {
        path = p;
        if (!r.open(p, true)) return false;
        const int P = parse_threads(p);
        if (P >= 2) start_parallel(P); else start_plain(0);
        return true;
    }
    void start_plain(uint64_t skip) {
        th = std::thread([this, skip] {
            std::string id;
            std::vector<uint8_t> scratch;
            for (uint64_t k = 0; k < skip; k++) { scratch.clear(); if (!r.next(id, scratch)) break; }     // handed out before the restart
            for (;;) {
                Block *b = nullptr;
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&] { return stop || ready.size() < DEPTH; });
                    if (stop) return;
                    if (!spare.empty()) { b = spare.front(); spare.pop_front(); }
                }
                if (!b) b = new Block();
                b->clear();
                bool more = true;
                while (b->n() < BLOCK_RECS && b->seq.size() < BLOCK_BYTES) {
                    if (!r.next(id, b->seq)) { more = false; break; }
                    b->ids.insert(b->ids.end(), id.begin(), id.end());
                    b->id_end.push_back((uint32_t)b->ids.size());
                    b->seq_end.push_back((uint32_t)b->seq.size());
                }
                std::lock_guard<std::mutex> lk(mu);
                if (b->n()) ready.push_back(b); else spare.push_back(b);
                if (!more) done = true;
                cv.notify_all();
                if (!more) return;
            }
        });
    }

    // ---- parallel mode ------------------------------------------------------------------------------------------------------
    // where to cut d[0, n): the start of the last line that begins a record and has its two following line starts inside the
    // text; 0 = no such line
    static size_t cut_point(const char *d, size_t n) {
        size_t e = n;                                 // lines are looked at from the end: [ls, e) is the current one
        size_t s1 = 0, s2 = 0;                        // starts of the next line and the one after it
        int have = 0;
        while (e > 0) {
            const char *q = e >= 2 ? (const char *)memrchr(d, '\n', e - 1) : nullptr;      // the line end in front of this line
            const size_t ls = q ? (size_t)(q - d) + 1 : 0;
            if (have >= 2 && d[ls] == '@' && d[s2] == '+' && ls > 0) return ls;
            s2 = s1; s1 = ls; have++;
            if (n - ls > (4u << 20) && have > 64) break;          // far from the end and still nothing: not this kind of file
            e = ls ? ls : 0;
            if (!ls) break;
        }
        return 0;
    }
    // a piece → records, or bad
    static void parse_piece(const Text &t, Block &b, std::vector<uint32_t> &nl) {
        b.clear();
        nl.clear();
        scan_newlines(t.d.data(), t.n, 0, nl);
        const char *d = t.d.data();
        if (nl.empty() || nl.size() % 4 != 0 || nl.back() + 1 != t.n) { b.bad = true; return; }
        b.id_end.reserve(nl.size() / 4); b.seq_end.reserve(nl.size() / 4);
        b.seq.reserve(t.n / 2); b.ids.reserve(t.n / 8);
        size_t h0 = 0;
        for (size_t k = 0; k < nl.size(); k += 4) {
            const size_t s0 = (size_t)nl[k] + 1, p0 = (size_t)nl[k + 1] + 1, q0 = (size_t)nl[k + 2] + 1;
            size_t h1 = nl[k], s1 = nl[k + 1], q1 = nl[k + 3];
            if (d[h0] != '@' || d[p0] != '+') { b.bad = true; return; }
            while (h1 > h0 && d[h1 - 1] == '\r') h1--;
            while (s1 > s0 && d[s1 - 1] == '\r') s1--;
            while (q1 > q0 && d[q1 - 1] == '\r') q1--;
            if (q1 - q0 != s1 - s0) { b.bad = true; return; }
            size_t e = h0 + 1;
            while (e < h1 && d[e] != ' ' && d[e] != '\t') e++;
            b.ids.insert(b.ids.end(), d + h0 + 1, d + e);
            b.seq.insert(b.seq.end(), (const uint8_t *)d + s0, (const uint8_t *)d + s1);
            b.id_end.push_back((uint32_t)b.ids.size());
            b.seq_end.push_back((uint32_t)b.seq.size());
            h0 = (size_t)nl[k + 3] + 1;
        }
    }
    void start_parallel(int P) {
        par = true;
        max_inflight = (size_t)P * 2 + 2;
        for (int w = 0; w < P; w++)
            workers.emplace_back([this] {
                std::vector<uint32_t> nl;
                for (;;) {
                    std::pair<uint64_t, Text *> job;
                    Block *b = nullptr;
                    {
                        std::unique_lock<std::mutex> lk(mu);
                        cv.wait(lk, [&] { return stop || !work_q.empty(); });
                        if (stop) return;
                        job = work_q.front(); work_q.pop_front();
                        if (!spare.empty()) { b = spare.front(); spare.pop_front(); }
                    }
                    if (!b) b = new Block();
                    if (job.second->n == 0) { b->clear(); b->bad = true; }        // the cutter found no record boundary
                    else parse_piece(*job.second, *b, nl);
                    std::lock_guard<std::mutex> lk(mu);
                    parsed[job.first] = b;
                    text_pool.push_back(job.second);
                    cv.notify_all();
                }
            });
        th = std::thread([this] {
            const size_t PIECE = std::max<size_t>(g_parse_piece, 64);
            std::vector<char> carry;
            bool eof = false;
            auto give = [&](Text *t) {
                std::lock_guard<std::mutex> lk(mu);
                work_q.push_back({cut_n++, t});
                cv.notify_all();
            };
            while (!eof) {
                Text *t = nullptr;
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&] { return stop || cut_n - next_n < max_inflight; });
                    if (stop) return;
                    if (!text_pool.empty()) { t = text_pool.front(); text_pool.pop_front(); }
                }
                if (!t) t = new Text();
                if (t->d.size() < PIECE + (1u << 20)) t->d.resize(PIECE + (1u << 20));
                t->n = carry.size();
                if (t->n > t->d.size()) t->d.resize(t->n + PIECE);
                if (t->n) memcpy(t->d.data(), carry.data(), t->n);
                carry.clear();
                size_t want = PIECE;
                size_t cut = 0;
                for (;;) {
                    while (t->n < want) {                             // fill up to the piece size
                        if (t->d.size() < want + 1) t->d.resize(want + 1);
                        const int got = r.read_text(t->d.data() + t->n, want - t->n);
                        if (got < 0) die("read error in %s: %s", path.c_str(), r.error());
                        if (got == 0) { eof = true; break; }
                        t->n += (size_t)got;
                    }
                    if (eof) {                                        // the rest of the file is the last piece
                        if (t->n && t->d[t->n - 1] != '\n') t->d[t->n++] = '\n';
                        cut = t->n;
                        break;
                    }
                    cut = cut_point(t->d.data(), t->n);
                    if (cut) break;
                    if (want >= PIECE_MAX) { cut = 0; break; }        // no record boundary in 256 MB of text
                    want *= 2;                                        // very long records: look at more text
                }
                if (!eof && !cut) { t->n = 0; give(t); break; }      // an empty piece tells the consumer to fall back
                if (eof && !t->n) {
                    std::lock_guard<std::mutex> lk(mu);
                    text_pool.push_back(t);
                    break;
                }
                carry.assign(t->d.data() + cut, t->d.data() + t->n);
                t->n = cut;
                give(t);
            }
            std::lock_guard<std::mutex> lk(mu);
            done = true;
            cv.notify_all();
        });
    }
    void stop_threads() {
        { std::lock_guard<std::mutex> lk(mu); stop = true; }
        cv.notify_all();
        if (th.joinable()) th.join();
        for (auto &w : workers) w.join();
        workers.clear();
        for (auto &j : work_q) delete j.second;
        for (Text *t : text_pool) delete t;
        for (auto &kv : parsed) delete kv.second;
        work_q.clear(); text_pool.clear(); parsed.clear();
    }

    // the next block of records, nullptr at the end of the file; the previous one goes back to the parser
    Block *next_block() {
        for (;;) {
            std::unique_lock<std::mutex> lk(mu);
            if (cur) { spare.push_back(cur); cur = nullptr; cv.notify_all(); }
            if (!par) {
                cv.wait(lk, [&] { return !ready.empty() || done; });
                if (ready.empty()) return nullptr;
                cur = ready.front(); ready.pop_front(); i = 0;
                delivered += cur->n();
                cv.notify_all();
                return cur;
            }
            cv.wait(lk, [&] { return parsed.count(next_n) || (done && next_n == cut_n); });
            auto it = parsed.find(next_n);
            if (it == parsed.end()) return nullptr;
            Block *b = it->second;
            parsed.erase(it);
            next_n++;
            cv.notify_all();
            if (!b->bad) {
                g_stat_pieces++;
                if (!b->n()) { spare.push_back(b); continue; }
                cur = b; i = 0;
                delivered += b->n();
                return cur;
            }
            // not (only) four-line records: the general reader takes over behind the records handed out so far
            delete b;
            g_stat_fallbacks++;
            lk.unlock();
            stop_threads();
            r.close();
            if (!r.open(path, true)) die("%s: no such file", path.c_str());
            { std::lock_guard<std::mutex> g(mu); stop = false; done = false; par = false; }
            start_plain(delivered);
        }
    }
    // record by record: false at the end of the file
    bool next(const char *&id, size_t &idn, const uint8_t *&sq, size_t &sn) {
        if (!cur || i == cur->n()) { if (!next_block()) return false; }
        const size_t a = i ? cur->id_end[i - 1] : 0, c = i ? cur->seq_end[i - 1] : 0;
        id = cur->ids.data() + a; idn = cur->id_end[i] - a;
        sq = cur->seq.data() + c; sn = cur->seq_end[i] - c;
        i++;
        return true;
    }
    void close() {
        stop_threads();
        for (Block *b : ready) delete b;
        for (Block *b : spare) delete b;
        delete cur;
        ready.clear(); spare.clear(); cur = nullptr;
        r.close();
    }
};

struct ReaderConfig {
    bool paired = false, whole_file = false, use_filename = false;
    std::string read1, read2, query_id;
    std::vector<std::string> files;
    size_t batch_reads = 1u << 18, batch_bytes = 256u << 20;
    int kmax = 21;
};

// S:793-1000: the input files as batches, in order; `emit` takes the batch over
void read_batches(const ReaderConfig &c, const std::function<void(Batch *)> &emit_fn) {
    Batch *cur = new Batch();
    auto emit = [&]() {
        if (cur->n_ids() == 0) return;
        if (cur->seq.empty()) cur->seq.push_back(0);
        emit_fn(cur);
        cur = new Batch();
    };
    auto full = [&]() { return cur->n_ids() >= c.batch_reads || cur->seq.size() >= c.batch_bytes; };
    if (c.paired) {
        RecordStream r1, r2;
        if (!r1.open(c.read1)) die("%s: no such file", c.read1.c_str());
        if (!r2.open(c.read2)) die("%s: no such file", c.read2.c_str());
        logf("INFO", "reading from paired-end files: %s, %s", c.read1.c_str(), c.read2.c_str());
        const char *id1, *id2; const uint8_t *s1, *s2; size_t n1, n2, l1, l2;
        for (;;) {                                                    // S:806-867: ID of read1; ends with the shorter file
            if (!r1.next(id1, n1, s1, l1)) break;
            if (!r2.next(id2, n2, s2, l2)) break;
            cur->add_id(id1, n1);
            cur->seq.insert(cur->seq.end(), s1, s1 + l1); cur->off.push_back(cur->seq.size());
            cur->seq.insert(cur->seq.end(), s2, s2 + l2); cur->off.push_back(cur->seq.size());
            if (full()) emit();
        }
        r1.close(); r2.close();
    } else {
        std::string id;
        for (auto &file : c.files) {
            logf("INFO", "reading sequence file: %s", file.c_str());
            if (c.whole_file) {                                       // S:885-937 (the N-run follows every record after the second)
                Reader r;
                if (!r.open(file, true)) die("%s: no such file", file.c_str());
                std::string qid;
                bool first = true;
                const size_t mark = cur->seq.size();
                while (r.next(id, cur->seq)) {
                    if (first) { qid = c.use_filename ? trim_ext(file) : (!c.query_id.empty() ? c.query_id : id); first = false; }
                    else cur->seq.insert(cur->seq.end(), (size_t)(c.kmax - 1), (uint8_t)'N');
                }
                r.close();
                if (first) { logf("WARN", "no valid sequences in file: %s", file.c_str()); cur->seq.resize(mark); continue; }
                cur->add_id(qid); cur->off.push_back(cur->seq.size());
                if (cur->seq.size() >= c.batch_bytes) emit();
            } else {
                RecordStream rs;
                if (!rs.open(file)) die("%s: no such file", file.c_str());
                bool any = false;
                while (RecordStream::Block *b = rs.next_block()) {   // whole blocks are appended: three copies and two offset loops
                    any = true;
                    size_t at = 0;                                    // records of the block already taken
                    while (at < b->n()) {
                        const size_t room = c.batch_reads > cur->n_ids() ? c.batch_reads - cur->n_ids() : 1;
                        const size_t take = std::min(room, b->n() - at);
                        const size_t i0 = at ? b->id_end[at - 1] : 0, i1 = b->id_end[at + take - 1];
                        const size_t s0 = at ? b->seq_end[at - 1] : 0, s1 = b->seq_end[at + take - 1];
                        const uint64_t ib = cur->id_buf.size() - i0, sb = cur->seq.size() - s0;
                        cur->id_buf.insert(cur->id_buf.end(), b->ids.begin() + (ptrdiff_t)i0, b->ids.begin() + (ptrdiff_t)i1);
                        cur->seq.insert(cur->seq.end(), b->seq.begin() + (ptrdiff_t)s0, b->seq.begin() + (ptrdiff_t)s1);
                        for (size_t k = at; k < at + take; k++) { cur->id_off.push_back(ib + b->id_end[k]); cur->off.push_back(sb + b->seq_end[k]); }
                        at += take;
                        if (full()) emit();
                    }
                }
                rs.close();
                if (!any) logf("WARN", "no valid sequences in file: %s", file.c_str());
            }
        }
    }
    emit();
    delete cur;
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
        else if (a == "--inflate-threads" && i + 1 < argc) g_inflate_threads = atoi(argv[++i]);
        else if (a == "--parse-threads" && i + 1 < argc) g_parse_threads = atoi(argv[++i]);
        else if (a == "--parse-piece" && i + 1 < argc) g_parse_piece = (size_t)atol(argv[++i]);
        else if (a == "--inflate-chunk" && i + 1 < argc) g_inflate_chunk = (size_t)atol(argv[++i]);
        else if (a == "--inflate-cap" && i + 1 < argc) g_inflate_cap = (size_t)atol(argv[++i]);
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
        ReaderConfig rc;
        rc.paired = !r1.empty() && !r2.empty(); rc.read1 = r1; rc.read2 = r2; rc.files = files; rc.whole_file = whole; rc.batch_reads = batch_reads;
        const auto t0 = std::chrono::steady_clock::now();
        uint64_t nq = 0;
        read_batches(rc, [&](Batch *bt) {
            const size_t step = rc.paired ? 2 : 1;
            if (!count_only) {
                out += "# batch of " + std::to_string(bt->n_ids()) + "\n";
                for (size_t q = 0; q < bt->n_ids(); q++) {
                    out.append(bt->id_buf.data() + bt->id_off[q], (size_t)(bt->id_off[q + 1] - bt->id_off[q]));
                    for (size_t m = 0; m < step; m++) {
                        const uint64_t a = bt->off[q * step + m], b = bt->off[q * step + m + 1];
                        int n = snprintf(line, sizeof(line), "\t%llu\t%08lx", (unsigned long long)(b - a), (unsigned long)crc32(0L, bt->seq.data() + a, (uInt)(b - a)));
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
        if (count_only)
            fprintf(stderr, "%llu queries, %.3f s, %.2f M queries/s (pieces parsed in parallel: %llu, files handed back to the general reader: %llu)\n",
                    (unsigned long long)nq, dt, nq / dt / 1e6, (unsigned long long)g_stat_pieces.load(), (unsigned long long)g_stat_fallbacks.load());
        return 0;
    }
    if (count_only) {
        const auto t0 = std::chrono::steady_clock::now();
        uint64_t n = 0, bases = 0;
        std::vector<char> ids;                               // what the search pipeline keeps of a record: ID and sequence, back to back
        std::vector<uint64_t> id_off, off;
        for (auto &f : files) {
            Reader r;
            if (!r.open(f, ahead)) die("%s: no such file", f.c_str());
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
        if (!a.open(r1, ahead) || !b.open(r2, ahead)) die("no such file");
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
            if (!r.open(f, ahead)) die("%s: no such file", f.c_str());
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
        else if (a == "--par-cap" && i + 1 < argc) g_inflate_cap = (size_t)atol(argv[++i]);
        else if (a == "--stats") stats = true;
        else file = a;
    }
    if (file.empty() || !read_size || !chunk) { fputs("usage: kmcp-gpu gunzip [--read-size N] [--chunk N] [--threads T [--par-chunk B] [--stats]] file|-\n", stderr); return 2; }
    const int fd = file == "-" ? 0 : ::open(file.c_str(), O_RDONLY);
    if (fd < 0) die("%s: no such file", file.c_str());
    if (threads > 0) {
        if (!fastgz::ParallelInflater::usable(fd)) { fprintf(stderr, "kmcp-gpu gunzip: %s: not a seekable gzip file\n", file.c_str()); return 2; }
        fastgz::ParallelInflater inf(fd, threads, par_chunk, g_inflate_cap);
        std::vector<char> buf(chunk);
        for (;;) {
            const ssize_t r = inf.read(buf.data(), buf.size());
            if (r < 0) { fprintf(stderr, "kmcp-gpu gunzip: %s: %s\n", file.c_str(), inf.error()); return inf.too_big() ? 4 : 1; }
            if (r == 0) break;
            if (fwrite(buf.data(), 1, (size_t)r, stdout) != (size_t)r) return 3;
        }
        if (stats) fprintf(stderr, "chunks used %llu, stretches decoded again in order %llu, symbols resolved %llu\n", (unsigned long long)inf.chunks_used(),
                           (unsigned long long)inf.chunks_redone(), (unsigned long long)inf.symbols_resolved());
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

int main(int argc, char **argv) {
    if (argc > 1 && !strcmp(argv[1], "index")) return index_main(argc, argv);
    if (argc > 1 && !strcmp(argv[1], "gzip-write")) return gzip_write_main(argc, argv);
    if (argc > 1 && !strcmp(argv[1], "gunzip")) return gunzip_main(argc, argv);
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
        else if (a == "--compression-level") { g_compression_level = atoi(sval().c_str()); if (g_compression_level < 1 || g_compression_level > 9) die("--compression-level should be in range [1, 9]"); }
        else if (a == "--inflate-threads") g_inflate_threads = atoi(sval().c_str());
        else if (a == "--parse-threads") g_parse_threads = atoi(sval().c_str());
        else if (a == "--parse-piece") g_parse_piece = (size_t)atol(sval().c_str());
        else if (a == "--inflate-chunk") g_inflate_chunk = (size_t)atol(sval().c_str());
        else if (a == "--inflate-cap") g_inflate_cap = (size_t)atol(sval().c_str());
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
    if (o.db_dirs.empty()) die("flag -d/--db-dir needed");
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
    if (o.all_devices) {                  // every visible device: probe the ordinals until one is refused
        for (int d = 0; d < 64; d++) {
            kmcpg_ctx *probe = nullptr;
            if (kmcpg_create(d, &probe)) break;
            kmcpg_close(probe);
            o.devices.push_back(d);
        }
        if (o.devices.empty()) die("%s", kmcpg_last_error(nullptr));
    }
    for (auto &db : dbs) {
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

    // ---- three-stage pipeline: reader thread (inflate + parse + pack) → this thread (GPU engine, every database) →
    //      writer thread (merge across databases, TSV formatting on several threads, parallel gzip), all in input order ----
    struct Job { Batch *batch = nullptr; std::vector<kmcpg_results> res; };
    std::mutex mu;
    std::condition_variable cv;
    std::deque<Batch *> in_q;
    std::deque<Job *> out_q;
    bool in_done = false, out_done = false;
    uint64_t total = 0, matched = 0;
    const int kmax = dbs[0].info.ks[0];

    std::thread reader([&] {
        ReaderConfig rc;
        rc.paired = paired; rc.read1 = o.read1; rc.read2 = o.read2; rc.files = files;
        rc.whole_file = o.whole_file; rc.use_filename = o.use_filename; rc.query_id = o.query_id;
        rc.batch_reads = o.batch_reads; rc.batch_bytes = o.batch_bytes; rc.kmax = kmax;
        uint64_t next_base = 0;
        read_batches(rc, [&](Batch *bt) {
            bt->base = next_base;
            next_base += bt->n_ids();
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return in_q.size() < 2; });
            in_q.push_back(bt);
            cv.notify_all();
        });
        std::lock_guard<std::mutex> lk(mu);
        in_done = true;
        cv.notify_all();
    });

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
            const Batch &bt = *job->batch;
            const uint32_t nq = (uint32_t)bt.n_ids();
            std::vector<std::string> text(FT);
            std::vector<uint64_t> nmatched(FT, 0);
            auto fmt = [&](int t) {
                char line[512];
                std::string &out = text[t];
                const uint32_t lo = (uint32_t)((uint64_t)nq * t / FT), hi = (uint32_t)((uint64_t)nq * (t + 1) / FT);
                out.reserve((size_t)(hi - lo) * 96);
                std::vector<std::pair<kmcpg_match, int>> merged;      // (match, database) of one query when several databases are searched
                for (uint32_t q = lo; q < hi; q++) {
                    const char *const idp = bt.id_buf.data() + bt.id_off[q];
                    const size_t idn = (size_t)(bt.id_off[q + 1] - bt.id_off[q]);
                    const kmcpg_results &r0 = job->res[0];
                    uint64_t hits = 0;
                    for (auto &r : job->res) hits += r.match_off[q + 1] - r.match_off[q];
                    if (hits == 0) {
                        if (!o.keep_unmatched) continue;
                        int n = snprintf(line, sizeof(line), "\t%d\t%d\t0\t0\t\t-1\t0\t0\t%d\t0\t0\t0\t0\t%llu\n", r0.query_len[q], r0.n_kmers[q], r0.k_used[q],     // S:460-511
                                         (unsigned long long)(bt.base + q));
                        out.append(idp, idn); out.append(line, (size_t)n);
                        continue;
                    }
                    nmatched[t]++;
                    auto put = [&](const kmcpg_results &r, const Db &db, const kmcpg_match &m) {
                        const kmcpg_target_t &tg = db.targets[m.target];
                        const std::string *mp = db.mapped[m.target];
                        int n1 = snprintf(line, sizeof(line), "\t%d\t%d\t%.4e\t%llu\t", r.query_len[q], r.n_kmers[q], m.fpr, (unsigned long long)hits);
                        out.append(idp, idn); out.append(line, (size_t)n1);
                        if (mp) out.append(*mp); else out.append(tg.name);
                        int n2 = snprintf(line, sizeof(line), "\t%u\t%u\t%llu\t%d\t%u\t%.4f\t%.4f\t%.4f\t%llu\n", tg.index & 0xFFFFu, tg.index >> 16,          // S:532-539
                                          (unsigned long long)tg.genome_size, r.k_used[q], m.count, m.qcov, m.tcov, m.jacc, (unsigned long long)(bt.base + q));
                        out.append(line, (size_t)n2);
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
            std::string all;
            size_t sz = 0;
            for (auto &t : text) sz += t.size();
            all.reserve(sz);
            for (auto &t : text) all += t;
            w.write(std::move(all));
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
            const uint32_t ns = (uint32_t)(bt->off.size() - 1);
            const int nc = (int)dbs[d].ctxs.size();
            const int rc = nc == 1           ? kmcpg_engine_search(dbs[d].ctxs[0], &eo, bt->seq.data(), bt->off.data(), ns, &job->res[d])
                           : dbs[d].replicas ? kmcpg_engine_search_replicas(dbs[d].ctxs.data(), nc, &eo, bt->seq.data(), bt->off.data(), ns, &job->res[d])
                                             : kmcpg_engine_search_sharded(dbs[d].ctxs.data(), nc, &eo, bt->seq.data(), bt->off.data(), ns, &job->res[d]);
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

// fastx_reader.h — the reader stage of `kmcp search` (reference search.go S:793-1000, sequences through
// bio/seqio/fastx over xopen/pgzip): FASTA/Q files → packed batches of queries (sequence bytes back to back + offsets, IDs in
// one arena) in input order; paired-end files are zipped pair by pair and end with the shorter one (S:806-867), `-g` makes one
// query of a whole file (S:885-937).  Everything around the parser is built for throughput: own gzip decoders (fastgz.h,
// pargz.h), decoding and parsing of every input on threads of their own, a table-driven path for four-line FASTQ records, an
// optional parallel parse of one input.  Host-only, header-only; errors are ReaderError exceptions.
// Used by reader.cpp (the C ABI: kmcpg_reader_*) and by cli_search.cpp (`kmcp-gpu parse`, the host-only test driver).
#pragma once
#include <errno.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <sys/wait.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "fastgz.h"
#include "pargz.h"
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace fastx {

struct ReaderError : std::runtime_error { using std::runtime_error::runtime_error; };

[[noreturn]] inline void fail(const char *fmt, ...) __attribute__((format(printf, 1, 2)));
[[noreturn]] inline void fail(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    throw ReaderError(buf);
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

// how hard one input is worked on (every reader carries its own copy)
struct Tuning {
    int inflate_threads = 0;                   // 0 = decide per file, 1 = always the sequential decoder, N = N workers on one .gz stream
    size_t inflate_chunk = 2u << 20;           // compressed bytes per task of the chunk-parallel decoder
    size_t inflate_cap = (size_t)64 << 20;     // most bytes a chunk may decode to before the sequential decoder takes over
    int parse_threads = 0;                     // 0/1 = one parser thread per input, N = N workers on one FASTQ text
    size_t parse_piece = 8u << 20;             // bytes of text per task of the parallel parser
};

// offsets (base + i) of every '\n' in p[0, n): 64 bytes per step with AVX2 where the CPU has it
#if defined(__x86_64__)
__attribute__((target("avx2"))) inline void scan_newlines_avx2(const char *p, size_t n, uint32_t base, std::vector<uint32_t> &out) {
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
inline void scan_newlines(const char *p, size_t n, uint32_t base, std::vector<uint32_t> &out) {
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
    Tuning tune;
    // line ends of buf[0, end) (offsets of '\n'), kept for the four-line FASTQ fast path: found 64 bytes at a time when the
    // buffer is filled instead of one memchr call per (short) line
    std::vector<uint32_t> nl;
    size_t nl_i = 0;
    // callers that filter on the whole header line (the index builder's -B): the last record's header without '>' / '@'
    bool keep_header = false;
    std::string header;
    bool open(const std::string &p, bool inflate_ahead = false, const Tuning &t = Tuning()) {
        tune = t;
        path = p;
        fd = p == "-" ? 0 : ::open(p.c_str(), O_RDONLY);
        ext_err.clear(); child = -1; tool = nullptr;
        if (fd > 0 && (tool = external_decoder(fd)) != nullptr) {
            // xz / zstd / bzip2 input (the reference's xopen reads these too): the system's decompressor writes into a pipe
            int pfd[2];
            if (pipe(pfd) != 0) { ::close(fd); fd = -1; return false; }
            child = fork();
            if (child == 0) {
                dup2(pfd[1], 1);
                ::close(pfd[0]); ::close(pfd[1]);
                execlp(tool, tool, "-dc", "--", p.c_str(), (char *)nullptr);
                _exit(127);
            }
            ::close(pfd[1]);
            ::close(fd);
            fd = pfd[0];
            if (child < 0) { ::close(fd); fd = -1; return false; }
        }
        if (fd >= 0) {
            const int h = fd, T = inflate_threads(h);
            if (T >= 2) pf = new fastgz::ParallelInflater(h, T, tune.inflate_chunk, tune.inflate_cap);
            else f = sequential(h);
        }
        buf.resize(16u << 20);
        pos = end = 0; eof = false; raw_total = 0;
        nl.clear(); nl_i = 0;
        if (fd >= 0 && inflate_ahead) { ahead = new InflateAhead(); ahead->start([this](void *dst, size_t n) { return raw_read(dst, n); }); }
        return fd >= 0;
    }
    Reader() = default;
    Reader(const Reader &) = delete;
    Reader &operator=(const Reader &) = delete;
    ~Reader() { close(); }
    void close() {
        if (ahead) { ahead->finish(); delete ahead; ahead = nullptr; }
        delete f; delete pf;
        f = nullptr; pf = nullptr;
        if (fd > 0) ::close(fd);
        fd = -1;
        if (child > 0) { int st; waitpid(child, &st, 0); child = -1; }     // its pipe is closed: it ends by itself
    }
    // "xz" / "zstd" / "bzip2" when the file starts with that format's magic bytes
    static const char *external_decoder(int h) {
        uint8_t m[6] = {0, 0, 0, 0, 0, 0};
        if (pread(h, m, 6, 0) < 4) return nullptr;
        if (m[0] == 0xFD && m[1] == '7' && m[2] == 'z' && m[3] == 'X' && m[4] == 'Z' && m[5] == 0) return "xz";
        if (m[0] == 0x28 && m[1] == 0xB5 && m[2] == 0x2F && m[3] == 0xFD) return "zstd";
        if (m[0] == 'B' && m[1] == 'Z' && m[2] == 'h' && m[3] >= '1' && m[3] <= '9') return "bzip2";
        return nullptr;
    }
    pid_t child = -1;
    const char *tool = nullptr;
    std::string ext_err;
    static fastgz::Inflater *sequential(int h) {
        return new fastgz::Inflater([h](void *dst, size_t n) -> ssize_t {
            for (;;) { const ssize_t r = ::read(h, dst, n); if (r >= 0 || errno != EINTR) return r; }
        });
    }
    uint64_t raw_total = 0;              // decoded bytes handed out so far
    ssize_t raw_read(void *dst, size_t n) {
        if (pf) {
            const ssize_t r = pf->read(dst, n);
            // a damaged or overwritten multi-member file is an error, as for the reference's reader (Go's gzip: "gzip: invalid header"), not a shorter input
            if (r == 0 && pf->trailing_garbage()) { ext_err = "gzip: invalid header (bytes that are not a gzip member follow the last member)"; return -1; }
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
        if (r == 0 && f->trailing_garbage()) { ext_err = "gzip: invalid header (bytes that are not a gzip member follow the last member)"; return -1; }
        if (r == 0 && child > 0) {              // the decompressor's verdict on the file
            int st = 0;
            waitpid(child, &st, 0);
            child = -1;
            if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) {
                ext_err = std::string(tool) + (WIFEXITED(st) && WEXITSTATUS(st) == 127 ? " is not installed: cannot read this file" : " could not decompress the file");
                return -1;
            }
        }
        return r;
    }
    // the decoded text itself (for callers that cut it into records themselves); like read(2)
    int read_text(char *dst, size_t n) { n = std::min<size_t>(n, 1u << 30); return ahead ? ahead->read(dst, n) : (int)raw_read(dst, n); }
    const char *error() const { return !ext_err.empty() ? ext_err.c_str() : pf ? pf->error() : (f ? f->error() : ""); }
    // worker threads for one input: --inflate-threads N, or by itself on big machines for big seekable gzip files
    int inflate_threads(int h) const {
        if (tune.inflate_threads == 1 || !fastgz::ParallelInflater::usable(h)) return 1;
        if (tune.inflate_threads > 1) return tune.inflate_threads;
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
        if (r < 0) fail("read error in %s: %s", path.c_str(), error());
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
        if (keep_header) header.assign(b + h0 + 1, h1 - h0 - 1);
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
        if (l[0] != '>' && l[0] != '@') fail("invalid FASTA/Q record in %s", path.c_str());
        if (keep_header) header.assign(l + 1, n - 1);
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

inline std::string trim_ext(const std::string &file) {     // filepathTrimExtension: strip dir, .gz/.xz/.zst/.bz2, then one extension
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
    void clear() { id_buf.clear(); id_off.assign(1, 0); seq.clear(); off.assign(1, 0); base = 0; }      // the arrays keep their memory
};

inline std::atomic<uint64_t> g_stat_pieces{0}, g_stat_fallbacks{0};      // pieces parsed by the workers / files handed back to the general reader

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
    bool stop = false, done = false, failed = false;
    std::string err;
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

    int parse_threads(const std::string &p) const {
        if (tune.parse_threads <= 1 || p == "-") return 1;
        struct stat st;
        if (stat(p.c_str(), &st) != 0 || !S_ISREG(st.st_mode)) return 1;
        // opt-in until it has been timed on the GPU box: on the 8 vCPUs it was written on the extra copies eat the gain
        return tune.parse_threads;
    }
    Tuning tune;
    bool open(const std::string &p, const Tuning &t = Tuning()) {
        path = p; tune = t;
        if (!r.open(p, true, tune)) return false;
        const int P = parse_threads(p);
        if (P >= 2) start_parallel(P); else start_plain(0);
        return true;
    }
    void start_plain(uint64_t skip) {
        th = std::thread([this, skip] { try { plain_loop(skip); } catch (const std::exception &e) { broke(e.what()); } });
    }
    // the plain parser: record by record into blocks, a few blocks ahead of the consumer
    void plain_loop(uint64_t skip) {
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
    }
    // a parser thread ends with an error: the consumer gets it after the blocks that are complete
    void broke(const char *what) {
        std::lock_guard<std::mutex> lk(mu);
        err = what; failed = true; done = true;
        cv.notify_all();
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
        th = std::thread([this] { try { cutter_loop(); } catch (const std::exception &e) { broke(e.what()); } });
    }
    // cuts the text into pieces that end on record boundaries and hands them to the workers
    void cutter_loop() {
        const size_t PIECE = std::max<size_t>(tune.parse_piece, 64);
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
                    if (got < 0) fail("read error in %s: %s", path.c_str(), r.error());
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
                if (ready.empty()) { if (failed) throw ReaderError(err); return nullptr; }
                cur = ready.front(); ready.pop_front(); i = 0;
                delivered += cur->n();
                cv.notify_all();
                return cur;
            }
            cv.wait(lk, [&] { return parsed.count(next_n) || (done && next_n == cut_n) || (failed && work_q.empty() && parsed.empty()); });
            auto it = parsed.find(next_n);
            if (it == parsed.end()) { if (failed) throw ReaderError(err); return nullptr; }
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
            if (!r.open(path, true, tune)) fail("%s: no such file", path.c_str());
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
    ~RecordStream() { close(); }             // an error on its way out must not meet a running thread
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
    Tuning tune;
    std::function<Batch *()> new_batch;                              // where empty batches come from (a pool of used ones); may be empty
    std::function<void(const char *level, const char *msg)> log;      // the reference's log lines (S:800, 878, 920); may be empty
    void say(const char *level, const char *fmt, ...) const __attribute__((format(printf, 3, 4))) {
        if (!log) return;
        char buf[1024];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof(buf), fmt, ap);
        va_end(ap);
        log(level, buf);
    }
};

// S:793-1000: the input files as batches, in order; `emit` takes the batch over
inline void read_batches(const ReaderConfig &c, const std::function<void(Batch *)> &emit_fn) {
    auto fresh = [&]() { return c.new_batch ? c.new_batch() : new Batch(); };
    std::unique_ptr<Batch> cur(fresh());
    auto emit = [&]() {
        if (cur->n_ids() == 0) return;
        if (cur->seq.empty()) cur->seq.push_back(0);
        if (cur->id_buf.empty()) cur->id_buf.push_back(0);        // the arrays of a batch are never NULL
        Batch *full_batch = cur.release();
        cur.reset(fresh());
        emit_fn(full_batch);
    };
    auto full = [&]() { return cur->n_ids() >= c.batch_reads || cur->seq.size() >= c.batch_bytes; };
    if (c.paired) {
        RecordStream r1, r2;
        if (!r1.open(c.read1, c.tune)) fail("%s: no such file", c.read1.c_str());
        if (!r2.open(c.read2, c.tune)) fail("%s: no such file", c.read2.c_str());
        c.say("INFO", "reading from paired-end files: %s, %s", c.read1.c_str(), c.read2.c_str());
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
            c.say("INFO", "reading sequence file: %s", file.c_str());
            if (c.whole_file) {                                       // S:885-937 (the N-run follows every record after the second)
                Reader r;
                if (!r.open(file, true, c.tune)) fail("%s: no such file", file.c_str());
                std::string qid;
                bool first = true;
                const size_t mark = cur->seq.size();
                while (r.next(id, cur->seq)) {
                    if (first) { qid = c.use_filename ? trim_ext(file) : (!c.query_id.empty() ? c.query_id : id); first = false; }
                    else cur->seq.insert(cur->seq.end(), (size_t)(c.kmax - 1), (uint8_t)'N');
                }
                r.close();
                if (first) { c.say("WARN", "no valid sequences in file: %s", file.c_str()); cur->seq.resize(mark); continue; }
                cur->add_id(qid); cur->off.push_back(cur->seq.size());
                if (cur->seq.size() >= c.batch_bytes) emit();
            } else {
                RecordStream rs;
                if (!rs.open(file, c.tune)) fail("%s: no such file", file.c_str());
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
                if (!any) c.say("WARN", "no valid sequences in file: %s", file.c_str());
            }
        }
    }
    emit();
}


}  // namespace fastx

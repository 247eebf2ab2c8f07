// pargz.h — chunk-parallel decoding of ONE gzip stream ("parallel FASTQ inflate"; the reference's pgzip
// reader only decompresses ahead on one goroutine, so a single .fq.gz caps `kmcp search` at the rate of one inflate).
//
// A DEFLATE stream can only be decoded from a block boundary, and a block may copy from the 32 KB in front of it.  So:
//   * the file is cut into chunks of CHUNK compressed bytes; a worker looks for the first block header at or behind its
//     chunk's first bit (dynamic-Huffman headers are self-validating: complete code-length code, complete literal/length
//     code with an end-of-block symbol, code lengths that add up exactly), and decodes from there up to the first block
//     boundary at or behind the next chunk's first bit; a gzip MEMBER that starts there (bgzip / pgzip files, `cat a.gz b.gz`)
//     is a place to start as well, and a better one: nothing in front of it can be copied from;
//   * bytes copied from in front of the chunk are not known yet: the worker writes 16-bit symbols, 0..255 for a known byte
//     and 0x8000+j for "byte j of the 32 KB window in front of this chunk"; markers are copied like any other symbol.  Once
//     the newest 32 KB hold no marker the worker carries on with plain bytes;
//   * the consumer takes the chunks in file order, checks that each one starts at the very bit the data before it ended at
//     (otherwise the chunk is discarded and the stretch is decoded again in order — a false block start, or a boundary the
//     search cannot see such as stored and fixed-Huffman blocks), replaces the markers from the real window, and verifies
//     CRC-32 and ISIZE of every member over the final bytes.  Streams where the workers keep missing (stored blocks: incompressible
//     data) are decoded in order after a few misses; a chunk that decodes to more than a cap ends the parallel mode (memory stays
//     bounded for any compression ratio) and the caller falls back to the sequential decoder.
// The bytes delivered are exactly those of fastgz::Inflater / gzread(); tests/test_fastgz.py compares the two on every stream.
//
// Host-only code.  Needs a seekable file (pread); pipes use fastgz::Inflater.
#pragma once
#include <stdlib.h>
#include <sys/stat.h>
#include <unistd.h>

#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <new>
#include <thread>

#include "fastgz.h"

namespace fastgz {

namespace pargz_detail {

template <class T>
struct RawBuf {                       // growable, uninitialised
    T *p = nullptr;
    size_t cap = 0;
    RawBuf() = default;
    RawBuf(const RawBuf &) = delete;
    RawBuf &operator=(const RawBuf &) = delete;
    RawBuf(RawBuf &&o) noexcept : p(o.p), cap(o.cap) { o.p = nullptr; o.cap = 0; }
    RawBuf &operator=(RawBuf &&o) noexcept { if (this != &o) { free(p); p = o.p; cap = o.cap; o.p = nullptr; o.cap = 0; } return *this; }
    ~RawBuf() { free(p); }
    void reserve(size_t n) {
        if (n <= cap) return;
        T *q = (T *)realloc(p, n * sizeof(T));
        if (!q) throw std::bad_alloc();
        p = q; cap = n;
    }
};

struct MemberEnd { uint64_t out_off; uint32_t crc, isize; };      // a member ended after out_off bytes of this chunk's output

struct ChunkResult {
    bool found = false;               // a place to start was found (chunk 0: the gzip header was read)
    bool not_gzip = false;            // chunk 0 only
    bool at_eof = false;              // the last member ended and nothing (or only garbage) follows
    bool garbage = false;             // ... and what follows is not a gzip member
    bool failed = false;
    bool too_big = false;             // failed because the chunk decodes to more than the cap (a stream for the sequential decoder)
    std::string err;
    uint64_t start_bit = 0, end_bit = 0;
    // output: first the symbols written while the window in front of the chunk was unknown, then plain bytes.  Both arrays
    // carry a 32768-element prefix (markers / the window at the switch) so that copies never need a special case.
    RawBuf<uint16_t> sym;
    RawBuf<uint8_t> bytes;
    size_t n_sym = 0, n_bytes = 0;
    std::vector<MemberEnd> members;
};

static constexpr size_t WIN = 32768;
using I = Inflater;

// compressed bytes of the file around a position, loaded on demand; 64 zero bytes always follow the valid bytes
struct Source {
    int fd;
    uint64_t file_size;
    std::vector<uint8_t> z;
    uint64_t z_off = 0;               // file offset of z[0]
    size_t z_len = 0;
    bool io_error = false;
    Source(int f, uint64_t size) : fd(f), file_size(size) {}
    bool eof() const { return z_off + z_len >= file_size; }
    void load(uint64_t off, size_t len) {
        z_off = std::min(off, file_size); z_len = 0;
        z.assign(64, 0);
        more(len);
    }
    // appends up to len more bytes; false when nothing could be added
    bool more(size_t len) {
        const uint64_t at = z_off + z_len;
        if (at >= file_size) return false;
        len = (size_t)std::min<uint64_t>(len, file_size - at);
        z.resize(z_len + len + 64);
        size_t got = 0;
        while (got < len) {
            const ssize_t r = pread(fd, z.data() + z_len + got, len - got, (off_t)(at + got));
            if (r < 0) { if (errno == EINTR) continue; io_error = true; break; }
            if (r == 0) break;
            got += (size_t)r;
        }
        if (got < len) file_size = at + got;          // shorter than fstat said
        z_len += got;
        memset(z.data() + z_len, 0, 64);
        return got != 0;
    }
};

#define PARGZ_REFILL()                  \
    do {                                \
        uint64_t w_;                    \
        memcpy(&w_, in, 8);             \
        bb |= w_ << bc;                 \
        in += (63 - bc) >> 3;           \
        bc |= 56;                       \
    } while (0)

// the decoder of one chunk (or of a stretch that has to be decoded again in order)
class ChunkDecoder {
  public:
    ChunkDecoder(int fd, uint64_t file_size, size_t load_bytes, size_t max_out = (size_t)256 << 20)
        : src_(fd, file_size), load_(load_bytes), max_out_(max_out) {
        static const I::FixedTables fixed;
        fixed_ = &fixed;
        lt_dyn_.resize(I::LT_SIZE);
        dt_dyn_.resize(I::DT_SIZE);
    }

    // chunk 0: gzip header at the start of the file, then bytes
    void run_first(ChunkResult &R, uint64_t stop_bit) {
        src_.load(0, load_);
        in_ = src_.z.data(); bb_ = 0; bc_ = 0;
        if (src_.z_len < 2 || in_[0] != 0x1f || in_[1] != 0x8b) { R.not_gzip = true; return; }
        if (!member_header(R)) return;
        R.found = true;
        R.start_bit = bitpos();
        start_bytes(R, nullptr, 0);
        decode(R, stop_bit);
    }
    // a later chunk: look for a block header in [lo_bit, stop_bit), decode with an unknown window
    void run_search(ChunkResult &R, uint64_t lo_bit, uint64_t stop_bit) {
        const uint64_t lo_byte = lo_bit >> 3;
        src_.load(lo_byte > BACK ? lo_byte - BACK : 0, load_ + BACK);           // a member header may straddle the chunk's first bit
        uint64_t at = 0, member_at = 0;
        const bool have_member = find_member(lo_bit, stop_bit, member_at);
        const bool have_block = find_block(lo_bit, have_member ? member_at : stop_bit, at);
        if (!have_member && !have_block) return;
        R.found = true;
        if (!have_block) {
            // the first block boundary is the start of a gzip member (bgzip / pgzip files, `cat a.gz b.gz`): nothing in front of
            // it can be copied from, so this chunk is plain bytes from its first symbol
            R.start_bit = member_at;
            seek(member_at);
            start_bytes(R, nullptr, 0);
            decode(R, stop_bit);
            return;
        }
        R.start_bit = at;
        seek(at);
        R.sym.reserve(WIN + (load_ * 4) + SLACK);
        for (size_t j = 0; j < WIN; j++) R.sym.p[j] = (uint16_t)(0x8000u | j);
        R.n_sym = 0;
        markers_ = true;
        decode(R, stop_bit);
    }
    // in order, from a known position with a known window (hist = the bytes in front of start_bit, at most 32 KB)
    void run_from(ChunkResult &R, uint64_t start_bit, uint64_t stop_bit, const uint8_t *hist, size_t n_hist) {
        src_.load(start_bit >> 3, load_);
        R.found = true;
        R.start_bit = start_bit;
        seek(start_bit);
        start_bytes(R, hist, n_hist);
        decode(R, stop_bit);
    }

  private:
    static constexpr size_t SLACK = 320;
    static constexpr uint64_t SEARCH_BYTES = 256u << 10;
    static constexpr uint64_t BACK = 1024;        // how far in front of a chunk a member header that reaches into it may start

    uint64_t bitpos() const { return (src_.z_off + (uint64_t)(in_ - src_.z.data())) * 8 - bc_; }
    void seek(uint64_t bit) {
        in_ = src_.z.data() + ((bit >> 3) - src_.z_off);
        bb_ = 0; bc_ = 0;
        const uint8_t *in = in_; uint64_t bb = bb_; unsigned bc = bc_;
        PARGZ_REFILL();
        bb >>= (bit & 7); bc -= (unsigned)(bit & 7);
        in_ = in; bb_ = bb; bc_ = bc;
    }
    // true: more compressed bytes are available behind in_ than before
    bool more_input() {
        const size_t at = (size_t)(in_ - src_.z.data());
        const bool ok = src_.more(load_);
        in_ = src_.z.data() + at;
        return ok;
    }
    bool fail(ChunkResult &R, const char *msg) { R.failed = true; R.err = src_.io_error ? "read error" : msg; return false; }
    void start_bytes(ChunkResult &R, const uint8_t *hist, size_t n_hist) {
        markers_ = false;
        R.bytes.reserve(WIN + load_ * 4 + SLACK);
        n_hist = std::min(n_hist, WIN);
        if (n_hist && hist) memcpy(R.bytes.p + WIN - n_hist, hist, n_hist);
        R.n_bytes = 0;
        vstart_ = WIN - n_hist;
    }
    int byte_aligned() {                       // the next whole byte of the stream, -1 at the end; bit buffer must be empty
        if ((size_t)(in_ - src_.z.data()) >= src_.z_len && !more_input()) return -1;
        return *in_++;
    }
    bool align() {
        bb_ >>= (bc_ & 7); bc_ -= (bc_ & 7);
        in_ -= bc_ >> 3;
        bb_ = 0; bc_ = 0;
        return (size_t)(in_ - src_.z.data()) <= src_.z_len;
    }
    // gzip member header at in_ (bit buffer empty)
    bool member_header(ChunkResult &R) {
        auto gb = [&]() { return byte_aligned(); };
        if (gb() != 0x1f || gb() != 0x8b) return fail(R, "not a gzip header");
        const int cm = gb(), flg = gb();
        if (cm != 8 || flg < 0 || (flg & 0xE0)) return fail(R, cm < 0 || flg < 0 ? "truncated gzip header" : "unsupported gzip header");
        for (int i = 0; i < 6; i++) if (gb() < 0) return fail(R, "truncated gzip header");
        if (flg & 4) {
            const int a = gb(), b = gb();
            if (a < 0 || b < 0) return fail(R, "truncated gzip header");
            for (int n = a | (b << 8); n > 0; n--) if (gb() < 0) return fail(R, "truncated gzip header");
        }
        for (int bit : {8, 16})
            if (flg & bit) { int c; do { c = gb(); if (c < 0) return fail(R, "truncated gzip header"); } while (c); }
        if (flg & 2) { if (gb() < 0 || gb() < 0) return fail(R, "truncated gzip header"); }
        return true;
    }

    // ---- block headers -------------------------------------------------------------------------------------------------
    // Reads HLIT/HDIST/HCLEN and the code lengths of a dynamic block from a 64-bit reader over memory.  strict: only accept what
    // a compressor writes (complete codes), as the search for a block start needs.  On success lens[0..hlit+hdist) are set.
    struct Peek {                               // bit reader for the search: nothing shared with the decoder's state
        const uint8_t *p;
        uint64_t bb = 0;
        unsigned bc = 0;
        explicit Peek(const uint8_t *at) : p(at) {}
        void fill() { uint64_t w; memcpy(&w, p, 8); bb |= w << bc; p += (63 - bc) >> 3; bc |= 56; }
        uint32_t peek(unsigned n) const { return (uint32_t)(bb & ((1u << n) - 1)); }
        void drop(unsigned n) { bb >>= n; bc -= n; }
        uint32_t bits(unsigned n) { fill(); const uint32_t v = peek(n); drop(n); return v; }      // n <= 32
    };
    static bool kraft_complete(const uint8_t *lens, int n, int maxlen, bool allow_single) {
        int count[16] = {0};
        for (int i = 0; i < n; i++) count[lens[i]]++;
        int nz = n - count[0];
        if (nz == 0) return allow_single;
        int left = 1;
        for (int l = 1; l <= maxlen; l++) { left = left * 2 - count[l]; if (left < 0) return false; }
        return left == 0 || (allow_single && nz == 1);
    }
    // is there a plausible non-final dynamic block header at bit `bit` of memory z (z must be readable for 400 bytes behind it)?
    static bool plausible_block(const uint8_t *z, uint64_t bit, bool allow_final = false) {
        uint64_t w;
        memcpy(&w, z + (bit >> 3), 8);
        w >>= (bit & 7);
        if ((w & 6) != 4) return false;                                   // BTYPE 2 (bits 1-2: 0 then 1)
        if ((w & 1) && !allow_final) return false;                        // BFINAL 0 unless the caller knows better
        const uint32_t hlit = (uint32_t)((w >> 3) & 31), hdist = (uint32_t)((w >> 8) & 31), hclen = (uint32_t)((w >> 13) & 15) + 4;
        if (hlit > 29 || hdist > 29) return false;
        Peek pk(z + (bit >> 3));
        pk.bits((unsigned)(bit & 7)); pk.bits(17);
        static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
        uint8_t pl[19] = {0};
        for (uint32_t i = 0; i < hclen; i++) pl[order[i]] = (uint8_t)pk.bits(3);
        if (!kraft_complete(pl, 19, 7, false)) return false;
        uint16_t pt[128];
        build_precode(pl, pt);
        uint8_t lens[286 + 30 + 138];
        const uint32_t total = hlit + 257 + hdist + 1;
        uint32_t n = 0;
        while (n < total) {
            pk.fill();
            const uint16_t e = pt[pk.peek(7)];
            if (e == 0xFFFF) return false;
            pk.drop(e >> 8);
            const int sym = e & 0xFF;
            if (sym < 16) { lens[n++] = (uint8_t)sym; continue; }
            uint32_t rep; uint8_t val = 0;
            if (sym == 16) { if (!n) return false; val = lens[n - 1]; rep = 3 + pk.peek(2); pk.drop(2); }
            else if (sym == 17) { rep = 3 + pk.peek(3); pk.drop(3); }
            else { rep = 11 + pk.peek(7); pk.drop(7); }
            if (n + rep > total) return false;
            memset(lens + n, val, rep);
            n += rep;
        }
        if (lens[256] == 0) return false;
        if (!kraft_complete(lens, (int)hlit + 257, 15, false)) return false;
        if (!kraft_complete(lens + hlit + 257, (int)hdist + 1, 15, true)) return false;
        return true;
    }
    static void build_precode(const uint8_t *pl, uint16_t *pt) {
        int count[8] = {0};
        for (int i = 0; i < 19; i++) count[pl[i]]++;
        uint32_t next[8], code = 0;
        count[0] = 0;
        for (int l = 1; l <= 7; l++) { code = (code + (uint32_t)count[l - 1]) << 1; next[l] = code; }
        for (int i = 0; i < 128; i++) pt[i] = 0xFFFF;
        for (int s = 0; s < 19; s++) {
            const int l = pl[s];
            if (!l) continue;
            for (uint32_t i = I::rev_bits(next[l]++, l); i < 128; i += 1u << l) pt[i] = (uint16_t)(s | (l << 8));
        }
    }
    // the first gzip member whose first block starts in [lo_bit, hi_bit): magic, method and flag bytes, the optional header fields,
    // then a first block that makes sense (a plausible dynamic header, final or not; a stored block whose two length fields agree;
    // fixed-Huffman blocks cannot be told from noise and are taken as they are — the member's bytes have to decode anyway, and
    // the consumer only uses the chunk if the data in front of it ends exactly here)
    bool find_member(uint64_t lo_bit, uint64_t hi_bit, uint64_t &at) {
        const uint64_t need_to = std::min<uint64_t>((hi_bit >> 3) + 1024, src_.file_size);
        while (src_.z_off + src_.z_len < need_to && src_.more(load_)) {}
        const uint8_t *z = src_.z.data();
        const size_t n = src_.z_len;
        size_t i = 0;
        while (i + 18 <= n) {
            const uint8_t *q = (const uint8_t *)memchr(z + i, 0x1f, n - i - 17);
            if (!q) break;
            i = (size_t)(q - z);
            size_t p = i + 10;
            const uint8_t flg = z[i + 3];
            bool ok = z[i + 1] == 0x8b && z[i + 2] == 8 && !(flg & 0xE0);
            if (ok && (flg & 4)) { if (p + 2 <= n) { const size_t xl = z[p] | (z[p + 1] << 8); p += 2 + xl; } else ok = false; }
            for (int bit : {8, 16})
                if (ok && (flg & bit)) { while (p < n && p < i + BACK && z[p]) p++; if (p >= n || z[p]) ok = false; else p++; }
            if (ok && (flg & 2)) p += 2;
            if (ok && p + 640 <= n) {
                const uint64_t b = (src_.z_off + p) * 8;                  // the member's first block
                if (b >= hi_bit) return false;
                if (b >= lo_bit) {
                    const uint32_t type = (z[p] >> 1) & 3;
                    bool good = false;
                    if (type == 2) good = plausible_block(z, (uint64_t)p * 8, true);
                    else if (type == 1) good = true;
                    else if (type == 0) good = ((z[p + 1] | (z[p + 2] << 8)) ^ (z[p + 3] | (z[p + 4] << 8))) == 0xFFFF;
                    if (good) { at = b; return true; }
                }
            }
            i++;
        }
        return false;
    }
    // first plausible block header in [lo_bit, hi_bit)
    bool find_block(uint64_t lo_bit, uint64_t hi_bit, uint64_t &at) {
        // the header of a candidate may reach ~600 bytes past it
        const uint64_t need_to = std::min<uint64_t>((hi_bit >> 3) + 1024, src_.file_size);
        while (src_.z_off + src_.z_len < need_to && src_.more(load_)) {}
        const uint64_t base_bit = src_.z_off * 8;
        // compressors close a block after 16-64 K symbols (tens of KB); a quarter of a megabyte without a header means this
        // stretch is stored blocks or one huge block, and looking further only burns time the consumer waits for
        hi_bit = std::min(hi_bit, lo_bit + SEARCH_BYTES * 8);
        for (uint64_t b = lo_bit; b < hi_bit; b++) {
            const uint64_t rel = b - base_bit;
            if ((rel >> 3) + 640 > src_.z_len) return false;       // a header may reach ~570 bytes past its first bit
            // cheap test on the first three bits before anything else
            const uint8_t *q = src_.z.data() + (rel >> 3);
            const uint32_t three = ((uint32_t)q[0] | ((uint32_t)q[1] << 8)) >> (rel & 7);
            if ((three & 7) != 4) continue;
            if (plausible_block(src_.z.data(), rel)) { at = b; return true; }
        }
        return false;
    }

    // BFINAL/BTYPE at the current position; sets up the tables or the stored length.  type_ = 0 stored, 1 compressed
    bool block_header(ChunkResult &R) {
        if (src_.z_len - (size_t)(in_ - src_.z.data()) < 1536) more_input();
        const uint8_t *in = in_;
        uint64_t bb = bb_;
        unsigned bc = bc_;
        auto bits = [&](unsigned n) -> uint32_t {
            PARGZ_REFILL();
            const uint32_t v = (uint32_t)(bb & ((1u << n) - 1));
            bb >>= n; bc -= n;
            return v;
        };
        auto store = [&]() { in_ = in; bb_ = bb; bc_ = bc; };
        auto overrun = [&]() { return (size_t)(in - src_.z.data()) - (bc >> 3) > src_.z_len; };
        final_ = bits(1) != 0;
        const uint32_t type = bits(2);
        if (type == 0) {
            store();
            if (!align()) return fail(R, "truncated deflate stream");
            uint8_t h[4];
            for (int i = 0; i < 4; i++) { const int c = byte_aligned(); if (c < 0) return fail(R, "truncated stored block"); h[i] = (uint8_t)c; }
            const uint32_t len = h[0] | (h[1] << 8), nlen = h[2] | (h[3] << 8);
            if ((len ^ nlen) != 0xFFFF) return fail(R, "invalid stored block lengths");
            stored_left_ = len;
            type_ = 0;
            return true;
        }
        if (type == 1) {
            lt_ = fixed_->lt.data(); dt_ = fixed_->dt.data();
        } else if (type == 2) {
            const uint32_t hlit = bits(5) + 257, hdist = bits(5) + 1, hclen = bits(4) + 4;
            if (hlit > 286 || hdist > 30) return fail(R, "too many length or distance symbols");
            static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
            uint8_t pl[19] = {0};
            for (uint32_t i = 0; i < hclen; i++) pl[order[i]] = (uint8_t)bits(3);
            {
                int count[8] = {0};
                for (int i = 0; i < 19; i++) count[pl[i]]++;
                int left = 1;
                for (int l = 1; l <= 7; l++) { left = left * 2 - count[l]; if (left < 0) return fail(R, "invalid code lengths set"); }
            }
            uint16_t pt[128];
            build_precode(pl, pt);
            uint8_t lens[286 + 30 + 138];
            const uint32_t total = hlit + hdist;
            uint32_t n = 0;
            while (n < total) {
                PARGZ_REFILL();
                const uint16_t e = pt[bb & 127];
                if (e == 0xFFFF) return fail(R, "invalid code lengths set");
                bb >>= (e >> 8); bc -= (e >> 8);
                const int sym = e & 0xFF;
                if (sym < 16) { lens[n++] = (uint8_t)sym; continue; }
                uint32_t rep; uint8_t val = 0;
                if (sym == 16) {
                    if (!n) return fail(R, "invalid bit length repeat");
                    val = lens[n - 1]; rep = 3 + (uint32_t)(bb & 3); bb >>= 2; bc -= 2;
                } else if (sym == 17) { rep = 3 + (uint32_t)(bb & 7); bb >>= 3; bc -= 3; }
                else { rep = 11 + (uint32_t)(bb & 127); bb >>= 7; bc -= 7; }
                if (n + rep > total) return fail(R, "invalid bit length repeat");
                memset(lens + n, val, rep);
                n += rep;
            }
            if (overrun()) return fail(R, "truncated deflate stream");
            if (lens[256] == 0) return fail(R, "invalid code -- missing end-of-block");
            if (!I::build_table(lt_dyn_.data(), I::LT_SIZE, I::LBITS, lens, (int)hlit, true)) return fail(R, "invalid literal/lengths set");
            if (!I::build_table(dt_dyn_.data(), I::DT_SIZE, I::DBITS, lens + hlit, (int)hdist, false)) return fail(R, "invalid distances set");
            lt_ = lt_dyn_.data(); dt_ = dt_dyn_.data();
        } else
            return fail(R, "invalid block type");
        if (overrun()) return fail(R, "truncated deflate stream");
        store();
        type_ = 1;
        return true;
    }

    // ---- the symbols of one compressed block, written as T (bytes, or 16-bit symbols while the window is unknown) ----------
    template <class T>
    bool huff(ChunkResult &R, RawBuf<T> &buf, size_t &n_out) {
        const uint8_t *in = in_;
        uint64_t bb = bb_;
        unsigned bc = bc_;
        T *out = buf.p + WIN + n_out;
        T *out_limit = buf.p + buf.cap - SLACK;
        const uint32_t *const LT = lt_, *const DT = dt_;
        const uint32_t LM = (1u << I::LBITS) - 1, DM = (1u << I::DBITS) - 1;
        const uint8_t *in_safe = safe_end();
        bool ok = false;
        // `e` is the table entry of the next symbol, looked up before the previous symbol's copy (see fastgz.h)
        PARGZ_REFILL();
        uint32_t e = LT[bb & LM];
        for (;;) {
            if (__builtin_expect(in >= in_safe || out >= out_limit, 0)) {
                if (out >= out_limit) {
                    const size_t n = (size_t)(out - buf.p);
                    if (n - WIN > max_out_) { R.too_big = true; fail(R, "chunk decodes to more than the cap of the parallel decoder"); break; }
                    buf.reserve(buf.cap + buf.cap / 2);
                    out = buf.p + n; out_limit = buf.p + buf.cap - SLACK;
                    continue;
                }
                if (!src_.eof()) {
                    in_ = in;
                    more_input();
                    in = in_; in_safe = safe_end();
                    if (src_.io_error) { fail(R, "read error"); break; }
                    continue;
                }
                if ((size_t)(in - src_.z.data()) - (bc >> 3) > src_.z_len) { fail(R, "truncated deflate stream"); break; }
            }
            if (e & I::E_LIT) {
#define PARGZ_LITERALS()                                        \
    do {                                                        \
        out[0] = (T)((e >> 16) & 0xFF);                         \
        out[1] = (T)(e >> 24);                                  \
        out += 1 + ((e >> 13) & 1);                             \
        bb >>= (e & 63); bc -= (e & 63);                        \
        e = LT[bb & LM];                                        \
    } while (0)
                PARGZ_LITERALS();
                if (e & I::E_LIT) {
                    PARGZ_LITERALS();
                    if (e & I::E_LIT) {
                        PARGZ_LITERALS();
                        if (e & I::E_LIT) PARGZ_LITERALS();
                    }
                }
#undef PARGZ_LITERALS
                PARGZ_REFILL();
                if (e & I::E_LIT) continue;
            }
            if (__builtin_expect(e & I::E_EXC, 0)) {
                if (e & I::E_SUB) {
                    bb >>= I::LBITS; bc -= I::LBITS;
                    e = LT[(e >> 16) + ((uint32_t)bb & ((1u << ((e >> 8) & 15)) - 1))];
                    if (e & I::E_LIT) {
                        bb >>= (e & 63); bc -= (e & 63); *out++ = (T)((e >> 16) & 0xFF);
                        PARGZ_REFILL();
                        e = LT[bb & LM];
                        continue;
                    }
                }
                if (e & I::E_EXC) {
                    if ((e >> 16) == 0) { bb >>= (e & 63); bc -= (e & 63); ok = true; break; }
                    fail(R, "invalid literal/length code"); break;
                }
            }
            uint64_t saved = bb;
            unsigned tot = e & 63;
            bb >>= tot; bc -= tot;
            const uint32_t len = (e >> 16) + (((uint32_t)saved & ((1u << tot) - 1)) >> ((e >> 8) & 15));
            e = DT[bb & DM];
            if (__builtin_expect(e & I::E_EXC, 0)) {
                if (e & I::E_SUB) {
                    bb >>= I::DBITS; bc -= I::DBITS;
                    e = DT[(e >> 16) + ((uint32_t)bb & ((1u << ((e >> 8) & 15)) - 1))];
                }
                if (e & I::E_EXC) { fail(R, "invalid distance code"); break; }
            }
            saved = bb;
            tot = e & 63;
            bb >>= tot; bc -= tot;
            const uint32_t dist = (e >> 16) + (((uint32_t)saved & ((1u << tot) - 1)) >> ((e >> 8) & 15));
            e = LT[bb & LM];
            // 16-bit symbols: the marker prefix makes every distance valid here, the consumer checks it when it resolves them
            if (__builtin_expect((size_t)(out - buf.p) - vstart_of<T>() < dist, 0)) { fail(R, "invalid distance too far back"); break; }
            const T *src = out - dist;
            T *dst = out;
            out += len;
            constexpr uint32_t W = 16 / sizeof(T);
            if (dist >= W) {
                do { memcpy(dst, src, 16); src += W; dst += W; } while (dst < out);
            } else if (dist == 1) {
                const T v = src[0];
                do { for (uint32_t i = 0; i < W; i++) dst[i] = v; dst += W; } while (dst < out);
            } else {
                do { *dst++ = *src++; } while (dst < out);
            }
            PARGZ_REFILL();
        }
        in_ = in; bb_ = bb; bc_ = bc;
        n_out = (size_t)(out - buf.p) - WIN;
        return ok;
    }
    template <class T> size_t vstart_of() const { return sizeof(T) == 1 ? vstart_ : 0; }
    const uint8_t *safe_end() const {
        const uint8_t *end = src_.z.data() + src_.z_len;
        if (src_.eof()) return end;
        return end - in_ > 16 ? end - 16 : in_;
    }
    template <class T>
    bool stored(ChunkResult &R, RawBuf<T> &buf, size_t &n_out) {
        while (stored_left_) {
            if ((size_t)(in_ - src_.z.data()) >= src_.z_len && !more_input()) return fail(R, "truncated stored block");
            const size_t n = std::min<size_t>(stored_left_, src_.z_len - (size_t)(in_ - src_.z.data()));
            if (n_out + n > max_out_ + (1u << 20)) { R.too_big = true; return fail(R, "chunk decodes to more than the cap of the parallel decoder"); }
            if (WIN + n_out + n + SLACK > buf.cap) buf.reserve(std::max(buf.cap + buf.cap / 2, WIN + n_out + n + SLACK));
            T *out = buf.p + WIN + n_out;
            if (sizeof(T) == 1) memcpy(out, in_, n);
            else for (size_t i = 0; i < n; i++) out[i] = in_[i];
            n_out += n; in_ += n; stored_left_ -= (uint32_t)n;
        }
        return true;
    }

    // blocks from the current position up to the first block boundary at or behind stop_bit, or the end of the stream
    void decode(ChunkResult &R, uint64_t stop_bit) {
        for (;;) {
            const uint64_t at = bitpos();
            if (at >= stop_bit && at > R.start_bit) { R.end_bit = at; return; }
            if (!block_header(R)) return;
            bool ok;
            if (markers_) ok = type_ ? huff<uint16_t>(R, R.sym, R.n_sym) : stored<uint16_t>(R, R.sym, R.n_sym);
            else ok = type_ ? huff<uint8_t>(R, R.bytes, R.n_bytes) : stored<uint8_t>(R, R.bytes, R.n_bytes);
            if (!ok) return;
            if (final_) {
                if (!align()) { fail(R, "truncated deflate stream"); return; }
                uint32_t v[2] = {0, 0};
                for (int i = 0; i < 8; i++) {
                    const int c = byte_aligned();
                    if (c < 0) { fail(R, "truncated gzip trailer"); return; }
                    v[i >> 2] |= (uint32_t)c << (8 * (i & 3));
                }
                R.members.push_back({(uint64_t)(R.n_sym + R.n_bytes), v[0], v[1]});
                // another member?  (the bit buffer is empty here)
                const size_t off = (size_t)(in_ - src_.z.data());
                if (src_.z_len - off < 2) more_input();
                const size_t left = src_.z_len - (size_t)(in_ - src_.z.data());
                if (left < 2 || in_[0] != 0x1f || in_[1] != 0x8b) { R.at_eof = true; R.garbage = left > 0; R.end_bit = bitpos(); return; }   // end, or garbage that gzread ignores too
                if (!member_header(R)) return;
                // a new member starts with an empty window: plain bytes from here on, whatever came before
                if (markers_) start_bytes(R, nullptr, 0);
                else vstart_ = WIN + R.n_bytes;
            } else if (markers_ && R.n_sym >= WIN) {
                // the newest 32 KB free of markers: they are the window, carry on with bytes
                const uint16_t *t = R.sym.p + WIN + R.n_sym - WIN;
                uint16_t any = 0;
                for (size_t i = 0; i < WIN; i++) any |= t[i];
                if (!(any & 0x8000)) {
                    uint8_t hist[WIN];
                    for (size_t i = 0; i < WIN; i++) hist[i] = (uint8_t)t[i];
                    start_bytes(R, hist, WIN);
                }
            }
        }
    }

    Source src_;
    size_t load_, max_out_;
    const uint8_t *in_ = nullptr;
    uint64_t bb_ = 0;
    unsigned bc_ = 0;
    bool final_ = false, markers_ = false;
    int type_ = 0;
    uint32_t stored_left_ = 0;
    size_t vstart_ = 0;               // first valid index of the byte buffer (history in front of it does not exist)
    const I::FixedTables *fixed_ = nullptr;
    std::vector<uint32_t> lt_dyn_, dt_dyn_;
    const uint32_t *lt_ = nullptr, *dt_ = nullptr;
};

#undef PARGZ_REFILL

}  // namespace pargz_detail

// ---------------------------------------------------------------------------------------------------------------------
// read() like gzread(), decoded by `threads` workers.  The file must be seekable and start with a gzip member.
//
// Workers run two kinds of task: (1) decode a chunk from a block start they find themselves — 16-bit symbols while the
// window is unknown; (2) turn the symbols of a chunk into bytes once the window in front of it is known.  The thread
// calling read() does the little that has to happen in order: it checks that chunk k starts at the bit chunk k-1 ended at,
// derives the window behind chunk k from its newest 32 KB (so chunk k+1 can be resolved while chunk k still is), and
// finally runs CRC-32 / ISIZE over the finished bytes and hands them out.
// ---------------------------------------------------------------------------------------------------------------------
class ParallelInflater {
  public:
    static constexpr size_t WIN = pargz_detail::WIN;

    // usable(fd): a regular file that starts with the gzip magic
    static bool usable(int fd) {
        struct stat st;
        if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode) || st.st_size < 18) return false;
        uint8_t m[2];
        return pread(fd, m, 2, 0) == 2 && m[0] == 0x1f && m[1] == 0x8b;
    }

    // max_chunk_out: a chunk (or a stretch decoded again) that decodes to more bytes than this ends the stream with too_big() set —
    // memory stays bounded by (2 * threads + 2) * 3 * max_chunk_out whatever the compression ratio; callers fall back to the
    // sequential decoder, which streams
    ParallelInflater(int fd, int threads, size_t chunk_bytes = 2u << 20, size_t max_chunk_out = (size_t)256 << 20)
        : fd_(fd), chunk_(std::max<size_t>(chunk_bytes, 65536)), max_out_(std::max<size_t>(max_chunk_out, 1u << 16)) {
        struct stat st;
        file_size_ = fstat(fd, &st) == 0 ? (uint64_t)st.st_size : 0;
        n_chunks_ = (size_t)((file_size_ + chunk_ - 1) / chunk_);
        if (!n_chunks_) n_chunks_ = 1;
        threads = std::max(1, threads);
        lookahead_ = 2 * (size_t)threads + 2;
        slots_.resize(n_chunks_);
        tail_.assign(WIN, 0);
        for (int t = 0; t < threads; t++) workers_.emplace_back([this] { work(); });
    }
    ~ParallelInflater() {
        { std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
        cv_.notify_all();
        for (auto &t : workers_) t.join();
    }

    ssize_t read(void *dst, size_t cap) {
        uint8_t *d = (uint8_t *)dst;
        size_t got = 0;
        while (got < cap) {
            if (pos_ == cur_len_) {
                if (failed_) return got ? (ssize_t)got : -1;
                if (done_) break;
                if (!advance()) failed_ = true;
                continue;
            }
            const size_t n = std::min(cap - got, cur_len_ - pos_);
            // the stretch is the resolved symbols followed by the plain bytes
            const size_t na = cur_->r->n_sym;
            size_t c = 0;
            if (pos_ < na) { c = std::min(n, na - pos_); memcpy(d + got, cur_->out.p + pos_, c); }
            if (c < n) memcpy(d + got + c, cur_->r->bytes.p + WIN + (pos_ + c - na), n - c);
            pos_ += n; got += n;
        }
        return (ssize_t)got;
    }
    const char *error() const { return err_.c_str(); }
    bool too_big() const { return too_big_; }
    bool trailing_garbage() const { return garbage_; }     // see fastgz::Inflater::trailing_garbage
    // how the stream was put together (tests and logs)
    uint64_t chunks_used() const { return used_; }
    uint64_t chunks_redone() const { return redone_; }
    bool gave_up() const { return unprofitable_; }
    uint64_t symbols_resolved() const { return resolved_; }

  private:
    using Result = pargz_detail::ChunkResult;
    struct Slot { std::unique_ptr<Result> r; bool ready = false; };
    // a stretch of the stream whose place is settled, waiting to be handed out
    struct Item {
        std::unique_ptr<Result> r;
        std::vector<uint8_t> win;          // the 32 KB in front of it (when it has symbols to resolve)
        uint64_t member_out_before = 0;    // bytes of the current member in front of it
        pargz_detail::RawBuf<uint8_t> out; // the symbols as bytes
        bool resolved = false, bad = false, oom = false;
    };

    void work() {
        pargz_detail::ChunkDecoder dec(fd_, file_size_, chunk_ + 65536, max_out_);
        for (;;) {
            size_t k = 0;
            std::shared_ptr<Item> job;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return stop_ || !resolve_q_.empty() || (!unprofitable_ && next_ < n_chunks_ && next_ < delivered_ + lookahead_); });
                if (stop_) return;
                if (!resolve_q_.empty()) { job = resolve_q_.front(); resolve_q_.pop_front(); }
                else k = next_++;
            }
            if (job) {
                {
                    std::lock_guard<std::mutex> lk(mu_);
                    if (!byte_pool_.empty()) { job->out = std::move(byte_pool_.back()); byte_pool_.pop_back(); }
                }
                resolve(*job);
                std::lock_guard<std::mutex> lk(mu_);
                job->resolved = true;
                cv_.notify_all();
                continue;
            }
            std::unique_ptr<Result> r(new Result());
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (!sym_pool_.empty()) { r->sym = std::move(sym_pool_.back()); sym_pool_.pop_back(); }
                if (!byte_pool_.empty()) { r->bytes = std::move(byte_pool_.back()); byte_pool_.pop_back(); }
            }
            try {
                const uint64_t lo = (uint64_t)k * chunk_ * 8, hi = (uint64_t)(k + 1) * chunk_ * 8;
                if (k == 0) dec.run_first(*r, hi);
                else dec.run_search(*r, lo, hi);
            } catch (const std::bad_alloc &) { r->failed = true; r->err = "out of memory"; }
            std::lock_guard<std::mutex> lk(mu_);
            slots_[k].r = std::move(r);
            slots_[k].ready = true;
            cv_.notify_all();
        }
    }
    // symbols → bytes.  Most symbols are plain bytes even in text that never gets rid of its markers, so 16 symbols at a time are
    // tested for a marker and packed; only groups with a marker go through the window look-up.
    static void resolve(Item &it) {
        const Result &r = *it.r;
        try { it.out.reserve(r.n_sym + 16); } catch (const std::bad_alloc &) { it.bad = true; it.oom = true; return; }
        const uint16_t *s = r.sym.p + WIN;
        const uint8_t *win = it.win.data();
        uint8_t *o = it.out.p;
        uint16_t lowest = 0xFFFF;
        size_t i = 0;
        const size_t n = r.n_sym;
#if defined(__x86_64__)
        i = resolve_sse(s, n, win, o, lowest);
#endif
        for (; i < n; i++) {
            const uint16_t v = s[i];
            if (v & 0x8000) { o[i] = win[v & 0x7FFF]; lowest = v < lowest ? v : lowest; }
            else o[i] = (uint8_t)v;
        }
        // a marker for window byte j stands for a copy from WIN - j bytes in front of the chunk: the member must be that old
        if (lowest != 0xFFFF && (uint64_t)(WIN - (lowest & 0x7FFF)) > it.member_out_before) it.bad = true;
    }
#if defined(__x86_64__)
    // SSE2 (part of x86-64): returns how many symbols were done (a multiple of 16)
    static size_t resolve_sse(const uint16_t *s, size_t n, const uint8_t *win, uint8_t *o, uint16_t &lowest) {
        size_t i = 0;
        for (; i + 16 <= n; i += 16) {
            const __m128i a = _mm_loadu_si128((const __m128i *)(s + i)), b = _mm_loadu_si128((const __m128i *)(s + i + 8));
            if (_mm_movemask_epi8(_mm_or_si128(a, b)) & 0xAAAA) {                   // a high bit set in some symbol
                for (size_t j = i; j < i + 16; j++) {
                    const uint16_t v = s[j];
                    if (v & 0x8000) { o[j] = win[v & 0x7FFF]; lowest = v < lowest ? v : lowest; }
                    else o[j] = (uint8_t)v;
                }
            } else
                _mm_storeu_si128((__m128i *)(o + i), _mm_packus_epi16(a, b));
        }
        return i;
    }
#endif
    // the arrays of a stretch that has been handed out (or dropped) go back to the workers
    void recycle(const std::shared_ptr<Item> &it) {
        if (!it) return;
        std::lock_guard<std::mutex> lk(mu_);
        if (it->r && it->r->sym.p && sym_pool_.size() < lookahead_) sym_pool_.push_back(std::move(it->r->sym));
        if (it->r && it->r->bytes.p && byte_pool_.size() < 2 * lookahead_) byte_pool_.push_back(std::move(it->r->bytes));
        if (it->out.p && byte_pool_.size() < 2 * lookahead_) byte_pool_.push_back(std::move(it->out));
    }
    bool ready_now(size_t k) { std::lock_guard<std::mutex> lk(mu_); return slots_[k].ready; }
    std::unique_ptr<Result> take(size_t k) {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return slots_[k].ready; });
        return std::move(slots_[k].r);
    }
    bool fail(const std::string &m) { err_ = m; return false; }

    // hands out the next stretch: settles the place of as many decoded chunks as are ready, then waits for the oldest one's bytes
    bool advance() {
        recycle(cur_);
        cur_.reset(); cur_len_ = 0; pos_ = 0;
        for (;;) {
            while (!settled_all_ && queue_.size() < lookahead_) {
                const bool must = queue_.empty();
                if (k_ < n_chunks_ && !must && !ready_now(k_)) break;
                if (!settle_next()) return false;
            }
            if (queue_.empty()) { done_ = true; return true; }
            std::shared_ptr<Item> it = queue_.front();
            queue_.pop_front();
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return it->resolved; });
                delivered_++;
                cv_.notify_all();
            }
            if (it->bad) return fail(it->oom ? "out of memory" : "invalid distance too far back");
            if (!check_members(*it)) return false;
            const size_t total = it->r->n_sym + it->r->n_bytes;
            if (it->r->failed) {                                           // what it decoded before it broke is handed out first
                too_big_ = it->r->too_big;
                if (!total) return fail(it->r->err);
                failed_ = true; err_ = it->r->err;
            }
            if (!total) continue;
            cur_ = it; cur_len_ = total;
            return true;
        }
    }
    // CRC-32 and ISIZE of every member that ends inside the stretch
    bool check_members(const Item &it) {
        const Result &r = *it.r;
        const size_t na = r.n_sym;
        auto crc_range = [&](uint64_t a, uint64_t b) {            // [a, b) of the stretch: symbols part, then bytes part
            if (a < na) { const uint64_t e = std::min<uint64_t>(b, na); crc_ = Crc32::update(crc_, it.out.p + a, (size_t)(e - a)); a = e; }
            if (a < b) crc_ = Crc32::update(crc_, r.bytes.p + WIN + (a - na), (size_t)(b - a));
        };
        uint64_t from = 0;
        for (const auto &m : r.members) {
            crc_range(from, m.out_off);
            member_out_ += m.out_off - from;
            if (crc_ != m.crc) return fail("gzip CRC-32 mismatch");
            if ((uint32_t)member_out_ != m.isize) return fail("gzip length mismatch");
            crc_ = 0; member_out_ = 0; from = m.out_off;
        }
        const uint64_t total = r.n_sym + r.n_bytes;
        crc_range(from, total);
        member_out_ += total - from;
        return true;
    }
    // decides what follows the settled data: chunk k_ if it starts at that very bit, otherwise the stretch decoded again in order
    bool settle_next() {
        if (at_eof_) { settled_all_ = true; return true; }
        if (k_ >= n_chunks_) {
            // the chunks are used up but the stream goes on (only when the last chunks had to be discarded)
            std::unique_ptr<Result> rec(new Result());
            if (!redo(*rec, ~(uint64_t)0)) return false;
            settle(std::move(rec));
            if (!at_eof_ && !queue_.back()->r->failed) return fail("truncated deflate stream");
            settled_all_ = true;
            return true;
        }
        // a stream whose block starts the workers keep missing (stored or fixed-Huffman blocks: incompressible or tiny data) is
        // decoded in order from here on; the chunks already under way are ignored
        if (!unprofitable_ && redone_ >= 4 && redone_ > used_) {
            std::lock_guard<std::mutex> lk(mu_);
            unprofitable_ = true;
        }
        std::unique_ptr<Result> r;
        if (unprofitable_ && k_ > 0) r.reset(new Result());
        else r = take(k_);
        const size_t k = k_++;
        if (k == 0) {
            if (r->not_gzip) return fail("not a gzip file");
            if (!r->found) return fail(r->failed ? r->err : "no gzip header");
            cur_bit_ = r->start_bit;
        }
        if (r->found && r->start_bit == cur_bit_) { used_++; settle(std::move(r)); return true; }
        const uint64_t stop = (r->found && r->start_bit > cur_bit_) ? r->start_bit : (uint64_t)(k + 1) * chunk_ * 8;
        if (cur_bit_ < stop) {
            std::unique_ptr<Result> rec(new Result());
            if (!redo(*rec, stop)) return false;
            const bool broke = rec->failed;
            settle(std::move(rec));
            if (broke) { settled_all_ = true; return true; }
        }
        if (!at_eof_ && r->found && r->start_bit == cur_bit_) { used_++; settle(std::move(r)); }
        else {                                                     // discarded: its slot in the look-ahead is free again
            std::lock_guard<std::mutex> lk(mu_);
            delivered_++;
            cv_.notify_all();
        }
        return true;
    }
    // decodes again, in order, from the end of the settled data up to the first block boundary at or behind stop
    bool redo(Result &rec, uint64_t stop) {
        redone_++;
        if (!redo_dec_) redo_dec_.reset(new pargz_detail::ChunkDecoder(fd_, file_size_, chunk_ + 65536, max_out_));   // kept: its buffers stay warm
        pargz_detail::ChunkDecoder &dec = *redo_dec_;
        {
            std::lock_guard<std::mutex> lk(mu_);
            if (!byte_pool_.empty()) { rec.bytes = std::move(byte_pool_.back()); byte_pool_.pop_back(); }
        }
        const size_t n_hist = (size_t)std::min<uint64_t>(s_member_out_, WIN);
        try {
            dec.run_from(rec, cur_bit_, stop, tail_.data() + WIN - n_hist, n_hist);
        } catch (const std::bad_alloc &) { return fail("out of memory"); }
        {   // a stretch decoded here takes a place in the look-ahead like a chunk
            std::lock_guard<std::mutex> lk(mu_);
            if (delivered_ > 0) delivered_--;
        }
        return true;
    }
    // a stretch that starts at cur_bit_: moves the end of the settled data behind it and queues it for delivery
    void settle(std::unique_ptr<Result> r) {
        std::shared_ptr<Item> it(new Item());
        const size_t ns = r->n_sym, nb = r->n_bytes, total = ns + nb;
        it->member_out_before = s_member_out_;
        if (ns) it->win = tail_;
        // the window behind it: its newest 32 KB, symbols looked up in the window in front of it
        if (total) {
            std::vector<uint8_t> nt(WIN);
            const size_t keep = total >= WIN ? 0 : WIN - total;       // bytes of the old window that stay
            if (keep) memcpy(nt.data(), tail_.data() + total, keep);
            const size_t first = total - (WIN - keep);                  // first output of the stretch that goes into the window
            for (size_t i = first; i < total; i++) {
                uint8_t b;
                if (i < ns) { const uint16_t v = r->sym.p[WIN + i]; b = (v & 0x8000) ? tail_[v & 0x7FFF] : (uint8_t)v; }
                else b = r->bytes.p[WIN + (i - ns)];
                nt[keep + (i - first)] = b;
            }
            tail_.swap(nt);
        }
        if (r->members.empty()) s_member_out_ += total;
        else s_member_out_ = total - r->members.back().out_off;
        if (!r->failed) cur_bit_ = r->end_bit;
        if (r->at_eof || r->failed) at_eof_ = true;                    // nothing can follow a broken stretch
        if (r->garbage) garbage_ = true;
        resolved_ += ns;
        it->r = std::move(r);
        std::lock_guard<std::mutex> lk(mu_);
        if (ns) resolve_q_.push_back(it);
        else it->resolved = true;
        queue_.push_back(it);
        cv_.notify_all();
    }

    int fd_;
    size_t chunk_, max_out_;
    bool too_big_ = false;
    std::unique_ptr<pargz_detail::ChunkDecoder> redo_dec_;
    bool unprofitable_ = false;              // guarded by mu_ (read by the workers), written by the thread that calls read()
    uint64_t file_size_ = 0;
    size_t n_chunks_ = 0, lookahead_ = 4;
    std::vector<Slot> slots_;
    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable cv_;
    size_t next_ = 0, delivered_ = 0;          // chunks handed to workers / chunks whose place in the look-ahead is free again
    std::deque<std::shared_ptr<Item>> resolve_q_;
    // the big arrays of finished stretches go round: a fresh 17 MB allocation per chunk costs its page faults again every time
    std::vector<pargz_detail::RawBuf<uint16_t>> sym_pool_;
    std::vector<pargz_detail::RawBuf<uint8_t>> byte_pool_;
    bool stop_ = false;
    // state of the thread that calls read()
    size_t k_ = 0;                             // next chunk to settle
    std::deque<std::shared_ptr<Item>> queue_;  // settled, not handed out yet
    std::vector<uint8_t> tail_;                // the 32 KB in front of cur_bit_, right-aligned
    uint64_t cur_bit_ = 0, s_member_out_ = 0;  // end of the settled data; bytes of the current member in front of it
    bool at_eof_ = false, settled_all_ = false, garbage_ = false;
    std::shared_ptr<Item> cur_;                // being handed out
    size_t cur_len_ = 0, pos_ = 0;
    uint64_t member_out_ = 0;                  // delivery side: bytes and CRC of the current member so far
    uint32_t crc_ = 0;
    bool done_ = false, failed_ = false;
    uint64_t used_ = 0, redone_ = 0, resolved_ = 0;
    std::string err_;
};

}  // namespace fastgz

// fastgz.h — streaming gzip (RFC 1952) / DEFLATE (RFC 1951) decoder for the read ingest of `kmcp-gpu search`
// (the reference reads FASTA/Q through xopen + pgzip readers, S:793-1000; once the search
// itself runs on the GPU the single-stream inflate of the input file is the end-to-end limiter).
//
// Written for throughput on FASTQ text: 64-bit bit buffer refilled once per symbol group, 11-bit (literal/length) and
// 8-bit (distance) first-level tables with second-level tables for longer codes, up to three literals per refill,
// word-wise match copies, carry-less-multiply CRC-32.  Same observable behaviour as zlib's gzread(): concatenated
// members are decoded back to back, input that does not start with the gzip magic is passed through unchanged,
// trailing garbage after a complete member is ignored, and CRC-32 / ISIZE of every member are verified.
//
// Host-only code (no CUDA).  Header-only; used by cli_search.cpp.  Tests: tests/test_fastgz.py (`kmcp-gpu gunzip`).
#pragma once
#include <stdint.h>
#include <string.h>
#include <sys/types.h>

#include <algorithm>
#include <functional>
#include <string>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace fastgz {

// ---------------------------------------------------------------------------------------------------------------------
// CRC-32 (IEEE 802.3, reflected, as in gzip).  update() takes and returns the conventional (inverted) value, like zlib.
// ---------------------------------------------------------------------------------------------------------------------
class Crc32 {
  public:
    static uint32_t update(uint32_t crc, const uint8_t *p, size_t n) {
        const Tables &t = tables();
        uint32_t s = ~crc;
#if defined(__x86_64__)
        if (n >= 128 && t.have_clmul) {
            const size_t body = n & ~(size_t)15;
            s = fold_clmul(t, s, p, body);
            p += body; n -= body;
        }
#endif
        s = bytes(t, s, p, n);
        return ~s;
    }
    // the table-driven routine alone (tests compare the two)
    static uint32_t update_portable(uint32_t crc, const uint8_t *p, size_t n) { return ~bytes(tables(), ~crc, p, n); }
    static bool accelerated() { return tables().have_clmul; }

  private:
    struct Tables {
        uint32_t t[8][256];
        uint64_t k512_lo, k512_hi, k128_lo, k128_hi;      // fold constants (see fold_clmul)
        bool have_clmul = false;
    };
    // x^n mod P in the ordinary (non-reflected) bit order: bit d = coefficient of x^d
    static uint32_t xpow_mod(unsigned n) {
        uint32_t r = 1;
        for (unsigned i = 0; i < n; i++) r = (r << 1) ^ ((r & 0x80000000u) ? 0x04C11DB7u : 0u);
        return r;
    }
    static uint64_t rev64(uint64_t v) {
        uint64_t r = 0;
        for (int i = 0; i < 64; i++) r |= ((v >> i) & 1ull) << (63 - i);
        return r;
    }
    static const Tables &tables() {
        static const Tables T = [] {
            Tables t;
            for (uint32_t i = 0; i < 256; i++) {
                uint32_t c = i;
                for (int k = 0; k < 8; k++) c = (c >> 1) ^ ((c & 1) ? 0xEDB88320u : 0u);
                t.t[0][i] = c;
            }
            for (uint32_t i = 0; i < 256; i++)
                for (int k = 1; k < 8; k++) t.t[k][i] = (t.t[k - 1][i] >> 8) ^ t.t[0][t.t[k - 1][i] & 0xFF];
            // A 64-bit lane holds a polynomial with bit i = coefficient of x^(63-i); a carry-less product of two lanes, read
            // as a 128-bit register in the same convention, is A(x)·B(x)·x.  Folding a register (lo lane = upper 64
            // coefficients) across a distance of D bits therefore multiplies lo by x^(D+63) and hi by x^(D-1), both mod P.
            t.k512_lo = rev64(xpow_mod(512 + 63)); t.k512_hi = rev64(xpow_mod(512 - 1));
            t.k128_lo = rev64(xpow_mod(128 + 63)); t.k128_hi = rev64(xpow_mod(128 - 1));
#if defined(__x86_64__)
            t.have_clmul = __builtin_cpu_supports("pclmul") && __builtin_cpu_supports("sse4.1");
            if (t.have_clmul) {      // self-check against the table routine; any doubt → tables only
                uint8_t buf[400];
                for (size_t i = 0; i < sizeof(buf); i++) buf[i] = (uint8_t)(i * 131u + 17u);
                for (size_t n : {128u, 144u, 256u, 400u})
                    if (fold_clmul(t, 0x1234567u, buf, n) != bytes(t, 0x1234567u, buf, n)) t.have_clmul = false;
            }
#endif
            return t;
        }();
        return T;
    }
    // state in, state out (not inverted); slicing-by-8
    static uint32_t bytes(const Tables &T, uint32_t s, const uint8_t *p, size_t n) {
        while (n >= 8) {
            uint64_t w;
            memcpy(&w, p, 8);
            w ^= s;
            s = T.t[7][w & 0xFF] ^ T.t[6][(w >> 8) & 0xFF] ^ T.t[5][(w >> 16) & 0xFF] ^ T.t[4][(w >> 24) & 0xFF] ^ T.t[3][(w >> 32) & 0xFF] ^
                T.t[2][(w >> 40) & 0xFF] ^ T.t[1][(w >> 48) & 0xFF] ^ T.t[0][w >> 56];
            p += 8; n -= 8;
        }
        while (n--) s = T.t[0][(s ^ *p++) & 0xFF] ^ (s >> 8);
        return s;
    }
#if defined(__x86_64__)
    // n is a multiple of 16 and >= 64.  Four 128-bit accumulators are folded across 512 bits per step, then into one, which
    // is finally reduced by running its 16 bytes through the table routine from state 0 (the register is congruent to the
    // message read so far, so its raw CRC is the CRC state).
    __attribute__((target("pclmul,sse4.1"))) static uint32_t fold_clmul(const Tables &T, uint32_t s, const uint8_t *p, size_t n) {
        const __m128i k512 = _mm_set_epi64x((long long)T.k512_hi, (long long)T.k512_lo);
        const __m128i k128 = _mm_set_epi64x((long long)T.k128_hi, (long long)T.k128_lo);
        __m128i x0 = _mm_loadu_si128((const __m128i *)p), x1 = _mm_loadu_si128((const __m128i *)(p + 16)),
                x2 = _mm_loadu_si128((const __m128i *)(p + 32)), x3 = _mm_loadu_si128((const __m128i *)(p + 48));
        x0 = _mm_xor_si128(x0, _mm_cvtsi32_si128((int)s));
        p += 64; n -= 64;
#define FASTGZ_FOLD(x, k, d) _mm_xor_si128(_mm_xor_si128(_mm_clmulepi64_si128(x, k, 0x00), _mm_clmulepi64_si128(x, k, 0x11)), d)
        while (n >= 64) {
            x0 = FASTGZ_FOLD(x0, k512, _mm_loadu_si128((const __m128i *)p));
            x1 = FASTGZ_FOLD(x1, k512, _mm_loadu_si128((const __m128i *)(p + 16)));
            x2 = FASTGZ_FOLD(x2, k512, _mm_loadu_si128((const __m128i *)(p + 32)));
            x3 = FASTGZ_FOLD(x3, k512, _mm_loadu_si128((const __m128i *)(p + 48)));
            p += 64; n -= 64;
        }
        x1 = FASTGZ_FOLD(x0, k128, x1);
        x2 = FASTGZ_FOLD(x1, k128, x2);
        x3 = FASTGZ_FOLD(x2, k128, x3);
        while (n >= 16) {
            x3 = FASTGZ_FOLD(x3, k128, _mm_loadu_si128((const __m128i *)p));
            p += 16; n -= 16;
        }
#undef FASTGZ_FOLD
        uint8_t last[16];
        _mm_storeu_si128((__m128i *)last, x3);
        return bytes(T, 0, last, 16);
    }
#endif
};

// ---------------------------------------------------------------------------------------------------------------------
// The decoder.  Pull interface: read() behaves like gzread().
// ---------------------------------------------------------------------------------------------------------------------
class Inflater {
  public:
    using ReadFn = std::function<ssize_t(void *, size_t)>;      // like read(2): > 0 bytes, 0 at end of input, < 0 on error

    explicit Inflater(ReadFn rd, bool verify_crc = true) : rd_(std::move(rd)), verify_(verify_crc) {
        ibuf_.resize(IN_CAP + PAD);
        obuf_.resize(HIST + OUT_CAP + SLACK);
        in_next_ = in_end_ = ibuf_.data();
        memset(ibuf_.data(), 0, PAD);
        out_base_ = obuf_.data() + HIST;
        out_limit_ = out_base_ + OUT_CAP;
        out_next_ = drain_ = crc_from_ = out_base_;
        win_start_ = out_base_;
        static const FixedTables fixed;
        fixed_ = &fixed;
        lt_dyn_.resize(LT_SIZE);
        dt_dyn_.resize(DT_SIZE);
    }

    // > 0: bytes stored at dst; 0: end of the stream; < 0: error (error() says which)
    ssize_t read(void *dst, size_t cap) {
        uint8_t *d = (uint8_t *)dst;
        size_t got = 0;
        while (got < cap) {
            if (state_ == ST_PLAIN && drain_ == out_next_ && !failed_) {
                // not gzip: the bytes go straight to the caller, past the window
                if (in_next_ < in_end_) {
                    const size_t n = std::min(cap - got, (size_t)(in_end_ - in_next_));
                    memcpy(d + got, in_next_, n);
                    in_next_ += n; got += n;
                    continue;
                }
                if (in_eof_) { if (io_error_) { fail("read error"); failed_ = true; continue; } state_ = ST_END; break; }
                const ssize_t r = rd_(d + got, cap - got);
                if (r < 0) { io_error_ = true; in_eof_ = true; continue; }
                if (r == 0) { in_eof_ = true; continue; }
                got += (size_t)r;
                if (got) break;                  // like read(2): hand over what there is
                continue;
            }
            if (drain_ == out_next_) {
                if (failed_) return got ? (ssize_t)got : -1;
                if (state_ == ST_END) break;
                if (out_next_ >= out_limit_) slide();
                if (!decode()) { failed_ = true; if (drain_ == out_next_) return got ? (ssize_t)got : -1; }
                if (drain_ == out_next_) continue;
            }
            const size_t n = std::min(cap - got, (size_t)(out_next_ - drain_));
            memcpy(d + got, drain_, n);
            drain_ += n; got += n;
        }
        return (ssize_t)got;
    }
    const char *error() const { return err_.c_str(); }
    bool is_gzip() const { return saw_gzip_; }
    // bytes that are not a gzip member followed the last member (ignored here; Go's gzip reader — the reference's xopen — fails with
    // "gzip: invalid header" on such a file)
    bool trailing_garbage() const { return garbage_; }
    uint64_t members() const { return members_; }

    // ---- table construction (also used by the chunk-parallel decoder in pargz.h) ------------------------------------------
    static constexpr int LBITS = 11, DBITS = 8;
    static constexpr size_t LT_SIZE = (1u << LBITS) + (1u << 15), DT_SIZE = (1u << DBITS) + (1u << 15);
    // Table entry.  bits 0-5: bits to take from the stream for this symbol, code AND extra bits together, so the bit buffer
    // is shifted once per symbol and the extra bits are cut out of a copy off the critical path; bits 8-11: code length alone;
    // bits 16-31: value (literal, length base, distance base).  E_LIT: a literal in bits 16-23, with E_LIT2 a second one in
    // bits 24-31 (both codes counted in bits 0-5).  E_EXC without E_SUB: value 0 = end of block, 1 = unused code.
    // E_EXC|E_SUB: bits 0-5 = first-level index bits to skip, bits 8-11 = index bits of the second level, value = its start.
    static constexpr uint32_t E_LIT = 0x8000, E_EXC = 0x4000, E_LIT2 = 0x2000, E_SUB = 0x1000;
    static constexpr uint32_t E_EOB = E_EXC | (0u << 16), E_BAD = E_EXC | (1u << 16);
    static uint32_t code_bits(uint32_t codelen, uint32_t extra) { return (codelen + extra) | (codelen << 8); }

    static uint32_t rev_bits(uint32_t code, int len) {
        uint32_t r = 0;
        for (int i = 0; i < len; i++) r |= ((code >> i) & 1u) << (len - 1 - i);
        return r;
    }
    // entry of a symbol whose code has `l` bits
    static uint32_t litlen_entry(int sym, uint32_t l) {
        static const uint16_t base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
        static const uint8_t extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
        if (sym < 256) return E_LIT | ((uint32_t)sym << 16) | code_bits(l, 0);
        if (sym == 256) return E_EOB | code_bits(l, 0);
        if (sym > 285) return E_BAD;
        return ((uint32_t)base[sym - 257] << 16) | code_bits(l, extra[sym - 257]);
    }
    static uint32_t dist_entry(int sym, uint32_t l) {
        static const uint16_t base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
        static const uint8_t extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
        if (sym > 29) return E_BAD;
        return ((uint32_t)base[sym] << 16) | code_bits(l, extra[sym]);
    }
    // canonical Huffman code → two-level lookup table.  false: over-subscribed set of lengths, or the table space ran out.
    static bool build_table(uint32_t *T, size_t cap, int P, const uint8_t *lens, int nsyms, bool litlen) {
        int count[16] = {0};
        for (int i = 0; i < nsyms; i++) count[lens[i]]++;
        int left = 1;
        for (int l = 1; l <= 15; l++) { left = left * 2 - count[l]; if (left < 0) return false; }
        uint32_t next[16], code = 0;
        count[0] = 0;                            // unused symbols take no code space
        for (int l = 1; l <= 15; l++) { code = (code + (uint32_t)count[l - 1]) << 1; next[l] = code; }
        const uint32_t PN = 1u << P, PM = PN - 1;
        for (uint32_t i = 0; i < PN; i++) T[i] = E_BAD;
        uint8_t submax[1u << LBITS];
        bool any_long = false;
        for (int l = P + 1; l <= 15; l++) any_long |= count[l] != 0;
        if (any_long) memset(submax, 0, PN);
        // codes in symbol order within each length (canonical): reversed for LSB-first lookup
        uint32_t rev[288];
        for (int s = 0; s < nsyms; s++) {
            const int l = lens[s];
            if (!l) continue;
            rev[s] = rev_bits(next[l]++, l);
            if (l > P) { uint8_t &m = submax[rev[s] & PM]; if (l - P > m) m = (uint8_t)(l - P); }
        }
        size_t used = PN;
        for (int s = 0; s < nsyms; s++) {
            const int l = lens[s];
            if (!l) continue;
            if (l <= P) {
                const uint32_t e = litlen ? litlen_entry(s, (uint32_t)l) : dist_entry(s, (uint32_t)l);
                for (uint32_t i = rev[s]; i < PN; i += 1u << l) T[i] = e;
            } else {
                const uint32_t pre = rev[s] & PM;
                const int sb = submax[pre];
                if (!(T[pre] & E_SUB)) {                                     // first long code with this prefix: open its second level
                    if (used + ((size_t)1 << sb) > cap) return false;
                    T[pre] = E_EXC | E_SUB | ((uint32_t)used << 16) | ((uint32_t)sb << 8) | (uint32_t)P;
                    for (size_t i = 0; i < ((size_t)1 << sb); i++) T[used + i] = E_BAD;
                    used += (size_t)1 << sb;
                }
                const uint32_t base = T[pre] >> 16;
                const uint32_t e = litlen ? litlen_entry(s, (uint32_t)(l - P)) : dist_entry(s, (uint32_t)(l - P));
                for (uint32_t i = rev[s] >> P; i < (1u << sb); i += 1u << (l - P)) T[base + i] = e;
            }
        }
        if (litlen) {
            // two literals per lookup where both codes fit into the first-level index (bases and qualities of FASTQ text have
            // 2-4 bit codes): the index bits behind the first code select the second symbol if its code is short enough
            uint32_t one[1u << LBITS];
            memcpy(one, T, sizeof(one));
            for (uint32_t i = 0; i < PN; i++) {
                const uint32_t a = one[i];
                if (!(a & E_LIT)) continue;
                const uint32_t la = a & 63, b = one[i >> la];
                if ((b & E_LIT) && la + (b & 63) <= (uint32_t)P) T[i] = (a & 0x00FF0000u) | ((b & 0x00FF0000u) << 8) | E_LIT | E_LIT2 | (la + (b & 63));
            }
        }
        return true;
    }
    struct FixedTables {
        std::vector<uint32_t> lt, dt;
        FixedTables() : lt(LT_SIZE), dt(DT_SIZE) {
            uint8_t l[288], d[32];
            for (int i = 0; i < 144; i++) l[i] = 8;
            for (int i = 144; i < 256; i++) l[i] = 9;
            for (int i = 256; i < 280; i++) l[i] = 7;
            for (int i = 280; i < 288; i++) l[i] = 8;
            for (int i = 0; i < 32; i++) d[i] = 5;
            build_table(lt.data(), LT_SIZE, LBITS, l, 288, true);
            build_table(dt.data(), DT_SIZE, DBITS, d, 32, false);
        }
    };

  private:
    static constexpr size_t IN_CAP = 1u << 20, PAD = 64, HIST = 32768, OUT_CAP = 1u << 20, SLACK = 512;
    enum State { ST_MEMBER, ST_BLOCK, ST_HUFF, ST_STORED, ST_TRAILER, ST_PLAIN, ST_END };

    bool fail(const char *msg) { err_ = msg; return false; }

    // ---- input -----------------------------------------------------------------------------------------------------
    // keeps [in_next, in_end) and reads more behind it; false when nothing was added (end of input or a read error)
    bool fill() {
        if (in_eof_) return false;
        uint8_t *b = ibuf_.data();
        const size_t keep = (size_t)(in_end_ - in_next_);
        if (keep && in_next_ != b) memmove(b, in_next_, keep);
        in_next_ = b; in_end_ = b + keep;
        size_t added = 0;
        while ((size_t)(in_end_ - b) < IN_CAP) {
            const ssize_t r = rd_(const_cast<uint8_t *>(in_end_), IN_CAP - (size_t)(in_end_ - b));
            if (r < 0) { io_error_ = true; in_eof_ = true; break; }
            if (r == 0) { in_eof_ = true; break; }
            in_end_ += r; added += (size_t)r;
            if ((size_t)(in_end_ - b) >= IN_CAP / 2) break;
        }
        memset(const_cast<uint8_t *>(in_end_), 0, PAD);
        return added != 0;
    }
    bool ensure(size_t n) {                     // at least n bytes buffered, or everything up to the end of the input
        while ((size_t)(in_end_ - in_next_) < n && !in_eof_) fill();
        return (size_t)(in_end_ - in_next_) >= n;
    }
    int get_byte() {                            // bit buffer must be empty
        if (in_next_ >= in_end_ && !fill()) return -1;
        return *in_next_++;
    }
    // drops the bits up to the next byte boundary and gives the whole bytes still in the bit buffer back to the input
    bool align_to_byte() {
        bitbuf_ >>= (bitcnt_ & 7); bitcnt_ -= (bitcnt_ & 7);
        in_next_ -= bitcnt_ >> 3;
        bitbuf_ = 0; bitcnt_ = 0;
        return in_next_ <= in_end_;             // false: bits past the end of the input were used
    }

    // ---- output window ---------------------------------------------------------------------------------------------
    void crc_upto_here() {
        if (verify_ && out_next_ > crc_from_) crc_ = Crc32::update(crc_, crc_from_, (size_t)(out_next_ - crc_from_));
        member_out_ += (uint64_t)(out_next_ - crc_from_);
        crc_from_ = out_next_;
    }
    void slide() {                              // everything has been handed out: keep the last 32 KB as match history
        crc_upto_here();
        const size_t keep = std::min<size_t>(HIST, (size_t)(out_next_ - win_start_));
        memmove(out_base_ - keep, out_next_ - keep, keep);
        win_start_ = out_base_ - keep;
        out_next_ = drain_ = crc_from_ = out_base_;
    }

    // ---- the state machine: runs until the window is full, the stream ends, or an error -------------------------------
    bool decode() {
        for (;;) {
            switch (state_) {
            case ST_MEMBER: {
                if (!ensure(2)) {
                    if (io_error_) return fail("read error");
                    if (in_end_ == in_next_) { state_ = ST_END; return true; }                  // clean end (or an empty file)
                    if (members_) { garbage_ = true; state_ = ST_END; return true; }            // one stray byte after the last member
                    state_ = ST_PLAIN; break;
                }
                if (in_next_[0] != 0x1f || in_next_[1] != 0x8b) {
                    if (members_) { garbage_ = true; state_ = ST_END; return true; }            // trailing garbage is ignored, as gzread does (callers may ask)
                    state_ = ST_PLAIN; break;
                }
                saw_gzip_ = true;
                in_next_ += 2;
                const int cm = get_byte(), flg = get_byte();
                if (cm != 8 || flg < 0 || (flg & 0xE0)) return fail(cm < 0 || flg < 0 ? "truncated gzip header" : "unsupported gzip header");
                for (int i = 0; i < 6; i++) if (get_byte() < 0) return fail("truncated gzip header");   // mtime, xfl, os
                if (flg & 4) {
                    const int a = get_byte(), b = get_byte();
                    if (a < 0 || b < 0) return fail("truncated gzip header");
                    for (int n = a | (b << 8); n > 0; n--) if (get_byte() < 0) return fail("truncated gzip header");
                }
                for (int bit : {8, 16})
                    if (flg & bit) { int c; do { c = get_byte(); if (c < 0) return fail("truncated gzip header"); } while (c); }
                if (flg & 2) { if (get_byte() < 0 || get_byte() < 0) return fail("truncated gzip header"); }
                crc_upto_here();                 // output of the previous member is accounted for
                crc_ = 0; member_out_ = 0;
                win_start_ = out_next_;          // a member cannot refer back into the previous one
                bitbuf_ = 0; bitcnt_ = 0;
                state_ = ST_BLOCK;
                break;
            }
            case ST_BLOCK:
                if (!block_header()) return false;
                break;
            case ST_HUFF: {
                const int r = huff();
                if (r < 0) return false;
                if (r == 0) return true;          // window full
                state_ = final_ ? ST_TRAILER : ST_BLOCK;
                break;
            }
            case ST_STORED: {
                if (stored_left_ == 0) { state_ = final_ ? ST_TRAILER : ST_BLOCK; break; }
                if (out_next_ >= out_limit_) return true;
                if (in_next_ >= in_end_ && !fill()) return fail(io_error_ ? "read error" : "truncated stored block");
                size_t n = std::min<size_t>(stored_left_, (size_t)(in_end_ - in_next_));
                n = std::min<size_t>(n, (size_t)(out_limit_ - out_next_));
                memcpy(out_next_, in_next_, n);
                out_next_ += n; in_next_ += n; stored_left_ -= (uint32_t)n;
                break;
            }
            case ST_TRAILER: {
                if (!align_to_byte()) return fail("truncated deflate stream");
                uint32_t v[2] = {0, 0};
                for (int i = 0; i < 8; i++) {
                    const int c = get_byte();
                    if (c < 0) return fail(io_error_ ? "read error" : "truncated gzip trailer");
                    v[i >> 2] |= (uint32_t)c << (8 * (i & 3));
                }
                crc_upto_here();
                if (verify_ && v[0] != crc_) return fail("gzip CRC-32 mismatch");
                if (v[1] != (uint32_t)member_out_) return fail("gzip length mismatch");
                members_++;
                state_ = ST_MEMBER;
                break;
            }
            case ST_PLAIN: {
                if (out_next_ >= out_limit_) return true;
                if (in_next_ >= in_end_ && !fill()) {
                    if (io_error_) return fail("read error");
                    state_ = ST_END; return true;
                }
                const size_t n = std::min<size_t>((size_t)(in_end_ - in_next_), (size_t)(out_limit_ - out_next_));
                memcpy(out_next_, in_next_, n);
                out_next_ += n; in_next_ += n;
                if (out_next_ >= out_limit_) return true;
                break;
            }
            case ST_END:
                return true;
            }
        }
    }

#define FASTGZ_REFILL()                                         \
    do {                                                        \
        uint64_t w_;                                            \
        memcpy(&w_, in, 8);                                     \
        bb |= w_ << bc;                                         \
        in += (63 - bc) >> 3;                                   \
        bc |= 56;                                               \
    } while (0)

    // BFINAL/BTYPE and, for dynamic blocks, the two code-length sets
    bool block_header() {
        ensure(1536);                            // the longest dynamic header is well under this
        const uint8_t *in = in_next_;
        uint64_t bb = bitbuf_;
        unsigned bc = bitcnt_;
        auto bits = [&](unsigned n) -> uint32_t {   // n <= 16
            FASTGZ_REFILL();
            const uint32_t v = (uint32_t)(bb & ((1u << n) - 1));
            bb >>= n; bc -= n;
            return v;
        };
        auto store = [&]() { in_next_ = in; bitbuf_ = bb; bitcnt_ = bc; };
        auto overrun = [&]() { return in - (bc >> 3) > in_end_; };
        final_ = bits(1) != 0;
        const uint32_t type = bits(2);
        if (type == 0) {
            store();
            if (!align_to_byte()) return fail("truncated deflate stream");
            uint8_t h[4];
            for (int i = 0; i < 4; i++) { const int c = get_byte(); if (c < 0) return fail("truncated stored block"); h[i] = (uint8_t)c; }
            const uint32_t len = h[0] | (h[1] << 8), nlen = h[2] | (h[3] << 8);
            if ((len ^ nlen) != 0xFFFF) return fail("invalid stored block lengths");
            stored_left_ = len;
            state_ = ST_STORED;
            return true;
        }
        if (type == 1) {
            lt_ = fixed_->lt.data(); dt_ = fixed_->dt.data();
        } else if (type == 2) {
            const uint32_t hlit = bits(5) + 257, hdist = bits(5) + 1, hclen = bits(4) + 4;
            if (hlit > 286 || hdist > 30) return fail("too many length or distance symbols");
            static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
            uint8_t pl[19] = {0};
            for (uint32_t i = 0; i < hclen; i++) pl[order[i]] = (uint8_t)bits(3);
            // the code-length code: at most 7 bits, one flat table
            uint16_t pt[128];
            {
                int count[8] = {0};
                for (int i = 0; i < 19; i++) count[pl[i]]++;
                int left = 1;
                for (int l = 1; l <= 7; l++) { left = left * 2 - count[l]; if (left < 0) return fail("invalid code lengths set"); }
                uint32_t next[8], code = 0;
                count[0] = 0;
                for (int l = 1; l <= 7; l++) { code = (code + (uint32_t)count[l - 1]) << 1; next[l] = code; }
                for (int i = 0; i < 128; i++) pt[i] = 0xFFFF;
                for (int s = 0; s < 19; s++) {
                    const int l = pl[s];
                    if (!l) continue;
                    for (uint32_t i = rev_bits(next[l]++, l); i < 128; i += 1u << l) pt[i] = (uint16_t)(s | (l << 8));
                }
            }
            uint8_t lens[286 + 30 + 138];
            const uint32_t total = hlit + hdist;
            uint32_t n = 0;
            while (n < total) {
                FASTGZ_REFILL();
                const uint16_t e = pt[bb & 127];
                if (e == 0xFFFF) return fail("invalid code lengths set");
                bb >>= (e >> 8); bc -= (e >> 8);
                const int sym = e & 0xFF;
                if (sym < 16) { lens[n++] = (uint8_t)sym; continue; }
                uint32_t rep; uint8_t val = 0;
                if (sym == 16) {
                    if (!n) return fail("invalid bit length repeat");
                    val = lens[n - 1]; rep = 3 + (uint32_t)(bb & 3); bb >>= 2; bc -= 2;
                } else if (sym == 17) { rep = 3 + (uint32_t)(bb & 7); bb >>= 3; bc -= 3; }
                else { rep = 11 + (uint32_t)(bb & 127); bb >>= 7; bc -= 7; }
                if (n + rep > total) return fail("invalid bit length repeat");
                memset(lens + n, val, rep);
                n += rep;
            }
            if (overrun()) return fail("truncated deflate stream");
            if (lens[256] == 0) return fail("invalid code -- missing end-of-block");
            if (!build_table(lt_dyn_.data(), LT_SIZE, LBITS, lens, (int)hlit, true)) return fail("invalid literal/lengths set");
            if (!build_table(dt_dyn_.data(), DT_SIZE, DBITS, lens + hlit, (int)hdist, false)) return fail("invalid distances set");
            lt_ = lt_dyn_.data(); dt_ = dt_dyn_.data();
        } else
            return fail("invalid block type");
        if (overrun()) return fail("truncated deflate stream");
        store();
        state_ = ST_HUFF;
        return true;
    }

    // one compressed block's symbols.  1: end of block, 0: the output window is full (call again), -1: error
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__) && !defined(__SANITIZE_THREAD__)      // (ifunc resolvers run before TSan is up)
    __attribute__((target_clones("bmi2", "default")))      // shrx/bzhi where the CPU has them (one variable shift per symbol)
#endif
    int huff() {
        const uint8_t *in = in_next_;
        uint64_t bb = bitbuf_;
        unsigned bc = bitcnt_;
        uint8_t *out = out_next_;
        const uint32_t *const LT = lt_, *const DT = dt_;
        uint8_t *const out_limit = out_limit_;
        const uint8_t *win_start = win_start_;
        const uint8_t *in_safe = in_eof_ ? in_end_ : (in_end_ - in_next_ > 16 ? in_end_ - 16 : in_next_);
        int ret = -1;
        const uint32_t LM = (1u << LBITS) - 1, DM = (1u << DBITS) - 1;
        // `e` is always the table entry of the next symbol, looked up before the previous symbol's bytes were copied: after a
        // refill all 64 bits of the buffer are stream bits, a symbol takes at most 48, so the 11 index bits behind it are valid
        FASTGZ_REFILL();
        uint32_t e = LT[bb & LM];
        for (;;) {
            if (__builtin_expect(in >= in_safe || out >= out_limit, 0)) {
                if (out >= out_limit) { ret = 0; break; }
                if (!in_eof_) {
                    in_next_ = in;
                    fill();
                    in = in_next_;
                    in_safe = in_eof_ ? in_end_ : (in_end_ - in_next_ > 16 ? in_end_ - 16 : in_next_);
                    if (io_error_) { fail("read error"); break; }
                    continue;
                }
                // the tail of the input: zero padding follows; stop as soon as bits past the real end were used
                if (in - (bc >> 3) > in_end_) { fail("truncated deflate stream"); break; }
            }
            if (e & E_LIT) {
                // up to four lookups (eight literals, at most 44 bits) on one refill; both bytes are always stored, the
                // second one is overwritten when the entry holds a single literal
#define FASTGZ_LITERALS()                                   \
    do {                                                    \
        const uint16_t two_ = (uint16_t)(e >> 16);          \
        memcpy(out, &two_, 2);                              \
        out += 1 + ((e >> 13) & 1);                         \
        bb >>= (e & 63); bc -= (e & 63);                    \
        e = LT[bb & LM];                                    \
    } while (0)
                FASTGZ_LITERALS();
                if (e & E_LIT) {
                    FASTGZ_LITERALS();
                    if (e & E_LIT) {
                        FASTGZ_LITERALS();
                        if (e & E_LIT) FASTGZ_LITERALS();
                    }
                }
#undef FASTGZ_LITERALS
                FASTGZ_REFILL();
                if (e & E_LIT) continue;
            }
            if (__builtin_expect(e & E_EXC, 0)) {
                if (e & E_SUB) {
                    bb >>= LBITS; bc -= LBITS;
                    e = LT[(e >> 16) + ((uint32_t)bb & ((1u << ((e >> 8) & 15)) - 1))];
                    if (e & E_LIT) {
                        bb >>= (e & 63); bc -= (e & 63); *out++ = (uint8_t)(e >> 16);
                        FASTGZ_REFILL();
                        e = LT[bb & LM];
                        continue;
                    }
                }
                if (e & E_EXC) {
                    if ((e >> 16) == 0) { bb >>= (e & 63); bc -= (e & 63); ret = 1; break; }
                    fail("invalid literal/length code"); break;
                }
            }
            // length: one shift of the bit buffer for code + extra bits; the extra bits come out of the copy
            uint64_t saved = bb;
            unsigned tot = e & 63;
            bb >>= tot; bc -= tot;
            const uint32_t len = (e >> 16) + (((uint32_t)saved & ((1u << tot) - 1)) >> ((e >> 8) & 15));
            e = DT[bb & DM];
            if (__builtin_expect(e & E_EXC, 0)) {
                if (e & E_SUB) {
                    bb >>= DBITS; bc -= DBITS;
                    e = DT[(e >> 16) + ((uint32_t)bb & ((1u << ((e >> 8) & 15)) - 1))];
                }
                if (e & E_EXC) { fail("invalid distance code"); break; }
            }
            saved = bb;
            tot = e & 63;
            bb >>= tot; bc -= tot;
            const uint32_t dist = (e >> 16) + (((uint32_t)saved & ((1u << tot) - 1)) >> ((e >> 8) & 15));
            e = LT[bb & LM];                                   // the next symbol's entry travels while the bytes are copied
            if (__builtin_expect((size_t)(out - win_start) < dist, 0)) { fail("invalid distance too far back"); break; }
            const uint8_t *src = out - dist;
            uint8_t *dst = out;
            out += len;
            if (dist >= 16) {
                // 16 bytes per step; the overshoot lands in the slack behind the window
                do {
                    uint64_t a, b;
                    memcpy(&a, src, 8); memcpy(&b, src + 8, 8);
                    memcpy(dst, &a, 8); memcpy(dst + 8, &b, 8);
                    src += 16; dst += 16;
                } while (dst < out);
            } else if (dist == 1) {
                const uint64_t v = 0x0101010101010101ull * src[0];
                do { memcpy(dst, &v, 8); dst += 8; } while (dst < out);
            } else if (dist >= 8) {
                do { uint64_t a; memcpy(&a, src, 8); memcpy(dst, &a, 8); src += 8; dst += 8; } while (dst < out);
            } else {
                do { *dst++ = *src++; } while (dst < out);
            }
            FASTGZ_REFILL();
        }
        in_next_ = in; bitbuf_ = bb; bitcnt_ = bc; out_next_ = out;
        return ret;
    }
#undef FASTGZ_REFILL

    ReadFn rd_;
    bool verify_;
    std::vector<uint8_t> ibuf_, obuf_;
    const uint8_t *in_next_, *in_end_;
    bool in_eof_ = false, io_error_ = false, failed_ = false, saw_gzip_ = false, final_ = false, garbage_ = false;
    uint64_t bitbuf_ = 0;
    unsigned bitcnt_ = 0;
    uint8_t *out_base_, *out_limit_, *out_next_, *drain_, *crc_from_;
    const uint8_t *win_start_;
    State state_ = ST_MEMBER;
    uint32_t stored_left_ = 0, crc_ = 0;
    uint64_t member_out_ = 0, members_ = 0;
    const FixedTables *fixed_ = nullptr;
    std::vector<uint32_t> lt_dyn_, dt_dyn_;
    const uint32_t *lt_ = nullptr, *dt_ = nullptr;
    std::string err_;
};

}  // namespace fastgz

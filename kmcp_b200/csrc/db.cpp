// db.cpp — on-disk formats of a kmcp database, host side.
//   `.uniki` block header/rows : reference kmcp/cmd/index/serialization.go (X:) 153-304 (write), 383-593 (read)
//   `__db.yml`                 : reference kmcp/cmd/util-db-info.go:46-130
// plus the small exact-arithmetic helpers the search path needs on the host (fastmod reciprocal, FPR).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>

#include <errno.h>
#include <fcntl.h>
#include <unistd.h>

#include <atomic>
#include <thread>

#include "common.h"

namespace kmcpg {

bool pread_parallel(int fd, void *dst, size_t bytes, uint64_t off, int threads) {
    if (!bytes) return true;
    const size_t SLICE = 4u << 20;                            // every stream takes the next 4 MB slice until none is left
    const size_t n_slices = (bytes + SLICE - 1) / SLICE;
    threads = (int)std::max<size_t>(1, std::min<size_t>((size_t)std::max(1, threads), n_slices));
    std::atomic<size_t> next{0};
    std::atomic<bool> ok{true};
    auto run = [&]() {
        for (;;) {
            const size_t s = next.fetch_add(1);
            if (s >= n_slices || !ok.load()) return;
            const size_t a = s * SLICE, e = std::min(bytes, a + SLICE);
            size_t got = a;
            while (got < e) {
                const ssize_t r = pread(fd, (char *)dst + got, e - got, (off_t)(off + got));
                if (r < 0 && errno == EINTR) continue;
                if (r <= 0) { ok.store(false); return; }       // error, or the file ends inside the range
                got += (size_t)r;
            }
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < threads; t++) th.emplace_back(run);
    run();
    for (auto &t : th) t.join();
    return ok.load();
}

namespace {

struct FileReader {
    FILE *f = nullptr;
    explicit FileReader(const std::string &p) { f = fopen(p.c_str(), "rb"); }
    ~FileReader() { if (f) fclose(f); }
    bool ok() const { return f != nullptr; }
    bool read(void *dst, size_t n) { return fread(dst, 1, n, f) == n; }
    bool u32(uint32_t &v) {
        unsigned char b[4];
        if (!read(b, 4)) return false;
        v = (uint32_t(b[0]) << 24) | (uint32_t(b[1]) << 16) | (uint32_t(b[2]) << 8) | b[3];
        return true;
    }
    bool u64(uint64_t &v) {
        unsigned char b[8];
        if (!read(b, 8)) return false;
        v = 0;
        for (int i = 0; i < 8; i++) v = (v << 8) | b[i];
        return true;
    }
};

std::string strip(const std::string &s) {
    size_t a = 0, b = s.size();
    while (a < b && isspace((unsigned char)s[a])) a++;
    while (b > a && isspace((unsigned char)s[b - 1])) b--;
    std::string r = s.substr(a, b - a);
    if (r.size() >= 2 && ((r.front() == '"' && r.back() == '"') || (r.front() == '\'' && r.back() == '\''))) r = r.substr(1, r.size() - 2);
    return r;
}

bool to_bool(const std::string &v) { return v == "true" || v == "True" || v == "yes" || v == "on"; }

void put32(std::string &s, uint32_t v) { for (int i = 3; i >= 0; i--) s.push_back(char((v >> (8 * i)) & 0xFF)); }
void put64(std::string &s, uint64_t v) { for (int i = 7; i >= 0; i--) s.push_back(char((v >> (8 * i)) & 0xFF)); }

}  // namespace

int read_block_header(const std::string &path, BlockMeta &m, std::string &err) {
    FileReader r(path);
    if (!r.ok()) { err = "cannot open index file: " + path; return KMCPG_EIO; }
    m = BlockMeta();
    m.path = path;
    unsigned char b[8];
    if (!r.read(b, 8)) { err = "kmcp: truncated index file: " + path; return KMCPG_EIO; }
    if (memcmp(b, ".kmcpidx", 8) != 0) { err = "kmcp: invalid index format: " + path; return KMCPG_EFORMAT; }    // X:393-407
    if (!r.read(b, 4)) { err = "kmcp: truncated index file: " + path; return KMCPG_EIO; }
    if (b[0] != 4) { err = "kmcp: version mismatch: " + path; return KMCPG_EFORMAT; }                            // X:415-417
    m.k = b[1];
    m.canonical = (b[2] & 1) != 0;
    m.compact = (b[2] & 2) != 0;
    m.num_hashes = b[3];
    uint32_t n = 0, cnt = 0, len = 0;
    bool ok = r.u64(m.num_sigs) && r.u32(n);
    if (!ok) { err = "kmcp: truncated index file: " + path; return KMCPG_EIO; }
    m.n_names = (int)n;
    m.row_bytes = (int)((n + 7) / 8);
    m.names.resize(n); m.indices.assign(n, 0); m.gsizes.assign(n, 0); m.sizes.assign(n, 0);
    for (uint32_t i = 0; i < n; i++) {
        if (!r.u32(len)) { err = "kmcp: truncated index file: " + path; return KMCPG_EIO; }
        std::string s(len, '\0');
        if (len && !r.read(&s[0], len)) { err = "kmcp: truncated index file: " + path; return KMCPG_EIO; }
        size_t nl = s.find('\n');                        // names joined by '\n'; search prints Target[0] (S:520)
        m.names[i] = nl == std::string::npos ? s : s.substr(0, nl);
    }
    uint32_t groups = 0;
    if (!r.u32(groups)) { err = "kmcp: truncated index file: " + path; return KMCPG_EIO; }
    for (uint32_t i = 0; i < groups; i++) {
        if (!r.u32(cnt)) { err = "kmcp: truncated index file: " + path; return KMCPG_EIO; }
        for (uint32_t j = 0; j < cnt; j++) {
            uint64_t v;
            if (!r.u64(v)) { err = "kmcp: truncated index file: " + path; return KMCPG_EIO; }
            if (j == 0 && i < n) m.gsizes[i] = v;
        }
    }
    if (!r.u32(groups)) { err = "kmcp: truncated index file: " + path; return KMCPG_EIO; }
    for (uint32_t i = 0; i < groups; i++) {
        if (!r.u32(cnt)) { err = "kmcp: truncated index file: " + path; return KMCPG_EIO; }
        for (uint32_t j = 0; j < cnt; j++) {
            uint32_t v;
            if (!r.u32(v)) { err = "kmcp: truncated index file: " + path; return KMCPG_EIO; }
            if (j == 0 && i < n) m.indices[i] = v;
        }
    }
    for (uint32_t i = 0; i < n; i++)
        if (!r.u64(m.sizes[i])) { err = "kmcp: truncated index file: " + path; return KMCPG_EIO; }
    m.data_offset = (uint64_t)ftell(r.f);
    // the row payload must be complete (X:307-349 ErrTruncateIndexFile)
    if (fseek(r.f, 0, SEEK_END) != 0) { err = "seek failed: " + path; return KMCPG_EIO; }
    uint64_t fsize = (uint64_t)ftell(r.f);
    if (fsize < m.data_offset + m.num_sigs * (uint64_t)m.row_bytes) { err = "kmcp: truncated index file: " + path; return KMCPG_EIO; }
    return KMCPG_OK;
}

int read_db_meta(const std::string &dir, DbMeta &db, std::string &err) {
    db = DbMeta();
    db.dir = dir;
    std::string yml = dir + "/__db.yml";
    std::ifstream in(yml);
    if (!in) { err = "fail to open kmcp database info file: " + yml; return KMCPG_EIO; }
    std::string line, list_key;
    int k_single = 0;
    auto add_item = [&](const std::string &key, const std::string &v) {
        if (key == "ks") db.ks.push_back(atoi(v.c_str()));
        else if (key == "files") db.files.push_back(v);
    };
    while (std::getline(in, line)) {
        std::string t = strip(line);
        if (t.empty() || t[0] == '#') continue;
        if (t[0] == '-') { add_item(list_key, strip(t.substr(1))); continue; }
        size_t c = t.find(':');
        if (c == std::string::npos) continue;
        std::string key = strip(t.substr(0, c)), v = strip(t.substr(c + 1));
        list_key = key;
        if (!v.empty() && v[0] == '[') {
            std::string inner = v.substr(1, v.find(']') == std::string::npos ? std::string::npos : v.find(']') - 1);
            std::stringstream ss(inner);
            std::string item;
            while (std::getline(ss, item, ',')) { item = strip(item); if (!item.empty()) add_item(key, item); }
            continue;
        }
        if (key == "version") db.version = atoi(v.c_str());
        else if (key == "unikiVersion") db.index_version = atoi(v.c_str());
        else if (key == "k") k_single = atoi(v.c_str());
        else if (key == "canonical") db.canonical = to_bool(v);
        else if (key == "scaled") db.scaled = to_bool(v);
        else if (key == "scale") db.scale = (uint32_t)strtoul(v.c_str(), nullptr, 10);
        else if (key == "minimizer") db.minimizer = to_bool(v);
        else if (key == "minimizer-w") db.minimizer_w = (uint32_t)strtoul(v.c_str(), nullptr, 10);
        else if (key == "syncmer") db.syncmer = to_bool(v);
        else if (key == "syncmer-s") db.syncmer_s = (uint32_t)strtoul(v.c_str(), nullptr, 10);
        else if (key == "hashes") db.num_hashes = atoi(v.c_str());
        else if (key == "fpr") db.fpr = strtod(v.c_str(), nullptr);
    }
    if (db.version != 4) { err = "kmcp/index: version mismatch"; return KMCPG_EFORMAT; }          // util-db-info.go:118-120
    if (db.ks.empty()) db.ks.push_back(k_single);                                               // :124-126
    std::sort(db.ks.begin(), db.ks.end(), [](int a, int b) { return a > b; });                   // U:752-759
    if (db.files.empty()) { err = "no index files"; return KMCPG_EFORMAT; }                      // U:654-656
    db.blocks.resize(db.files.size());
    for (size_t i = 0; i < db.files.size(); i++) {
        int rc = read_block_header(dir + "/" + db.files[i], db.blocks[i], err);
        if (rc) return rc;
        BlockMeta &b = db.blocks[i];
        // U:689-695 / X:84-94 Compatible
        if (b.k != db.ks.front() || b.canonical != db.canonical || b.num_hashes != db.num_hashes) {
            err = "index files not compatible";
            return KMCPG_EFORMAT;
        }
        b.target_base = db.n_targets;
        db.n_targets += b.n_names;
    }
    return KMCPG_OK;
}

int write_block_file(const std::string &path, const BlockMeta &m, const uint8_t *rows, std::string &err) {
    std::string h;
    h.append(".kmcpidx", 8);
    h.push_back(4); h.push_back((char)m.k); h.push_back((char)((m.canonical ? 1 : 0) | (m.compact ? 2 : 0))); h.push_back((char)m.num_hashes);
    put64(h, m.num_sigs);
    put32(h, (uint32_t)m.n_names);
    for (auto &nm : m.names) { put32(h, (uint32_t)nm.size() + 1); h += nm; h.push_back('\n'); }
    put32(h, (uint32_t)m.n_names);
    for (auto g : m.gsizes) { put32(h, 1); put64(h, g); }
    put32(h, (uint32_t)m.n_names);
    for (auto ix : m.indices) { put32(h, 1); put32(h, ix); }
    for (auto s : m.sizes) put64(h, s);
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) { err = "cannot create " + path; return KMCPG_EIO; }
    bool ok = fwrite(h.data(), 1, h.size(), f) == h.size();
    size_t bytes = (size_t)m.num_sigs * (size_t)m.row_bytes;
    ok = ok && fwrite(rows, 1, bytes, f) == bytes;
    fclose(f);
    if (!ok) { err = "short write: " + path; return KMCPG_EIO; }
    return KMCPG_OK;
}

// ---- exact x % d by 128-bit reciprocal (Lemire, Kaser, Kurz 2019: with a 2N-bit magic the N-bit remainder is exact)
FastMod make_fastmod(uint64_t d) {
    FastMod f;
    f.d = d ? d : 1;
    f.b64 = ~0ull / f.d;
    f.b32 = f.d <= 0xFFFFFFFFull ? (uint32_t)(0xFFFFFFFFull / f.d) : 0;
    if (f.d == 1) { f.m_hi = 0; f.m_lo = 0; return f; }       // x % 1 == 0: lowbits = 0 → result 0
    unsigned __int128 M = ~(unsigned __int128)0;
    M /= f.d;
    M += 1;
    f.m_hi = (uint64_t)(M >> 64);
    f.m_lo = (uint64_t)M;
    return f;
}

uint64_t fastmod_host(uint64_t a, const FastMod &f) {
    unsigned __int128 M = ((unsigned __int128)f.m_hi << 64) | f.m_lo;
    unsigned __int128 low = M * a;
    unsigned __int128 bottom = ((low & 0xFFFFFFFFFFFFFFFFULL) * (unsigned __int128)f.d) >> 64;
    unsigned __int128 top = (low >> 64) * (unsigned __int128)f.d;
    return (uint64_t)((bottom + top) >> 64);
}

// ---- Go-exact floating point --------------------------------------------------------------------------
// math.Pow (go/src/math/pow.go): integral exponents run a frexp-normalised square-and-multiply loop.
double go_pow(double x, double y) {
    if (y == 0 || x == 1) return 1;
    if (y == 1) return x;
    if (std::isnan(x) || std::isnan(y)) return NAN;
    if (x == 0) return y < 0 ? INFINITY : 0.0;
    if (std::isinf(y)) {
        if (x == -1) return 1;
        return ((std::fabs(x) < 1) == (y > 0)) ? 0.0 : INFINITY;
    }
    if (std::isinf(x)) return y < 0 ? 0.0 : INFINITY;
    if (y == 0.5) return std::sqrt(x);
    if (y == -0.5) return 1 / std::sqrt(x);
    double yi, yf = std::modf(std::fabs(y), &yi);
    if (yf != 0 && x < 0) return NAN;
    if (yi >= 9.223372036854775808e18) {
        if (x == -1) return 1;
        return ((std::fabs(x) < 1) == (y > 0)) ? 0.0 : INFINITY;
    }
    double a1 = 1.0;
    long ae = 0;
    if (yf != 0) {
        if (yf > 0.5) { yf--; yi++; }
        a1 = std::exp(yf * std::log(x));
    }
    int e0;
    double x1 = std::frexp(x, &e0);
    long xe = e0;
    for (int64_t i = (int64_t)yi; i != 0; i >>= 1) {
        if (xe < -(1L << 12) || (1L << 12) < xe) { ae += xe; break; }
        if (i & 1) { a1 *= x1; ae += xe; }
        x1 *= x1;
        xe <<= 1;
        if (x1 < .5) { x1 += x1; xe--; }
    }
    if (y < 0) { a1 = 1 / a1; ae = -ae; }
    ae = std::max(-200000L, std::min(200000L, ae));
    return std::ldexp(a1, (int)ae);
}

namespace {
// F:54-71 BinomialCoeff: big.Float with 53-bit mantissa and unbounded exponent → (mantissa, exponent) pair
double binomial(int n, int k) {
    if (k > n - k) k = n - k;
    double m = 1.0;
    long e = 0;
    int t;
    for (int i = 0; i < k; i++) {
        m = std::frexp(m * (double)(n - i), &t); e += t;
        m = std::frexp(m / (double)(i + 1), &t); e += t;
    }
    if (e > 1024) return INFINITY;
    return std::ldexp(m, (int)e);
}
}  // namespace

double query_fpr(int n, int c, double p) {
    double r = 1;
    for (int i = 0; i <= c; i++) {
        double coeff = binomial(n, i);
        if (coeff > 1.7976931348623157e308) return 0;                           // F:38-40
        r -= coeff * go_pow(p, (double)i) * go_pow(1 - p, (double)(n - i));     // F:42
        if (r < 0) return 0;                                                    // F:44-46
    }
    return r;
}

uint64_t calc_signature_size(uint64_t ne, int h, double fpr) {
    double ratio = (double)(-h) / std::log(1 - go_pow(fpr, 1 / (double)h));
    return (uint64_t)std::ceil((double)ne * ratio);
}

}  // namespace kmcpg

// test hook (host only, tests/test_abi.py): the chunk reader of kmcpg_open_db on any file
extern "C" int kmcpg_internal_pread_selftest(const char *path, uint64_t off, uint64_t bytes, int threads, void *dst) {
    const int fd = ::open(path, O_RDONLY);
    if (fd < 0) return KMCPG_EIO;
    const bool ok = kmcpg::pread_parallel(fd, dst, (size_t)bytes, off, threads);
    ::close(fd);
    return ok ? KMCPG_OK : KMCPG_EIO;
}

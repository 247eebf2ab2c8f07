/*
 * kmcp_oracle.c — CPU ORACLE (test infrastructure; see kmcp_oracle.h for the rules and citations).
 * Compile WITHOUT floating-point contraction (-ffp-contract=off): the FPR column must reproduce Go's
 * plain IEEE-double arithmetic digit for digit.
 */
#define _GNU_SOURCE
#include "kmcp_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------------
 * ntHash1 — restated from will-rowe/nthash v0.4.0 (go.mod:47, not in /root/reference), which is a Go
 * port of bcgsc ntHash 1.x.  Forward seed = seedTab[base]; reverse seed = seedTab[base & 7].
 * Validated end to end by golden vectors G1-G4.
 * ---------------------------------------------------------------------------------------------- */
#define SEED_A 0x3c8bfbb395c60474ULL
#define SEED_C 0x3193c18562a02b4cULL
#define SEED_G 0x20323ed082572324ULL
#define SEED_T 0x295549f54be24456ULL

static uint64_t SEED_TAB[256];
static int g_init = 0;

static void ko_init(void) {
    if (g_init) return;
    memset(SEED_TAB, 0, sizeof(SEED_TAB));
    /* entries 0..7 serve the reverse strand lookup seedTab[b & 7]: {N,T,N,G,A,A,N,C} */
    SEED_TAB[1] = SEED_T; SEED_TAB[3] = SEED_G; SEED_TAB[4] = SEED_A; SEED_TAB[5] = SEED_A; SEED_TAB[7] = SEED_C;
    SEED_TAB['A'] = SEED_TAB['a'] = SEED_A;
    SEED_TAB['C'] = SEED_TAB['c'] = SEED_C;
    SEED_TAB['G'] = SEED_TAB['g'] = SEED_G;
    SEED_TAB['T'] = SEED_TAB['t'] = SEED_T;
    SEED_TAB['U'] = SEED_TAB['u'] = SEED_T;   /* parity unpinned (no reference artefact has U) */
    g_init = 1;
}

static inline uint64_t rol64(uint64_t x, unsigned r) { r &= 63; return r ? (x << r) | (x >> (64 - r)) : x; }
static inline uint64_t ror64(uint64_t x, unsigned r) { r &= 63; return r ? (x >> r) | (x << (64 - r)) : x; }

int64_t ko_nthash_all(const uint8_t *s, int64_t len, int k, int canonical, uint64_t *out) {
    ko_init();
    if (k < 1 || k > 64 || len < k) return 0;
    uint64_t fh = 0, rh = 0;
    for (int j = 0; j < k; j++) {
        fh ^= rol64(SEED_TAB[s[j]], (unsigned)(k - 1 - j));
        rh ^= rol64(SEED_TAB[s[j] & 7], (unsigned)j);
    }
    int64_t n = len - k + 1;
    out[0] = canonical ? (fh < rh ? fh : rh) : fh;
    for (int64_t i = 1; i < n; i++) {
        uint8_t cout = s[i - 1], cin = s[i + k - 1];
        fh = rol64(fh, 1) ^ rol64(SEED_TAB[cout], (unsigned)k) ^ SEED_TAB[cin];
        rh = ror64(rh, 1) ^ ror64(SEED_TAB[cout & 7], 1) ^ rol64(SEED_TAB[cin & 7], (unsigned)(k - 1));
        out[i] = canonical ? (fh < rh ? fh : rh) : fh;
    }
    return n;
}

static uint64_t max_hash_for(const ko_sketch_params *p) {
    if (!p->scaled) return ~0ULL;
    /* U:1040-1043: uint64(float64(^uint64(0)) / float64(scale)); float64(2^64-1) rounds to 2^64 */
    double v = 18446744073709551616.0 / (double)p->scale;
    if (v >= 18446744073709551616.0) return ~0ULL; /* Go's conversion of an out-of-range float is implementation-defined; scale=1 keeps all */
    return (uint64_t)v;
}

/* leftmost argmin over a[lo..hi) */
static inline int64_t argmin_left(const uint64_t *a, int64_t lo, int64_t hi) {
    int64_t m = lo;
    for (int64_t i = lo + 1; i < hi; i++) if (a[i] < a[m]) m = i;
    return m;
}

int64_t ko_generate_kmers(const uint8_t *seq, int64_t len, const ko_sketch_params *p, uint64_t *out) {
    ko_init();
    int k = p->k;
    if (len < k) return 0;                                  /* sketches.ErrShortSeq → no k-mers (U:1059-1062) */
    uint64_t maxh = max_hash_for(p);
    int64_t nk = len - k + 1, n = 0;
    uint64_t *ck = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)nk);
    ko_nthash_all(seq, len, k, p->canonical, ck);
    if (p->syncmer) {
        /* SURVEY A.4 (bio/sketches NextSyncmer, windowed closed syncmer; pinned by G4) */
        int s = (int)p->syncmer_s;
        int64_t L = 2LL * k - s - 1, W = 2LL * (k - s), kms = k - s;
        if (s >= 1 && s < k && len >= L) {
            int64_t ns = len - s + 1;
            uint64_t *cs = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)ns);
            ko_nthash_all(seq, len, s, p->canonical, cs);
            int64_t prev = -1;
            for (int64_t idx = 0; idx + L <= len; idx++) {
                int64_t t = argmin_left(cs, idx, idx + W) - idx;
                int64_t pos = t < kms ? idx + t : idx + t - kms;
                if (pos == prev) continue;
                prev = pos;
                uint64_t code = ck[pos];
                if (p->scaled && code > maxh) continue;
                if (code > 0) out[n++] = code;
            }
            free(cs);
        }
    } else if (p->minimizer) {
        /* SURVEY A.5 (bio/sketches NextMinimizer) — PARITY UNPINNED */
        int64_t w = (int64_t)p->minimizer_w;
        if (w < 1) w = 1;
        if (nk >= w) {
            int64_t prev = -1;
            for (int64_t i = w - 1; i < nk; i++) {
                int64_t m = argmin_left(ck, i - w + 1, i + 1);
                if (m == prev) continue;
                prev = m;
                uint64_t code = ck[m];
                if (p->scaled && code > maxh) continue;
                if (code > 0) out[n++] = code;
            }
        }
    } else {
        for (int64_t i = 0; i < nk; i++) {
            uint64_t code = ck[i];
            if (p->scaled && code > maxh) continue;          /* U:1097-1099 */
            if (code > 0) out[n++] = code;                   /* U:1100-1102 */
        }
    }
    free(ck);
    return n;
}

static int cmp_u64(const void *a, const void *b) {
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}

int64_t ko_dedup(uint64_t *codes, int64_t n, int64_t thr) {
    if (n <= thr) return n;                                   /* strict > (U:874) */
    qsort(codes, (size_t)n, sizeof(uint64_t), cmp_u64);
    int64_t j = 1;
    for (int64_t i = 1; i < n; i++) if (codes[i] != codes[i - 1]) codes[j++] = codes[i];
    return j;
}

void ko_hash_values(uint64_t code, int h, uint64_t *out) {
    if (h == 1) { out[0] = code; return; }
    uint32_t a = (uint32_t)(code >> 32), b = (uint32_t)code;    /* H:61-63 */
    for (uint32_t i = 0; i < (uint32_t)h; i++) out[i] = (uint64_t)(uint32_t)(a + b * i);  /* H:137-139, uint32 wraparound */
}

/* ------------------------------------------------------------------------------------------------
 * FPR (F:32-50, 54-71, 140-193).  Go's math.Pow for integral y is the frexp/square-and-multiply loop
 * below (go/src/math/pow.go); BinomialCoeff runs in big.Float with 53-bit mantissa and unbounded
 * exponent — emulated with a (mantissa, exponent) pair so that large n never overflows mid-way.
 * ---------------------------------------------------------------------------------------------- */
double ko_go_pow(double x, double y) {
    if (y == 0 || x == 1) return 1;
    if (y == 1) return x;
    if (isnan(x) || isnan(y)) return NAN;
    if (x == 0) {
        if (y < 0) return INFINITY;
        return 0;
    }
    if (isinf(y)) {
        if (x == -1) return 1;
        if ((fabs(x) < 1) == (y > 0)) return 0;
        return INFINITY;
    }
    if (isinf(x)) { if (y < 0) return 0; return INFINITY; }
    if (y == 0.5) return sqrt(x);
    if (y == -0.5) return 1 / sqrt(x);
    double yi, yf = modf(fabs(y), &yi);
    if (yf != 0 && x < 0) return NAN;
    if (yi >= 9.223372036854775808e18) {
        if (x == -1) return 1;
        if ((fabs(x) < 1) == (y > 0)) return 0;
        return INFINITY;
    }
    double a1 = 1.0; long ae = 0;
    if (yf != 0) {
        if (yf > 0.5) { yf--; yi++; }
        a1 = exp(yf * log(x));
    }
    int xe_i; double x1 = frexp(x, &xe_i); long xe = xe_i;
    for (int64_t i = (int64_t)yi; i != 0; i >>= 1) {
        if (xe < -(1L << 12) || (1L << 12) < xe) { ae += xe; break; }   /* catastrophic overflow */
        if (i & 1) { a1 *= x1; ae += xe; }
        x1 *= x1; xe <<= 1;
        if (x1 < .5) { x1 += x1; xe--; }
    }
    if (y < 0) { a1 = 1 / a1; ae = -ae; }
    if (ae > 100000) ae = 100000;
    if (ae < -100000) ae = -100000;
    return ldexp(a1, (int)ae);
}

/* C(n,k) as Go's big.Float(prec 53) loop; returns value or +inf when it exceeds MaxFloat64 */
static double binomial_coeff(int n, int k) {
    if (k > n - k) k = n - k;
    double m = 1.0; long e = 0;                       /* value = m * 2^e, m in [0.5,1) or 1.0 initially */
    for (int i = 0; i < k; i++) {
        int t;
        m = m * (double)(n - i); m = frexp(m, &t); e += t;   /* Mul: rounded to 53 bits, exponent free */
        m = m / (double)(i + 1); m = frexp(m, &t); e += t;   /* Quo */
    }
    if (e > 1024) return INFINITY;                    /* res.Float64() → +Inf → coeff > MaxFloat64 */
    return ldexp(m, (int)e);
}

double ko_query_fpr(int n, int c, double p) {
    double r = 1;
    for (int i = 0; i <= c; i++) {
        double coeff = binomial_coeff(n, i);
        if (coeff > 1.79769313486231570814527423731704356798070e+308) return 0;
        r -= coeff * ko_go_pow(p, (double)i) * ko_go_pow(1 - p, (double)(n - i));
        if (r < 0) return 0;
    }
    return r;
}

uint64_t ko_calc_signature_size(uint64_t ne, int h, double fpr) {
    double ratio = (double)(-h) / log(1 - ko_go_pow(fpr, 1 / (double)h));
    return (uint64_t)ceil((double)ne * ratio);
}

/* ------------------------------------------------------------------------------------------------
 * Database loading
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int k, canonical, num_hashes;
    uint64_t num_sigs;
    int32_t n_names, row_bytes;
    char **names; uint32_t *indices; uint64_t *gsizes; uint64_t *sizes;
    uint8_t *rows;    /* num_sigs * row_bytes */
    int64_t target_base;
} ko_block;

struct ko_db {
    ko_db_info info;
    ko_block *blocks;
};

static uint64_t be64(const uint8_t *p) { uint64_t v = 0; for (int i = 0; i < 8; i++) v = (v << 8) | p[i]; return v; }
static uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

#define FAIL(...) do { if (err) snprintf(err, (size_t)errlen, __VA_ARGS__); goto fail; } while (0)

/* X:383-593 */
static int load_block(const char *path, ko_block *b, char *err, int errlen) {
    FILE *f = fopen(path, "rb");
    uint8_t buf[16];
    memset(b, 0, sizeof(*b));
    if (!f) { if (err) snprintf(err, (size_t)errlen, "cannot open %s", path); return -1; }
    if (fread(buf, 1, 8, f) != 8 || memcmp(buf, ".kmcpidx", 8)) FAIL("kmcp: invalid index format: %s", path);
    if (fread(buf, 1, 4, f) != 4) FAIL("truncated: %s", path);
    if (buf[0] != 4) FAIL("kmcp: version mismatch: %s", path);
    b->k = buf[1]; b->canonical = buf[2] & 1; b->num_hashes = buf[3];
    if (fread(buf, 1, 8, f) != 8) FAIL("truncated: %s", path);
    b->num_sigs = be64(buf);
    if (fread(buf, 1, 4, f) != 4) FAIL("truncated: %s", path);
    uint32_t n = be32(buf);
    b->n_names = (int32_t)n; b->row_bytes = (int32_t)((n + 7) / 8);
    b->names = (char **)calloc(n ? n : 1, sizeof(char *));
    b->indices = (uint32_t *)calloc(n ? n : 1, 4);
    b->gsizes = (uint64_t *)calloc(n ? n : 1, 8);
    b->sizes = (uint64_t *)calloc(n ? n : 1, 8);
    for (uint32_t i = 0; i < n; i++) {
        if (fread(buf, 1, 4, f) != 4) FAIL("truncated: %s", path);
        uint32_t l = be32(buf);
        char *s = (char *)malloc(l + 1);
        if (l && fread(s, 1, l, f) != l) { free(s); FAIL("truncated: %s", path); }
        s[l] = 0;
        char *nl = strchr(s, '\n'); if (nl) *nl = 0;      /* Target[0] (S:520) */
        b->names[i] = s;
    }
    if (fread(buf, 1, 4, f) != 4) FAIL("truncated: %s", path);
    uint32_t ng = be32(buf);
    for (uint32_t i = 0; i < ng; i++) {
        if (fread(buf, 1, 4, f) != 4) FAIL("truncated: %s", path);
        uint32_t c = be32(buf);
        for (uint32_t j = 0; j < c; j++) {
            if (fread(buf, 1, 8, f) != 8) FAIL("truncated: %s", path);
            if (j == 0 && i < n) b->gsizes[i] = be64(buf);
        }
    }
    if (fread(buf, 1, 4, f) != 4) FAIL("truncated: %s", path);
    uint32_t ni = be32(buf);
    for (uint32_t i = 0; i < ni; i++) {
        if (fread(buf, 1, 4, f) != 4) FAIL("truncated: %s", path);
        uint32_t c = be32(buf);
        for (uint32_t j = 0; j < c; j++) {
            if (fread(buf, 1, 4, f) != 4) FAIL("truncated: %s", path);
            if (j == 0 && i < n) b->indices[i] = be32(buf);
        }
    }
    for (uint32_t i = 0; i < n; i++) {
        if (fread(buf, 1, 8, f) != 8) FAIL("truncated: %s", path);
        b->sizes[i] = be64(buf);
    }
    size_t bytes = (size_t)b->num_sigs * (size_t)b->row_bytes;
    b->rows = (uint8_t *)malloc(bytes ? bytes : 1);
    if (fread(b->rows, 1, bytes, f) != bytes) FAIL("kmcp: truncated index file: %s", path);
    fclose(f);
    return 0;
fail:
    fclose(f);
    return -1;
}

static char *trim(char *s) {
    while (*s && isspace((unsigned char)*s)) s++;
    char *e = s + strlen(s);
    while (e > s && isspace((unsigned char)e[-1])) *--e = 0;
    if (*s == '"' || *s == '\'') { s++; size_t l = strlen(s); if (l && (s[l - 1] == '"' || s[l - 1] == '\'')) s[l - 1] = 0; }
    return s;
}
static int yaml_bool(const char *v) { return !strcmp(v, "true") || !strcmp(v, "True") || !strcmp(v, "yes"); }

ko_db *ko_db_open(const char *dir, char *err, int errlen) {
    ko_init();
    char path[4096];
    snprintf(path, sizeof(path), "%s/__db.yml", dir);
    FILE *f = fopen(path, "r");
    if (!f) { if (err) snprintf(err, (size_t)errlen, "fail to open kmcp database info file: %s", path); return NULL; }
    ko_db *db = (ko_db *)calloc(1, sizeof(ko_db));
    char **files = NULL; int nfiles = 0, version = -1, kk = 0;
    char line[8192], list_key[64] = "";
    while (fgets(line, sizeof(line), f)) {
        char *s = line;
        char *t = trim(s);
        if (!*t || *t == '#') continue;
        if (*t == '-') {                                   /* sequence item of the last key */
            char *v = trim(t + 1);
            if (!strcmp(list_key, "ks") && db->info.n_ks < 8) db->info.ks[db->info.n_ks++] = atoi(v);
            else if (!strcmp(list_key, "files")) { files = (char **)realloc(files, sizeof(char *) * (size_t)(nfiles + 1)); files[nfiles++] = strdup(v); }
            continue;
        }
        char *colon = strchr(t, ':');
        if (!colon) continue;
        *colon = 0;
        char *key = trim(t), *v = trim(colon + 1);
        snprintf(list_key, sizeof(list_key), "%s", key);
        if (*v == '[') {                                   /* flow sequence */
            char *q = v + 1; char *tok;
            while ((tok = strsep(&q, ",]")) != NULL) {
                tok = trim(tok); if (!*tok) continue;
                if (!strcmp(key, "ks") && db->info.n_ks < 8) db->info.ks[db->info.n_ks++] = atoi(tok);
                else if (!strcmp(key, "files")) { files = (char **)realloc(files, sizeof(char *) * (size_t)(nfiles + 1)); files[nfiles++] = strdup(tok); }
            }
            continue;
        }
        if (!strcmp(key, "version")) version = atoi(v);
        else if (!strcmp(key, "k")) kk = atoi(v);
        else if (!strcmp(key, "canonical")) db->info.canonical = yaml_bool(v);
        else if (!strcmp(key, "scaled")) db->info.scaled = yaml_bool(v);
        else if (!strcmp(key, "scale")) db->info.scale = (uint32_t)strtoul(v, NULL, 10);
        else if (!strcmp(key, "minimizer")) db->info.minimizer = yaml_bool(v);
        else if (!strcmp(key, "minimizer-w")) db->info.minimizer_w = (uint32_t)strtoul(v, NULL, 10);
        else if (!strcmp(key, "syncmer")) db->info.syncmer = yaml_bool(v);
        else if (!strcmp(key, "syncmer-s")) db->info.syncmer_s = (uint32_t)strtoul(v, NULL, 10);
        else if (!strcmp(key, "hashes")) db->info.num_hashes = atoi(v);
        else if (!strcmp(key, "fpr")) db->info.fpr = strtod(v, NULL);
    }
    fclose(f);
    if (version != 4) { if (err) snprintf(err, (size_t)errlen, "kmcp/index: version mismatch"); goto bad; }
    if (db->info.n_ks == 0) { db->info.ks[0] = kk; db->info.n_ks = 1; }
    /* U:752-759: ks sorted descending */
    for (int i = 0; i < db->info.n_ks; i++) for (int j = i + 1; j < db->info.n_ks; j++)
        if (db->info.ks[j] > db->info.ks[i]) { int t = db->info.ks[i]; db->info.ks[i] = db->info.ks[j]; db->info.ks[j] = t; }
    if (nfiles == 0) { if (err) snprintf(err, (size_t)errlen, "no index files"); goto bad; }
    db->blocks = (ko_block *)calloc((size_t)nfiles, sizeof(ko_block));
    db->info.n_blocks = nfiles;
    for (int i = 0; i < nfiles; i++) {
        snprintf(path, sizeof(path), "%s/%s", dir, files[i]);
        if (load_block(path, &db->blocks[i], err, errlen)) { db->info.n_blocks = i; goto bad; }
        ko_block *b = &db->blocks[i];
        /* U:689-695, 731 */
        if (b->k != db->info.ks[0] || b->canonical != db->info.canonical || b->num_hashes != db->info.num_hashes) {
            if (err) snprintf(err, (size_t)errlen, "index files not compatible");
            db->info.n_blocks = i + 1; goto bad;
        }
        b->target_base = db->info.n_targets;
        db->info.n_targets += b->n_names;
        db->info.sum_row_bytes += b->row_bytes;
        db->info.total_bytes += (int64_t)b->num_sigs * b->row_bytes;
    }
    for (int i = 0; i < nfiles; i++) free(files[i]);
    free(files);
    return db;
bad:
    for (int i = 0; i < nfiles; i++) free(files[i]);
    free(files);
    ko_db_close(db);
    return NULL;
}

void ko_db_close(ko_db *db) {
    if (!db) return;
    for (int i = 0; i < db->info.n_blocks; i++) {
        ko_block *b = &db->blocks[i];
        if (b->names) for (int j = 0; j < b->n_names; j++) free(b->names[j]);
        free(b->names); free(b->indices); free(b->gsizes); free(b->sizes); free(b->rows);
    }
    free(db->blocks);
    free(db);
}

void ko_db_get_info(const ko_db *db, ko_db_info *out) { *out = db->info; }

int ko_db_target(const ko_db *db, int64_t g, ko_target *out) {
    for (int i = 0; i < db->info.n_blocks; i++) {
        const ko_block *b = &db->blocks[i];
        if (g >= b->target_base && g < b->target_base + b->n_names) {
            int c = (int)(g - b->target_base);
            out->name = b->names[c]; out->index = b->indices[c]; out->genome_size = b->gsizes[c];
            out->n_kmers = b->sizes[c]; out->block = i; out->col = c;
            return 0;
        }
    }
    return -1;
}

int ko_db_block(const ko_db *db, int bi, uint64_t *num_sigs, int32_t *row_bytes, int32_t *n_names, const uint8_t **rows) {
    if (bi < 0 || bi >= db->info.n_blocks) return -1;
    const ko_block *b = &db->blocks[bi];
    *num_sigs = b->num_sigs; *row_bytes = b->row_bytes; *n_names = b->n_names; *rows = b->rows;
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Block probe.  algo 0: per-bit counting, the simplest statement of U:6613-7408.
 * algo 1: the reference's shape — up to 64 (AND-ed) rows are buffered, then for every byte column the
 * 64 bytes are gathered (the "transpose", U:6824-6966) and fed to a positional popcount (pospop.Count8).
 * Both produce counts[target]; bit (7-j) of byte i ↔ target 8i+j (I:1157, U:7415-7733).
 * ---------------------------------------------------------------------------------------------- */
#define POSPOP_BUF 64   /* U:1164 */

#ifdef __AVX2__
#include <immintrin.h>
/* positional popcount of one 64-byte column buffer, the AVX2 form pospop.Count8 takes for short inputs: the most significant bit of
 * every byte is collected by vpmovmskb and counted, then the bytes are shifted left by one (vpaddb) for the next position */
static inline void count8(uint32_t *cnt /* 8 targets of this column, cnt[j] ↔ bit 7-j */, const uint8_t *buf /* 64 bytes */, int n) {
    uint8_t pad[POSPOP_BUF];
    if (n < POSPOP_BUF) { memset(pad, 0, sizeof(pad)); memcpy(pad, buf, (size_t)n); buf = pad; }
    __m256i a = _mm256_loadu_si256((const __m256i *)buf), b = _mm256_loadu_si256((const __m256i *)(buf + 32));
    for (int j = 0; j < 8; j++) {                 /* j = 0 ↔ bit 7 */
        cnt[j] += (uint32_t)(__builtin_popcount((unsigned)_mm256_movemask_epi8(a)) + __builtin_popcount((unsigned)_mm256_movemask_epi8(b)));
        a = _mm256_add_epi8(a, a);
        b = _mm256_add_epi8(b, b);
    }
}
#else
static inline void count8(uint32_t *cnt /* 8 targets of this column, cnt[j] ↔ bit 7-j */, const uint8_t *buf, int n) {
    /* positional popcount over n<=64 bytes, SWAR on 64-bit words */
    uint64_t w[8];
    memset(w, 0, sizeof(w));
    memcpy(w, buf, (size_t)n);
    for (int b = 0; b < 8; b++) {
        uint64_t m = 0x0101010101010101ULL << b;
        int c = 0;
        for (int i = 0; i < 8; i++) c += __builtin_popcountll(w[i] & m);
        cnt[7 - b] += (uint32_t)c;
    }
}
#endif

static void probe_block(const ko_block *b, const uint64_t *codes, int64_t n, int algo, uint32_t *counts /* row_bytes*8 */, uint8_t *scratch /* 64*row_bytes + row_bytes */) {
    int rb = b->row_bytes, h = b->num_hashes;
    memset(counts, 0, sizeof(uint32_t) * (size_t)rb * 8);
    uint64_t hv[8];
    if (algo == 0) {
        uint8_t *acc = scratch;
        for (int64_t q = 0; q < n; q++) {
            ko_hash_values(codes[q], h, hv);
            const uint8_t *r0 = b->rows + (size_t)(hv[0] % b->num_sigs) * (size_t)rb;   /* U:6811 fastdiv.Mod == % */
            memcpy(acc, r0, (size_t)rb);
            for (int i = 1; i < h; i++) {
                const uint8_t *ri = b->rows + (size_t)(hv[i] % b->num_sigs) * (size_t)rb;
                for (int x = 0; x < rb; x++) acc[x] &= ri[x];                             /* U:6639-6645 */
            }
            for (int x = 0; x < rb; x++) {
                uint8_t v = acc[x];
                while (v) { int bit = __builtin_ctz(v); counts[x * 8 + (7 - bit)]++; v &= (uint8_t)(v - 1); }
            }
        }
        return;
    }
    const uint8_t *ptr[POSPOP_BUF];
    uint8_t *anded = scratch;                    /* 64 * rb, only used when h>1 */
    uint8_t col[POSPOP_BUF];
    int nbuf = 0;
    for (int64_t q = 0; q <= n; q++) {
        if (q < n) {
            ko_hash_values(codes[q], h, hv);
            const uint8_t *r0 = b->rows + (size_t)(hv[0] % b->num_sigs) * (size_t)rb;
            if (h == 1) ptr[nbuf] = r0;           /* zero-copy slice (U:6813-6816) */
            else {
                uint8_t *dst = anded + (size_t)nbuf * (size_t)rb;
                const uint8_t *r1 = b->rows + (size_t)(hv[1] % b->num_sigs) * (size_t)rb;
                for (int x = 0; x < rb; x++) dst[x] = r0[x] & r1[x];
                for (int i = 2; i < h; i++) {
                    const uint8_t *ri = b->rows + (size_t)(hv[i] % b->num_sigs) * (size_t)rb;
                    for (int x = 0; x < rb; x++) dst[x] &= ri[x];
                }
                ptr[nbuf] = dst;
            }
            nbuf++;
        }
        if (nbuf == POSPOP_BUF || (q == n && nbuf > 0)) {
            for (int x = 0; x < rb; x++) {
                for (int r = 0; r < nbuf; r++) col[r] = ptr[r][x];     /* the byte-column gather */
                count8(counts + (size_t)x * 8, col, nbuf);
            }
            nbuf = 0;
        }
    }
}

int ko_count_codes(const ko_db *db, const uint64_t *codes, int64_t n, uint32_t *counts) {
    for (int i = 0; i < db->info.n_blocks; i++) {
        const ko_block *b = &db->blocks[i];
        uint32_t *c = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)b->row_bytes * 8);
        uint8_t *scr = (uint8_t *)malloc((size_t)b->row_bytes * 65 + 64);
        probe_block(b, codes, n, 0, c, scr);
        memcpy(counts + b->target_base, c, sizeof(uint32_t) * (size_t)b->n_names);
        free(c); free(scr);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * search
 * ---------------------------------------------------------------------------------------------- */
void ko_default_opts(ko_search_opts *o) {
    memset(o, 0, sizeof(*o));
    o->min_query_len = 30; o->min_matched = 10; o->dedup_threshold = 256;
    o->min_query_cov = 0.55; o->min_target_cov = 0; o->max_fpr = 0.01;
}

typedef struct { ko_hit *v; size_t n, cap; } hitvec;
static void hv_push(hitvec *h, const ko_hit *x) {
    if (h->n == h->cap) { h->cap = h->cap ? h->cap * 2 : 16; h->v = (ko_hit *)realloc(h->v, h->cap * sizeof(ko_hit)); }
    h->v[h->n++] = *x;
}

static int g_sort_by = 0;
#pragma omp threadprivate(g_sort_by)
/* U:105-145 Less functions; ties (reference: unstable quicksort, any order) broken by target index */
static int cmp_hit(const void *pa, const void *pb) {
    const ko_hit *a = (const ko_hit *)pa, *b = (const ko_hit *)pb;
    if (g_sort_by == 0) {
        if (a->qcov > b->qcov) return -1;
        if (a->qcov < b->qcov) return 1;
        if (a->tcov > b->tcov) return -1;
        if (a->tcov < b->tcov) return 1;
    } else if (g_sort_by == 1) {
        if (a->tcov > b->tcov) return -1;
        if (a->tcov < b->tcov) return 1;
        if (a->count > b->count) return -1;
        if (a->count < b->count) return 1;
    } else {
        if (a->jacc > b->jacc) return -1;
        if (a->jacc < b->jacc) return 1;
        if (a->count > b->count) return -1;
        if (a->count < b->count) return 1;
    }
    return a->target < b->target ? -1 : (a->target > b->target ? 1 : 0);
}

/* probes codes against all blocks, appends matches (U:7412-7741) in block order, column order */
static void match_codes(const ko_db *db, const ko_search_opts *o, const uint64_t *codes, int64_t n, int algo,
                        uint32_t q, hitvec *out, uint32_t **cbuf, uint8_t **sbuf, size_t *ccap) {
    double nh = (double)n, thr = nh * o->min_query_cov;       /* nHashesThr (U:6625) */
    for (int bi = 0; bi < db->info.n_blocks; bi++) {
        const ko_block *b = &db->blocks[bi];
        size_t need = (size_t)b->row_bytes;
        if (need > *ccap) {
            *ccap = need;
            *cbuf = (uint32_t *)realloc(*cbuf, sizeof(uint32_t) * need * 8);
            *sbuf = (uint8_t *)realloc(*sbuf, need * 65 + 64);
        }
        probe_block(b, codes, n, algo, *cbuf, *sbuf);
        for (int t = 0; t < b->n_names; t++) {
            uint32_t cnt = (*cbuf)[t];
            if ((int64_t)cnt < o->min_matched) continue;       /* U:7466 */
            double c = (double)cnt;
            if (!(c > thr)) continue;                          /* U:7469 strict */
            double tcov = c / (double)b->sizes[t];
            if (!(tcov >= o->min_target_cov)) continue;        /* U:7473-7474 */
            double fpr = ko_query_fpr((int)n, (int)cnt, db->info.fpr);
            if (!(fpr <= o->max_fpr)) continue;                /* U:7477-7478 */
            ko_hit hit;
            hit.query = q; hit.target = (uint32_t)(b->target_base + t); hit.count = cnt; hit._pad = 0;
            hit.fpr = fpr; hit.qcov = c / nh; hit.tcov = tcov;
            hit.jacc = c / (nh + (double)b->sizes[t] - c);
            hv_push(out, &hit);
        }
    }
}

int ko_search(const ko_db *db, const ko_search_opts *o, const uint8_t *seqs, const uint64_t *off, uint32_t n_seqs,
              int paired, int threads, int algo, ko_results *res) {
    uint32_t nq = paired ? n_seqs / 2 : n_seqs;
    memset(res, 0, sizeof(*res));
    res->n_queries = nq;
    res->query_len = (int32_t *)calloc(nq ? nq : 1, 4);
    res->n_kmers = (int32_t *)calloc(nq ? nq : 1, 4);
    res->k_used = (int32_t *)calloc(nq ? nq : 1, 4);
    res->hit_off = (uint64_t *)calloc((size_t)nq + 1, 8);
    hitvec *per = (hitvec *)calloc(nq ? nq : 1, sizeof(hitvec));
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#else
    threads = 1;
#endif
#pragma omp parallel num_threads(threads)
    {
        uint64_t *codes = NULL; size_t codes_cap = 0;
        uint32_t *cbuf = NULL; uint8_t *sbuf = NULL; size_t ccap = 0;
#pragma omp for schedule(dynamic, 256)
        for (uint32_t q = 0; q < nq; q++) {
            const uint8_t *s1, *s2 = NULL; int64_t l1, l2 = 0;
            if (paired) {
                s1 = seqs + off[2 * q]; l1 = (int64_t)(off[2 * q + 1] - off[2 * q]);
                s2 = seqs + off[2 * q + 1]; l2 = (int64_t)(off[2 * q + 2] - off[2 * q + 1]);
            } else { s1 = seqs + off[q]; l1 = (int64_t)(off[q + 1] - off[q]); }
            res->query_len[q] = (int32_t)(l1 + l2);
            res->k_used[q] = db->info.ks[db->info.n_ks - 1];
            if (l1 < o->min_query_len && !(s2 && l2 >= o->min_query_len)) { res->k_used[q] = db->info.ks[0]; continue; }   /* U:778-786 */
            size_t need = (size_t)(l1 + l2 + 2);
            if (need > codes_cap) { codes_cap = need; codes = (uint64_t *)realloc(codes, sizeof(uint64_t) * codes_cap); }
            for (int ik = 0; ik < db->info.n_ks; ik++) {
                ko_sketch_params sp;
                sp.k = db->info.ks[ik]; sp.canonical = db->info.canonical; sp.scaled = db->info.scaled; sp.scale = db->info.scale;
                sp.minimizer = db->info.minimizer; sp.minimizer_w = db->info.minimizer_w;
                sp.syncmer = db->info.syncmer; sp.syncmer_s = db->info.syncmer_s;
                res->k_used[q] = sp.k;
                int64_t n1 = ko_generate_kmers(s1, l1, &sp, codes);
                int64_t nall = n1 + (s2 ? ko_generate_kmers(s2, l2, &sp, codes + n1) : 0);
                int found = 0, stop = 0;
                int tries_max = (o->try_se && s2) ? 3 : 1;
                for (int tries = 0; tries < tries_max && !found; tries++) {
                    const uint64_t *src = codes; int64_t n = nall;
                    if (tries == 1) { src = codes; n = n1; res->query_len[q] = (int32_t)l1; }
                    else if (tries == 2) { src = codes + n1; n = nall - n1; res->query_len[q] = (int32_t)l2; }
                    if (n < o->min_matched) { stop = 1; break; }                 /* U:854-869: returns for good */
                    uint64_t *work = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(n ? n : 1));
                    memcpy(work, src, sizeof(uint64_t) * (size_t)n);
                    n = ko_dedup(work, n, o->dedup_threshold);
                    res->n_kmers[q] = (int32_t)n;
                    size_t before = per[q].n;
                    match_codes(db, o, work, n, algo, q, &per[q], &cbuf, &sbuf, &ccap);
                    free(work);
                    if (per[q].n > before) found = 1;
                }
                if (found || stop) break;                                           /* else try smaller k (U:1018-1023) */
            }
            hitvec *hv = &per[q];
            if (hv->n > 1 && !o->do_not_sort) { g_sort_by = o->sort_by; qsort(hv->v, hv->n, sizeof(ko_hit), cmp_hit); }
            if (hv->n > 0 && o->top_n_scores > 0 && !o->do_not_sort) {              /* U:285-311, kept verbatim incl. the [:i+1] cut */
                int nsc = 0; size_t i; double pscore = 1024, score; int broke = 0;
                for (i = 0; i < hv->n; i++) {
                    score = o->sort_by == 0 ? hv->v[i].qcov : (o->sort_by == 1 ? hv->v[i].tcov : hv->v[i].jacc);
                    if (score < pscore) { nsc++; if (nsc > o->top_n_scores) { broke = 1; break; } pscore = score; }
                }
                if (broke) hv->n = i + 1;
            }
        }
        free(codes); free(cbuf); free(sbuf);
    }
    uint64_t tot = 0;
    for (uint32_t q = 0; q < nq; q++) { res->hit_off[q] = tot; tot += per[q].n; }
    res->hit_off[nq] = tot; res->n_hits = tot;
    res->hits = (ko_hit *)malloc(sizeof(ko_hit) * (size_t)(tot ? tot : 1));
    for (uint32_t q = 0; q < nq; q++) { if (per[q].n) memcpy(res->hits + res->hit_off[q], per[q].v, per[q].n * sizeof(ko_hit)); free(per[q].v); }
    free(per);
    return 0;
}

void ko_free_results(ko_results *r) {
    free(r->query_len); free(r->n_kmers); free(r->k_used); free(r->hit_off); free(r->hits);
    memset(r, 0, sizeof(*r));
}

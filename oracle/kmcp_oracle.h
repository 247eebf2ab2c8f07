/*
 * kmcp_oracle.h — CPU ORACLE for the `kmcp search` hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C restatement of the reference algorithm (shenwei356/kmcp v0.9.5, pure Go) used
 * as the parity checker for the CUDA path and as the labelled "port" CPU baseline in bench.py.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it.  The product (kmcp_b200/, libkmcp_gpu.so) never links, imports or calls anything here.
 *
 * Parity status: PINNED against the reference's published outputs G1..G5 by
 * tests/test_oracle_golden.py, which runs in the build container where /root/reference exists.
 * Unpinned pieces (no reference artefact): Minimizer sketch, IUPAC/'U' seeds, multi-k DBs, --try-se.
 *
 * Reference citations use: U: = kmcp/cmd/util-db-search.go, S: = kmcp/cmd/search.go,
 * X: = kmcp/cmd/index/serialization.go, H: = kmcp/cmd/util-hash.go, F: = kmcp/cmd/util-fpr.go.
 * Third-party arithmetic that is NOT in /root/reference (go.mod pins): will-rowe/nthash v0.4.0,
 * shenwei356/bio v0.9.0 (sketches), bmkessler/fastdiv, shenwei356/pospop v1.2.3, pand v0.0.7 —
 * restated here from their published algorithms.
 */
#ifndef KMCP_ORACLE_H
#define KMCP_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- sketching (U:1037-1107 generateKmers; external nthash + bio/sketches) ---- */
typedef struct {
    int32_t k;            /* k-mer size, <= 64 */
    int32_t canonical;    /* index.Header.Canonical (always 1 for kmcp index) */
    int32_t scaled;       /* FracMinHash on/off */
    uint32_t scale;       /* maxHash = uint64(float64(2^64-1)/float64(scale)) */
    int32_t minimizer;    /* minimizer on/off */
    uint32_t minimizer_w;
    int32_t syncmer;      /* closed syncmer on/off */
    uint32_t syncmer_s;
} ko_sketch_params;

/* raw per-position ntHash1 (canonical=min(fwd,rev)); out has len-k+1 entries; returns that count (0 if len<k) */
int64_t ko_nthash_all(const uint8_t *seq, int64_t len, int k, int canonical, uint64_t *out);
/* generateKmers: appends the codes the search path would probe; returns count (<= len-k+1) */
int64_t ko_generate_kmers(const uint8_t *seq, int64_t len, const ko_sketch_params *p, uint64_t *out);
/* U:874-908: if n > dedup_threshold sort ascending + unique in place; returns new n */
int64_t ko_dedup(uint64_t *codes, int64_t n, int64_t dedup_threshold);
/* H:125-141 hashValues */
void ko_hash_values(uint64_t code, int num_hashes, uint64_t *out);
/* F:32-50 + 140-193 (uncached function; bit-exact emulation of Go math.Pow / big.Float prec 53) */
double ko_query_fpr(int n, int c, double p);
double ko_go_pow(double x, double y);
/* H:46-50 */
uint64_t ko_calc_signature_size(uint64_t num_elements, int num_hashes, double fpr);

/* ---- database (X:383-593 block header; util-db-info.go:46-130 __db.yml) ---- */
typedef struct ko_db ko_db;
typedef struct {
    int32_t n_ks; int32_t ks[8];   /* sorted descending as U:752-759 */
    int32_t canonical, num_hashes;
    int32_t scaled; uint32_t scale;
    int32_t minimizer; uint32_t minimizer_w;
    int32_t syncmer; uint32_t syncmer_s;
    double fpr;
    int32_t n_blocks;
    int64_t n_targets;             /* sum over blocks of len(Names) */
    int64_t sum_row_bytes;         /* sum over blocks of numRowBytes (algorithmic bytes per probed k-mer per hash) */
    int64_t total_bytes;           /* sum numSigs*numRowBytes */
} ko_db_info;
typedef struct {
    const char *name; uint32_t index; /* chunkIdx | nChunks<<16 */
    uint64_t genome_size; uint64_t n_kmers;
    int32_t block; int32_t col;
} ko_target;

ko_db *ko_db_open(const char *dir, char *err, int errlen);   /* dir contains __db.yml */
void ko_db_close(ko_db *);
void ko_db_get_info(const ko_db *, ko_db_info *);
int ko_db_target(const ko_db *, int64_t global_target, ko_target *out);
/* block accessors for tests */
int ko_db_block(const ko_db *, int b, uint64_t *num_sigs, int32_t *row_bytes, int32_t *n_names,
                const uint8_t **rows);

/* ---- search (U:763-1025 handleQuery, U:6613-7741 block worker, U:260-345 sort/top-N) ---- */
typedef struct {
    int32_t min_query_len;   /* -m 30 */
    int32_t min_matched;     /* -c 10 */
    int32_t dedup_threshold; /* -u 256 */
    double min_query_cov;    /* -t 0.55 */
    double min_target_cov;   /* -T 0 */
    double max_fpr;          /* -f 0.01 */
    int32_t sort_by;         /* 0 qcov, 1 tcov, 2 jacc */
    int32_t do_not_sort;
    int32_t top_n_scores;    /* -n 0 */
    int32_t try_se;
} ko_search_opts;
typedef struct {
    uint32_t query; uint32_t target;  /* target = global index (blocks in __db.yml order, columns in order) */
    uint32_t count; uint32_t _pad;
    double fpr, qcov, tcov, jacc;
} ko_hit;
typedef struct {
    uint32_t n_queries;
    int32_t *query_len;   /* per query */
    int32_t *n_kmers;     /* per query (after dedup); 0 when skipped */
    int32_t *k_used;      /* per query */
    uint64_t *hit_off;    /* n_queries+1 */
    ko_hit *hits;         /* per query, in final (sorted / truncated) order */
    uint64_t n_hits;
} ko_results;

void ko_default_opts(ko_search_opts *);
/* seqs: concatenated ASCII; off: n_seqs+1 offsets.  paired: seq 2q,2q+1 are mates of query q.
 * threads: OpenMP threads (<=0: all).  algo: 0 = straightforward per-bit counting (checker),
 * 1 = reference algorithm shape (64-row buffer, byte-column transpose, positional popcount; U:6824-6969). */
int ko_search(const ko_db *, const ko_search_opts *, const uint8_t *seqs, const uint64_t *off,
              uint32_t n_seqs, int paired, int threads, int algo, ko_results *out);
void ko_free_results(ko_results *);
/* dense per-target counts of one code list against the whole DB (for kernel-level parity tests) */
int ko_count_codes(const ko_db *, const uint64_t *codes, int64_t n, uint32_t *counts /* n_targets */);

#ifdef __cplusplus
}
#endif
#endif

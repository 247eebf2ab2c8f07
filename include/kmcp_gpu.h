/*
 * kmcp_gpu.h — C ABI of libkmcp_gpu.so, the B200 (sm_100a) implementation of the `kmcp search` hot path.
 *
 * The reference (shenwei356/kmcp v0.9.5) is pure Go and has no FFI; the seam this library occupies is the
 * search-engine object used by kmcp/cmd/search.go.  Every entry point below names the
 * reference interface it replaces (paths relative to /root/reference/kmcp/cmd/):
 *   U: = util-db-search.go   S: = search.go   X: = index/serialization.go   H: = util-hash.go   F: = util-fpr.go
 * INTEGRATION.md shows the cgo binding a kmcp maintainer would add on top of these symbols.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success or a negative KMCPG_E*
 * code and never aborts; kmcpg_last_error() gives the message.  Inputs are caller-owned and not retained
 * after the call returns (cgo pointer rule); outputs are library-owned until the matching free call.
 * There is NO CPU fallback anywhere: without a CUDA device kmcpg_create fails with KMCPG_ECUDA.
 */
#ifndef KMCP_GPU_H
#define KMCP_GPU_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define KMCPG_ABI_VERSION 2

enum {
    KMCPG_OK = 0,
    KMCPG_EINVAL = -1,      /* bad argument */
    KMCPG_EIO = -2,         /* file missing / truncated */
    KMCPG_EFORMAT = -3,     /* X:38-56 ErrInvalidIndexFileFormat / ErrVersionMismatch / incompatible blocks (U:689-695) */
    KMCPG_ECUDA = -4,       /* CUDA runtime error, or no device */
    KMCPG_ENOMEM = -5,      /* HBM or host allocation failed */
    KMCPG_EUNSUPPORTED = -6
};

typedef struct kmcpg_ctx kmcpg_ctx;

/* ---- lifecycle: replaces NewUnikIndexDBSearchEngine / sg.Close (U:222, U:591) ------------------------- */
/* One context drives ONE device (one process per GPU, or several contexts in one process). */
int kmcpg_create(int device, kmcpg_ctx **out);
int kmcpg_close(kmcpg_ctx *ctx);
const char *kmcpg_last_error(const kmcpg_ctx *ctx); /* ctx may be NULL: error of the failed kmcpg_create */
int kmcpg_abi_version(void);
/* run the probe launches of this context on the caller's CUDA stream (e.g. torch's current stream) so the caller's
 * events bracket them; NULL restores the context's own stream.  Query preparation, staging and result return always run
 * on the library's own streams, so inputs handed over in device memory must be complete when they are submitted (or
 * carry a ready_event, see kmcpg_batch). */
int kmcpg_set_stream(kmcpg_ctx *ctx, void *cuda_stream);

/* ---- database: replaces NewUnikIndexDB + NewUnikIndex + index.NewReader (U:648-760, U:1196-1280, X:372-593) */
typedef struct {
    int32_t shard_rank;   /* this context keeps the blocks assigned to shard_rank of shard_world (greedy by bytes, */
    int32_t shard_world;  /* largest first); 0/1 = keep everything.  Target numbering stays global in every shard. */
                          /* A DB with FEWER blocks than shards is cut by column (= target) ranges instead: see kmcpg_shard_pieces. */
    int64_t max_resident_bytes; /* 0 = no limit; otherwise fail with KMCPG_ENOMEM instead of oversubscribing HBM */
} kmcpg_db_opts;

typedef struct {
    int32_t n_ks; int32_t ks[8];        /* descending, U:752-759 */
    int32_t canonical, num_hashes;
    int32_t scaled; uint32_t scale;
    int32_t minimizer; uint32_t minimizer_w;
    int32_t syncmer; uint32_t syncmer_s;
    double fpr;                          /* __db.yml fpr: p of one k-mer (F:140) */
    int32_t n_blocks;                    /* all blocks of the DB */
    int32_t n_resident_blocks;           /* blocks in this context's HBM */
    int64_t n_targets;                   /* all targets of the DB */
    int64_t sum_row_bytes;               /* Σ numRowBytes over RESIDENT blocks: algorithmic bytes per probed row set */
    int64_t resident_bytes;              /* HBM bytes of the re-pitched resident rows */
    int64_t disk_bytes;                  /* Σ numSigs·numRowBytes over resident blocks */
} kmcpg_db_info_t;

typedef struct {
    const char *name;        /* library-owned, valid until kmcpg_close */
    uint32_t index;          /* chunkIdx | nChunks<<16 (I:1096, S:532-533) */
    uint64_t genome_size;
    uint64_t n_kmers;        /* Sizes[t] */
    int32_t block, col;
    int32_t resident;        /* 1 if the target's block is in this context */
} kmcpg_target_t;

/* the block → shard assignment kmcpg_open_db uses (host only, no device needed): blocks sorted by re-pitched
 * bytes, largest first onto the least loaded shard; owner[i] = shard of block i in __db.yml order */
int kmcpg_shard_plan(const char *dir, int shard_world, int32_t *owner, int32_t n_owner);
/* the same plan as pieces (host only): with at least as many blocks as shards every piece is a whole block; with fewer
 * blocks than shards (e.g. one wide block on 8 GPUs) blocks are cut by column range — the columns of all blocks in units
 * of 128 targets (16 row bytes), weighted by numSigs, laid end to end and dealt out in shard_world equal-cost stretches.
 * Rows stay aligned and counts are per target, so the shards' hit lists are still disjoint by target and are only
 * concatenated.  Returns the number of pieces (every column of every block in exactly one piece). */
typedef struct {
    int32_t block, shard;      /* block in __db.yml order; owning shard */
    uint32_t col0, n_cols;     /* resident columns of the block: [col0, col0 + n_cols), col0 a multiple of 128 */
    uint64_t resident_bytes;   /* numSigs x re-pitched row bytes of the piece */
} kmcpg_shard_piece;
int kmcpg_shard_pieces(const char *dir, int shard_world, kmcpg_shard_piece *out, int32_t cap);
/* dir = the directory holding __db.yml (normally <db>/R001, S:299-324) */
int kmcpg_open_db(kmcpg_ctx *ctx, const char *dir, const kmcpg_db_opts *opts);
int kmcpg_db_info(const kmcpg_ctx *ctx, kmcpg_db_info_t *out);
int kmcpg_target(const kmcpg_ctx *ctx, int64_t global_target, kmcpg_target_t *out);
/* Sizes[t] (k-mers of target t) of all n_targets targets as float64 (U:1393-1396 sizesFloat), n >= n_targets */
int kmcpg_target_sizes(const kmcpg_ctx *ctx, double *out, int64_t n);

/* ---- the hot path: replaces UnikIndexDB.handleQuery k-mer generation + every UnikIndex worker `fn`
 *      (U:763-941 and U:6613-7741) for a whole batch of queries ------------------------------------------- */
typedef struct {
    int32_t min_query_len;    /* -m, SearchOptions.MinQLen (U:778) */
    int32_t min_matched;      /* -c, MinMatched (U:854, U:7466) */
    int32_t dedup_threshold;  /* -u, DeduplicateThreshold (U:874) */
    int32_t paired;           /* 1: sequences 2q and 2q+1 are Seq / Seq2 of query q (U:797-805) */
    double min_query_cov;     /* -t, MinQueryCov: count must be > n*t in float64 (U:6625, U:7469) */
    int32_t k;                /* 0 = largest k of the DB; multi-k DBs are driven by the host engine (U:763) */
    int32_t mate_select;      /* 0 both mates, 1 Seq only, 2 Seq2 only (--try-se retries, U:826-842) */
} kmcpg_search_params;

typedef struct {
    uint32_t query;   /* index within the batch */
    uint32_t target;  /* global target index */
    uint32_t count;   /* matched k-mers (Match.NumKmers) */
} kmcpg_hit;

typedef struct {
    uint32_t n_queries;
    uint64_t n_hits;
    int32_t *n_kmers;      /* per query: k-mers probed after dedup (QueryResult.NumKmers); 0 = skipped (U:778-786, 854-869) */
    int32_t *query_len;    /* per query (QueryResult.QueryLen) */
    kmcpg_hit *hits;       /* every (query,target) with count >= min_matched and count > n*min_query_cov, sorted by (query,target) */
    /* device-side timing of this call (CUDA events on the library's streams), milliseconds */
    float ms_hash;    /* slot scan + hash (+ sort/unique) + per-query verdict */
    float ms_locs;    /* always 0 since ABI 2: the row-index kernels run on the query-preparation stream and are part of ms_hash's stage */
    float ms_probe;   /* Σ durations of the probe kernel launches ONLY (events bracketing each launch) */
    float ms_total;   /* host wall clock of the call */
    uint32_t probe_launches;
    uint64_t probe_row_bytes;   /* algorithmic bytes the probe kernels had to fetch: Σ_q n_q · h · Σ_b numRowBytes_b */
    uint32_t kernel_launches;   /* kernels of this library launched for the call */
    void *_priv;
} kmcpg_hits;

void kmcpg_default_params(kmcpg_search_params *p);
/* host buffers: seq = concatenated ASCII, off = n_seqs+1 offsets into seq */
int kmcpg_search_batch(kmcpg_ctx *ctx, const kmcpg_search_params *p, const uint8_t *seq, const uint64_t *off,
                       uint32_t n_seqs, kmcpg_hits *out);
/* same with seq/off already in this device's HBM (e.g. a batch that arrived by ncclBroadcast) */
int kmcpg_search_batch_device(kmcpg_ctx *ctx, const kmcpg_search_params *p, const uint8_t *d_seq,
                              const uint64_t *d_off, uint32_t n_seqs, uint64_t seq_bytes, kmcpg_hits *out);
void kmcpg_free_hits(kmcpg_hits *h);

/* streaming form: the batch is processed in parts (≈ 250 k reads); `cb` is called on the context's executor thread (one
 * call at a time, in query order) as soon as a part's hits are in host memory — while the GPU is already probing the
 * next parts — so a host can post-process (tCov/FPR/sort, TSV formatting) in the shadow of the device work.  The
 * pointers are valid until the call returns; `hits[i].query` is the index inside the batch.  The callback must not
 * call back into this ctx. */
typedef struct {
    uint32_t first_query, n_queries;
    const int32_t *n_kmers;      /* [n_queries], of first_query.. */
    const int32_t *query_len;    /* [n_queries] */
    const kmcpg_hit *hits;       /* sorted by (query, target) */
    uint64_t n_hits;
} kmcpg_part;
typedef void (*kmcpg_part_cb)(void *user, const kmcpg_part *part);
int kmcpg_search_batch_cb(kmcpg_ctx *ctx, const kmcpg_search_params *p, const uint8_t *seq, const uint64_t *off,
                          uint32_t n_seqs, kmcpg_part_cb cb, void *user, kmcpg_hits *summary /* timings and totals; arrays stay valid until freed */);

/* asynchronous form (what the three calls above are made of): kmcpg_search_submit queues a batch and returns; the context's
 * executor thread feeds the GPU with the parts of all queued batches back to back, so with a second batch submitted before
 * the first is waited for, the boundary between batches costs no GPU time (the Go engine keeps two batches in flight:
 * one being filled from InCh while the other is searched, U:209-243).  Inputs must stay valid until kmcpg_search_wait
 * returns; jobs of one context complete in submission order; submit may be called from any thread. */
typedef struct kmcpg_job kmcpg_job;
typedef struct {
    const uint8_t *seq;          /* concatenated ASCII: host memory (pinned or not), or this context's device memory */
    const uint64_t *off;         /* n_seqs + 1 offsets into seq, in the same memory space as seq */
    uint32_t n_seqs;
    int32_t on_device;           /* 1: seq / off are device pointers (e.g. a batch that arrived by ncclBroadcast) */
    const uint64_t *host_off;    /* on_device: optional host copy of off (saves fetching it back to cut the batch into parts) */
    void *ready_event;           /* on_device: optional cudaEvent_t after which seq / off are complete; NULL: complete now */
    kmcpg_hit *hits_dst;         /* optional caller-owned destination of the hit list (pinned or cudaHostRegister-ed host memory, */
    uint64_t hits_cap;           /*   e.g. a shared-memory segment another process reads); more than hits_cap hits: KMCPG_ENOMEM */
    kmcpg_part_cb cb;            /* optional: parts are handed over as they land (see kmcpg_search_batch_cb) */
    void *user;
    uint32_t first_query;        /* added to the query index of every hit and part reported (a batch submitted as several jobs); */
    uint32_t _pad;               /*   n_kmers / query_len of kmcpg_search_wait stay indexed from 0 */
} kmcpg_batch;
int kmcpg_search_submit(kmcpg_ctx *ctx, const kmcpg_search_params *p, const kmcpg_batch *b, kmcpg_job **job);
/* blocks until the job is done and releases it; out as for kmcpg_search_batch (out->hits == hits_dst when one was given);
 * out may be NULL to drop the results */
int kmcpg_search_wait(kmcpg_job *job, kmcpg_hits *out);

/* ---- the reader stage (host only, no device needed): replaces the reader loop of search.go (S:793-1000, fastx.Reader over
 *      xopen/pgzip on one goroutine).  FASTA/Q files, plain or gzip → packed batches of queries in input order, ready for
 *      kmcpg_search_batch / kmcpg_engine_search: paired-end files are zipped pair by pair and end with the shorter file
 *      (S:806-867), `-g` makes one query of a whole file, its records joined as S:899-913 does.  Every input is decoded and
 *      parsed on threads of its own (own gzip decoders: one .gz stream can be decoded by several threads). ------------------ */
typedef struct kmcpg_reader kmcpg_reader;
typedef struct {
    const char *read1, *read2;        /* -1 / -2: paired-end files (both or neither) */
    const char *const *files;         /* single-end files when read1/read2 are NULL; "-" = stdin */
    int32_t n_files;
    int32_t whole_file;               /* -g: every file is ONE query */
    int32_t use_filename;             /* -G: the file name (extensions cut) is the ID of a -g query */
    const char *query_id;             /* --query-id, or NULL */
    int32_t k;                        /* largest k of the database: -g joins records with k-1 'N' (S:881) */
    uint32_t batch_reads;             /* queries per batch, 0 = 262144 */
    uint64_t batch_bytes;             /* sequence bytes per batch, 0 = 256 MB */
    int32_t inflate_threads;          /* 0 = decided per file, 1 = sequential decoder, N = N threads on ONE .gz stream */
    int32_t parse_threads;            /* 0, 1 = one parser thread per input, N = N threads on ONE FASTQ text */
    void (*log)(void *user, const char *level, const char *msg);   /* the reference's log lines (S:800, 878, 920), or NULL */
    void *log_user;
    uint64_t inflate_chunk, inflate_cap, parse_piece;              /* test knobs, 0 = defaults */
} kmcpg_reader_opts;
typedef struct {
    uint32_t n_queries, n_seqs;       /* n_seqs = n_queries, or 2 x n_queries for paired-end input */
    const uint8_t *seq;               /* the sequences back to back */
    const uint64_t *off;              /* n_seqs + 1 offsets into seq */
    const char *ids;                  /* ID of query q = ids[id_off[q] .. id_off[q+1]) (ID of read 1 for pairs) */
    const uint64_t *id_off;           /* n_queries + 1 */
    uint64_t first_query;             /* position of the batch's first query in the whole input (Query.Idx, S:862) */
    void *_priv;
} kmcpg_read_batch;
void kmcpg_default_reader_opts(kmcpg_reader_opts *o);
int kmcpg_reader_open(const kmcpg_reader_opts *o, kmcpg_reader **out);
/* 1: *out is the next batch (release it with kmcpg_reader_free_batch), 0: end of the input, < 0: error (kmcpg_reader_error) */
int kmcpg_reader_next(kmcpg_reader *r, kmcpg_read_batch *out);
void kmcpg_reader_free_batch(kmcpg_read_batch *b);
const char *kmcpg_reader_error(const kmcpg_reader *r);
int kmcpg_reader_close(kmcpg_reader *r);

/* host memory shared between the processes of a one-process-per-GPU run (the read batch is broadcast, every rank probes
 * its blocks, "per-GPU hit lists concatenated on the host"): a named segment (a file under /dev/shm, or /tmp where that is missing),
 * mapped by every process that opens the name; with cuda_register it is page-locked for this process's device, so a rank's hit
 * list can be copied device→host straight into it (kmcpg_batch.hits_dst) and the gathering rank reads it in place.  bytes must be
 * the same in every process; create = 1 in the process that owns the segment (the others open it afterwards). */
int kmcpg_shm_open(const char *name, size_t bytes, int create, int cuda_register, void **ptr);
int kmcpg_shm_close(const char *name, void *ptr, size_t bytes, int cuda_registered, int unlink_it);

/* pinned host memory for batch buffers (so the H2D copy of kmcpg_search_batch runs at full PCIe speed) */
int kmcpg_host_alloc(void **p, size_t bytes);
int kmcpg_host_free(void *p);
/* device memory helpers for callers without a CUDA binding (tests, bench, NCCL staging) */
int kmcpg_device_memory(kmcpg_ctx *ctx, size_t *free_bytes, size_t *total_bytes);   /* cudaMemGetInfo of the context's device */
int kmcpg_device_alloc(kmcpg_ctx *ctx, void **p, size_t bytes);
int kmcpg_device_free(kmcpg_ctx *ctx, void *p);
int kmcpg_memcpy_h2d(kmcpg_ctx *ctx, void *d, const void *h, size_t bytes);
int kmcpg_memcpy_d2h(kmcpg_ctx *ctx, void *h, const void *d, size_t bytes);

/* ---- stage-level entry points (each replaces one reference function; used by the parity tests) --------- */
/* UnikIndexDB.generateKmers (U:1037-1107) for a batch: codes of sequence i are out_codes[out_off[i] .. out_off[i+1]) */
typedef struct {
    int32_t k, canonical, scaled; uint32_t scale;
    int32_t minimizer; uint32_t minimizer_w;
    int32_t syncmer; uint32_t syncmer_s;
} kmcpg_sketch_params;
int kmcpg_generate_kmers(kmcpg_ctx *ctx, const kmcpg_sketch_params *sp, const uint8_t *seq, const uint64_t *off,
                         uint32_t n_seqs, uint64_t **out_codes, uint64_t **out_off);
/* one UnikIndex worker pass without thresholds (U:6613-7408): dense counts[n_targets] of one code list
 * against all resident blocks (non-resident targets get 0); dedup is NOT applied */
int kmcpg_count_codes(kmcpg_ctx *ctx, const uint64_t *codes, uint64_t n, uint32_t *counts);
void kmcpg_free(void *p);

/* ---- host engine: C++ mirror of UnikIndexDBSearchEngine result handling (U:260-345, U:7466-7491) -------- */
typedef struct {
    int32_t min_query_len, min_matched, dedup_threshold;
    double min_query_cov, min_target_cov, max_fpr;
    int32_t sort_by;        /* 0 qcov, 1 tcov, 2 jacc (S:1093) */
    int32_t do_not_sort;    /* -S */
    int32_t top_n_scores;   /* -n */
    int32_t try_se;         /* --try-se */
    int32_t paired;
    int32_t threads;        /* host threads for the post-filter; <=0: all */
} kmcpg_engine_opts;

typedef struct {
    uint32_t query, target, count, _pad;
    double fpr, qcov, tcov, jacc;      /* Match.FPR/QCov/TCov/JaccardIndex (U:83-93) */
} kmcpg_match;

typedef struct {
    uint32_t n_queries;
    uint64_t n_matches;
    int32_t *query_len, *n_kmers, *k_used;   /* per query */
    uint64_t *match_off;                     /* n_queries+1 */
    kmcpg_match *matches;                    /* per query in output order (sorted / truncated as U:273-311) */
    float ms_gpu_total;                      /* Σ wall time of the kmcpg_search_batch calls */
    float ms_post;                           /* host post-filter (tCov, FPR, sort, top-N) */
    float ms_total;                          /* wall time of the whole engine call */
    uint64_t probe_row_bytes;
    uint32_t kernel_launches;
    void *_priv;
} kmcpg_results;

void kmcpg_default_engine_opts(kmcpg_engine_opts *o);
int kmcpg_engine_search(kmcpg_ctx *ctx, const kmcpg_engine_opts *o, const uint8_t *seq, const uint64_t *off,
                        uint32_t n_seqs, kmcpg_results *out);
/* the same engine over ONE database sharded across several contexts of this process (normally one per GPU of the box,
 * each opened with kmcpg_open_db(shard_rank = i, shard_world = n_ctx)): every context receives the whole batch (its own
 * H2D copy from the caller's buffer) and probes it against its resident blocks on its own host thread; the per-shard hit
 * lists — disjoint by target — are merged part by part in (query, target) order while the GPUs work on the next parts, and
 * the thresholds / sort / top-N / --try-se / multi-k logic then runs on the union exactly as for one context, so the result
 * is identical to kmcpg_engine_search on a context holding the whole database (block fan-out + gather of U:939-964). */
int kmcpg_engine_search_sharded(kmcpg_ctx *const *ctxs, int n_ctx, const kmcpg_engine_opts *o, const uint8_t *seq,
                                const uint64_t *off, uint32_t n_seqs, kmcpg_results *out);
/* the other way to use several GPUs, for databases that fit every one of them: each context holds the WHOLE database
 * (kmcpg_open_db without shard options) and the READS are split — contiguous query ranges with about the same number of
 * sequence bytes, one per context, each searched by kmcpg_engine_search on its own host thread (so hashing, the input copy
 * and the host post-filter are divided as well, which block shards cannot do); queries are independent, so the result is
 * the concatenation of the ranges' results and equals kmcpg_engine_search on one context.  Contexts holding only a shard are
 * refused with KMCPG_EINVAL. */
int kmcpg_engine_search_replicas(kmcpg_ctx *const *ctxs, int n_ctx, const kmcpg_engine_opts *o, const uint8_t *seq,
                                 const uint64_t *off, uint32_t n_seqs, kmcpg_results *out);
/* the engine's result handling alone (host only, no device): hit lists that were produced by other processes — the per-rank lists of
 * a one-process-per-GPU run after kmcpg_merge_hits — go through the same filter / sort / top-N code as behind kmcpg_engine_search.
 * hits sorted by (query, target); target_sizes = kmcpg_target_sizes of the whole database; fpr, k = the database's; one k, no
 * --try-se (retries need the device); part_queries 0 = default. */
int kmcpg_engine_postfilter(const kmcpg_engine_opts *o, uint32_t n_queries, const int32_t *n_kmers, const int32_t *query_len, const kmcpg_hit *hits,
                            uint64_t n_hits, const double *target_sizes, int64_t n_targets, double fpr, int k, uint32_t part_queries, kmcpg_results *out);
/* union of per-shard hit lists (each sorted by (query, target), disjoint by target) in (query, target) order — the gather of
 * U:939-964 across processes; queries within [first_query, first_query + n_queries); out holds the sum of n[] records */
int kmcpg_merge_hits(const kmcpg_hit *const *lists, const uint64_t *n, int n_lists, uint32_t first_query, uint32_t n_queries, int threads, kmcpg_hit *out);
/* order-sensitive 64-bit digest of a hit list (sum over i of a mix of (first_index + i, query, target, count)): equal for runs that
 * return the same hits in the same order, e.g. one search on 1, 2, 4 and 8 GPUs */
uint64_t kmcpg_hits_digest(const kmcpg_hit *hits, uint64_t n, uint64_t first_index);
void kmcpg_free_results(kmcpg_results *r);
/* QueryFPRWithCacheWithConstantFPR's underlying function (F:32-50, F:140-193), bit-exact with Go */
double kmcpg_query_fpr(int n, int c, double p);

/* ---- database build on the GPU: `kmcp compute` + `kmcp index` fused (compute.go C:577-1033, index.go I:117-1400;
 *      SURVEY.md §8 f1).  FASTA/Q(.gz) files → <out_dir>/R001/{_blockNNN.uniki, __db.yml, __name_mapping.tsv}; the
 *      database is left open in this context.  One file = one reference genome. ----------------------------------- */
typedef struct {
    int32_t k;                    /* compute -k */
    int32_t num_hashes;           /* index -n/--num-hash (1..4) */
    double fpr;                   /* index -f/--false-positive-rate */
    int32_t split_number;         /* compute -n/--split-number (<=1: no splitting) */
    int32_t split_overlap;        /* compute -l/--split-overlap (<0: k-1, C:268-270) */
    int32_t split_min_ref;        /* compute -m/--split-min-ref */
    uint32_t scale;               /* compute -D/--scale (>1: FracMinHash) */
    uint32_t minimizer_w;         /* compute -W */
    uint32_t syncmer_s;           /* compute -S */
    int32_t block_size;           /* index -b/--block-size; 0: (nFiles/threads+7)/8*8 clamped to [8,nFiles] (I:670-682) */
    int32_t threads;              /* the -j the block-size rule divides by */
    const char *ref_name_regexp;  /* compute -N (first capture group of the file name); NULL: file name without extension */
    const char *const *seq_name_filters; /* compute -B regexps (case-insensitive), matched against the whole header */
    int32_t n_seq_name_filters;
} kmcpg_index_params;
void kmcpg_default_index_params(kmcpg_index_params *p);
int kmcpg_index_fasta(kmcpg_ctx *ctx, const kmcpg_index_params *p, const char *const *files, int n_files, const char *out_dir);

/* ---- `kmcp profile` stage 1/4 straight from the result stream ----------------------------------
 * Replaces the first pass of kmcp/cmd/profile.go:761-990 (+ parseMatchResult, util-profile.go:94-182) over the
 * search TSV: per reference genome and chunk, Match / UniqMatch / UniqMatchHic, without the text round trip.
 * qCov and FPR are compared after the same rounding the TSV applies (%.4f / %.4e), so the counters are the ones the
 * reference computes from the file.  Taxonomy (`--level species`) is not applied.  Rows of a query must be sorted
 * by qCov (the search default), as `profile` assumes. */
typedef struct {
    double min_query_cov;     /* profile -t/--min-query-cov (0.55) */
    double max_fpr;           /* profile -f/--max-fpr (0.01) */
    int32_t top_n_scores;     /* profile -n/--keep-top-qcovs (0: off) */
    int32_t keep_perfect;     /* --keep-perfect-matches */
    int32_t keep_main;        /* --keep-main-matches */
    double max_qcov_gap;      /* --max-qcov-gap (0.4) */
    double hic_min_qcov;      /* -H/--min-hic-ureads-qcov (0.75) */
} kmcpg_refcount_params;
typedef struct kmcpg_refcounts kmcpg_refcounts;
typedef struct {
    const char *name;         /* reference (target name); library-owned */
    uint64_t genome_size;
    uint32_t n_chunks, _pad;
    const double *match, *uniq_match, *uniq_match_hic;   /* n_chunks each: Target.Match/UniqMatch/UniqMatchHic */
} kmcpg_refcount_row;
typedef struct {
    uint64_t n_reads;         /* queries with at least one kept row (nReads) */
    uint32_t n_refs, _pad;    /* references with at least one kept row, in database order */
    const kmcpg_refcount_row *rows;   /* valid until the next add/get/free on the accumulator */
} kmcpg_refcount_table;
void kmcpg_default_refcount_params(kmcpg_refcount_params *p);
/* targets from the database open in ctx, or (ctx NULL) from the block headers under db_dir (…/R001; host only) */
int kmcpg_refcounts_create(kmcpg_ctx *ctx, const char *db_dir, const kmcpg_refcount_params *p, kmcpg_refcounts **out);
int kmcpg_refcounts_add(kmcpg_refcounts *rc, const kmcpg_results *r);     /* one batch of kmcpg_engine_search, in input order */
int kmcpg_refcounts_get(kmcpg_refcounts *rc, kmcpg_refcount_table *out);
void kmcpg_refcounts_free(kmcpg_refcounts *rc);

/* ---- synthetic workloads (bench/test tooling; seeded pure functions, mirrored in oracle/oracle.py) ------ */
/* d_out[i*read_len .. ) = read (first+i) of the seeded read set; returns device pointers */
int kmcpg_synth_reads(kmcpg_ctx *ctx, uint64_t seed, uint64_t first, uint32_t n_reads, uint32_t read_len,
                      uint64_t genome_seed, uint32_t n_genomes, uint32_t genome_len, uint8_t *d_out);
/* build an in-HBM DB from seeded random genomes exactly as `kmcp compute`+`kmcp index` would (C:577-826,
 * I:667-1309, non-circular split mode), without touching disk; replaces any open DB of ctx */
typedef struct {
    uint64_t genome_seed; uint32_t n_genomes; uint32_t genome_len;
    int32_t k; int32_t n_chunks; int32_t overlap;
    int32_t num_hashes; double fpr; int32_t block_size; /* targets per block */
    uint32_t scale;                                      /* > 1: FracMinHash sketch database (compute -D) */
    int32_t shard_rank, shard_world;                     /* > 1 shards: only the blocks kmcpg_open_db(shard_rank, shard_world) would keep are built */
} kmcpg_synth_db;
int kmcpg_build_synth_db(kmcpg_ctx *ctx, const kmcpg_synth_db *spec);
/* d_out[i*genome_len ..) = seeded genome (first+i), ASCII */
int kmcpg_synth_genomes(kmcpg_ctx *ctx, uint64_t genome_seed, uint32_t first, uint32_t n_genomes, uint32_t genome_len, uint8_t *d_out);
/* dump resident block b in .uniki format (X:153-304), for parity checks of the device builder */
int kmcpg_write_block(kmcpg_ctx *ctx, int resident_block, const char *path);

#ifdef __cplusplus
}
#endif
#endif

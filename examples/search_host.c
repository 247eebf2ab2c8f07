/* search_host.c — a plain C99 host of libkmcp_gpu.so: what any FFI (cgo, JNI, ctypes) does, without the language runtime.
 *
 *   search_host <db>/R001 "ACGT...read1" ["ACGT...read2" ...]
 *
 * prints one line per (query, target) hit: query index, target name, chunk index, matched k-mers, query k-mers.
 * Build: gcc -std=c99 examples/search_host.c -Iinclude -Lkmcp_b200 -lkmcp_gpu -Wl,-rpath,$PWD/kmcp_b200 -o search_host
 * Exit codes: 0 ok, 1 usage, 2 no CUDA device (the library has no CPU fallback), 3 any other library error. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "kmcp_gpu.h"

int main(int argc, char **argv) {
    kmcpg_ctx *ctx = NULL;
    kmcpg_search_params p;
    kmcpg_hits hits;
    uint64_t *off;
    uint8_t *seq;
    size_t total = 0;
    uint32_t n, i;
    uint64_t h;
    int rc;

    if (argc < 3) {
        fprintf(stderr, "usage: %s <db>/R001 <sequence> [<sequence> ...]\n", argv[0]);
        return 1;
    }
    rc = kmcpg_create(0, &ctx);
    if (rc != KMCPG_OK) {
        fprintf(stderr, "kmcpg_create: %s\n", kmcpg_last_error(NULL));
        return rc == KMCPG_ECUDA ? 2 : 3;
    }
    rc = kmcpg_open_db(ctx, argv[1], NULL);
    if (rc != KMCPG_OK) {
        fprintf(stderr, "kmcpg_open_db: %s\n", kmcpg_last_error(ctx));
        kmcpg_close(ctx);
        return 3;
    }
    /* the batch: concatenated ASCII sequences + n+1 offsets; caller-owned, not retained by the library */
    n = (uint32_t)(argc - 2);
    for (i = 0; i < n; i++) total += strlen(argv[2 + i]);
    seq = (uint8_t *)malloc(total ? total : 1);
    off = (uint64_t *)malloc(((size_t)n + 1) * sizeof(uint64_t));
    if (!seq || !off) return 3;
    off[0] = 0;
    for (i = 0; i < n; i++) {
        const size_t len = strlen(argv[2 + i]);
        memcpy(seq + off[i], argv[2 + i], len);
        off[i + 1] = off[i] + len;
    }
    kmcpg_default_params(&p);               /* kmcp search defaults: -m 30 -c 10 -t 0.55 -u 256 */
    rc = kmcpg_search_batch(ctx, &p, seq, off, n, &hits);
    if (rc != KMCPG_OK) {
        fprintf(stderr, "kmcpg_search_batch: %s\n", kmcpg_last_error(ctx));
        kmcpg_close(ctx);
        return 3;
    }
    for (h = 0; h < hits.n_hits; h++) {
        kmcpg_target_t t;
        const kmcpg_hit *x = &hits.hits[h];
        if (kmcpg_target(ctx, (int64_t)x->target, &t) != KMCPG_OK) continue;
        printf("%u\t%s\t%u\t%u\t%d\n", x->query, t.name, t.index & 0xFFFFu, x->count, hits.n_kmers[x->query]);
    }
    kmcpg_free_hits(&hits);                 /* results are library-owned until this call */
    free(seq);
    free(off);
    kmcpg_close(ctx);
    return 0;
}

/* read_host.c — the reader stage of libkmcp_gpu.so from plain C99 (no device needed): FASTA/Q(.gz) files → batches.
 *
 *   read_host reads.fq.gz [more files ...]            single-end
 *   read_host -p reads_1.fq.gz reads_2.fq.gz          paired-end (zipped pair by pair, ends with the shorter file)
 *
 * prints one line per batch (first query, queries, sequence bytes) and the ID and length(s) of the first query of every batch;
 * the arrays of a batch are exactly what kmcpg_search_batch / kmcpg_engine_search take (see search_host.c).
 * Build: gcc -std=c99 examples/read_host.c -Iinclude -Lkmcp_b200 -lkmcp_gpu -Wl,-rpath,$PWD/kmcp_b200 -o read_host
 * Exit codes: 0 ok, 1 usage, 3 reader error (message on stderr). */
#include <stdio.h>
#include <string.h>

#include "kmcp_gpu.h"

int main(int argc, char **argv) {
    kmcpg_reader_opts o;
    kmcpg_reader *rd = NULL;
    kmcpg_read_batch b;
    unsigned long long total = 0;
    int rc, paired;

    paired = argc == 4 && strcmp(argv[1], "-p") == 0;
    if (argc < 2 || (argv[1][0] == '-' && argv[1][1] == 'p' && !paired)) {
        fprintf(stderr, "usage: %s <file> [<file> ...] | -p <read1> <read2>\n", argv[0]);
        return 1;
    }
    kmcpg_default_reader_opts(&o);
    o.batch_reads = 1000;
    if (paired) { o.read1 = argv[2]; o.read2 = argv[3]; }
    else { o.files = (const char *const *)(argv + 1); o.n_files = argc - 1; }
    rc = kmcpg_reader_open(&o, &rd);
    if (rc != KMCPG_OK) { fprintf(stderr, "kmcpg_reader_open: %d\n", rc); return 3; }
    while ((rc = kmcpg_reader_next(rd, &b)) == 1) {
        const unsigned step = b.n_seqs / b.n_queries;
        printf("batch first=%llu queries=%u bytes=%llu id=%.*s len=%llu", (unsigned long long)b.first_query, b.n_queries,
               (unsigned long long)b.off[b.n_seqs], (int)(b.id_off[1] - b.id_off[0]), b.ids, (unsigned long long)(b.off[1] - b.off[0]));
        if (step == 2) printf(",%llu", (unsigned long long)(b.off[2] - b.off[1]));
        printf("\n");
        total += b.n_queries;
        kmcpg_reader_free_batch(&b);
    }
    if (rc < 0) { fprintf(stderr, "reader error: %s\n", kmcpg_reader_error(rd)); kmcpg_reader_close(rd); return 3; }
    printf("queries: %llu\n", total);
    kmcpg_reader_close(rd);
    return 0;
}

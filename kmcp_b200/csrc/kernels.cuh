// kernels.cuh — launch interface of the sm_100a kernels of the kmcp search path.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "common.h"

namespace kmcpg {

// ---- kernel 1: sequences → k-mer codes (UnikIndexDB.generateKmers, U:1037-1107) -------------------------
struct HashArgs {
    const uint8_t *seq;         // concatenated ASCII
    const uint64_t *seq_off;    // n_seqs+1
    const uint64_t *slot_off;   // n_seqs+1: start of the code region of every sequence (upper bound layout)
    uint64_t *codes;            // [slot_off[n_seqs]]
    uint32_t *n_codes;          // per QUERY: codes written (mates concatenated)
    int32_t *query_len;         // per QUERY: len(Seq)+len(Seq2)
    uint32_t n_queries;
    int paired;                 // 1: query q = sequences 2q, 2q+1
    int mate_select;            // 0 both, 1 first, 2 second
    int k;
    int canonical;
    int scaled;
    uint64_t max_hash;
    int minimizer; uint32_t minimizer_w;
    int syncmer; uint32_t syncmer_s;
    int min_query_len;
    int raw;                    // 1: every sequence on its own, hash of position p written to codes[slot_off[s]+p],
                                //    no filter, no compaction, n_codes/query_len untouched (input of the sketch selection)
};
// minimizer / closed-syncmer selection (bio/sketches NextMinimizer / NextSyncmer)
struct SelectArgs {
    const uint64_t *seq_off;    // n_seqs+1 (lengths)
    const uint64_t *ck;         // canonical k-mer hash of every position, region slot_off[s]
    const uint64_t *slot_off;   // n_seqs+1
    const uint64_t *cs;         // syncmer: canonical s-mer hash of every position, region cs_off[s]
    const uint64_t *cs_off;     // n_seqs+1
    uint64_t *codes;            // out: selected codes, compacted at slot_off[first sequence of the query]
    uint32_t *n_codes;          // per query
    int32_t *query_len;         // per query
    uint32_t n_queries;
    int paired, mate_select;
    int k;
    int syncmer_s;              // > 0: closed syncmer with this s; 0: minimizer
    uint32_t minimizer_w;
    int scaled;
    uint64_t max_hash;
    int min_query_len;
};
cudaError_t launch_select(const SelectArgs &a, cudaStream_t st);
cudaError_t launch_slot_bounds(const uint64_t *seq_off, uint32_t n_seqs, int k, uint64_t *slot_cnt, cudaStream_t st);
cudaError_t launch_hash(const HashArgs &a, cudaStream_t st);
// plain k-mers of short reads (every mate at most HASH_GROUP_MAX_KMERS k-mers, no FracMinHash cut, not raw): eight lanes per query
constexpr uint32_t HASH_GROUP_MAX_KMERS = 512;
cudaError_t launch_hash_groups(const HashArgs &a, cudaStream_t st);
// long sequences (genomes, long reads): warp per 4096-position tile + gather, same results in the same order
constexpr uint32_t HASH_TILE_POS = 4096;
cudaError_t launch_tiles_per_seq(const uint64_t *seq_off, uint32_t n_seqs, int k, uint64_t *cnt, cudaStream_t st);
cudaError_t launch_hash_tiles(const HashArgs &a, uint32_t n_seqs, const uint64_t *tile_off, uint64_t max_tiles, uint64_t *tmp, uint32_t *tile_cnt,
                              cudaStream_t st);
cudaError_t launch_gather_tiles(const HashArgs &a, uint32_t n_seqs, const uint64_t *tile_off, const uint64_t *tile_pre, const uint32_t *tile_cnt,
                                const uint64_t *tmp, uint64_t max_tiles, cudaStream_t st);

// in-place unique of sorted regions longer than the dedup threshold (U:874-908), and the per-query verdict
struct FinalizeArgs {
    uint64_t *codes;
    const uint64_t *slot_off;   // per sequence
    uint32_t *n_codes;          // in: codes per query; out: after dedup
    int32_t *n_kmers_out;       // reported NumKmers (0 when skipped)
    uint32_t *n_eff;            // codes to probe (0 = skip)
    uint32_t *thresh;           // smallest count that passes min_matched and count > n*min_query_cov
    unsigned long long *n_sum;  // += Σ n_eff (algorithmic probe volume of the sub-batch)
    uint32_t n_queries;
    int paired;
    int dedup_threshold;
    int do_unique;              // regions with n > dedup_threshold are already sorted: unique them
    int min_matched;
    double min_query_cov;
};
cudaError_t launch_finalize(const FinalizeArgs &a, cudaStream_t st);
// segment descriptors for cub::DeviceSegmentedSort: begin/end of regions with n > min_n (else empty)
cudaError_t launch_sort_segments(const uint64_t *slot_off, const uint32_t *n_codes, uint32_t n_queries, int paired,
                                 int min_n, int *seg_begin, int *seg_end, cudaStream_t st);
// queries with dedup_threshold < n <= SMALL_DEDUP_MAX k-mers: sort + unique inside one warp (bitonic network in
// registers), in place; n_codes updated.  Covers paired-end 2x150 bp (n = 260) and reads up to ~2 kb.
constexpr int SMALL_DEDUP_MAX = 2048;
cudaError_t launch_small_dedup(uint64_t *codes, const uint64_t *slot_off, uint32_t *n_codes, uint32_t n_queries, int paired,
                               int dedup_threshold, int min_matched, uint64_t max_query_slots, cudaStream_t st);

// ---- kernel 2: the COBS probe of one block (hashValues + fastdiv.Mod + U:6613-7741) ----------------------------------------------
struct ProbeArgs {
    const uint8_t *rows;        // re-pitched bit matrix of the block in HBM
    uint32_t pitch;             // bytes between rows
    uint32_t row_bytes;         // bytes per row that carry data (un-padded numRowBytes)
    uint32_t lanes_per_task_override;   // 0: a task = up to 128 bytes of a row (8 lanes × 16 B or 16 lanes × 8 B); dev knob otherwise
    uint32_t n_names;
    uint32_t target_base;
    int num_hashes;
    const uint64_t *codes;      // k-mer codes of every query at slot_off[its first sequence]; the kernel derives the row indices itself:
    FastMod fm;                 // hashValues (H:125-141) + code % numSigs (fastdiv.Mod, U:6811), numSigs up to 2^64-1
    const uint32_t *locs;       // [slot][h] row indices from launch_locs (blocks with numSigs < 2^32-1; wider blocks derive them in the kernel)
    const uint64_t *slot_off;   // per sequence
    const uint32_t *n_eff;      // per query
    const uint32_t *thresh;     // per query
    uint32_t n_queries;
    int paired;
    uint64_t *hit_keys;         // query << target_bits | global target: only as many key bits as the hit sort has to look at
    int target_bits;            // bits of the largest global target index of the database (<= 32)
    uint32_t *hit_vals;         // matched k-mers
    unsigned long long *hit_count;
    unsigned long long *task_counter;   // zeroed before the launch: next task of the long-query kernels (planes > 8)
    const uint32_t *order;              // optional, long-query kernels: queries by descending number of k-mers (draw i works on query order[i / chunks])
    uint64_t hit_cap;
    uint32_t *dense_counts;     // optional [n_queries=1][n_targets] dump of all counts (kmcpg_count_codes)
    int planes;                 // counter bits: 8, 16, 24, 32
    int long_mode;              // set by launch_probe: one CTA per (query, chunk) for few long queries
};
cudaError_t launch_probe(const ProbeArgs &a, int sm_count, cudaStream_t st);
// the row indices of one block, locs[slot*H + h] (hashValues H:125-141 + fastdiv.Mod U:6811); the 64-bit form is the arithmetic test hook
cudaError_t launch_locs(const uint64_t *codes, uint64_t n_slots, int num_hashes, FastMod fm, uint32_t *locs, cudaStream_t st);
// same, walking the queries (sketch databases leave most slots empty)
cudaError_t launch_locs_by_query(const uint64_t *codes, const uint64_t *slot_off, const uint32_t *n_eff, uint32_t n_queries, int paired,
                                 int num_hashes, FastMod fm, uint32_t *locs, cudaStream_t st);
cudaError_t launch_locs64(const uint64_t *codes, uint64_t n_slots, int num_hashes, FastMod fm, uint64_t *locs, cudaStream_t st);

// ---- small utilities ------------------------------------------------------------------------------------
// rows (unpadded, row_bytes each) → dst with `pitch` bytes per row, zero padded
cudaError_t launch_repitch(const uint8_t *src, uint8_t *dst, uint64_t n_rows, uint32_t row_bytes, uint32_t pitch,
                           cudaStream_t st);
// the bytes [src_off, src_off + row_bytes) of every src row (src_stride bytes apart) → dst rows of `pitch` bytes
cudaError_t launch_repitch_cols(const uint8_t *src, uint8_t *dst, uint64_t n_rows, uint32_t src_stride, uint32_t src_off, uint32_t row_bytes,
                                uint32_t pitch, cudaStream_t st);
cudaError_t launch_unpitch(const uint8_t *src, uint8_t *dst, uint64_t n_rows, uint32_t row_bytes, uint32_t pitch,
                           cudaStream_t st);
cudaError_t launch_iota(uint32_t *v, uint32_t n, cudaStream_t st);
// sorted (key,val) pairs → kmcpg_hit records, query index rebased by query_base
cudaError_t launch_pack_hits(const uint64_t *keys, const uint32_t *vals, uint64_t n, uint32_t query_base, int target_bits, kmcpg_hit *out,
                             cudaStream_t st);

// ---- synthetic data + device index builder (synth.cu) ---------------------------------------------------
cudaError_t launch_synth_reads(uint64_t seed, uint64_t first, uint32_t n_reads, uint32_t read_len, uint64_t genome_seed,
                               uint32_t n_genomes, uint32_t genome_len, uint8_t *out, cudaStream_t st);
cudaError_t launch_synth_genome(uint64_t genome_seed, uint32_t genome, uint64_t start, uint64_t len, uint8_t *out,
                                cudaStream_t st);
// sets bit (7-(col&7)) of byte col>>3 in row loc for every code (I:1157 / I:1188)
cudaError_t launch_set_bits(const uint64_t *codes, uint64_t n, int num_hashes, FastMod fm, uint8_t *rows, uint32_t pitch,
                            uint32_t col, cudaStream_t st);

}  // namespace kmcpg

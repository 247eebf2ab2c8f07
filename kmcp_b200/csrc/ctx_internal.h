// ctx_internal.h — internals of libkmcp_gpu shared by ctx.cu and synth.cu (not part of the ABI).
#pragma once
#include <cuda_runtime.h>

#include <mutex>
#include <string>
#include <vector>

#include "common.h"
#include "kernels.cuh"

struct kmcpg_ctx;

namespace kmcpg {

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { (void)cudaGetLastError(); e = cudaMalloc(&p, bytes); want = bytes; }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return (T *)p; }
};

struct HostBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMallocHost(&p, bytes + bytes / 8 + 256);
        if (e == cudaSuccess) cap = bytes + bytes / 8 + 256;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return (T *)p; }
};

// one resident piece of a block: all of its columns (the usual case) or, when a DB has fewer blocks than shards, a
// column range [col0, col0+n_cols) of it (rows stay aligned, counts are per target: no reduction between shards)
struct DeviceBlock {
    int meta_idx = -1;
    uint8_t *d_rows = nullptr;
    size_t bytes = 0;
    uint32_t pitch = 0, row16 = 0, G = 1, chunks = 1;
    uint32_t col0 = 0, n_cols = 0;   // resident columns (targets) of block meta_idx; col0 is a multiple of 8
    uint32_t row_bytes = 0;          // (n_cols+7)/8: the un-padded bytes of one resident row
    FastMod fm;
    bool whole = true;               // every column of the block is resident here
};

struct ShardPiece { int block; int shard; uint32_t col0, n_cols; };

struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
};

// results of one search call live in pinned host memory borrowed from the context's pool
struct HitsPriv {
    kmcpg_ctx *ctx = nullptr;
    PinBuf nk, ql, hits;
    bool ext_hits = false;      // hits.p is the caller's buffer (kmcpg_batch.hits_dst): never grown, never returned to the pool
    uint32_t nq = 0;
    uint32_t first_query = 0;   // kmcpg_batch.first_query: added to the query index of every hit and part the job reports
    uint64_t nh = 0;
};

struct Executor;

// one sub-batch, inputs already on the device
struct SubBatch {
    const uint8_t *d_seq;
    const uint64_t *d_off;     // n_seqs+1, offsets into d_seq
    uint32_t n_seqs;
    uint64_t total_slots;      // sum of max(0, len-k+1)
    uint64_t max_query_slots;  // max over queries of its slot bound
    uint32_t query_base;       // index of the first query inside the caller's batch
};

// device + pinned buffers of one in-flight sub-batch; two sets let the host-side work of part i
// (hit count round trip, result copies) hide behind the kernels of part i+1
struct WorkSet {
    DevBuf seq, off, slot_cnt, slot_off, codes, codes2, locs[2], ncodes, qlen, nk, neff, thresh;
    DevBuf hkeys, hvals, hkeys2, hvals2, hits, counters, tmp, tmp2, segb, sege;
    DevBuf ck, cs, cs_cnt, cs_off;      // sketch selection: per-position k-mer / s-mer hashes
    DevBuf tile_n, tile_off, tile_cnt, tile_pre;   // long sequences: tiles per sequence, their scan, codes per tile, their scan
    DevBuf order, order_in, neff_sorted;           // long queries: query indices by descending number of k-mers (task order of the probe)
    bool order_valid = false;
    HostBuf h_off, h_cnt;
    cudaEvent_t ev_in = nullptr, ev_a0 = nullptr, ev_hash = nullptr, ev_a = nullptr, ev_cnt = nullptr, ev_sorted = nullptr, ev_b = nullptr;
    std::vector<cudaEvent_t> probe_ev;   // 3 slots per resident block: [0] its row indices are ready, [1] before, [2] after its probe launch
    // state of the part currently in flight
    bool busy = false;
    SubBatch sb{};
    uint32_t nq = 0;
    uint64_t cap = 0, n_hits = 0, hit_dst = 0;
    int planes = 8;
    uint64_t *codes_ptr = nullptr;
    void release();
};

}  // namespace kmcpg

struct kmcpg_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t st = nullptr;        // compute stream (own_st or the caller's)
    cudaStream_t own_st = nullptr;
    cudaStream_t copy_st = nullptr;   // device→host transfers that overlap the kernels
    cudaStream_t cnt_st = nullptr;    // the 16-byte hit counters of a part: its own stream, so that the wait for part i's probes does not
                                      // sit in front of part i-1's result copies (streams are FIFO)
    cudaStream_t post_st = nullptr;   // hit-list sort + pack of a finished part (tiny kernels, concurrent with the next probe)
    cudaStream_t in_st = nullptr;     // host→device input staging (its own queue, so it never waits behind result copies)
    cudaStream_t hash_st = nullptr;   // query preparation (slot scan, hash, dedup, verdict) of the NEXT part, beside the probes of the current one
    kmcpg::Executor *exec = nullptr;  // the thread that feeds the GPU (executor.cu); started by the first kmcpg_search_submit
    bool has_db = false;
    kmcpg::DbMeta meta;
    std::vector<kmcpg::DeviceBlock> blocks;
    std::vector<int> resident_of;     // meta block index -> index in `blocks` or -1
    std::vector<double> target_sizes; // Sizes[t] of every target as float64 (U:1393-1396 sizesFloat)
    int64_t sum_row_bytes = 0, resident_bytes = 0, disk_bytes = 0;
    int64_t part_row_bytes = 0;       // Σ row bytes of the WIDEST shard of the plan this context belongs to (part sizing, executor.cu)
    kmcpg::WorkSet ws[2];
    kmcpg::DevBuf d_tmp, d_dense, d_scal, d_genome;
    kmcpg::HostBuf h_stage, h_small;
    std::vector<kmcpg::PinBuf> pin_pool;
    std::vector<kmcpg::DevBuf> stage_pool;   // device buffers of batches staged for several contexts (kmcpg_engine_search_sharded)
    std::mutex pin_mu;
    std::string err;
    std::mutex mu;
    uint32_t launches = 0;
};

namespace kmcpg {

int fail(kmcpg_ctx *c, int code, const std::string &msg);
uint32_t pitch_for(uint32_t row_bytes);
void layout_block(DeviceBlock &b, const BlockMeta &m, uint32_t col0 = 0, uint32_t n_cols = 0);
void plan_pieces(const DbMeta &m, int world, std::vector<ShardPiece> &pieces, std::vector<uint64_t> &load);
int64_t widest_shard_row_bytes(const std::vector<ShardPiece> &pieces, int world);
void free_db(kmcpg_ctx *ctx);
void executor_drain(kmcpg_ctx *ctx);   // waits until no search job is in flight (call with ctx->mu held)
void executor_stop(kmcpg_ctx *ctx);
int run_hash_stage(kmcpg_ctx *ctx, WorkSet &w, const kmcpg_search_params &p, int k, const SubBatch &sb, uint32_t nq, uint64_t **codes_out, cudaStream_t st = nullptr);
int planes_for(uint64_t max_n);
int pin_acquire(kmcpg_ctx *ctx, size_t bytes, PinBuf &out);
void pin_release(kmcpg_ctx *ctx, PinBuf &b);

#define CU(call)                                                                                      \
    do {                                                                                              \
        cudaError_t _e = (call);                                                                      \
        if (_e != cudaSuccess) {                                                                      \
            (void)cudaGetLastError();                                                                 \
            return kmcpg::fail(ctx, _e == cudaErrorMemoryAllocation ? KMCPG_ENOMEM : KMCPG_ECUDA,     \
                        std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " + __FILE__ + ":" + std::to_string(__LINE__)); \
        }                                                                                             \
    } while (0)

}  // namespace kmcpg

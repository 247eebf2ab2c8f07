// ctx_internal.h — internals of libkmcp_gpu shared by ctx.cu and synth.cu (not part of the ABI).
#pragma once
#include <cuda_runtime.h>

#include <mutex>
#include <string>
#include <vector>

#include "common.h"
#include "kernels.cuh"

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

struct DeviceBlock {
    int meta_idx = -1;
    uint8_t *d_rows = nullptr;
    size_t bytes = 0;
    uint32_t pitch = 0, row16 = 0, G = 1, chunks = 1;
    FastMod fm;
};

struct HitsPriv {
    std::vector<int32_t> n_kmers, query_len;
    std::vector<kmcpg_hit> hits;
};

// one sub-batch, inputs already on the device
struct SubBatch {
    const uint8_t *d_seq;
    const uint64_t *d_off;     // n_seqs+1, offsets into d_seq
    uint32_t n_seqs;
    uint64_t total_slots;      // sum of max(0, len-k+1)
    uint64_t max_query_slots;  // max over queries of its slot bound
    uint32_t query_base;       // index of the first query inside the caller's batch
};

}  // namespace kmcpg

struct kmcpg_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t st = nullptr;
    cudaStream_t own_st = nullptr;
    std::vector<cudaEvent_t> probe_ev;   // 2 per probe launch of a sub-batch (+2 around the locs kernels)
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    bool has_db = false;
    kmcpg::DbMeta meta;
    std::vector<kmcpg::DeviceBlock> blocks;
    std::vector<int> resident_of;     // meta block index -> index in `blocks` or -1
    int64_t sum_row_bytes = 0, resident_bytes = 0, disk_bytes = 0;
    kmcpg::DevBuf d_seq, d_off, d_slot_cnt, d_slot_off, d_codes, d_codes2, d_locs, d_ncodes, d_qlen, d_nk, d_neff, d_thresh;
    kmcpg::DevBuf d_hkeys, d_hvals, d_hkeys2, d_hvals2, d_hits, d_hitcount, d_tmp, d_segb, d_sege, d_dense, d_scal, d_genome;
    kmcpg::HostBuf h_stage, h_off, h_small;
    std::string err;
    std::mutex mu;
    uint32_t launches = 0;
};

namespace kmcpg {

int fail(kmcpg_ctx *c, int code, const std::string &msg);
uint32_t pitch_for(uint32_t row_bytes);
void layout_block(DeviceBlock &b, const BlockMeta &m);
void free_db(kmcpg_ctx *ctx);
int run_hash_stage(kmcpg_ctx *ctx, const kmcpg_search_params &p, int k, const SubBatch &sb, uint32_t nq, uint64_t **codes_out);

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

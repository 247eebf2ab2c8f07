// common.h — shared host-side declarations of libkmcp_gpu (no CUDA types here).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/kmcp_gpu.h"

namespace kmcpg {

// One `.uniki` block header (reference: kmcp/cmd/index/serialization.go:60-80 Header, read by X:383-593).
struct BlockMeta {
    std::string path;
    int k = 0;
    bool canonical = false, compact = false;
    int num_hashes = 0;
    uint64_t num_sigs = 0;
    int n_names = 0;
    int row_bytes = 0;           // (n_names+7)/8  (X:379)
    uint64_t data_offset = 0;    // file position of row 0 (U:1207)
    std::vector<std::string> names;     // Target[0] of every column
    std::vector<uint32_t> indices;      // chunkIdx | nChunks<<16
    std::vector<uint64_t> gsizes, sizes;
    int64_t target_base = 0;     // global index of column 0
};

// `__db.yml` (reference: kmcp/cmd/util-db-info.go:46-79) + every block header.
struct DbMeta {
    std::string dir;
    int version = 0, index_version = 0;
    std::vector<int> ks;         // descending
    bool canonical = false, scaled = false, minimizer = false, syncmer = false;
    uint32_t scale = 0, minimizer_w = 0, syncmer_s = 0;
    int num_hashes = 0;
    double fpr = 0;
    std::vector<std::string> files;
    std::vector<BlockMeta> blocks;
    int64_t n_targets = 0;
};

// returns 0 or a KMCPG_E* code, message in err
int read_block_header(const std::string &path, BlockMeta &out, std::string &err);
int read_db_meta(const std::string &dir, DbMeta &out, std::string &err);
// X:153-304 writer (used by kmcpg_write_block and the device index builder)
int write_block_file(const std::string &path, const BlockMeta &m, const uint8_t *rows_unpadded, std::string &err);

// exactly `bytes` bytes at offset `off` of fd into dst, read as `threads` concurrent pread streams (the rows of a block come from
// the page cache or an NVMe drive: one stream reaches neither's bandwidth); false on an I/O error or a short file
bool pread_parallel(int fd, void *dst, size_t bytes, uint64_t off, int threads);

// FASTA/Q records of one file, plain or gzip, through the reader stage's parser and decoders (reader.cpp; usable from .cu files,
// which do not see fastx_reader.h): header = the whole header line without '>' / '@', seq = the sequence lines joined.
struct FastxFile;
FastxFile *fastx_open(const std::string &path);                                            // nullptr: cannot open
int fastx_next(FastxFile *f, std::string &header, std::string &seq, std::string &err);     // 1 a record, 0 end of file, -1 error (err)
void fastx_close(FastxFile *f);

// H:46-50 CalcSignatureSize
uint64_t calc_signature_size(uint64_t n_elements, int num_hashes, double fpr);
// F:32-50 QueryFPR, bit-exact with Go (math.Pow loop, big.Float prec-53 binomials)
double query_fpr(int n, int c, double p);
double go_pow(double x, double y);

// 128-bit reciprocal for exact x % d (replaces bmkessler/fastdiv Uint64.Mod; U:6611, U:6811)
struct FastMod {
    uint64_t d = 1, m_hi = 0, m_lo = 0;   // m = ceil(2^128 / d): Lemire's exact remainder (host side, fastmod_host)
    uint64_t b64 = ~0ull;                 // floor((2^64-1) / d): Barrett quotient estimate for 64-bit x, off by at most one (device side)
    uint32_t b32 = ~0u, _pad = 0;         // floor((2^32-1) / d) for 32-bit x when d < 2^32 (the h > 1 hash values are 32-bit, H:137-139)
};
FastMod make_fastmod(uint64_t d);
uint64_t fastmod_host(uint64_t a, const FastMod &f);

}  // namespace kmcpg

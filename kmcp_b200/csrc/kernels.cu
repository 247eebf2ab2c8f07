// kernels.cu — hand-written sm_100a kernels of the kmcp search hot path.
//
//   hash_kernel   : ASCII reads → canonical ntHash1 codes (+ FracMinHash cut, code>0), one warp per query.
//                   Replaces sketches.Iterator/nthash.NTHi driven by UnikIndexDB.generateKmers (U:1037-1107).
//   locs_kernel   : code → Bloom row index of one block: hashValues (H:125-141) + exact 64-bit modulo
//                   (fastdiv.Mod, U:6811) by a 128-bit reciprocal.
//   probe_kernel  : the COBS probe (U:6613-7741).  A task = (query, 16·G-byte column chunk of the block);
//                   G lanes own one 16-byte column slab each and keep the per-target match counters
//                   BIT-SLICED in registers (plane p = bit p of 128 targets' counts).  Rows are fetched with
//                   independent, coalesced 16-byte loads (8 rows in flight per lane), h rows are AND-ed in
//                   registers (pand), and 8 rows at a time go through a carry-save adder tree (Harley–Seal),
//                   which replaces the reference's byte transpose + pospop.Count8.  Thresholding
//                   (count >= min_matched, float64(count) > n·t) is evaluated bit-serially on the planes;
//                   only surviving targets are unpacked and appended to the hit list.
//
// Integer/bitset work, HBM-bandwidth bound: no tensor cores by design.
#include <cub/cub.cuh>

#include "kernels.cuh"

namespace kmcpg {

// ------------------------------------------------------------------------------------------------------
// ntHash1 seeds (will-rowe/nthash v0.4.0 == bcgsc ntHash 1.x)
// ------------------------------------------------------------------------------------------------------
#define SEED_A 0x3c8bfbb395c60474ULL
#define SEED_C 0x3193c18562a02b4cULL
#define SEED_G 0x20323ed082572324ULL
#define SEED_T 0x295549f54be24456ULL

__device__ __forceinline__ uint64_t rol1(uint64_t x) { return (x << 1) | (x >> 63); }
__device__ __forceinline__ uint64_t ror1(uint64_t x) { return (x >> 1) | (x << 63); }
__device__ __forceinline__ uint64_t rolv(uint64_t x, unsigned r) {
    r &= 63;
    return r ? (x << r) | (x >> (64 - r)) : x;
}

__device__ __forceinline__ uint64_t seed_fwd(uint8_t b) {
    // forward table: A/a C/c G/g T/t U/u, everything else 0
    switch (b | 0x20) {
        case 'a': return SEED_A;
        case 'c': return SEED_C;
        case 'g': return SEED_G;
        case 't': return SEED_T;
        case 'u': return SEED_T;
        default: return 0;
    }
}

__global__ void slot_bounds_kernel(const uint64_t *__restrict__ seq_off, uint32_t n_seqs, int k, uint64_t *__restrict__ cnt) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_seqs) {
        uint64_t len = seq_off[i + 1] - seq_off[i];
        cnt[i] = len >= (uint64_t)k ? len - k + 1 : 0;
    }
    if (i == n_seqs) cnt[i] = 0;
}

cudaError_t launch_slot_bounds(const uint64_t *seq_off, uint32_t n_seqs, int k, uint64_t *slot_cnt, cudaStream_t st) {
    uint32_t n = n_seqs + 1;
    slot_bounds_kernel<<<(n + 255) / 256, 256, 0, st>>>(seq_off, n_seqs, k, slot_cnt);
    return cudaGetLastError();
}

// shared seed tables: F[256], Fk[256] = rol(F,k); R[8], R1[8] = ror(R,1), Rk1[8] = rol(R,k-1)
struct SeedTables {
    uint64_t F[256], Fk[256], R[8], R1[8], Rk1[8];
};

constexpr int HASH_WARPS = 8;
constexpr int HASH_RUN = 8;   // max consecutive positions rolled by one lane

__global__ void __launch_bounds__(HASH_WARPS * 32) hash_kernel(HashArgs a) {
    __shared__ SeedTables T;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        uint64_t f = seed_fwd((uint8_t)i);
        // entries 0..7 double as the reverse-strand table seedTab[b & 7] = {N,T,N,G,A,A,N,C}
        if (i < 8) {
            const uint64_t r8[8] = {0, SEED_T, 0, SEED_G, SEED_A, SEED_A, 0, SEED_C};
            f = r8[i];
            T.R[i] = f;
            T.R1[i] = ror1(f);
            T.Rk1[i] = rolv(f, (unsigned)(a.k - 1));
        }
        T.F[i] = f;
        T.Fk[i] = rolv(f, (unsigned)a.k);
    }
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    const int k = a.k;

    for (uint32_t q = warp; q < a.n_queries; q += n_warps) {
        const uint32_t s0 = (a.paired && !a.raw) ? 2 * q : q;
        const int n_mates = (a.paired && !a.raw) ? 2 : 1;
        const uint64_t len0 = a.seq_off[s0 + 1] - a.seq_off[s0];
        const uint64_t len1 = (a.paired && !a.raw) ? a.seq_off[s0 + 2] - a.seq_off[s0 + 1] : 0;
        uint64_t *out = a.codes + a.slot_off[s0];
        uint32_t written = 0;
        // U:778-786: skip when Seq is shorter than min-query-len unless Seq2 is long enough
        bool skip = !a.raw && (int64_t)len0 < a.min_query_len && !(a.paired && (int64_t)len1 >= a.min_query_len);
        int32_t qlen = (int32_t)(len0 + len1);
        if (a.mate_select == 1) qlen = (int32_t)len0;
        if (a.mate_select == 2) qlen = (int32_t)len1;
        if (!skip) {
            for (int m = 0; m < n_mates; m++) {
                if (a.mate_select == 1 && m == 1) continue;
                if (a.mate_select == 2 && m == 0) continue;
                const uint8_t *s = a.seq + a.seq_off[s0 + m];
                const uint64_t len = m == 0 ? len0 : len1;
                if (len < (uint64_t)k) continue;                       // sketches.ErrShortSeq (U:1059-1062)
                const uint64_t nk = len - k + 1;
                uint64_t run = (nk + 31) / 32;
                if (run > HASH_RUN) run = HASH_RUN;
                for (uint64_t base = 0; base < nk; base += 32 * run) {
                    const uint64_t p0 = base + (uint64_t)lane * run;
                    int cnt = 0;
                    if (p0 < nk) cnt = (int)((nk - p0) < run ? (nk - p0) : run);
                    uint64_t c[HASH_RUN];
                    uint32_t valid = 0;
                    if (cnt > 0) {
                        // one pass over the k bases: fwd = XOR_j rol(F[s_j], k-1-j) by Horner (rotate left),
                        // rev = XOR_j rol(R[s_j&7], j) by rotating the accumulator right and adding rol(R, k-1)
                        uint64_t fh = 0, rh = 0;
#pragma unroll 4
                        for (int j = 0; j < k; j++) {
                            const uint8_t b = s[p0 + j];
                            fh = rol1(fh) ^ T.F[b];
                            rh = ror1(rh) ^ T.Rk1[b & 7];
                        }
#pragma unroll
                        for (int r = 0; r < HASH_RUN; r++) {
                            if (r < cnt) {
                                uint64_t code = a.canonical ? (fh < rh ? fh : rh) : fh;
                                bool ok = a.raw || (code != 0 && !(a.scaled && code > a.max_hash));       // U:1097-1102
                                c[r] = code;
                                if (ok) valid |= 1u << r;
                                if (r + 1 < cnt) {
                                    uint8_t bo = s[p0 + r], bi = s[p0 + r + k];
                                    fh = rol1(fh) ^ T.Fk[bo] ^ T.F[bi];
                                    rh = ror1(rh) ^ T.R1[bo & 7] ^ T.Rk1[bi & 7];
                                }
                            }
                        }
                    }
                    if (a.raw) {                                    // position-indexed output for the sketch selection
#pragma unroll
                        for (int r = 0; r < HASH_RUN; r++)
                            if (valid & (1u << r)) out[p0 + r] = c[r];
                        continue;
                    }
                    // in-order compaction inside the query's code region
                    int mine = __popc(valid);
                    int incl = mine;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        int t = __shfl_up_sync(0xffffffffu, incl, d);
                        if (lane >= d) incl += t;
                    }
                    int total = __shfl_sync(0xffffffffu, incl, 31);
                    uint32_t w = written + (uint32_t)(incl - mine);
#pragma unroll
                    for (int r = 0; r < HASH_RUN; r++)
                        if (valid & (1u << r)) out[w++] = c[r];
                    written += (uint32_t)total;
                }
            }
        }
        if (lane == 0 && !a.raw) {
            a.n_codes[q] = skip ? 0xFFFFFFFFu : written;    // 0xFFFFFFFF marks "skipped by length" (NumKmers = 0)
            a.query_len[q] = qlen;
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// short reads, plain k-mers (no FracMinHash cut, no sketch selection): EIGHT lanes per query, four queries per warp.  With a whole
// warp on a 150 bp read every lane pays the k-step Horner initialisation for a run of only 5 k-mers (21 of 26 steps are set-up);
// eight lanes roll runs of 17, so a warp spends 2.3x fewer instructions per read.  Every k-mer code is kept unless it is 0
// (U:1097-1102), which a 64-bit hash practically never is: codes go straight to their position, and a mate that did produce a
// zero code is compacted afterwards by its first lane.  Same codes in the same order as hash_kernel.
// ------------------------------------------------------------------------------------------------------
constexpr int HG_LANES = 8;
constexpr uint32_t HG_MAX_KMERS = 512;      // per mate: runs of at most 64 k-mers per lane

__global__ void __launch_bounds__(HASH_WARPS * 32) hash_group_kernel(HashArgs a) {
    __shared__ SeedTables T;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        uint64_t f = seed_fwd((uint8_t)i);
        if (i < 8) {
            const uint64_t r8[8] = {0, SEED_T, 0, SEED_G, SEED_A, SEED_A, 0, SEED_C};
            f = r8[i];
            T.R[i] = f; T.R1[i] = ror1(f); T.Rk1[i] = rolv(f, (unsigned)(a.k - 1));
        }
        T.F[i] = f;
        T.Fk[i] = rolv(f, (unsigned)a.k);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, gl = lane & (HG_LANES - 1), gbase = lane & ~(HG_LANES - 1);
    const uint32_t gmask = ((1u << HG_LANES) - 1u) << gbase;
    const uint32_t group = (blockIdx.x * blockDim.x + threadIdx.x) / HG_LANES;
    const uint32_t n_groups = (gridDim.x * blockDim.x) / HG_LANES;
    const int k = a.k;
    for (uint32_t q = group; q < a.n_queries; q += n_groups) {
        const uint32_t s0 = a.paired ? 2 * q : q;
        const int n_mates = a.paired ? 2 : 1;
        const uint64_t len0 = a.seq_off[s0 + 1] - a.seq_off[s0];
        const uint64_t len1 = a.paired ? a.seq_off[s0 + 2] - a.seq_off[s0 + 1] : 0;
        uint64_t *out = a.codes + a.slot_off[s0];
        uint32_t written = 0;
        const bool skip = (int64_t)len0 < a.min_query_len && !(a.paired && (int64_t)len1 >= a.min_query_len);      // U:778-786
        int32_t qlen = (int32_t)(len0 + len1);
        if (a.mate_select == 1) qlen = (int32_t)len0;
        if (a.mate_select == 2) qlen = (int32_t)len1;
        if (!skip) {
            for (int m = 0; m < n_mates; m++) {
                if (a.mate_select == 1 && m == 1) continue;
                if (a.mate_select == 2 && m == 0) continue;
                const uint8_t *s = a.seq + a.seq_off[s0 + m];
                const uint64_t len = m == 0 ? len0 : len1;
                if (len < (uint64_t)k) continue;                       // sketches.ErrShortSeq (U:1059-1062)
                const uint32_t nk = (uint32_t)(len - k + 1);
                const uint32_t run = (nk + HG_LANES - 1) / HG_LANES;
                const uint32_t p0 = (uint32_t)gl * run;
                const uint32_t cnt = p0 < nk ? (nk - p0 < run ? nk - p0 : run) : 0;
                bool zero = false;
                if (cnt > 0) {
                    uint64_t fh = 0, rh = 0;
#pragma unroll 4
                    for (int j = 0; j < k; j++) {
                        const uint8_t b = s[p0 + j];
                        fh = rol1(fh) ^ T.F[b];
                        rh = ror1(rh) ^ T.Rk1[b & 7];
                    }
                    uint64_t *o = out + written + p0;
#pragma unroll 4
                    for (uint32_t r = 0; r < cnt; r++) {
                        const uint64_t code = a.canonical ? (fh < rh ? fh : rh) : fh;
                        o[r] = code;
                        zero |= code == 0;
                        if (r + 1 < cnt) {
                            const uint8_t bo = s[p0 + r], bi = s[p0 + r + k];
                            fh = rol1(fh) ^ T.Fk[bo] ^ T.F[bi];
                            rh = ror1(rh) ^ T.R1[bo & 7] ^ T.Rk1[bi & 7];
                        }
                    }
                }
                uint32_t kept = nk;
                if (__any_sync(gmask, zero)) {                          // practically never: drop the zero codes, in order
                    __syncwarp(gmask);
                    uint32_t w = 0;
                    if (gl == 0) {
                        uint64_t *o = out + written;
                        for (uint32_t i = 0; i < nk; i++) { const uint64_t c = o[i]; if (c) o[w++] = c; }
                    }
                    kept = __shfl_sync(gmask, w, gbase);
                }
                written += kept;
            }
        }
        if (gl == 0) {
            a.n_codes[q] = skip ? 0xFFFFFFFFu : written;                // 0xFFFFFFFF marks "skipped by length" (NumKmers = 0)
            a.query_len[q] = qlen;
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// minimizer / closed syncmer selection: one warp per query, 32 windows per round, every lane scans its window
// for the LEFTMOST minimum (bio/sketches keeps a sorted buffer whose ties stay in arrival order)
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) select_kernel(SelectArgs a) {
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    const int k = a.k, s = a.syncmer_s;
    for (uint32_t q = warp; q < a.n_queries; q += n_warps) {
        const uint32_t s0 = a.paired ? 2 * q : q;
        const int n_mates = a.paired ? 2 : 1;
        const uint64_t len0 = a.seq_off[s0 + 1] - a.seq_off[s0];
        const uint64_t len1 = a.paired ? a.seq_off[s0 + 2] - a.seq_off[s0 + 1] : 0;
        uint64_t *out = a.codes + a.slot_off[s0];
        uint32_t written = 0;
        bool skip = (int64_t)len0 < a.min_query_len && !(a.paired && (int64_t)len1 >= a.min_query_len);   // U:778-786
        int32_t qlen = (int32_t)(len0 + len1);
        if (a.mate_select == 1) qlen = (int32_t)len0;
        if (a.mate_select == 2) qlen = (int32_t)len1;
        if (!skip) {
            for (int m = 0; m < n_mates; m++) {
                if (a.mate_select == 1 && m == 1) continue;
                if (a.mate_select == 2 && m == 0) continue;
                const uint64_t len = m == 0 ? len0 : len1;
                if (len < (uint64_t)k) continue;
                const uint64_t *ck = a.ck + a.slot_off[s0 + m];
                uint64_t n_win, W;
                int64_t kms = 0;
                const uint64_t *src;
                if (s > 0) {                       // windows of L = 2k-s-1 bases holding W = 2(k-s) s-mers (A.4)
                    const uint64_t L = 2ull * k - s - 1;
                    if (len < L) continue;
                    n_win = len - L + 1; W = 2ull * (k - s); kms = k - s;
                    src = a.cs + a.cs_off[s0 + m];
                } else {                           // windows of w consecutive k-mers (A.5)
                    const uint64_t nk = len - k + 1;
                    W = a.minimizer_w < 1 ? 1 : a.minimizer_w;
                    if (nk < W) continue;
                    n_win = nk - W + 1;
                    src = ck;
                }
                uint64_t prev_last = ~0ull;        // selected position of the window before this round
                for (uint64_t base = 0; base < n_win; base += 32) {
                    const uint64_t idx = base + lane;
                    uint64_t pos = ~0ull;
                    if (idx < n_win) {
                        uint64_t best = src[idx], bi = 0;
                        for (uint64_t t = 1; t < W; t++) {
                            uint64_t v = src[idx + t];
                            if (v < best) { best = v; bi = t; }     // strict <: leftmost minimum
                        }
                        pos = s > 0 ? ((int64_t)bi < kms ? idx + bi : idx + bi - (uint64_t)kms) : idx + bi;
                    }
                    uint64_t before = __shfl_up_sync(0xffffffffu, pos, 1);
                    if (lane == 0) before = prev_last;
                    bool emit = idx < n_win && pos != before;         // runs of equal positions are contiguous
                    uint64_t code = 0;
                    if (emit) {
                        code = ck[pos];
                        if (code == 0 || (a.scaled && code > a.max_hash)) emit = false;     // U:1071-1076 / 1084-1089
                    }
                    const int last_active = (int)((n_win - base) < 32 ? (n_win - base - 1) : 31);
                    prev_last = __shfl_sync(0xffffffffu, pos, last_active);
                    __syncwarp();
                    uint32_t mask = __ballot_sync(0xffffffffu, emit);
                    if (emit) out[written + __popc(mask & ((1u << lane) - 1))] = code;
                    written += __popc(mask);
                }
            }
        }
        if (lane == 0) {
            a.n_codes[q] = skip ? 0xFFFFFFFFu : written;
            a.query_len[q] = qlen;
        }
    }
}

cudaError_t launch_select(const SelectArgs &a, cudaStream_t st) {
    if (a.n_queries == 0) return cudaSuccess;
    uint32_t blocks = (a.n_queries + 7) / 8;
    if (blocks > 148u * 32u) blocks = 148u * 32u;
    select_kernel<<<blocks, 256, 0, st>>>(a);
    return cudaGetLastError();
}

// ---- long sequences: one warp per TILE of positions instead of one warp per query --------------------------------
// tile t of sequence s covers k-mer positions [j*TILE_POS, (j+1)*TILE_POS); its codes are compacted at the start of the tile's
// own slot range of a scratch array, then gathered per query in tile order (= position order) after a scan of the tile counts.
constexpr uint32_t TILE_POS = 4096;

__global__ void tiles_per_seq_kernel(const uint64_t *__restrict__ seq_off, uint32_t n_seqs, int k, uint64_t *__restrict__ cnt) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_seqs) {
        uint64_t len = seq_off[i + 1] - seq_off[i];
        uint64_t nk = len >= (uint64_t)k ? len - k + 1 : 0;
        cnt[i] = (nk + TILE_POS - 1) / TILE_POS;
    }
    if (i == n_seqs) cnt[i] = 0;
}

__device__ __forceinline__ uint32_t seq_of_tile(const uint64_t *__restrict__ tile_off, uint32_t n_seqs, uint64_t t) {
    uint32_t lo = 0, hi = n_seqs;                      // largest s with tile_off[s] <= t
    while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (tile_off[mid] <= t) lo = mid; else hi = mid; }
    return lo;
}

__global__ void __launch_bounds__(HASH_WARPS * 32) hash_tile_kernel(HashArgs a, uint32_t n_seqs, const uint64_t *__restrict__ tile_off, uint64_t *__restrict__ tmp,
                                                                    uint32_t *__restrict__ tile_cnt) {
    __shared__ SeedTables T;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        uint64_t f = seed_fwd((uint8_t)i);
        if (i < 8) {
            const uint64_t r8[8] = {0, SEED_T, 0, SEED_G, SEED_A, SEED_A, 0, SEED_C};
            f = r8[i];
            T.R[i] = f; T.R1[i] = ror1(f); T.Rk1[i] = rolv(f, (unsigned)(a.k - 1));
        }
        T.F[i] = f;
        T.Fk[i] = rolv(f, (unsigned)a.k);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const int k = a.k;
    const uint64_t n_tiles = tile_off[n_seqs];
    for (uint64_t t = warp; t < n_tiles; t += n_warps) {
        const uint32_t sq = seq_of_tile(tile_off, n_seqs, t);
        const uint64_t j = t - tile_off[sq];
        const uint8_t *s = a.seq + a.seq_off[sq];
        const uint64_t len = a.seq_off[sq + 1] - a.seq_off[sq];
        const uint64_t nk = len - k + 1;                                // a sequence with tiles has len >= k
        const uint64_t p_begin = j * TILE_POS, p_end = (p_begin + TILE_POS < nk) ? p_begin + TILE_POS : nk;
        // query-level rules (U:778-786, mate selection of --try-se)
        bool drop = false;
        if (!a.raw) {
            const uint32_t s0 = a.paired ? (sq & ~1u) : sq;
            const uint64_t len0 = a.seq_off[s0 + 1] - a.seq_off[s0];
            const uint64_t len1 = a.paired ? a.seq_off[s0 + 2] - a.seq_off[s0 + 1] : 0;
            drop = (int64_t)len0 < a.min_query_len && !(a.paired && (int64_t)len1 >= a.min_query_len);
            if (a.paired && a.mate_select == 1 && (sq & 1)) drop = true;
            if (a.paired && a.mate_select == 2 && !(sq & 1)) drop = true;
        }
        uint64_t *out = (a.raw ? a.codes : tmp) + a.slot_off[sq] + p_begin;
        uint32_t written = 0;
        if (!drop) {
            for (uint64_t base = p_begin; base < p_end; base += 32 * HASH_RUN) {
                const uint64_t p0 = base + (uint64_t)lane * HASH_RUN;
                int cnt = 0;
                if (p0 < p_end) cnt = (int)((p_end - p0) < HASH_RUN ? (p_end - p0) : HASH_RUN);
                uint64_t c[HASH_RUN];
                uint32_t valid = 0;
                if (cnt > 0) {
                    uint64_t fh = 0, rh = 0;
#pragma unroll 4
                    for (int jj = 0; jj < k; jj++) {
                        const uint8_t b = s[p0 + jj];
                        fh = rol1(fh) ^ T.F[b];
                        rh = ror1(rh) ^ T.Rk1[b & 7];
                    }
#pragma unroll
                    for (int r = 0; r < HASH_RUN; r++) {
                        if (r < cnt) {
                            uint64_t code = a.canonical ? (fh < rh ? fh : rh) : fh;
                            bool ok = a.raw || (code != 0 && !(a.scaled && code > a.max_hash));
                            c[r] = code;
                            if (ok) valid |= 1u << r;
                            if (r + 1 < cnt) {
                                uint8_t bo = s[p0 + r], bi = s[p0 + r + k];
                                fh = rol1(fh) ^ T.Fk[bo] ^ T.F[bi];
                                rh = ror1(rh) ^ T.R1[bo & 7] ^ T.Rk1[bi & 7];
                            }
                        }
                    }
                }
                if (a.raw) {
#pragma unroll
                    for (int r = 0; r < HASH_RUN; r++)
                        if (valid & (1u << r)) out[(p0 - p_begin) + r] = c[r];
                    continue;
                }
                int mine = __popc(valid);
                int incl = mine;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    int tt = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += tt;
                }
                int total = __shfl_sync(0xffffffffu, incl, 31);
                uint32_t w = written + (uint32_t)(incl - mine);
#pragma unroll
                for (int r = 0; r < HASH_RUN; r++)
                    if (valid & (1u << r)) out[w++] = c[r];
                written += (uint32_t)total;
            }
        }
        if (lane == 0 && !a.raw) tile_cnt[t] = written;
    }
}

// tile-compacted codes → contiguous per query, in position order; n_codes / query_len of every query
__global__ void __launch_bounds__(256) gather_tiles_kernel(HashArgs a, uint32_t n_seqs, const uint64_t *__restrict__ tile_off, const uint64_t *__restrict__ tile_pre,
                                                           const uint32_t *__restrict__ tile_cnt, const uint64_t *__restrict__ tmp) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint64_t n_tiles = tile_off[n_seqs];
    for (uint64_t t = warp; t < n_tiles; t += n_warps) {
        const uint32_t sq = seq_of_tile(tile_off, n_seqs, t);
        const uint32_t s0 = a.paired ? (sq & ~1u) : sq;
        const uint64_t j = t - tile_off[sq];
        const uint64_t *src = tmp + a.slot_off[sq] + j * TILE_POS;
        uint64_t *dst = a.codes + a.slot_off[s0] + (tile_pre[t] - tile_pre[tile_off[s0]]);
        const uint32_t n = tile_cnt[t];
        for (uint32_t i = lane; i < n; i += 32) dst[i] = src[i];
    }
    // per query bookkeeping
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < a.n_queries; q += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t s0 = a.paired ? 2 * (uint32_t)q : (uint32_t)q;
        const uint32_t s1 = s0 + (a.paired ? 2 : 1);
        const uint64_t len0 = a.seq_off[s0 + 1] - a.seq_off[s0];
        const uint64_t len1 = a.paired ? a.seq_off[s0 + 2] - a.seq_off[s0 + 1] : 0;
        const bool skip = (int64_t)len0 < a.min_query_len && !(a.paired && (int64_t)len1 >= a.min_query_len);
        int32_t qlen = (int32_t)(len0 + len1);
        if (a.mate_select == 1) qlen = (int32_t)len0;
        if (a.mate_select == 2) qlen = (int32_t)len1;
        a.n_codes[q] = skip ? 0xFFFFFFFFu : (uint32_t)(tile_pre[tile_off[s1]] - tile_pre[tile_off[s0]]);
        a.query_len[q] = qlen;
    }
}

cudaError_t launch_tiles_per_seq(const uint64_t *seq_off, uint32_t n_seqs, int k, uint64_t *cnt, cudaStream_t st) {
    uint32_t n = n_seqs + 1;
    tiles_per_seq_kernel<<<(n + 255) / 256, 256, 0, st>>>(seq_off, n_seqs, k, cnt);
    return cudaGetLastError();
}

cudaError_t launch_hash_tiles(const HashArgs &a, uint32_t n_seqs, const uint64_t *tile_off, uint64_t max_tiles, uint64_t *tmp, uint32_t *tile_cnt, cudaStream_t st) {
    if (!n_seqs || !max_tiles) return cudaSuccess;
    uint64_t blocks64 = (max_tiles + HASH_WARPS - 1) / HASH_WARPS;
    uint32_t blocks = blocks64 > 148u * 32u ? 148u * 32u : (uint32_t)blocks64;
    hash_tile_kernel<<<blocks, HASH_WARPS * 32, 0, st>>>(a, n_seqs, tile_off, tmp, tile_cnt);
    return cudaGetLastError();
}

cudaError_t launch_gather_tiles(const HashArgs &a, uint32_t n_seqs, const uint64_t *tile_off, const uint64_t *tile_pre, const uint32_t *tile_cnt,
                                const uint64_t *tmp, uint64_t max_tiles, cudaStream_t st) {
    if (!n_seqs) return cudaSuccess;
    uint64_t blocks64 = (std::max<uint64_t>(max_tiles, a.n_queries / 8 + 1) + 7) / 8;
    uint32_t blocks = blocks64 > 148u * 32u ? 148u * 32u : (uint32_t)blocks64;
    gather_tiles_kernel<<<blocks, 256, 0, st>>>(a, n_seqs, tile_off, tile_pre, tile_cnt, tmp);
    return cudaGetLastError();
}

cudaError_t launch_hash_groups(const HashArgs &a, cudaStream_t st) {
    if (a.n_queries == 0) return cudaSuccess;
    const uint32_t per_block = HASH_WARPS * 32 / HG_LANES;
    uint32_t blocks = (a.n_queries + per_block - 1) / per_block;
    if (blocks > 148u * 64u) blocks = 148u * 64u;
    hash_group_kernel<<<blocks, HASH_WARPS * 32, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_hash(const HashArgs &a, cudaStream_t st) {
    if (a.n_queries == 0) return cudaSuccess;
    uint32_t blocks = (a.n_queries + HASH_WARPS - 1) / HASH_WARPS;
    if (blocks > 148u * 64u) blocks = 148u * 64u;
    hash_kernel<<<blocks, HASH_WARPS * 32, 0, st>>>(a);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------
// dedup + per-query verdict
// ------------------------------------------------------------------------------------------------------
__global__ void sort_segments_kernel(const uint64_t *__restrict__ slot_off, const uint32_t *__restrict__ n_codes, uint32_t nq,
                                     int paired, int thr, int *__restrict__ seg_begin, int *__restrict__ seg_end) {
    uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    uint64_t b = slot_off[paired ? 2 * q : q];
    uint32_t n = n_codes[q];
    if (n == 0xFFFFFFFFu || (n & 0x80000000u)) n = 0;                  // skipped, or already handled by small_dedup_kernel
    seg_begin[q] = (int)b;
    seg_end[q] = (int)((n > (uint32_t)thr) ? b + n : b);      // strict > (U:874)
}

cudaError_t launch_sort_segments(const uint64_t *slot_off, const uint32_t *n_codes, uint32_t nq, int paired, int thr,
                                 int *seg_begin, int *seg_end, cudaStream_t st) {
    if (!nq) return cudaSuccess;
    sort_segments_kernel<<<(nq + 255) / 256, 256, 0, st>>>(slot_off, n_codes, nq, paired, thr, seg_begin, seg_end);
    return cudaGetLastError();
}

// ---- warp-level sort + unique for mid-sized queries (U:874-908) ----
// element i of the N = 32*ITEMS keys lives in register i/32 of lane i%32; compare-exchange partners at distance
// j < 32 sit in another lane (shuffle), at distance j >= 32 in another register of the same lane.
template <int ITEMS>
__device__ __forceinline__ uint32_t warp_sort_unique(uint64_t *c, uint32_t n, int lane) {
    uint64_t v[ITEMS];
#pragma unroll
    for (int r = 0; r < ITEMS; r++) { const uint32_t i = r * 32 + lane; v[r] = i < n ? c[i] : ~0ull; }
    constexpr int N = ITEMS * 32;
#pragma unroll
    for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 32) {
                const int rr = j >> 5;
#pragma unroll
                for (int r = 0; r < ITEMS; r++) {
                    if ((r & rr) == 0) {
                        const bool asc = (((r * 32) & k) == 0);            // k >= 64 here: depends on the register index only
                        const uint64_t a = v[r], b = v[r | rr];
                        const uint64_t lo = a < b ? a : b, hi = a < b ? b : a;
                        v[r] = asc ? lo : hi; v[r | rr] = asc ? hi : lo;
                    }
                }
            } else {
                const bool lower = (lane & j) == 0;
#pragma unroll
                for (int r = 0; r < ITEMS; r++) {
                    const uint32_t i = r * 32 + lane;
                    const bool asc = ((i & k) == 0);
                    const uint64_t o = __shfl_xor_sync(0xffffffffu, v[r], j);
                    const uint64_t mn = v[r] < o ? v[r] : o, mx = v[r] < o ? o : v[r];
                    v[r] = (asc == lower) ? mn : mx;
                }
            }
        }
    }
    // unique, written back in ascending order
    __syncwarp();
    uint32_t w = 0;
#pragma unroll
    for (int r = 0; r < ITEMS; r++) {
        const uint32_t i = r * 32 + lane;
        uint64_t prev = __shfl_up_sync(0xffffffffu, v[r], 1);
        if (r > 0) { const uint64_t pl = __shfl_sync(0xffffffffu, v[r - 1], 31); if (lane == 0) prev = pl; }
        const bool keep = i < n && (i == 0 || v[r] != prev);
        const uint32_t m = __ballot_sync(0xffffffffu, keep);
        if (keep) c[w + __popc(m & ((1u << lane) - 1))] = v[r];
        w += __popc(m);
    }
    return w;
}

template <int ITEMS>
__global__ void __launch_bounds__(128) small_dedup_kernel(uint64_t *__restrict__ codes, const uint64_t *__restrict__ slot_off, uint32_t *__restrict__ n_codes,
                                                          uint32_t nq, int paired, uint32_t lo, uint32_t hi, uint32_t min_matched) {
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t q = warp; q < nq; q += n_warps) {
        const uint32_t n = n_codes[q];
        if (n == 0xFFFFFFFFu || n <= lo || n > hi) continue;              // this launch handles lo < n <= hi
        if (n < min_matched) continue;                                    // unmatched by U:854-869, which looks at the count BEFORE dedup
        uint64_t *c = codes + slot_off[paired ? 2 * q : q];
        const uint32_t m = warp_sort_unique<ITEMS>(c, n, lane);
        if (lane == 0) n_codes[q] = m | 0x80000000u;                      // flag: already deduplicated, and it passed the min-kmers gate
    }
}

cudaError_t launch_small_dedup(uint64_t *codes, const uint64_t *slot_off, uint32_t *n_codes, uint32_t nq, int paired, int thr, int min_matched,
                               uint64_t max_n, cudaStream_t st) {
    if (!nq) return cudaSuccess;
    uint32_t blocks = (nq + 3) / 4;
    if (blocks > 148u * 32u) blocks = 148u * 32u;
    const uint32_t t = (uint32_t)thr;                                      // strict > (U:874)
    // one launch per size class that can occur in this part (register budget differs by class)
    if (t < 512 && max_n > t) small_dedup_kernel<16><<<blocks, 128, 0, st>>>(codes, slot_off, n_codes, nq, paired, t, 512, (uint32_t)min_matched);
    if (t < 1024 && max_n > 512) small_dedup_kernel<32><<<blocks, 128, 0, st>>>(codes, slot_off, n_codes, nq, paired, t > 512 ? t : 512, 1024, (uint32_t)min_matched);
    if (t < 2048 && max_n > 1024) small_dedup_kernel<64><<<blocks, 128, 0, st>>>(codes, slot_off, n_codes, nq, paired, t > 1024 ? t : 1024, 2048, (uint32_t)min_matched);
    return cudaGetLastError();
}

// one warp per query: unique of an ascending region in place (U:878-908), then n_eff / threshold
__global__ void __launch_bounds__(256) finalize_kernel(FinalizeArgs a) {
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    unsigned long long local_sum = 0;
    for (uint32_t q = warp; q < a.n_queries; q += n_warps) {
        uint32_t n = a.n_codes[q];
        bool skipped = n == 0xFFFFFFFFu;
        if (skipped) n = 0;
        const bool deduped = !skipped && (n & 0x80000000u);                // small_dedup_kernel already sorted + uniqued this query
        n &= 0x7FFFFFFFu;
        // U:854-869: fewer k-mers than min-matched → unmatched, checked BEFORE dedup
        bool too_few = !deduped && (int64_t)n < (int64_t)a.min_matched;
        if (!skipped && !too_few && !deduped && a.do_unique && n > (uint32_t)a.dedup_threshold) {
            uint64_t *c = a.codes + a.slot_off[a.paired ? 2 * q : q];
            uint32_t w = 0;
            for (uint32_t base = 0; base < n; base += 32) {
                uint32_t i = base + lane;
                uint64_t v = 0, prev = 0;
                bool keep = false;
                if (i < n) {
                    v = c[i];
                    keep = (i == 0);
                    if (i > 0) { prev = c[i - 1]; keep = v != prev; }
                }
                __syncwarp();                                  // all reads of this round before any write
                uint32_t m = __ballot_sync(0xffffffffu, keep);
                if (keep) c[w + __popc(m & ((1u << lane) - 1))] = v;
                w += __popc(m);
                __syncwarp();
            }
            n = w;
        }
        if (lane == 0) {
            uint32_t eff = (skipped || too_few) ? 0 : n;
            a.n_codes[q] = n;
            a.n_eff[q] = eff;
            a.n_kmers_out[q] = (int32_t)eff;
            // smallest integer count with float64(count) > float64(n)*t (U:6625, U:7469), and >= min_matched (U:7466)
            double x = (double)n * a.min_query_cov;
            double fl = floor(x);
            uint32_t t = fl >= 4294967294.0 ? 0xFFFFFFFFu : (uint32_t)fl + 1;
            if ((int64_t)t < (int64_t)a.min_matched) t = (uint32_t)a.min_matched;
            if (t == 0) t = 1;
            a.thresh[q] = t;
            local_sum += eff;
        }
    }
    if (lane == 0 && local_sum && a.n_sum) atomicAdd(a.n_sum, local_sum);
}

cudaError_t launch_finalize(const FinalizeArgs &a, cudaStream_t st) {
    if (!a.n_queries) return cudaSuccess;
    uint32_t blocks = (a.n_queries + 7) / 8;
    if (blocks > 148u * 32u) blocks = 148u * 32u;
    finalize_kernel<<<blocks, 256, 0, st>>>(a);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------
// row indices: hashValues (H:125-141) + exact modulo (fastdiv.Mod, U:6811), computed inside the probe kernel
// ------------------------------------------------------------------------------------------------------
// x % d by Barrett reduction: q = floor(x * floor((2^W-1)/d) / 2^W) is floor(x/d) or one less (the estimate is short by
// x(1+s)/(d 2^W) < 1, s = (2^W-1) mod d), so one conditional subtraction makes the remainder exact for every x and every d >= 1.
__device__ __forceinline__ uint64_t mod64_dev(uint64_t x, const FastMod &fm) {
    const uint64_t r = x - __umul64hi(x, fm.b64) * fm.d;
    return r >= fm.d ? r - fm.d : r;
}
__device__ __forceinline__ uint32_t mod32_dev(uint32_t x, const FastMod &fm) {      // d < 2^32
    const uint32_t d = (uint32_t)fm.d;
    const uint32_t r = x - __umulhi(x, fm.b32) * d;
    return r >= d ? r - d : r;
}

// row index of hash number h of a k-mer code.  LocT = uint32_t for blocks with numSigs < 2^32-1, uint64_t beyond.
template <int H, typename LocT>
__device__ __forceinline__ LocT loc_of(uint64_t code, uint32_t h, const FastMod &fm) {
    if (H == 1) return (LocT)mod64_dev(code, fm);
    const uint32_t v = (uint32_t)(code >> 32) + (uint32_t)code * h;            // baseHashes (H:61-63), uint32 wrap-around (H:137-139)
    if (sizeof(LocT) == 8) return (LocT)((uint64_t)v >= fm.d ? (uint64_t)v - fm.d : (uint64_t)v);   // numSigs >= 2^32-1 >= v: at most one subtraction
    return (LocT)mod32_dev(v, fm);
}

// the row indices as a kernel of its own (u32, numSigs < 2^32-1): locs[slot*H + h].  It runs on the query-preparation stream beside
// the probe of the block before, so the probe kernel itself stays a pure load + count loop (deriving the indices inside the probe
// costs it 6 % of its bandwidth, profiles/README.md round 2); blocks with numSigs >= 2^32-1 derive them in the kernel.
template <int H>
__global__ void __launch_bounds__(256) locs_kernel(const uint64_t *__restrict__ codes, uint64_t n, FastMod fm, uint32_t *__restrict__ locs) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const uint64_t code = codes[i];
#pragma unroll
        for (uint32_t j = 0; j < (uint32_t)H; j++) locs[i * H + j] = loc_of<H, uint32_t>(code, j, fm);
    }
}
template <int H>
__global__ void __launch_bounds__(256) locs64_kernel(const uint64_t *__restrict__ codes, uint64_t n, FastMod fm, uint64_t *__restrict__ locs) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const uint64_t code = codes[i];
#pragma unroll
        for (uint32_t j = 0; j < (uint32_t)H; j++) {
            locs[i * H + j] = loc_of<H, uint64_t>(code, j, fm);
        }
    }
}

cudaError_t launch_locs(const uint64_t *codes, uint64_t n, int h, FastMod fm, uint32_t *locs, cudaStream_t st) {
    if (!n) return cudaSuccess;
    uint64_t blocks64 = (n + 255) / 256;
    uint32_t blocks = blocks64 > 148u * 64u ? 148u * 64u : (uint32_t)blocks64;
    switch (h) {
        case 1: locs_kernel<1><<<blocks, 256, 0, st>>>(codes, n, fm, locs); break;
        case 2: locs_kernel<2><<<blocks, 256, 0, st>>>(codes, n, fm, locs); break;
        case 3: locs_kernel<3><<<blocks, 256, 0, st>>>(codes, n, fm, locs); break;
        case 4: locs_kernel<4><<<blocks, 256, 0, st>>>(codes, n, fm, locs); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}
// sketch databases: the code regions are mostly empty (FracMinHash keeps ~2/scale of the positions), so walk the
// queries instead of the slots: one warp per query, only its n_eff codes
template <int H>
__global__ void __launch_bounds__(256) locs_query_kernel(const uint64_t *__restrict__ codes, const uint64_t *__restrict__ slot_off, const uint32_t *__restrict__ n_eff,
                                                         uint32_t nq, int paired, FastMod fm, uint32_t *__restrict__ locs) {
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t q = warp; q < nq; q += n_warps) {
        const uint32_t n = n_eff[q];
        const uint64_t base = slot_off[paired ? 2 * q : q];
        for (uint32_t i = lane; i < n; i += 32) {
            const uint64_t code = codes[base + i];
#pragma unroll
            for (uint32_t j = 0; j < (uint32_t)H; j++) locs[(base + i) * H + j] = loc_of<H, uint32_t>(code, j, fm);
        }
    }
}

cudaError_t launch_locs_by_query(const uint64_t *codes, const uint64_t *slot_off, const uint32_t *n_eff, uint32_t nq, int paired, int h, FastMod fm,
                                 uint32_t *locs, cudaStream_t st) {
    if (!nq) return cudaSuccess;
    uint32_t blocks = (nq + 7) / 8;
    if (blocks > 148u * 32u) blocks = 148u * 32u;
    switch (h) {
        case 1: locs_query_kernel<1><<<blocks, 256, 0, st>>>(codes, slot_off, n_eff, nq, paired, fm, locs); break;
        case 2: locs_query_kernel<2><<<blocks, 256, 0, st>>>(codes, slot_off, n_eff, nq, paired, fm, locs); break;
        case 3: locs_query_kernel<3><<<blocks, 256, 0, st>>>(codes, slot_off, n_eff, nq, paired, fm, locs); break;
        case 4: locs_query_kernel<4><<<blocks, 256, 0, st>>>(codes, slot_off, n_eff, nq, paired, fm, locs); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t launch_locs64(const uint64_t *codes, uint64_t n, int h, FastMod fm, uint64_t *locs, cudaStream_t st) {
    if (!n) return cudaSuccess;
    uint64_t blocks64 = (n + 255) / 256;
    uint32_t blocks = blocks64 > 148u * 64u ? 148u * 64u : (uint32_t)blocks64;
    switch (h) {
        case 1: locs64_kernel<1><<<blocks, 256, 0, st>>>(codes, n, fm, locs); break;
        case 2: locs64_kernel<2><<<blocks, 256, 0, st>>>(codes, n, fm, locs); break;
        case 3: locs64_kernel<3><<<blocks, 256, 0, st>>>(codes, n, fm, locs); break;
        case 4: locs64_kernel<4><<<blocks, 256, 0, st>>>(codes, n, fm, locs); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------
// probe
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t xor3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t maj3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// W 32-bit words of a bit-matrix row owned by one lane (W = 4: 16-byte slab, W = 2: 8-byte slab)
template <int W> struct Slab { uint32_t v[W]; };

template <int W>
__device__ __forceinline__ Slab<W> ld_slab(const uint8_t *p) {
    // streaming load of a row slab: read-only path, do not keep in L1
    Slab<W> s;
    if (W == 4) {
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(s.v[0]), "=r"(s.v[1]), "=r"(s.v[W > 2 ? 2 : 0]), "=r"(s.v[W > 3 ? 3 : 0]) : "l"(p));
    } else {
        asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(s.v[0]), "=r"(s.v[1]) : "l"(p));
    }
    return s;
}

constexpr int PROBE_THREADS = 256;
constexpr int PROBE_ROWS = 8;   // rows per carry-save tree

// Row indices of 8 consecutive k-mers (H each) of one task, from its k-mer codes; NONE past the end of the query.
// A FULL lane group (G == GF = 128 / slab bytes lanes, i.e. rows of at least 64 + 1 slab bytes) shares the arithmetic: lane j computes
// entry j (k-mer j / H, hash j % H) — and j + GF, ... when 8·H > GF — and the group exchanges the results by shuffle, so every
// lane pays one or two reciprocals per 8 rows.  Narrower groups (rows of a few bytes) compute every entry themselves.
template <typename LocT> struct LocNone { static constexpr LocT v = (LocT)~(LocT)0; };

template <int H, int GF, typename LocT, int SRC>
__device__ __forceinline__ void load_locs(LocT (&L)[PROBE_ROWS * H], const uint64_t *__restrict__ cp, const uint32_t *__restrict__ lp32, uint32_t i, uint32_t n,
                                          const FastMod &fm, bool full, uint32_t gl, uint32_t gmask, int gbase) {
    constexpr int NL = PROBE_ROWS * H;
    constexpr LocT NONE = LocNone<LocT>::v;
    if (SRC == 0) {                                                   // development alternative: indices precomputed by locs_kernel
#pragma unroll
        for (int u = 0; u < PROBE_ROWS; u++)
#pragma unroll
            for (int h = 0; h < H; h++) L[u * H + h] = (i + u < n) ? (LocT)__ldg(lp32 + (uint64_t)(i + u) * H + h) : NONE;
        return;
    }
    if (full) {
        constexpr int R = (NL + GF - 1) / GF;
        LocT mine[R];
#pragma unroll
        for (int r = 0; r < R; r++) {
            const uint32_t e = (uint32_t)r * GF + gl, u = e / H, h = e - u * H;
            mine[r] = NONE;
            if (e < (uint32_t)NL && i + u < n) mine[r] = loc_of<H, LocT>(__ldg(cp + i + u), h, fm);
        }
#pragma unroll
        for (int e = 0; e < NL; e++) L[e] = __shfl_sync(gmask, mine[e / GF], gbase + (e % GF));
    } else {
#pragma unroll
        for (int u = 0; u < PROBE_ROWS; u++) {
            const bool ok = i + u < n;
            const uint64_t code = ok ? __ldg(cp + i + u) : 0;
#pragma unroll
            for (int h = 0; h < H; h++) L[u * H + h] = ok ? loc_of<H, LocT>(code, (uint32_t)h, fm) : NONE;
        }
    }
}

// slabs of the 8 rows (h rows AND-ed: pand, U:6639-6645); zero for NONE and for lanes past the end of the row
template <int H, int W, typename LocT>
__device__ __forceinline__ void load_rows(Slab<W> (&r)[PROBE_ROWS], const LocT (&L)[PROBE_ROWS * H], const uint8_t *__restrict__ colbase, uint32_t pitch, bool lane_ok) {
#pragma unroll
    for (int u = 0; u < PROBE_ROWS; u++) {
        if (lane_ok && L[u * H] != LocNone<LocT>::v) {
            r[u] = ld_slab<W>(colbase + (uint64_t)L[u * H] * pitch);
#pragma unroll
            for (int h = 1; h < H; h++) {
                const Slab<W> t = ld_slab<W>(colbase + (uint64_t)L[u * H + h] * pitch);
#pragma unroll
                for (int w = 0; w < W; w++) r[u].v[w] &= t.v[w];
            }
        } else {
#pragma unroll
            for (int w = 0; w < W; w++) r[u].v[w] = 0;
        }
    }
}

// ---- h > 1, raw double buffering (VAR 3): the h rows of a k-mer must be AND-ed before they are counted, so keeping only the AND-ed rows of
// the next group in registers (VAR 2) makes every group wait for its own loads.  Here the RAW rows of the next group (8·h slabs) stay in
// flight while the current group's raw rows are AND-ed and counted.  The registers for that come from the row indices: a lane keeps only
// its share of them (entry e of the 8·h lives in lane e % GF, coalesced u32 loads) and the group hands them round by shuffle when the
// loads are issued — never a private copy of all 8·h.  Full lane groups only (GF lanes = rows of at least GF slabs).
template <int H, int GF>
__device__ __forceinline__ void load_shared_locs(uint32_t (&mine)[(PROBE_ROWS * H + GF - 1) / GF], const uint32_t *__restrict__ lp32, uint32_t i, uint32_t n, uint32_t gl) {
    constexpr int NL = PROBE_ROWS * H, R = (NL + GF - 1) / GF;
#pragma unroll
    for (int r = 0; r < R; r++) {
        const uint32_t e = (uint32_t)r * GF + gl;
        mine[r] = (e < (uint32_t)NL && i + e / H < n) ? __ldg(lp32 + (uint64_t)i * H + e) : 0xFFFFFFFFu;
    }
}
template <int H, int W, int GF>
__device__ __forceinline__ void issue_raw_rows(Slab<W> (&t)[PROBE_ROWS * H], const uint32_t (&mine)[(PROBE_ROWS * H + GF - 1) / GF], const uint8_t *__restrict__ colbase,
                                               uint32_t pitch, bool lane_ok, uint32_t gmask, int gbase) {
#pragma unroll
    for (int e = 0; e < PROBE_ROWS * H; e++) {
        const uint32_t loc = __shfl_sync(gmask, mine[e / GF], gbase + (e % GF));
        if (lane_ok && loc != 0xFFFFFFFFu) t[e] = ld_slab<W>(colbase + (uint64_t)loc * pitch);
        else {
#pragma unroll
            for (int w = 0; w < W; w++) t[e].v[w] = 0;
        }
    }
}
template <int H, int W>
__device__ __forceinline__ void and_raw_rows(Slab<W> (&r)[PROBE_ROWS], const Slab<W> (&t)[PROBE_ROWS * H]) {
#pragma unroll
    for (int u = 0; u < PROBE_ROWS; u++)
#pragma unroll
        for (int w = 0; w < W; w++) {
            uint32_t v = t[u * H].v[w];
#pragma unroll
            for (int h = 1; h < H; h++) v &= t[u * H + h].v[w];             // pand, U:6639-6645
            r[u].v[w] = v;
        }
}

// Harley–Seal: 8 one-bit inputs + planes 0..2 → planes 0..2 and one weight-8 carry rippled into planes 3..7
template <int W>
__device__ __forceinline__ void csa8(uint32_t (&c)[8][W], const Slab<W> (&r)[PROBE_ROWS]) {
#pragma unroll
    for (int w = 0; w < W; w++) {
        const uint32_t x0 = r[0].v[w], x1 = r[1].v[w], x2 = r[2].v[w], x3 = r[3].v[w];
        const uint32_t x4 = r[4].v[w], x5 = r[5].v[w], x6 = r[6].v[w], x7 = r[7].v[w];
        uint32_t ones = c[0][w], twos = c[1][w], fours = c[2][w];
        uint32_t t1a = maj3(ones, x0, x1); ones = xor3(ones, x0, x1);
        uint32_t t1b = maj3(ones, x2, x3); ones = xor3(ones, x2, x3);
        uint32_t t2a = maj3(twos, t1a, t1b); twos = xor3(twos, t1a, t1b);
        t1a = maj3(ones, x4, x5); ones = xor3(ones, x4, x5);
        t1b = maj3(ones, x6, x7); ones = xor3(ones, x6, x7);
        uint32_t t2b = maj3(twos, t1a, t1b); twos = xor3(twos, t1a, t1b);
        uint32_t carry = maj3(fours, t2a, t2b); fours = xor3(fours, t2a, t2b);
        c[0][w] = ones; c[1][w] = twos; c[2][w] = fours;
#pragma unroll
        for (int p = 3; p < 8; p++) {
            uint32_t t = c[p][w] & carry;
            c[p][w] ^= carry;
            carry = t;
        }
    }
}

// Counters: 8 bit-sliced planes in registers count up to 255 rows.  Queries with more k-mers (long reads, -g genomes)
// keep the full P = 8+PH planes per thread in shared memory and fold the register planes into them every 248 rows
// (one ripple add), so every query length runs the same register-lean inner loop.
template <int PH, int W>
__device__ __forceinline__ void fold_planes(uint32_t (&c)[8][W], uint32_t *T) {
    // T[(p*W+w)*blockDim + tid] += c (bit-sliced add), c = 0
#pragma unroll
    for (int w = 0; w < W; w++) {
        uint32_t carry = 0;
#pragma unroll
        for (int p = 0; p < 8; p++) {
            uint32_t *t = T + (p * W + w) * PROBE_THREADS;
            const uint32_t tv = *t, x = c[p][w];
            *t = xor3(tv, x, carry);
            carry = maj3(tv, x, carry);
            c[p][w] = 0;
        }
#pragma unroll
        for (int p = 8; p < 8 + PH; p++) {
            uint32_t *t = T + (p * W + w) * PROBE_THREADS;
            const uint32_t tv = *t;
            *t = tv ^ carry;
            carry &= tv;
        }
    }
}

// VAR 0: load locs, load rows, add (simple).  VAR 1: row indices of the next 8 k-mers are prefetched while the
// current rows are in flight.  VAR 2: additionally the rows are double-buffered in registers, so 8..16 rows per
// lane are always in flight while the carry-save tree of the previous 8 runs.
// W = words per lane: 4 (16-byte slabs, 8 lanes per 128-byte task) or 2 (8-byte slabs, 16 lanes per task — half the
// registers per row in flight, which is what the h>1 AND of several rows per k-mer needs).
template <int H, int PH, int VAR, int MINB, int W, typename LocT, int SRC>
__global__ void __launch_bounds__(PROBE_THREADS, MINB) probe_kernel(ProbeArgs a) {
    constexpr int P = 8 + PH;
    constexpr uint32_t SLAB = 4 * W;                                    // bytes per lane
    constexpr int GF = 128 / SLAB;                                      // lanes of a full task group
    extern __shared__ uint32_t smem_planes[];                          // PH > 0: P*W*PROBE_THREADS words
    uint32_t *T = smem_planes + threadIdx.x;
    // task geometry for this slab width: a task covers up to 128 bytes of a row
    const uint32_t row_units = (a.row_bytes + SLAB - 1) / SLAB;
    uint32_t G = 1;
    while (G < row_units && G < 128 / SLAB) G <<= 1;
    if (a.lanes_per_task_override) G = a.lanes_per_task_override;
    const uint32_t chunks = (row_units + G - 1) / G;
    const uint32_t gl = threadIdx.x & (G - 1);                          // lane inside the task group
    const uint64_t groups_per_grid = ((uint64_t)gridDim.x * blockDim.x) / G;
    const uint64_t total = (uint64_t)a.n_queries * chunks;
    const uint64_t first = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const uint64_t iters = (total + groups_per_grid - 1) / groups_per_grid;
    const int lane = threadIdx.x & 31;

    // Long queries (more than 255 k-mers: PH > 0) differ a lot in length and are few per launch, so a fixed share of the tasks per lane group
    // leaves much of the grid idle at the end; their lane groups draw the next task from a counter instead.  Short reads keep the fixed
    // assignment (ten million atomics on one address per batch would cost more than they save).
    constexpr bool DYNAMIC = PH > 0;
    const int gbase0 = lane & ~(int)(G - 1);
    const uint32_t gmask0 = (G >= 32 ? 0xFFFFFFFFu : ((1u << G) - 1u)) << gbase0;
    for (uint64_t it = 0;; it++) {
        uint64_t g;
        if (DYNAMIC) {
            // one draw per WARP: its 32/G lane groups take consecutive tasks — with the chunk index running fastest these are
            // slices of the same query, so the groups of a warp finish together (queries of very different lengths in one warp
            // left the shorter one's lanes idle: 0.79 of the short-read bandwidth on HiFi reads)
            unsigned long long t = 0;
            if (lane == 0) t = atomicAdd(a.task_counter, (unsigned long long)(32 / G));
            t = __shfl_sync(0xffffffffu, t, 0);
            if (t >= total) break;                                      // warp-uniform: the hit append below is warp-wide
            g = t + (uint64_t)(lane / G);
        } else {
            if (it >= iters) break;
            g = first + it * groups_per_grid;
        }
        uint32_t q = 0, chunk = 0, n = 0;
        if (g < total) {
            q = (uint32_t)(g / chunks);
            chunk = (uint32_t)(g - (uint64_t)q * chunks);               // chunk fastest: neighbouring groups share rows
            if (DYNAMIC && a.order) q = a.order[q];                     // longest queries first: the last draws are the short ones
            n = a.n_eff[q];
        }
        const uint32_t colu = chunk * G + gl;                           // this lane's slab of the row
        const bool lane_ok = colu < row_units;                          // the last chunk of a row may not fill its group
        const bool active = n > 0 && lane_ok;
        uint32_t c[8][W];
#pragma unroll
        for (int p = 0; p < 8; p++)
#pragma unroll
            for (int w = 0; w < W; w++) c[p][w] = 0;

        if (n > 0) {                                                    // the WHOLE lane group runs the loop (it shares the row-index arithmetic)
            if (PH > 0) {
#pragma unroll
                for (int i = 0; i < P * W; i++) T[i * PROBE_THREADS] = 0;
            }
            const uint64_t slot0 = a.slot_off[a.paired ? 2 * q : q];
            const uint64_t *lp = a.codes + slot0;
            const uint32_t *lp32 = SRC == 0 ? a.locs + slot0 * (uint64_t)H : nullptr;
            const uint8_t *colbase = a.rows + (uint64_t)(lane_ok ? colu : 0) * SLAB;
            const bool full = G == (uint32_t)GF;
            const int gbase = gbase0;
            const uint32_t gmask = gmask0;
            LocT L[PROBE_ROWS * H];
            uint32_t acc = 0;                                           // rows added to the register planes since the last fold
            auto add8 = [&](const Slab<W> (&r)[PROBE_ROWS]) {
                if (PH > 0) {
                    if (acc + PROBE_ROWS > 255) { fold_planes<PH, W>(c, T); acc = 0; }
                    acc += PROBE_ROWS;
                }
                csa8<W>(c, r);
            };
#define KMCPG_LOCS(i_) load_locs<H, GF, LocT, SRC>(L, lp, lp32, (i_), n, a.fm, full, gl, gmask, gbase)
#define KMCPG_ROWS(r_) load_rows<H, W, LocT>(r_, L, colbase, a.pitch, lane_ok)
            if (VAR == 0) {
                for (uint32_t i = 0; i < n; i += PROBE_ROWS) {
                    Slab<W> r[PROBE_ROWS];
                    KMCPG_LOCS(i);
                    KMCPG_ROWS(r);
                    add8(r);
                }
            } else if (VAR == 1) {
                KMCPG_LOCS(0);
                for (uint32_t i = 0; i < n; i += PROBE_ROWS) {
                    Slab<W> r[PROBE_ROWS];
                    KMCPG_ROWS(r);
                    KMCPG_LOCS(i + PROBE_ROWS);
                    add8(r);
                }
            } else if (VAR == 4) {
                // VAR 1 with shared row indices: the 8·h indices never sit in every lane's registers, which leaves room for a third CTA per SM
                if constexpr (SRC == 0 && sizeof(LocT) == 4) {
                    constexpr int RS = (PROBE_ROWS * H + GF - 1) / GF;
                    uint32_t mine[RS];
                    Slab<W> t[PROBE_ROWS * H], r[PROBE_ROWS];
                    load_shared_locs<H, GF>(mine, lp32, 0, n, gl);
                    for (uint32_t i = 0; i < n; i += PROBE_ROWS) {
                        issue_raw_rows<H, W, GF>(t, mine, colbase, a.pitch, lane_ok, gmask, gbase);
                        load_shared_locs<H, GF>(mine, lp32, i + PROBE_ROWS, n, gl);
                        and_raw_rows<H, W>(r, t);
                        add8(r);
                    }
                }
            } else if (VAR == 3) {
                if constexpr (SRC == 0 && sizeof(LocT) == 4) {           // (instantiated for u32 indices from the locs kernel only)
                    constexpr int RS = (PROBE_ROWS * H + GF - 1) / GF;
                    uint32_t mine[RS];
                    Slab<W> ta[PROBE_ROWS * H], tb[PROBE_ROWS * H], r[PROBE_ROWS];
                    load_shared_locs<H, GF>(mine, lp32, 0, n, gl);
                    issue_raw_rows<H, W, GF>(ta, mine, colbase, a.pitch, lane_ok, gmask, gbase);
                    load_shared_locs<H, GF>(mine, lp32, PROBE_ROWS, n, gl);
                    for (uint32_t i = 0; i < n; i += 2 * PROBE_ROWS) {
                        issue_raw_rows<H, W, GF>(tb, mine, colbase, a.pitch, lane_ok, gmask, gbase);
                        load_shared_locs<H, GF>(mine, lp32, i + 2 * PROBE_ROWS, n, gl);
                        and_raw_rows<H, W>(r, ta);
                        add8(r);
                        issue_raw_rows<H, W, GF>(ta, mine, colbase, a.pitch, lane_ok, gmask, gbase);
                        load_shared_locs<H, GF>(mine, lp32, i + 3 * PROBE_ROWS, n, gl);
                        if (i + PROBE_ROWS < n) { and_raw_rows<H, W>(r, tb); add8(r); }
                    }
                }
            } else {
                Slab<W> r0[PROBE_ROWS], r1[PROBE_ROWS];
                KMCPG_LOCS(0);
                KMCPG_ROWS(r0);
                KMCPG_LOCS(PROBE_ROWS);
                for (uint32_t i = 0; i < n; i += 2 * PROBE_ROWS) {
                    KMCPG_ROWS(r1);
                    KMCPG_LOCS(i + 2 * PROBE_ROWS);
                    add8(r0);
                    KMCPG_ROWS(r0);
                    KMCPG_LOCS(i + 3 * PROBE_ROWS);
                    if (i + PROBE_ROWS < n) add8(r1);
                }
            }
#undef KMCPG_LOCS
#undef KMCPG_ROWS
            if (PH > 0) fold_planes<PH, W>(c, T);
        }
        // plane p, word w of this thread's final counters
        auto plane = [&](int p, int w) -> uint32_t { return PH > 0 ? T[(p * W + w) * PROBE_THREADS] : c[p < 8 ? p : 0][w]; };

        // ---- thresholds on the bit-sliced counters: ge = (count >= T) per target bit ----
        uint32_t ge[W];
#pragma unroll
        for (int w = 0; w < W; w++) ge[w] = 0;
        int nhit = 0;
        if (active) {
            const uint32_t Tq = a.thresh[q];
            uint32_t high = (P < 32) ? (Tq >> P) : 0;                    // threshold does not fit in P bits → nothing passes
            if (!high) {
#pragma unroll
                for (int w = 0; w < W; w++) {
                    uint32_t gt = 0, eq = 0xFFFFFFFFu;
#pragma unroll
                    for (int p = P - 1; p >= 0; p--) {
                        const uint32_t v = plane(p, w);
                        if ((Tq >> p) & 1) { eq &= v; }
                        else { gt |= eq & v; eq &= ~v; }
                    }
                    ge[w] = gt | eq;
                    nhit += __popc(ge[w]);
                }
            }
            if (a.dense_counts) {                                        // test hook: dump every count of this slab
#pragma unroll
                for (int w = 0; w < W; w++)
                    for (int bit = 0; bit < 32; bit++) {
                        uint32_t t = (colu * SLAB + w * 4 + (bit >> 3)) * 8 + (7 - (bit & 7));
                        if (t < a.n_names) {
                            uint32_t cnt = 0;
#pragma unroll
                            for (int p = 0; p < P; p++) cnt |= ((plane(p, w) >> bit) & 1u) << p;
                            a.dense_counts[a.target_base + t] = cnt;
                        }
                    }
            }
        }
        // ---- append hits: one atomic per warp ----
        __syncwarp();
        int incl = nhit;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        int tot = __shfl_sync(0xffffffffu, incl, 31);
        if (tot > 0) {
            unsigned long long base = 0;
            if (lane == 31) base = atomicAdd(a.hit_count, (unsigned long long)tot);
            base = __shfl_sync(0xffffffffu, base, 31);
            unsigned long long slot = base + (unsigned long long)(incl - nhit);
            if (nhit > 0) {
#pragma unroll
                for (int w = 0; w < W; w++) {
                    uint32_t m = ge[w];
                    while (m) {
                        int bit = __ffs(m) - 1;
                        m &= m - 1;
                        // byte (colu*SLAB + w*4 + bit/8), bit 7-j ↔ target 8*byte + j  (I:1157, U:7415)
                        uint32_t t = (colu * SLAB + w * 4 + (bit >> 3)) * 8 + (7 - (bit & 7));
                        uint32_t cnt = 0;
#pragma unroll
                        for (int p = 0; p < P; p++) cnt |= ((plane(p, w) >> bit) & 1u) << p;
                        if (slot < a.hit_cap) {
                            a.hit_keys[slot] = ((uint64_t)q << a.target_bits) | (uint64_t)(a.target_base + t);
                            a.hit_vals[slot] = cnt;
                        }
                        slot++;
                    }
                }
            }
        }
    }
}

// Few, long queries (genome-sized -g queries against sketch DBs, a handful of HiFi reads): a task = (query, chunk) is given
// to a WHOLE CTA.  Its 256/G lane groups each take every (256/G)-th group of 8 k-mers, count them exactly like the kernel
// above (8 register planes folded into P-plane totals in shared memory), then the per-group totals are added pairwise in
// shared memory (bit-sliced ripple adders, log2(256/G) levels) and group 0 applies the threshold and emits the hits.
template <int H, int PH, int W, typename LocT>
__global__ void __launch_bounds__(PROBE_THREADS, 1) probe_long_kernel(ProbeArgs a) {
    static_assert(PH > 0, "long queries always carry full-width totals in shared memory");
    constexpr int P = 8 + PH;
    constexpr uint32_t SLAB = 4 * W;
    constexpr int GF = 128 / SLAB;
    extern __shared__ uint32_t smem_planes[];                          // P*W*PROBE_THREADS words
    uint32_t *T = smem_planes + threadIdx.x;
    const uint32_t row_units = (a.row_bytes + SLAB - 1) / SLAB;
    uint32_t G = 1;
    while (G < row_units && G < 128 / SLAB) G <<= 1;
    const uint32_t chunks = (row_units + G - 1) / G;
    const uint32_t gl = threadIdx.x & (G - 1);
    const uint32_t seg = threadIdx.x / G, nseg = PROBE_THREADS / G;     // k-mer segment of this lane group
    const uint64_t total = (uint64_t)a.n_queries * chunks;
    const int lane = threadIdx.x & 31;

    for (uint64_t g = blockIdx.x; g < total; g += gridDim.x) {
        const uint32_t q = (uint32_t)(g / chunks);
        const uint32_t chunk = (uint32_t)(g - (uint64_t)q * chunks);
        const uint32_t n = a.n_eff[q];
        const uint32_t colu = chunk * G + gl;
        const bool lane_ok = colu < row_units;
        const bool active = n > 0 && lane_ok;
        uint32_t c[8][W];
#pragma unroll
        for (int p = 0; p < 8; p++)
#pragma unroll
            for (int w = 0; w < W; w++) c[p][w] = 0;
#pragma unroll
        for (int i = 0; i < P * W; i++) T[i * PROBE_THREADS] = 0;
        if (n > 0) {                                                     // whole lane groups: they share the row-index arithmetic
            const uint64_t *lp = a.codes + a.slot_off[a.paired ? 2 * q : q];
            const uint8_t *colbase = a.rows + (uint64_t)(lane_ok ? colu : 0) * SLAB;
            const bool full = G == (uint32_t)GF;
            const int gbase = lane & ~(int)(G - 1);
            const uint32_t gmask = (G >= 32 ? 0xFFFFFFFFu : ((1u << G) - 1u)) << gbase;
            LocT L[PROBE_ROWS * H];
            uint32_t acc = 0;
            const uint32_t stride = nseg * PROBE_ROWS;
            uint32_t i = seg * PROBE_ROWS;
            if (i < n) load_locs<H, GF, LocT, 1>(L, lp, nullptr, i, n, a.fm, full, gl, gmask, gbase);
            for (; i < n; i += stride) {
                Slab<W> r[PROBE_ROWS];
                load_rows<H, W, LocT>(r, L, colbase, a.pitch, lane_ok);
                if (i + stride < n) load_locs<H, GF, LocT, 1>(L, lp, nullptr, i + stride, n, a.fm, full, gl, gmask, gbase);
                if (acc + PROBE_ROWS > 255) { fold_planes<PH, W>(c, T); acc = 0; }
                acc += PROBE_ROWS;
                csa8<W>(c, r);
            }
            fold_planes<PH, W>(c, T);
        }
        __syncthreads();
        // pairwise addition of the segments' totals: thread t += thread t + s*G
        for (uint32_t s2 = nseg >> 1; s2 >= 1; s2 >>= 1) {
            if (seg < s2) {
                const uint32_t *O = T + s2 * G;
#pragma unroll
                for (int w = 0; w < W; w++) {
                    uint32_t carry = 0;
#pragma unroll
                    for (int p = 0; p < P; p++) {
                        const uint32_t x = T[(p * W + w) * PROBE_THREADS], y = O[(p * W + w) * PROBE_THREADS];
                        T[(p * W + w) * PROBE_THREADS] = xor3(x, y, carry);
                        carry = maj3(x, y, carry);
                    }
                }
            }
            __syncthreads();
        }
        // ---- thresholds + hits: lane group 0 holds the totals ----
        uint32_t ge[W];
#pragma unroll
        for (int w = 0; w < W; w++) ge[w] = 0;
        int nhit = 0;
        const bool owner = active && seg == 0;
        if (owner) {
            const uint32_t Tq = a.thresh[q];
            uint32_t high = (P < 32) ? (Tq >> P) : 0;
            if (!high) {
#pragma unroll
                for (int w = 0; w < W; w++) {
                    uint32_t gt = 0, eq = 0xFFFFFFFFu;
#pragma unroll
                    for (int p = P - 1; p >= 0; p--) {
                        const uint32_t v = T[(p * W + w) * PROBE_THREADS];
                        if ((Tq >> p) & 1) { eq &= v; }
                        else { gt |= eq & v; eq &= ~v; }
                    }
                    ge[w] = gt | eq;
                    nhit += __popc(ge[w]);
                }
            }
            if (a.dense_counts) {
#pragma unroll
                for (int w = 0; w < W; w++)
                    for (int bit = 0; bit < 32; bit++) {
                        uint32_t t = (colu * SLAB + w * 4 + (bit >> 3)) * 8 + (7 - (bit & 7));
                        if (t < a.n_names) {
                            uint32_t cnt = 0;
#pragma unroll
                            for (int p = 0; p < P; p++) cnt |= ((T[(p * W + w) * PROBE_THREADS] >> bit) & 1u) << p;
                            a.dense_counts[a.target_base + t] = cnt;
                        }
                    }
            }
        }
        if (threadIdx.x < 32) {                                          // lane group 0 lives in warp 0 (G <= 32)
            int incl = nhit;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            int tot = __shfl_sync(0xffffffffu, incl, 31);
            if (tot > 0) {
                unsigned long long base = 0;
                if (lane == 31) base = atomicAdd(a.hit_count, (unsigned long long)tot);
                base = __shfl_sync(0xffffffffu, base, 31);
                unsigned long long slot = base + (unsigned long long)(incl - nhit);
                if (nhit > 0) {
#pragma unroll
                    for (int w = 0; w < W; w++) {
                        uint32_t m = ge[w];
                        while (m) {
                            int bit = __ffs(m) - 1;
                            m &= m - 1;
                            uint32_t t = (colu * SLAB + w * 4 + (bit >> 3)) * 8 + (7 - (bit & 7));
                            uint32_t cnt = 0;
#pragma unroll
                            for (int p = 0; p < P; p++) cnt |= ((T[(p * W + w) * PROBE_THREADS] >> bit) & 1u) << p;
                            if (slot < a.hit_cap) {
                                a.hit_keys[slot] = ((uint64_t)q << a.target_bits) | (uint64_t)(a.target_base + t);
                                a.hit_vals[slot] = cnt;
                            }
                            slot++;
                        }
                    }
                }
            }
        }
        __syncthreads();                                                 // the totals are reused by the next task
    }
}

template <int H, int PH, int W, typename LocT>
static cudaError_t launch_probe_long_k(const ProbeArgs &a, uint32_t blocks, cudaStream_t st) {
    const size_t smem = (size_t)(8 + PH) * W * PROBE_THREADS * sizeof(uint32_t);
    if (smem > 48 * 1024) {
        static bool done = false;
        if (!done) {
            cudaError_t e = cudaFuncSetAttribute(probe_long_kernel<H, PH, W, LocT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            done = true;
        }
    }
    probe_long_kernel<H, PH, W, LocT><<<blocks, PROBE_THREADS, smem, st>>>(a);
    return cudaGetLastError();
}

// Development knobs (tools/*.sh, builds with -DKMCPG_DEV only): KMCPG_PROBE_VAR / _MINB (h=1), KMCPG_PROBE_VARH / _MINBH / _WH (h>1),
// KMCPG_PROBE_CAP, KMCPG_PROBE_G.  The product library compiles the shipped variants only and reads no environment.
struct ProbeTune { int var = 2, minb = 2, cap = 16, var_h = 1, minb_h = 2, w_h = 2, g = 0; };
static ProbeTune probe_tune() {
#ifdef KMCPG_DEV
    static ProbeTune t = [] {
        ProbeTune x;
        if (const char *e = getenv("KMCPG_PROBE_VAR")) x.var = atoi(e);
        if (const char *e = getenv("KMCPG_PROBE_MINB")) x.minb = atoi(e);
        if (const char *e = getenv("KMCPG_PROBE_VARH")) x.var_h = atoi(e);
        if (const char *e = getenv("KMCPG_PROBE_MINBH")) x.minb_h = atoi(e);
        if (const char *e = getenv("KMCPG_PROBE_WH")) x.w_h = atoi(e) == 2 ? 2 : 4;
        if (const char *e = getenv("KMCPG_PROBE_CAP")) x.cap = atoi(e);
        if (const char *e = getenv("KMCPG_PROBE_G")) { int v = atoi(e); if (v == 1 || v == 2 || v == 4 || v == 8 || v == 16 || v == 32) x.g = v; }
        return x;
    }();
    return t;
#else
    return ProbeTune();
#endif
}

template <int H, int PH, int VAR, int MINB, int W, typename LocT, int SRC = 1>
static cudaError_t launch_probe_k(const ProbeArgs &a, uint32_t blocks, cudaStream_t st) {
    const size_t smem = PH > 0 ? (size_t)(8 + PH) * W * PROBE_THREADS * sizeof(uint32_t) : 0;
    if (smem > 48 * 1024) {
        static bool done = false;       // per instantiation
        if (!done) {
            cudaError_t e = cudaFuncSetAttribute(probe_kernel<H, PH, VAR, MINB, W, LocT, SRC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            done = true;
        }
    }
    probe_kernel<H, PH, VAR, MINB, W, LocT, SRC><<<blocks, PROBE_THREADS, smem, st>>>(a);
    return cudaGetLastError();
}

template <int H, int PH>
static cudaError_t launch_probe_hp(const ProbeArgs &a, uint32_t blocks, cudaStream_t st) {
    // blocks with numSigs >= 2^32-1 (the format's NumSigs is a uint64, X:173): 64-bit row indices, one CTA per SM (more registers)
    if (a.fm.d >= 0xFFFFFFFFull) {
        if constexpr (PH > 0) {
            if (a.long_mode) {
                if constexpr (H == 1) return launch_probe_long_k<H, PH, 4, uint64_t>(a, blocks, st);
                else return launch_probe_long_k<H, PH, 2, uint64_t>(a, blocks, st);
            }
        }
        if constexpr (H == 1) return launch_probe_k<H, PH, 2, 1, 4, uint64_t>(a, blocks, st);
        else return launch_probe_k<H, PH, 1, 1, 2, uint64_t>(a, blocks, st);
    }
    if constexpr (PH > 0) {
        if (a.long_mode) {                                   // few long queries: one CTA per (query, chunk)
            if constexpr (H == 1) return launch_probe_long_k<H, PH, 4, uint32_t>(a, blocks, st);
            else return launch_probe_long_k<H, PH, 2, uint32_t>(a, blocks, st);
        }
    }
#ifdef KMCPG_DEV
    // the development variants (documented in profiles/config_sweeps_r01.md) exist only for the 8-plane kernels
    const ProbeTune t = probe_tune();
    if constexpr (PH == 0) {
        if (!a.locs) {                                      // KMCPG_PROBE_LOCS=kernel: row indices derived in the probe kernel
            if constexpr (H == 1) return launch_probe_k<H, PH, 2, 2, 4, uint32_t, 1>(a, blocks, st);
            else return launch_probe_k<H, PH, 1, 2, 2, uint32_t, 1>(a, blocks, st);
        }
        if constexpr (H == 1) {
            if (t.var == 0) return launch_probe_k<H, PH, 0, 2, 4, uint32_t, 0>(a, blocks, st);
            if (t.var == 1) return t.minb >= 3 ? launch_probe_k<H, PH, 1, 3, 4, uint32_t, 0>(a, blocks, st) : launch_probe_k<H, PH, 1, 2, 4, uint32_t, 0>(a, blocks, st);
        } else {
            if (t.w_h == 4) return t.minb_h >= 3 ? launch_probe_k<H, PH, 1, 3, 4, uint32_t, 0>(a, blocks, st) : launch_probe_k<H, PH, 1, 2, 4, uint32_t, 0>(a, blocks, st);
            if (t.var_h == 3 && H < 4 && a.row_bytes > 8 * 8) return launch_probe_k<H, PH, 3, 2, 2, uint32_t, 0>(a, blocks, st);   // raw double buffering (full groups)
            if (t.var_h == 4 && a.row_bytes > 8 * 8) return t.minb_h >= 3 ? launch_probe_k<H, PH, 4, 3, 2, uint32_t, 0>(a, blocks, st) : launch_probe_k<H, PH, 4, 2, 2, uint32_t, 0>(a, blocks, st);
            if (t.var_h == 2) return launch_probe_k<H, PH, 2, 2, 2, uint32_t, 0>(a, blocks, st);
            if (t.minb_h >= 3) return launch_probe_k<H, PH, 1, 3, 2, uint32_t, 0>(a, blocks, st);
        }
    }
#endif
    if (!a.locs) return cudaErrorInvalidValue;              // 32-bit row indices come from launch_locs (the in-kernel form is a KMCPG_DEV variant)
    if constexpr (H == 1) {
        if constexpr (PH >= 24) return launch_probe_k<H, PH, 2, 1, 4, uint32_t, 0>(a, blocks, st);      // 128 KB of counter planes per CTA
        else return launch_probe_k<H, PH, 2, 2, 4, uint32_t, 0>(a, blocks, st);               // shipped: double-buffered 16-byte slabs, 2 CTAs/SM
    } else {
        return launch_probe_k<H, PH, 1, 2, 2, uint32_t, 0>(a, blocks, st);                    // shipped for h>1: 8-byte slabs, index prefetch, 2 CTAs/SM
    }
}

template <int H>
static cudaError_t launch_probe_h(const ProbeArgs &a, uint32_t blocks, cudaStream_t st) {
    switch (a.planes) {
        case 8: return launch_probe_hp<H, 0>(a, blocks, st);
        case 16: return launch_probe_hp<H, 8>(a, blocks, st);
        case 24: return launch_probe_hp<H, 16>(a, blocks, st);
        case 32: return launch_probe_hp<H, 24>(a, blocks, st);
        default: return cudaErrorInvalidValue;
    }
}

// ------------------------------------------------------------------------------------------------------
// probe, TMA form (wide rows): the rows travel global → shared memory by cp.async.bulk (the TMA unit's 1-D bulk copy), completion
// is signalled on mbarriers, the threads read them back with LDS.  A task = (query, slice of up to 1 KB of the row); thread t owns
// 4 bytes of the slice and 8 bit-sliced counter words.  Sixteen k-mers (x h rows) are in flight per CTA in a shared-memory ring —
// the memory-level parallelism that the register kernel above buys with 128 registers per thread.  One elected thread issues the
// copies of k-mer group g+2 while everybody adds up group g.  Short queries only (<= 255 k-mers: 8 counter planes), 32-bit row indices.
// ------------------------------------------------------------------------------------------------------
constexpr int BULK_THREADS = 256;
constexpr int BULK_STAGES = 16;          // k-mers in flight per CTA (two groups of PROBE_ROWS)
constexpr uint32_t BULK_SLICE = 4096;    // bytes of a row per task, at most (whole rows up to 4 KB: one bulk copy per row — the TMA unit
                                         // takes 60-100 cycles per copy, so slices of 640-900 bytes cap it at 2-4 TB/s, profiles/README.md)

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)), "l"(src), "r"(bytes),
                 "r"(smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_addr(bar)),
        "r"(parity)
        : "memory");
}

// WPT = 32-bit words of the slice per thread (thread t owns words t, t + 256, ...: conflict-free LDS)
template <int H, int WPT>
__global__ void __launch_bounds__(BULK_THREADS) probe_bulk_kernel(ProbeArgs a, uint32_t slice, uint32_t n_slices) {
    extern __shared__ __align__(128) uint8_t bulk_smem[];
    uint8_t *ring = bulk_smem;                                                     // [BULK_STAGES][H][slice]
    uint64_t *full = (uint64_t *)(bulk_smem + (size_t)BULK_STAGES * H * slice);    // [BULK_STAGES]
    uint32_t *sloc = (uint32_t *)(full + BULK_STAGES);                             // [2][PROBE_ROWS * H]: row indices of the two groups ahead
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < BULK_STAGES; s++) mbar_init(full + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    constexpr int NL = PROBE_ROWS * H;
    const uint64_t total = (uint64_t)a.n_queries * n_slices;
    uint32_t kglobal = 0;                                                          // k-mers this CTA has pushed through the ring so far (stage = kglobal % STAGES)
    for (uint64_t task = blockIdx.x; task < total; task += gridDim.x) {
        const uint32_t q = (uint32_t)(task / n_slices), sl = (uint32_t)(task - (uint64_t)q * n_slices);
        const uint32_t n = a.n_eff[q];
        const uint32_t byte0 = sl * slice;
        const uint32_t bytes = a.pitch - byte0 < slice ? a.pitch - byte0 : slice;  // the pitch is a multiple of 128 and zero padded behind row_bytes
        uint32_t c[8][WPT];
#pragma unroll
        for (int p = 0; p < 8; p++)
#pragma unroll
            for (int w = 0; w < WPT; w++) c[p][w] = 0;
        if (n > 0) {
            const uint32_t *lp = a.locs + a.slot_off[a.paired ? 2 * q : q] * (uint64_t)H;
            const uint8_t *col = a.rows + byte0;
            const uint32_t groups = (n + PROBE_ROWS - 1) / PROBE_ROWS;
            // issue the copies of one group of 8 k-mers (indices in sloc[buf]); k0 = ring position of its first k-mer
            auto issue = [&](uint32_t g, int buf, uint32_t k0) {
                const uint32_t cnt = n - g * PROBE_ROWS < (uint32_t)PROBE_ROWS ? n - g * PROBE_ROWS : (uint32_t)PROBE_ROWS;
                for (uint32_t u = 0; u < cnt; u++) {
                    const uint32_t s = (k0 + u) % BULK_STAGES;
                    mbar_expect_tx(full + s, (uint32_t)H * bytes);
#pragma unroll
                    for (int h = 0; h < H; h++)
                        bulk_g2s(ring + ((size_t)s * H + h) * slice, col + (uint64_t)sloc[buf * NL + u * H + h] * a.pitch, bytes, full + s);
                }
            };
            auto fetch_locs = [&](uint32_t g, int buf) {                           // by the first 8*H threads
                if (tid < NL) {
                    const uint32_t i = g * PROBE_ROWS + (uint32_t)tid / H;
                    sloc[buf * NL + tid] = i < n ? __ldg(lp + (uint64_t)g * NL + tid) : 0;
                }
            };
            fetch_locs(0, 0);
            if (groups > 1) fetch_locs(1, 1);
            __syncthreads();
            if (tid == 0) {
                issue(0, 0, kglobal);
                if (groups > 1) issue(1, 1, kglobal + PROBE_ROWS);
            }
            for (uint32_t g = 0; g < groups; g++) {
                uint32_t r[PROBE_ROWS][WPT];
#pragma unroll
                for (int u = 0; u < PROBE_ROWS; u++) {
                    const uint32_t i = g * PROBE_ROWS + u;
#pragma unroll
                    for (int w = 0; w < WPT; w++) r[u][w] = 0;
                    if (i < n) {
                        const uint32_t kk = kglobal + i, s = kk % BULK_STAGES;
                        mbar_wait(full + s, (kk / BULK_STAGES) & 1u);
#pragma unroll
                        for (int w = 0; w < WPT; w++) {
                            const uint32_t o = ((uint32_t)w * BULK_THREADS + (uint32_t)tid) * 4;
                            if (o < bytes) {
                                uint32_t v = *(const uint32_t *)(ring + (size_t)s * H * slice + o);
#pragma unroll
                                for (int h = 1; h < H; h++) v &= *(const uint32_t *)(ring + ((size_t)s * H + h) * slice + o);      // pand, U:6639-6645
                                r[u][w] = v;
                            }
                        }
                    }
                }
                if (g + 2 < groups) fetch_locs(g + 2, (int)(g & 1));               // group g's indices are not needed any more
                __syncthreads();                                                   // every thread has taken group g out of the ring
                if (tid == 0 && g + 2 < groups) issue(g + 2, (int)(g & 1), kglobal + (g + 2) * PROBE_ROWS);
                // Harley–Seal over the 8 rows
#pragma unroll
                for (int w = 0; w < WPT; w++) {
                    uint32_t ones = c[0][w], twos = c[1][w], fours = c[2][w];
                    uint32_t t1a = maj3(ones, r[0][w], r[1][w]); ones = xor3(ones, r[0][w], r[1][w]);
                    uint32_t t1b = maj3(ones, r[2][w], r[3][w]); ones = xor3(ones, r[2][w], r[3][w]);
                    uint32_t t2a = maj3(twos, t1a, t1b); twos = xor3(twos, t1a, t1b);
                    t1a = maj3(ones, r[4][w], r[5][w]); ones = xor3(ones, r[4][w], r[5][w]);
                    t1b = maj3(ones, r[6][w], r[7][w]); ones = xor3(ones, r[6][w], r[7][w]);
                    uint32_t t2b = maj3(twos, t1a, t1b); twos = xor3(twos, t1a, t1b);
                    uint32_t carry = maj3(fours, t2a, t2b); fours = xor3(fours, t2a, t2b);
                    c[0][w] = ones; c[1][w] = twos; c[2][w] = fours;
#pragma unroll
                    for (int p = 3; p < 8; p++) { const uint32_t t = c[p][w] & carry; c[p][w] ^= carry; carry = t; }
                }
            }
            kglobal += n;
            __syncthreads();                                                       // sloc is rewritten by the next task
        }
        // ---- threshold on the bit-sliced counters + hit append (as in probe_kernel) ----
        uint32_t ge[WPT];
        int nhit = 0;
#pragma unroll
        for (int w = 0; w < WPT; w++) ge[w] = 0;
        if (n > 0) {
            const uint32_t Tq = a.thresh[q];
            if (!(Tq >> 8)) {
#pragma unroll
                for (int w = 0; w < WPT; w++) {
                    uint32_t gt = 0, eq = 0xFFFFFFFFu;
#pragma unroll
                    for (int p = 7; p >= 0; p--) {
                        if ((Tq >> p) & 1) eq &= c[p][w];
                        else { gt |= eq & c[p][w]; eq &= ~c[p][w]; }
                    }
                    ge[w] = gt | eq;                                               // words behind the row stayed zero: they never pass (threshold >= 1)
                    nhit += __popc(ge[w]);
                }
            }
        }
        int incl = nhit;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        const int tot = __shfl_sync(0xffffffffu, incl, 31);
        if (tot > 0) {
            unsigned long long base = 0;
            if (lane == 31) base = atomicAdd(a.hit_count, (unsigned long long)tot);
            base = __shfl_sync(0xffffffffu, base, 31);
            unsigned long long slot = base + (unsigned long long)(incl - nhit);
#pragma unroll
            for (int w = 0; w < WPT; w++) {
                uint32_t m = ge[w];
                while (m) {
                    const int bit = __ffs(m) - 1;
                    m &= m - 1;
                    const uint32_t t = (byte0 + ((uint32_t)w * BULK_THREADS + (uint32_t)tid) * 4 + (bit >> 3)) * 8 + (7 - (bit & 7));
                    uint32_t cnt = 0;
#pragma unroll
                    for (int p = 0; p < 8; p++) cnt |= ((c[p][w] >> bit) & 1u) << p;
                    if (slot < a.hit_cap) {
                        a.hit_keys[slot] = ((uint64_t)q << a.target_bits) | (uint64_t)(a.target_base + t);
                        a.hit_vals[slot] = cnt;
                    }
                    slot++;
                }
            }
        }
    }
}

static size_t bulk_smem_bytes(int H, uint32_t slice) { return (size_t)BULK_STAGES * H * slice + BULK_STAGES * sizeof(uint64_t) + 2 * PROBE_ROWS * H * sizeof(uint32_t); }
static void bulk_geometry(const ProbeArgs &a, uint32_t &slice, uint32_t &n_slices) {
    n_slices = (a.pitch + BULK_SLICE - 1) / BULK_SLICE;
    slice = ((a.pitch + n_slices - 1) / n_slices + 127) / 128 * 128;               // equal slices, multiples of 128 bytes
    // the ring must fit the SM's shared memory: narrower slices when h rows of 16 k-mers do not
    while (bulk_smem_bytes(a.num_hashes, slice) > 200 * 1024) { n_slices++; slice = ((a.pitch + n_slices - 1) / n_slices + 127) / 128 * 128; }
}

template <int H, int WPT>
static cudaError_t launch_probe_bulk_hw(const ProbeArgs &a, int sm_count, uint32_t slice, uint32_t n_slices, cudaStream_t st) {
    const size_t smem = bulk_smem_bytes(H, slice);
    static size_t attr = 0;
    if (smem > attr) {
        cudaError_t e = cudaFuncSetAttribute(probe_bulk_kernel<H, WPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr = smem;
    }
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (220 * 1024) / (smem + 1024)));
    const uint64_t tasks = (uint64_t)a.n_queries * n_slices;
    const uint32_t blocks = (uint32_t)std::min<uint64_t>(tasks, (uint64_t)sm_count * per_sm);
    probe_bulk_kernel<H, WPT><<<blocks, BULK_THREADS, smem, st>>>(a, slice, n_slices);
    return cudaGetLastError();
}

template <int H>
static cudaError_t launch_probe_bulk_h(const ProbeArgs &a, int sm_count, cudaStream_t st) {
    uint32_t slice = 0, n_slices = 0;
    bulk_geometry(a, slice, n_slices);
    switch ((slice + 1023) / 1024) {
        case 1: return launch_probe_bulk_hw<H, 1>(a, sm_count, slice, n_slices, st);
        case 2: return launch_probe_bulk_hw<H, 2>(a, sm_count, slice, n_slices, st);
        case 3: return launch_probe_bulk_hw<H, 3>(a, sm_count, slice, n_slices, st);
        default: return launch_probe_bulk_hw<H, 4>(a, sm_count, slice, n_slices, st);
    }
}

// true when the TMA form applies: short queries, precomputed 32-bit row indices, rows of at least 512 bytes, no dense-count dump
static bool probe_bulk_applies(const ProbeArgs &a) {
    return a.planes == 8 && a.locs && !a.dense_counts && a.row_bytes >= 512 && a.fm.d < 0xFFFFFFFFull && a.pitch % 128 == 0;
}

cudaError_t launch_probe(const ProbeArgs &a_in, int sm_count, cudaStream_t st) {
    if (!a_in.n_queries) return cudaSuccess;
    ProbeArgs a = a_in;
    const ProbeTune t = probe_tune();
    a.lanes_per_task_override = (uint32_t)t.g;
#ifdef KMCPG_DEV
    // KMCPG_PROBE_BULK=1: wide rows through the cp.async.bulk + mbarrier form (A/B against the register kernel, profiles/README.md round 2)
    static const bool bulk = getenv("KMCPG_PROBE_BULK") && atoi(getenv("KMCPG_PROBE_BULK")) != 0;
    if (bulk && probe_bulk_applies(a)) {
        switch (a.num_hashes) {
            case 1: return launch_probe_bulk_h<1>(a, sm_count, st);
            case 2: return launch_probe_bulk_h<2>(a, sm_count, st);
            case 3: return launch_probe_bulk_h<3>(a, sm_count, st);
            case 4: return launch_probe_bulk_h<4>(a, sm_count, st);
            default: return cudaErrorInvalidValue;
        }
    }
#endif
    {   // few long queries leave the lane-per-slab mapping without parallelism: give every (query, chunk) a whole CTA
        const uint32_t slab = a.num_hashes == 1 ? 16 : 8;
        const uint32_t row_units = (a.row_bytes + slab - 1) / slab;
        uint32_t G = 1;
        while (G < row_units && G < 128 / slab) G <<= 1;
        const uint64_t tasks = (uint64_t)a.n_queries * ((row_units + G - 1) / G);
        a.long_mode = a.planes > 8 && tasks * G < (uint64_t)sm_count * 512;
        if (a.long_mode) {
            const uint64_t cap2 = (uint64_t)sm_count * (a.planes >= 32 ? 1 : 2);
            uint32_t blocks = (uint32_t)(tasks < cap2 ? tasks : cap2);
            switch (a.num_hashes) {
                case 1: return launch_probe_h<1>(a, blocks, st);
                case 2: return launch_probe_h<2>(a, blocks, st);
                case 3: return launch_probe_h<3>(a, blocks, st);
                case 4: return launch_probe_h<4>(a, blocks, st);
                default: return cudaErrorInvalidValue;
            }
        }
    }
    // upper bound of the thread demand (8-byte slabs need the most lanes); the grid is persistent-style anyway:
    // a multiple of the SM count, several CTAs per SM, each thread loops over tasks
    const uint64_t threads = (uint64_t)a.n_queries * ((a.row_bytes + 7) / 8 + 15);
    uint64_t blocks64 = (threads + PROBE_THREADS - 1) / PROBE_THREADS;
    const uint64_t cap = (uint64_t)sm_count * t.cap;
    uint32_t blocks = (uint32_t)(blocks64 < cap ? blocks64 : cap);
    switch (a.num_hashes) {
        case 1: return launch_probe_h<1>(a, blocks, st);
        case 2: return launch_probe_h<2>(a, blocks, st);
        case 3: return launch_probe_h<3>(a, blocks, st);
        case 4: return launch_probe_h<4>(a, blocks, st);
        default: return cudaErrorInvalidValue;
    }
}

// ------------------------------------------------------------------------------------------------------
// re-pitch rows on upload
// ------------------------------------------------------------------------------------------------------
__global__ void repitch_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, uint64_t n_rows, uint32_t row_bytes, uint32_t pitch) {
    const uint64_t total = n_rows * pitch;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        uint64_t r = i / pitch;
        uint32_t x = (uint32_t)(i - r * pitch);
        dst[i] = x < row_bytes ? src[r * row_bytes + x] : 0;
    }
}
// same for a column range of the rows: bytes [src_off, src_off + row_bytes) of every src row (stride src_stride)
__global__ void repitch_cols_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, uint64_t n_rows, uint32_t src_stride, uint32_t src_off,
                                    uint32_t row_bytes, uint32_t pitch) {
    const uint64_t total = n_rows * pitch;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        uint64_t r = i / pitch;
        uint32_t x = (uint32_t)(i - r * pitch);
        dst[i] = x < row_bytes ? src[r * src_stride + src_off + x] : 0;
    }
}
__global__ void unpitch_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, uint64_t n_rows, uint32_t row_bytes, uint32_t pitch) {
    const uint64_t total = n_rows * row_bytes;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        uint64_t r = i / row_bytes;
        uint32_t x = (uint32_t)(i - r * row_bytes);
        dst[i] = src[r * pitch + x];
    }
}

__global__ void iota_kernel(uint32_t *__restrict__ v, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = i;
}
cudaError_t launch_iota(uint32_t *v, uint32_t n, cudaStream_t st) {
    if (!n) return cudaSuccess;
    iota_kernel<<<(n + 255) / 256, 256, 0, st>>>(v, n);
    return cudaGetLastError();
}

__global__ void pack_hits_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals, uint64_t n, uint32_t qbase, int target_bits,
                                 kmcpg_hit *__restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t k = keys[i];
    kmcpg_hit h;
    h.query = (uint32_t)(k >> target_bits) + qbase; h.target = (uint32_t)(k & ((1ull << target_bits) - 1)); h.count = vals[i];
    out[i] = h;
}
cudaError_t launch_pack_hits(const uint64_t *keys, const uint32_t *vals, uint64_t n, uint32_t qbase, int target_bits, kmcpg_hit *out, cudaStream_t st) {
    if (!n) return cudaSuccess;
    pack_hits_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(keys, vals, n, qbase, target_bits, out);
    return cudaGetLastError();
}

cudaError_t launch_repitch(const uint8_t *src, uint8_t *dst, uint64_t n_rows, uint32_t row_bytes, uint32_t pitch, cudaStream_t st) {
    if (!n_rows) return cudaSuccess;
    repitch_kernel<<<148 * 16, 256, 0, st>>>(src, dst, n_rows, row_bytes, pitch);
    return cudaGetLastError();
}
cudaError_t launch_repitch_cols(const uint8_t *src, uint8_t *dst, uint64_t n_rows, uint32_t src_stride, uint32_t src_off, uint32_t row_bytes, uint32_t pitch,
                                cudaStream_t st) {
    if (!n_rows) return cudaSuccess;
    if (src_off == 0 && src_stride == row_bytes) return launch_repitch(src, dst, n_rows, row_bytes, pitch, st);
    repitch_cols_kernel<<<148 * 16, 256, 0, st>>>(src, dst, n_rows, src_stride, src_off, row_bytes, pitch);
    return cudaGetLastError();
}
cudaError_t launch_unpitch(const uint8_t *src, uint8_t *dst, uint64_t n_rows, uint32_t row_bytes, uint32_t pitch, cudaStream_t st) {
    if (!n_rows) return cudaSuccess;
    unpitch_kernel<<<148 * 16, 256, 0, st>>>(src, dst, n_rows, row_bytes, pitch);
    return cudaGetLastError();
}

}  // namespace kmcpg

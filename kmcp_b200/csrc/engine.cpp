// engine.cpp — host side of the search engine above the device path: the C++ mirror of what
// UnikIndexDBSearchEngine / UnikIndexDB.handleQuery do with the per-block match counts
// (reference kmcp/cmd/util-db-search.go: multi-k loop U:763-1025, --try-se U:808-842 + 995-1011,
//  Match fields and filters U:7466-7491, sort U:273-282 with Less functions U:105-145, top-N scores U:285-311).
// The counts themselves always come from the GPU (kmcpg_search_batch); nothing here probes an index.
//
// A "round" = one device search of the still-undecided queries with one k and one mate selection; its hit
// list (sorted by query) is filtered and sorted per query by a few host threads into one flat array.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <future>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "common.h"
#include "nvtx_ranges.h"

extern "C" const double *kmcpg_internal_target_sizes(const kmcpg_ctx *ctx);
struct kmcpg_stage;
extern "C" int kmcpg_internal_stage_begin(kmcpg_ctx *const *ctxs, int n_ctx, const uint8_t *seq, const uint64_t *off, uint32_t n_seqs, const uint32_t *cuts, int n_pieces,
                                          kmcpg_stage **out);
extern "C" int kmcpg_internal_stage_get(kmcpg_stage *st, int i, int piece, const uint8_t **d_seq, const uint64_t **d_off, void **ready);
extern "C" void kmcpg_internal_stage_end(kmcpg_stage *st);

namespace {

using namespace kmcpg;

// QueryFPRWithCacheWithConstantFPR (F:140-193) without its slot-sharing quirk: plain memo of the pure function.  One table per
// database FPR, shared by every thread that filters hits of such a database (worker pool, replica threads, several -d databases
// in turn): slots are relaxed atomics holding the double's bits, a racing duplicate store writes the same value.
struct FprCache {
    static constexpr int N = 1024;
    const double p;
    std::unique_ptr<std::atomic<uint64_t>[]> tab;
    static constexpr uint64_t EMPTY = ~0ull;             // a NaN pattern query_fpr never returns
    explicit FprCache(double fpr) : p(fpr), tab(new std::atomic<uint64_t>[(size_t)(N + 1) * (N + 1)]) {
        for (size_t i = 0; i < (size_t)(N + 1) * (N + 1); i++) tab[i].store(EMPTY, std::memory_order_relaxed);
    }
    double get(int n, int c) {
        if (n > N || c > n || n < 0 || c < 0) return query_fpr(n, c, p);
        std::atomic<uint64_t> &slot = tab[(size_t)n * (N + 1) + c];
        uint64_t bits = slot.load(std::memory_order_relaxed);
        double v;
        if (bits == EMPTY) {
            v = query_fpr(n, c, p);
            memcpy(&bits, &v, 8);
            slot.store(bits, std::memory_order_relaxed);
        } else memcpy(&v, &bits, 8);
        return v;
    }
};

struct Less {
    int sort_by;
    bool operator()(const kmcpg_match &a, const kmcpg_match &b) const {
        if (sort_by == 0) {                    // Matches.Less (U:105-114)
            if (a.qcov != b.qcov) return a.qcov > b.qcov;
            if (a.tcov != b.tcov) return a.tcov > b.tcov;
        } else if (sort_by == 1) {             // SortByTCov (U:123-131)
            if (a.tcov != b.tcov) return a.tcov > b.tcov;
            if (a.count != b.count) return a.count > b.count;
        } else {                               // SortByJacc (U:137-145)
            if (a.jacc != b.jacc) return a.jacc > b.jacc;
            if (a.count != b.count) return a.count > b.count;
        }
        return a.target < b.target;            // reference: unstable quicksort, any order among ties
    }
};

// Big result arrays are recycled through a small process-wide pool: a fresh 40 MB allocation per call costs
// more in page faults than the whole post-filter.
struct BigBuf { void *p = nullptr; size_t cap = 0; };
std::mutex g_pool_mu;
std::vector<BigBuf> g_pool;

BigBuf big_acquire(size_t bytes) {
    if (bytes < 64) bytes = 64;
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        int best = -1;
        for (size_t i = 0; i < g_pool.size(); i++)
            if (g_pool[i].cap >= bytes && (best < 0 || g_pool[i].cap < g_pool[best].cap)) best = (int)i;
        if (best >= 0 && g_pool[best].cap <= bytes * 4 + (1u << 20)) {
            BigBuf b = g_pool[best];
            g_pool.erase(g_pool.begin() + best);
            return b;
        }
    }
    BigBuf b;
    b.cap = bytes + bytes / 4;
    b.p = malloc(b.cap);
    return b;
}
void big_release(BigBuf &b) {
    if (!b.p) return;
    std::lock_guard<std::mutex> lk(g_pool_mu);
    g_pool.push_back(b);
    b = BigBuf();
    while (g_pool.size() > 64) {          // several engine calls may run at once (one per replica): keep their arrays too
        size_t m = 0;
        for (size_t i = 1; i < g_pool.size(); i++) if (g_pool[i].cap < g_pool[m].cap) m = i;
        free(g_pool[m].p);
        g_pool.erase(g_pool.begin() + m);
    }
}

// A few persistent workers: spawning threads per part costs more than filtering a part's hits.
class WorkerPool {
  public:
    explicit WorkerPool(int n) {
        for (int i = 0; i < n; i++) th_.emplace_back([this, i] { run(i); });
    }
    ~WorkerPool() {
        { std::lock_guard<std::mutex> lk(mu_); stop_ = true; gen_++; }
        cv_.notify_all();
        for (auto &t : th_) t.join();
    }
    int size() const { return (int)th_.size(); }
    // runs fn(0..n-1); the caller executes task 0, workers the rest (n-1 <= size())
    void parallel(int n, const std::function<void(int)> &fn) {
        if (n <= 1) { if (n == 1) fn(0); return; }
        // one job at a time: a caller that finds the pool busy (several engine calls on different contexts run at once)
        // does its tasks itself — those callers are already parallel to each other
        std::unique_lock<std::mutex> job(job_mu_, std::try_to_lock);
        if (!job.owns_lock()) { for (int i = 0; i < n; i++) fn(i); return; }
        {
            std::lock_guard<std::mutex> lk(mu_);
            fn_ = &fn; n_ = n; pending_ = n - 1; gen_++;
        }
        cv_.notify_all();
        fn(0);
        std::unique_lock<std::mutex> lk(mu_);
        done_.wait(lk, [this] { return pending_ == 0; });
        fn_ = nullptr;
    }

  private:
    void run(int id) {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(int)> *fn = nullptr;
            int task = -1;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
                if (fn_ && id + 1 < n_) { fn = fn_; task = id + 1; }
            }
            if (fn) {
                (*fn)(task);
                std::lock_guard<std::mutex> lk(mu_);
                if (--pending_ == 0) done_.notify_one();
            }
        }
    }
    std::vector<std::thread> th_;
    std::mutex mu_, job_mu_;
    std::condition_variable cv_, done_;
    const std::function<void(int)> *fn_ = nullptr;
    int n_ = 0, pending_ = 0;
    uint64_t gen_ = 0;
    bool stop_ = false;
};

WorkerPool &pool() {
    static WorkerPool p(std::max(1, std::min((int)std::thread::hardware_concurrency() - 1, 15)));
    return p;
}

struct ResPriv {
    BigBuf query_len, n_kmers, k_used, match_off, matches;
    ~ResPriv() { big_release(query_len); big_release(n_kmers); big_release(k_used); big_release(match_off); big_release(matches); }
};

// matches of one round, flat, in local-query order
struct Round {
    std::vector<uint32_t> queries;        // local index → global query
    std::vector<uint64_t> off;            // local index → range in `matches`
    uint64_t *offp = nullptr;             // = off.data(), or the result's own offset array when this round is the whole answer
    BigBuf buf;                           // kmcpg_match[n]
    size_t n = 0;
    kmcpg_match *matches() const { return (kmcpg_match *)buf.p; }
};

// filters + sorts the hits [h0, h1) of a part (U:7466-7491, U:273-311).  The range starts and ends on query boundaries;
// hits are sorted by query, so the groups are found on the fly.  Matches are written densely from dst; count[] (zeroed by
// the caller) receives the number of matches of every query that has any.
size_t post_filter(const kmcpg_engine_opts *o, const kmcpg_part &hits, uint64_t h0, uint64_t h1, const uint32_t *cur, uint32_t round_q0,
                   const double *tsize, FprCache *cache, kmcpg_match *dst, uint32_t *count) {
    NvtxRange nvtx("kmcpg:post-filter (tCov, FPR, sort, top-N)");
    const Less less{o->sort_by};
    const uint32_t q0 = hits.first_query;                              // base of hits[].query inside this device call
    size_t w = 0;
    uint64_t i = h0;
    while (i < h1) {
        const uint32_t lq = hits.hits[i].query - q0;                  // index inside the part
        const int n = hits.n_kmers[lq];
        const uint32_t q = cur ? cur[lq] : round_q0 + lq;
        const double nh = (double)n;
        const size_t start = w;
        for (; i < h1 && hits.hits[i].query - q0 == lq; i++) {
            const kmcpg_hit &h = hits.hits[i];
            const double c = (double)h.count, sz = tsize[h.target];
            const double tcov = c / sz;
            if (!(tcov >= o->min_target_cov)) continue;              // U:7473-7474
            const double fpr = cache->get(n, (int)h.count);
            if (!(fpr <= o->max_fpr)) continue;                      // U:7477-7478
            kmcpg_match &m = dst[w++];
            m.query = q; m.target = h.target; m.count = h.count; m._pad = 0;
            m.fpr = fpr; m.qcov = c / nh; m.tcov = tcov; m.jacc = c / (nh + sz - c);   // U:7470-7488
        }
        size_t cnt = w - start;
        if (cnt > 1 && !o->do_not_sort) std::sort(dst + start, dst + w, less);     // U:273-282
        if (cnt > 0 && o->top_n_scores > 0 && !o->do_not_sort) {     // U:285-311 (kept verbatim, including [:i+1])
            int nsc = 0; double pscore = 1024; size_t j = 0; bool broke = false;
            for (j = 0; j < cnt; j++) {
                const kmcpg_match &m = dst[start + j];
                double score = o->sort_by == 0 ? m.qcov : (o->sort_by == 1 ? m.tcov : m.jacc);
                if (score < pscore) { nsc++; if (nsc > o->top_n_scores) { broke = true; break; } pscore = score; }
            }
            if (broke) { cnt = j + 1; w = start + cnt; }
        }
        count[lq] = (uint32_t)cnt;
    }
    return w;
}

// ---- one database sharded over several contexts (one per GPU) ------------------------------------------------------
// what a shard's delivery callback keeps of a part (the executor's own arrays are only valid during the callback)
struct ShardPart {
    uint32_t first_query = 0, n_queries = 0;
    std::vector<kmcpg_hit> hits;
    std::vector<int32_t> n_kmers, query_len;      // kept by shard 0 only (identical in every shard: they depend on the reads alone)
};

// Union of per-shard hit lists — each sorted by (query, target), disjoint by target — in the same (query, target) order, for the
// queries [q_lo, q_hi) the lists hold.  A counting pass per query, a scatter in shard order (so every query's segment is a
// concatenation of sorted runs), then the segments that are not yet ascending by target (shards whose target ranges interleave,
// as whole-block plans produce) are sorted; column-range shards arrive in target order and only pay the check.
void merge_shard_hits(const kmcpg_hit *const *lists, const uint64_t *n, int k, kmcpg_hit *out, uint32_t q_lo, uint32_t q_hi, std::vector<uint64_t> &pos) {
    const size_t nq = (size_t)(q_hi - q_lo);
    pos.assign(nq + 1, 0);
    int contributing = 0;
    for (int s = 0; s < k; s++) {
        const kmcpg_hit *h = lists[s];
        for (uint64_t i = 0; i < n[s]; i++) pos[(size_t)(h[i].query - q_lo) + 1]++;
        contributing += n[s] > 0;
    }
    for (size_t q = 0; q < nq; q++) pos[q + 1] += pos[q];            // pos[q] = start of query q's segment
    if (contributing <= 1) {                                         // nothing to interleave
        for (int s = 0; s < k; s++) if (n[s]) memcpy(out, lists[s], n[s] * sizeof(kmcpg_hit));
        return;
    }
    for (int s = 0; s < k; s++) {
        const kmcpg_hit *h = lists[s];
        for (uint64_t i = 0; i < n[s]; i++) out[pos[(size_t)(h[i].query - q_lo)]++] = h[i];
    }
    // pos[q] is now the END of segment q; its start is the end of segment q-1
    uint64_t b = 0;
    for (size_t q = 0; q < nq; q++) {
        const uint64_t e = pos[q];
        if (e - b > 1) {
            bool asc = true;
            for (uint64_t i = b + 1; i < e; i++) if (out[i].target < out[i - 1].target) { asc = false; break; }
            if (!asc) std::sort(out + b, out + e, [](const kmcpg_hit &x, const kmcpg_hit &y) { return x.target < y.target; });
        }
        b = e;
    }
}

// the same split over query ranges: thread t takes the queries [first + nq·t/T, first + nq·(t+1)/T) of every list (found by
// binary search) and writes its stretch of the output, whose start is the number of hits in front of that range
void merge_shard_hits_mt(const kmcpg_hit *const *lists, const uint64_t *n, int k, kmcpg_hit *out, uint32_t first_query, uint32_t n_queries, int threads) {
    uint64_t total = 0;
    for (int s = 0; s < k; s++) total += n[s];
    int T = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)std::max(1, threads), total / 65536 + 1));
    T = std::min(T, pool().size() + 1);
    if ((uint32_t)T > n_queries) T = (int)std::max<uint32_t>(1, n_queries);
    if (T <= 1) {
        std::vector<uint64_t> pos;
        merge_shard_hits(lists, n, k, out, first_query, first_query + n_queries, pos);
        return;
    }
    std::vector<uint64_t> cut((size_t)(T + 1) * (size_t)k);          // cut[t*k + s]: first hit of list s at or after thread t's first query
    std::vector<uint32_t> qcut((size_t)T + 1);
    for (int t = 0; t <= T; t++) {
        const uint64_t q = (uint64_t)first_query + (uint64_t)n_queries * (uint64_t)t / (uint64_t)T;
        qcut[t] = (uint32_t)q;
        for (int s = 0; s < k; s++) {
            if (t == 0) { cut[s] = 0; continue; }
            if (t == T) { cut[(size_t)t * k + s] = n[s]; continue; }
            const kmcpg_hit *b = lists[s], *e = lists[s] + n[s];
            cut[(size_t)t * k + s] = (uint64_t)(std::lower_bound(b, e, q, [](const kmcpg_hit &h, uint64_t qq) { return (uint64_t)h.query < qq; }) - b);
        }
    }
    std::function<void(int)> work = [&](int t) {
        std::vector<const kmcpg_hit *> sub((size_t)k);
        std::vector<uint64_t> cnt((size_t)k), pos;
        uint64_t o = 0;
        for (int s = 0; s < k; s++) {
            sub[s] = lists[s] + cut[(size_t)t * k + s];
            cnt[s] = cut[(size_t)(t + 1) * k + s] - cut[(size_t)t * k + s];
            o += cut[(size_t)t * k + s];
        }
        merge_shard_hits(sub.data(), cnt.data(), k, out + o, qcut[t], qcut[t + 1], pos);
    };
    pool().parallel(T, work);
}

// One device round on every shard at once: a host thread per context runs the streamed search, the calling thread merges
// part p of all shards as soon as every shard has delivered it and hands the union to `absorb` — while the GPUs are already
// probing the next parts.  Parts are cut from the read lengths alone, so every shard delivers the same parts.
// `search(shard, cb, user, summary)` is the streamed device call of one shard (kmcpg_search_batch_cb; a stand-in in the host-only self-test).
using ShardSearch = std::function<int(int, kmcpg_part_cb, void *, kmcpg_hits *)>;

int sharded_round(int n_ctx, const ShardSearch &search, const std::function<void(const kmcpg_part &)> &absorb, kmcpg_results *out, int threads) {
    struct Shard {
        std::vector<std::unique_ptr<ShardPart>> parts;
        bool finished = false;
        int rc = KMCPG_OK;
        kmcpg_hits summary;
        std::function<void(const kmcpg_part &)> on_part;
    };
    std::vector<Shard> sh((size_t)n_ctx);
    std::mutex mu;
    std::condition_variable cv;
    std::vector<std::thread> th;
    for (int s = 0; s < n_ctx; s++) {
        memset(&sh[s].summary, 0, sizeof(kmcpg_hits));
        sh[s].on_part = [&, s](const kmcpg_part &pt) {
            std::unique_ptr<ShardPart> sp(new ShardPart());
            sp->first_query = pt.first_query; sp->n_queries = pt.n_queries;
            sp->hits.assign(pt.hits, pt.hits + pt.n_hits);
            if (s == 0) { sp->n_kmers.assign(pt.n_kmers, pt.n_kmers + pt.n_queries); sp->query_len.assign(pt.query_len, pt.query_len + pt.n_queries); }
            std::lock_guard<std::mutex> lk(mu);
            sh[s].parts.push_back(std::move(sp));
            cv.notify_all();
        };
        th.emplace_back([&, s] {
            int rc = search(s, [](void *user, const kmcpg_part *pt) { (*(std::function<void(const kmcpg_part &)> *)user)(*pt); }, &sh[s].on_part, &sh[s].summary);
            std::lock_guard<std::mutex> lk(mu);
            sh[s].rc = rc; sh[s].finished = true;
            cv.notify_all();
        });
    }
    int rc = KMCPG_OK;
    std::vector<kmcpg_hit> merged;
    std::vector<const kmcpg_hit *> lists((size_t)n_ctx);
    std::vector<uint64_t> counts((size_t)n_ctx);
    std::vector<ShardPart *> cur((size_t)n_ctx, nullptr);
    for (size_t pi = 0; rc == KMCPG_OK; pi++) {
        bool all_have = false;
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] {
                bool have = true, stuck = false;
                for (auto &x : sh) { if (x.parts.size() <= pi) { have = false; if (x.finished) stuck = true; } }
                all_have = have;
                return have || stuck;
            });
            if (!all_have) {
                // a shard ended without part pi: the regular end (every shard finished with exactly pi parts) or a failure
                for (auto &x : sh) if (x.finished && x.rc) rc = x.rc;
                if (rc == KMCPG_OK) {
                    cv.wait(lk, [&] { for (auto &x : sh) if (!x.finished) return false; return true; });
                    for (auto &x : sh) { if (x.rc) rc = x.rc; else if (x.parts.size() != pi) rc = KMCPG_EINVAL; }
                }
                break;
            }
            // the part objects are heap-stable; the vectors holding them may be regrown by the shard threads, so take the pointers here
            for (int s = 0; s < n_ctx; s++) { cur[s] = sh[s].parts[pi].get(); lists[s] = cur[s]->hits.data(); counts[s] = cur[s]->hits.size(); }
        }
        const ShardPart &p0 = *cur[0];
        uint64_t total = 0;
        for (int s = 0; s < n_ctx; s++) {
            const ShardPart &ps = *cur[s];
            if (ps.first_query != p0.first_query || ps.n_queries != p0.n_queries) rc = KMCPG_EINVAL;   // cannot happen: same reads, same cuts
            total += counts[s];
        }
        if (rc) break;
        merged.resize(std::max<uint64_t>(total, 1));
        merge_shard_hits_mt(lists.data(), counts.data(), n_ctx, merged.data(), p0.first_query, p0.n_queries, threads);
        kmcpg_part pt;
        pt.first_query = p0.first_query; pt.n_queries = p0.n_queries;
        pt.n_kmers = p0.n_kmers.data(); pt.query_len = p0.query_len.data();
        pt.hits = merged.data(); pt.n_hits = total;
        absorb(pt);
        for (int s = 0; s < n_ctx; s++) { std::lock_guard<std::mutex> lk(mu); sh[s].parts[pi].reset(); }
    }
    for (auto &t : th) t.join();
    float ms = 0;
    for (int s = 0; s < n_ctx; s++) {
        if (rc == KMCPG_OK && sh[s].rc) rc = sh[s].rc;
        if (sh[s].rc == KMCPG_OK) {
            ms = std::max(ms, sh[s].summary.ms_total);
            out->probe_row_bytes += sh[s].summary.probe_row_bytes; out->kernel_launches += sh[s].summary.kernel_launches;
            kmcpg_free_hits(&sh[s].summary);
        }
    }
    out->ms_gpu_total += ms;
    return rc;
}

int engine_search_impl(kmcpg_ctx *const *ctxs, int n_ctx, const kmcpg_engine_opts *o, const uint8_t *seq, const uint64_t *off, uint32_t n_seqs,
                       kmcpg_results *out, FprCache *shared_cache = nullptr);

// Host-only test hook (kmcpg_internal_engine_standin): what the engine asks a context for — database facts, target sizes, one
// streamed device call — answered from arrays, so the result handling (U:7466-7491, U:273-311) runs without a device.
struct StandIn {
    kmcpg_db_info_t info;
    const double *tsize;
    std::function<int(const kmcpg_search_params *, uint32_t, kmcpg_part_cb, void *, kmcpg_hits *)> device_call;
};
thread_local const StandIn *tl_standin = nullptr;

// the memo of this database's p (F:140-193): tables live as long as the process, one per distinct FPR (databases built with
// the same -f share one)
FprCache *thread_fpr_cache(double fpr) {
    static std::mutex mu;
    static std::vector<std::unique_ptr<FprCache>> all;
    std::lock_guard<std::mutex> lk(mu);
    for (auto &c : all) if (c->p == fpr) return c.get();
    all.emplace_back(new FprCache(fpr));
    return all.back().get();
}

}  // namespace

extern "C" {

void kmcpg_default_engine_opts(kmcpg_engine_opts *o) {
    if (!o) return;
    memset(o, 0, sizeof(*o));
    o->min_query_len = 30; o->min_matched = 10; o->dedup_threshold = 256;
    o->min_query_cov = 0.55; o->min_target_cov = 0; o->max_fpr = 0.01;      // S:1055-1075
}

double kmcpg_query_fpr(int n, int c, double p) { return query_fpr(n, c, p); }

int kmcpg_engine_search(kmcpg_ctx *ctx, const kmcpg_engine_opts *o, const uint8_t *seq, const uint64_t *off, uint32_t n_seqs, kmcpg_results *out) {
    if (!ctx || !o || !out) return KMCPG_EINVAL;
    return engine_search_impl(&ctx, 1, o, seq, off, n_seqs, out);
}

int kmcpg_engine_search_sharded(kmcpg_ctx *const *ctxs, int n_ctx, const kmcpg_engine_opts *o, const uint8_t *seq, const uint64_t *off, uint32_t n_seqs,
                                kmcpg_results *out) {
    if (!ctxs || n_ctx < 1 || n_ctx > 64 || !o || !out) return KMCPG_EINVAL;
    for (int i = 0; i < n_ctx; i++) {
        if (!ctxs[i]) return KMCPG_EINVAL;
        for (int j = 0; j < i; j++) if (ctxs[j] == ctxs[i]) return KMCPG_EINVAL;      // a context runs one call at a time
    }
    return engine_search_impl(ctxs, n_ctx, o, seq, off, n_seqs, out);
}

}  // extern "C"

namespace {

int engine_search_impl(kmcpg_ctx *const *ctxs, int n_ctx, const kmcpg_engine_opts *o, const uint8_t *seq, const uint64_t *off, uint32_t n_seqs,
                       kmcpg_results *out, FprCache *shared_cache) {
    kmcpg_ctx *ctx = ctxs[0];
    const StandIn *standin = tl_standin;
    kmcpg_db_info_t info;
    int rc = KMCPG_OK;
    if (standin) info = standin->info;
    else rc = kmcpg_db_info(ctx, &info);
    if (rc) return rc;
    for (int i = 1; i < n_ctx; i++) {                 // every shard must hold (a part of) the same database
        kmcpg_db_info_t oi;
        rc = kmcpg_db_info(ctxs[i], &oi);
        if (rc) return rc;
        if (oi.n_targets != info.n_targets || oi.n_blocks != info.n_blocks || oi.n_ks != info.n_ks || memcmp(oi.ks, info.ks, sizeof(info.ks)) ||
            oi.fpr != info.fpr || oi.num_hashes != info.num_hashes)
            return KMCPG_EINVAL;
    }
    memset(out, 0, sizeof(*out));
    auto T0 = std::chrono::steady_clock::now();
    auto ms_since = [](std::chrono::steady_clock::time_point t) { return std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t).count(); };
    const uint32_t step = o->paired ? 2 : 1;
    const uint32_t nq = n_seqs / step;
    FprCache *cache = shared_cache ? shared_cache : thread_fpr_cache(info.fpr);      // worker threads share THIS instance, not their own thread_local
    const double *tsize = standin ? standin->tsize : kmcpg_internal_target_sizes(ctx);

    ResPriv *priv = new ResPriv();
    priv->query_len = big_acquire((size_t)nq * 4); priv->n_kmers = big_acquire((size_t)nq * 4); priv->k_used = big_acquire((size_t)nq * 4);
    priv->match_off = big_acquire(((size_t)nq + 1) * 8);
    int32_t *r_qlen = (int32_t *)priv->query_len.p, *r_nk = (int32_t *)priv->n_kmers.p, *r_k = (int32_t *)priv->k_used.p;
    uint64_t *r_off = (uint64_t *)priv->match_off.p;
    const int tries_max = (o->try_se && o->paired) ? 3 : 1;
    int threads = o->threads > 0 ? o->threads : (int)std::thread::hardware_concurrency();
    threads = std::max(1, std::min(threads, pool().size() + 1));

    // scratch that survives between calls of this thread (no page faults after the first batch)
    static thread_local std::vector<uint32_t> tl_count;
    std::vector<uint32_t> &count = tl_count;      // a plain reference: the filter thread must use THIS thread's scratch
    std::vector<Round> rounds;
    std::vector<uint32_t> pending;                // empty + all_pending = every query
    bool all_pending = true;
    std::vector<int32_t> q_round;                 // only materialised when more than one round happens
    std::vector<uint32_t> q_local;
    std::vector<uint8_t> sub_seq;
    std::vector<uint64_t> sub_off;
    for (int ik = 0; ik < info.n_ks && (all_pending || !pending.empty()); ik++) {
        const int k = info.ks[ik];
        std::vector<uint32_t> next_k;                 // queries that found nothing with this k
        std::vector<uint32_t> cur = pending;
        bool cur_all = all_pending;
        for (int tries = 0; tries < tries_max && (cur_all || !cur.empty()); tries++) {
            // pack the pending subset (the first pass uses the caller's buffers untouched)
            const uint8_t *bs = seq; const uint64_t *bo = off;
            if (!cur_all) {
                sub_off.assign(1, 0); sub_seq.clear();
                for (uint32_t q : cur)
                    for (uint32_t m = 0; m < step; m++) {
                        uint64_t a = off[q * step + m], b = off[q * step + m + 1];
                        sub_seq.insert(sub_seq.end(), seq + a, seq + b);
                        sub_off.push_back(sub_seq.size());
                    }
                if (sub_seq.empty()) sub_seq.push_back(0);
                bs = sub_seq.data(); bo = sub_off.data();
            }
            kmcpg_search_params p;
            kmcpg_default_params(&p);
            p.min_query_len = o->min_query_len; p.min_matched = o->min_matched; p.dedup_threshold = o->dedup_threshold;
            p.paired = o->paired; p.min_query_cov = o->min_query_cov; p.k = k; p.mate_select = tries;

            const uint32_t ln_total = cur_all ? nq : (uint32_t)cur.size();
            const uint32_t *curp = cur_all ? nullptr : cur.data();
            rounds.emplace_back();
            Round &R = rounds.back();
            const int ridx = (int)rounds.size() - 1;
            const bool single = (ik == 0 && tries == 0 && cur_all);
            const bool whole_answer = single && tries_max == 1 && info.n_ks == 1;      // offsets go straight into the result
            if (whole_answer) { R.offp = r_off; r_off[0] = 0; }
            else { R.off.assign((size_t)ln_total + 1, 0); R.offp = R.off.data(); }
            if (!single && q_round.empty()) { q_round.assign(nq, -1); q_local.assign(nq, 0); }
            std::vector<uint32_t> retry;

            // filters the hits of one delivered part (local queries [q0, q0+ln) of this round) and appends the matches to R
            auto absorb = [&](const kmcpg_part &hits, uint32_t q0 /* position of the part's first query inside the round */) {
                auto Tp = std::chrono::steady_clock::now();
                const uint32_t ln = hits.n_queries;
                count.assign(ln, 0);
                const size_t need = (R.n + std::max<uint64_t>(hits.n_hits, 1)) * sizeof(kmcpg_match);
                if (R.buf.cap < need) {
                    // first part of a round: size the array for the whole round from this part's hit density
                    size_t want = need + need / 2;
                    if (R.n == 0 && ln > 0 && ln < ln_total)
                        want = std::max(want, (size_t)((double)hits.n_hits * ((double)ln_total / (double)ln) * 1.2 + 4096) * sizeof(kmcpg_match));
                    BigBuf nb = big_acquire(want);
                    if (R.n) memcpy(nb.p, R.buf.p, R.n * sizeof(kmcpg_match));
                    big_release(R.buf);
                    R.buf = nb;
                }
                kmcpg_match *M = R.matches() + R.n;
                int T = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)threads, hits.n_hits / 8192 + 1));
                // split the hit list evenly, moving every cut forward to the next query boundary
                std::vector<uint64_t> cut(T + 1, hits.n_hits);
                std::vector<size_t> produced(T, 0);
                cut[0] = 0;
                for (int t = 1; t < T; t++) {
                    uint64_t c = std::max<uint64_t>(cut[t - 1], hits.n_hits * (uint64_t)t / T);
                    while (c > 0 && c < hits.n_hits && hits.hits[c].query == hits.hits[c - 1].query) c++;
                    cut[t] = c;
                }
                // every thread writes at the index of its first hit: no overlap, compaction only if a filter dropped hits
                uint32_t *cp = count.data();
                const uint32_t *lcur = curp ? curp + q0 : nullptr;
                auto work = [&](int t) { produced[t] = post_filter(o, hits, cut[t], cut[t + 1], lcur, q0, tsize, cache, M + cut[t], cp); };
                pool().parallel(T, work);
                size_t w = produced[0];
                for (int t = 1; t < T; t++) {
                    if (produced[t] && w != cut[t]) memmove(M + w, M + cut[t], produced[t] * sizeof(kmcpg_match));
                    w += produced[t];
                }
                R.n += w;
                const bool more_rounds_possible = tries + 1 < tries_max || ik + 1 < info.n_ks;
                if (!curp && !more_rounds_possible && q_round.empty()) {
                    // the common case (one k, no --try-se, every query in the round): bulk copies, no per-query bookkeeping
                    memcpy(r_qlen + q0, hits.query_len, (size_t)ln * 4);
                    memcpy(r_nk + q0, hits.n_kmers, (size_t)ln * 4);
                    std::fill(r_k + q0, r_k + q0 + ln, (int32_t)k);
                    uint64_t acc = R.offp[q0];
                    for (uint32_t l = 0; l < ln; l++) { acc += cp[l]; R.offp[q0 + l + 1] = acc; }
                } else {
                    for (uint32_t l = 0; l < ln; l++) {
                        const uint32_t gl = q0 + l;                                    // index inside the round
                        const uint32_t q = curp ? curp[gl] : gl;
                        const int n = hits.n_kmers[l];
                        R.offp[gl + 1] = R.offp[gl] + cp[l];
                        r_qlen[q] = hits.query_len[l];
                        r_k[q] = k;
                        if (n == 0) { if (tries == 0) r_nk[q] = 0; continue; }         // U:778-786, U:854-869: final, unmatched
                        r_nk[q] = n;
                        if (cp[l]) { if (!q_round.empty()) { q_round[q] = ridx; q_local[q] = gl; } }
                        else retry.push_back(q);
                    }
                }
                out->ms_post += ms_since(Tp);
            };

            // one device call per round: the executor hands over every part (≈ 250 k reads) as soon as its hits have landed in
            // host memory and has the next part's kernels enqueued by then, so the filtering below runs beside the GPU work;
            // only the last part's filtering is exposed
            if (n_ctx > 1) {
                std::function<void(const kmcpg_part &)> fn = [&](const kmcpg_part &pt) { absorb(pt, pt.first_query); };
                const uint32_t round_seqs = ln_total * step;
                // The batch crosses PCIe once (into the first context's device) and reaches the others by peer copies, in a few pieces;
                // every shard searches piece j as a job of its own, so piece j+1 travels while piece j is probed.
                std::vector<uint32_t> cuts{0};
                {
                    const uint64_t bytes = round_seqs ? bo[round_seqs] - bo[0] : 0;
                    const int n_pieces = (int)std::max<uint64_t>(1, std::min<uint64_t>(4, bytes / (64ull << 20)));
                    for (int j = 1; j < n_pieces; j++) {
                        const uint64_t want = bo[0] + bytes * (uint64_t)j / (uint64_t)n_pieces;
                        uint32_t lo = cuts.back() / step, hi = ln_total;       // first query whose first byte is at or after `want`
                        while (lo < hi) { const uint32_t mid = lo + (hi - lo) / 2; if (bo[(size_t)mid * step] < want) lo = mid + 1; else hi = mid; }
                        if (lo * step > cuts.back() && lo < ln_total) cuts.push_back(lo * step);
                    }
                    cuts.push_back(round_seqs);
                }
                const int n_pieces = (int)cuts.size() - 1;
                kmcpg_stage *stage = nullptr;
                if (round_seqs) {
                    rc = kmcpg_internal_stage_begin(ctxs, n_ctx, bs, bo, round_seqs, cuts.data(), n_pieces, &stage);
                    if (rc) { for (auto &r : rounds) big_release(r.buf); delete priv; return rc; }
                }
                ShardSearch dev = [&](int s, kmcpg_part_cb cb, void *user, kmcpg_hits *summary) {
                    if (!stage) return kmcpg_search_batch_cb(ctxs[s], &p, bs, bo, round_seqs, cb, user, summary);
                    const auto t0 = std::chrono::steady_clock::now();
                    std::vector<kmcpg_job *> jobs((size_t)n_pieces, nullptr);
                    int r = KMCPG_OK;
                    for (int j = 0; j < n_pieces && !r; j++) {
                        kmcpg_batch b;
                        memset(&b, 0, sizeof(b));
                        void *ready = nullptr;
                        r = kmcpg_internal_stage_get(stage, s, j, &b.seq, &b.off, &ready);
                        if (r) break;
                        b.off += cuts[j]; b.host_off = bo + cuts[j]; b.n_seqs = cuts[j + 1] - cuts[j]; b.on_device = 1; b.ready_event = ready;
                        b.cb = cb; b.user = user; b.first_query = cuts[j] / step;
                        r = kmcpg_search_submit(ctxs[s], &p, &b, &jobs[j]);
                    }
                    memset(summary, 0, sizeof(*summary));
                    for (int j = 0; j < n_pieces; j++) {
                        if (!jobs[j]) continue;
                        kmcpg_hits h;
                        const int rw = kmcpg_search_wait(jobs[j], &h);
                        if (rw) { if (!r) r = rw; continue; }
                        summary->probe_row_bytes += h.probe_row_bytes; summary->kernel_launches += h.kernel_launches; summary->probe_launches += h.probe_launches;
                        summary->ms_probe += h.ms_probe; summary->ms_hash += h.ms_hash; summary->n_hits += h.n_hits; summary->n_queries += h.n_queries;
                        kmcpg_free_hits(&h);
                    }
                    summary->ms_total = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
                    return r;
                };
                rc = sharded_round(n_ctx, dev, fn, out, threads);
                kmcpg_internal_stage_end(stage);
                if (rc) { for (auto &r : rounds) big_release(r.buf); delete priv; return rc; }
            } else {
                std::function<void(const kmcpg_part &)> fn = [&](const kmcpg_part &pt) { absorb(pt, pt.first_query); };
                kmcpg_hits h;
                memset(&h, 0, sizeof(h));
                const kmcpg_part_cb tramp = [](void *user, const kmcpg_part *pt) { (*(std::function<void(const kmcpg_part &)> *)user)(*pt); };
                rc = standin ? standin->device_call(&p, ln_total * step, tramp, &fn, &h) : kmcpg_search_batch_cb(ctx, &p, bs, bo, ln_total * step, tramp, &fn, &h);
                if (rc) { for (auto &r : rounds) big_release(r.buf); delete priv; return rc; }
                out->ms_gpu_total += h.ms_total; out->probe_row_bytes += h.probe_row_bytes; out->kernel_launches += h.kernel_launches;
                if (!standin) kmcpg_free_hits(&h);
            }
            if (single && (tries_max > 1 || info.n_ks > 1) && !retry.empty()) {
                // more rounds will follow: remember where round 0 put every matched query
                q_round.assign(nq, -1); q_local.assign(nq, 0);
                for (uint32_t l = 0; l < ln_total; l++) if (R.offp[l + 1] > R.offp[l]) { q_round[l] = 0; q_local[l] = l; }
            }
            if (!cur_all) R.queries.swap(cur);
            cur_all = false;
            if (tries + 1 < tries_max) cur.swap(retry);          // --try-se: read1 only, then read2 only
            else { next_k.insert(next_k.end(), retry.begin(), retry.end()); cur.clear(); }
        }
        std::sort(next_k.begin(), next_k.end());
        pending.swap(next_k);                                    // U:1018-1023: try the next smaller k
        all_pending = false;
    }

    if (rounds.size() == 1) {
        // the common case: one round over all queries — its flat array already is the answer
        priv->matches = rounds[0].buf; rounds[0].buf = BigBuf();
        if (rounds[0].offp != r_off) memcpy(r_off, rounds[0].offp, ((size_t)nq + 1) * 8);
        out->n_matches = rounds[0].n;
    } else {
        r_off[0] = 0;
        for (uint32_t q = 0; q < nq; q++) {
            uint64_t c = 0;
            if (!q_round.empty() && q_round[q] >= 0) { const Round &R = rounds[q_round[q]]; c = R.offp[q_local[q] + 1] - R.offp[q_local[q]]; }
            r_off[q + 1] = r_off[q] + c;
        }
        priv->matches = big_acquire(std::max<uint64_t>(r_off[nq], 1) * sizeof(kmcpg_match));
        kmcpg_match *dst = (kmcpg_match *)priv->matches.p;
        for (uint32_t q = 0; q < nq; q++)
            if (!q_round.empty() && q_round[q] >= 0) {
                const Round &R = rounds[q_round[q]];
                memcpy(dst + r_off[q], R.matches() + R.offp[q_local[q]], (r_off[q + 1] - r_off[q]) * sizeof(kmcpg_match));
            }
        out->n_matches = r_off[nq];
        for (auto &R : rounds) big_release(R.buf);
    }
    out->n_queries = nq;
    out->query_len = r_qlen; out->n_kmers = r_nk; out->k_used = r_k;
    out->match_off = r_off; out->matches = (kmcpg_match *)priv->matches.p;
    out->_priv = priv;
    out->ms_total = ms_since(T0);
    return KMCPG_OK;
}

// ---- the whole database on every context, the READS split between them ------------------------------------------------
// For databases that fit every GPU this is the better split: a context hashes and probes only its share of the queries
// (block / column shards all hash every read and all copy the whole batch), and the host post-filter runs once per replica
// in parallel.  Queries are independent, so the answer is the concatenation of the per-range answers.
using ReplicaSearch = std::function<int(int, const uint64_t *, uint32_t, kmcpg_results *)>;   // (replica, off of its range, n_seqs, out)

kmcpg_results alloc_results(uint32_t nq, uint64_t n_matches) {
    kmcpg_results r;
    memset(&r, 0, sizeof(r));
    ResPriv *priv = new ResPriv();
    priv->query_len = big_acquire((size_t)nq * 4); priv->n_kmers = big_acquire((size_t)nq * 4); priv->k_used = big_acquire((size_t)nq * 4);
    priv->match_off = big_acquire(((size_t)nq + 1) * 8);
    priv->matches = big_acquire(std::max<uint64_t>(n_matches, 1) * sizeof(kmcpg_match));
    r.n_queries = nq; r.n_matches = n_matches;
    r.query_len = (int32_t *)priv->query_len.p; r.n_kmers = (int32_t *)priv->n_kmers.p; r.k_used = (int32_t *)priv->k_used.p;
    r.match_off = (uint64_t *)priv->match_off.p; r.matches = (kmcpg_match *)priv->matches.p;
    r.match_off[0] = 0;
    r._priv = priv;
    return r;
}

int replicas_impl(int n_rep, const ReplicaSearch &search, bool paired, const uint64_t *off, uint32_t n_seqs, kmcpg_results *out) {
    auto T0 = std::chrono::steady_clock::now();
    const uint32_t step = paired ? 2 : 1;
    if (paired && (n_seqs & 1)) return KMCPG_EINVAL;
    const uint32_t nq = n_seqs / step;
    // contiguous query ranges with about the same number of sequence bytes (by count when there are no bytes)
    std::vector<uint32_t> cut((size_t)n_rep + 1, nq);
    cut[0] = 0;
    const uint64_t b0 = n_seqs ? off[0] : 0, total = n_seqs ? off[n_seqs] - off[0] : 0;
    for (int r = 1; r < n_rep; r++) {
        uint32_t q;
        if (total == 0) q = (uint32_t)((uint64_t)nq * (uint64_t)r / (uint64_t)n_rep);
        else {
            const uint64_t want = b0 + (uint64_t)((long double)total * (long double)r / (long double)n_rep);
            uint32_t lo = cut[r - 1], hi = nq;                      // first query whose first byte is at or after `want`
            while (lo < hi) { const uint32_t mid = lo + (hi - lo) / 2; if (off[(size_t)mid * step] < want) lo = mid + 1; else hi = mid; }
            q = lo;
        }
        cut[r] = std::max(cut[r - 1], std::min(q, nq));
    }
    std::vector<kmcpg_results> res((size_t)n_rep);
    std::vector<int> rcs((size_t)n_rep, KMCPG_OK);
    for (auto &r : res) memset(&r, 0, sizeof(r));
    {
        std::vector<std::thread> th;
        auto run = [&](int r) { if (cut[r + 1] > cut[r]) rcs[r] = search(r, off + (size_t)cut[r] * step, (cut[r + 1] - cut[r]) * step, &res[r]); };
        for (int r = 1; r < n_rep; r++) th.emplace_back(run, r);
        run(0);
        for (auto &t : th) t.join();
    }
    int rc = KMCPG_OK;
    for (int r = 0; r < n_rep; r++) if (rcs[r] && !rc) rc = rcs[r];
    if (rc) {
        for (int r = 0; r < n_rep; r++) if (!rcs[r] && res[r]._priv) kmcpg_free_results(&res[r]);
        return rc;
    }
    std::vector<uint64_t> mbase((size_t)n_rep + 1, 0);
    for (int r = 0; r < n_rep; r++) {
        if (cut[r + 1] > cut[r] && res[r].n_queries != cut[r + 1] - cut[r]) rc = KMCPG_EINVAL;     // cannot happen
        mbase[r + 1] = mbase[r] + (cut[r + 1] > cut[r] ? res[r].n_matches : 0);
    }
    if (rc) { for (auto &r : res) if (r._priv) kmcpg_free_results(&r); return rc; }
    *out = alloc_results(nq, mbase[n_rep]);
    std::function<void(int)> copy = [&](int r) {                    // every replica's answer moves to its place, side by side
        const uint32_t a = cut[r], n = cut[r + 1] - cut[r];
        if (!n) return;
        const kmcpg_results &x = res[r];
        memcpy(out->query_len + a, x.query_len, (size_t)n * 4);
        memcpy(out->n_kmers + a, x.n_kmers, (size_t)n * 4);
        memcpy(out->k_used + a, x.k_used, (size_t)n * 4);
        for (uint32_t i = 0; i < n; i++) out->match_off[(size_t)a + i + 1] = mbase[r] + x.match_off[i + 1];
    };
    // the match arrays are the bulk (48 B per match): every replica's array is moved by several threads
    const int T = pool().size() + 1;                                // the pool runs at most size()+1 tasks of a job
    std::function<void(int)> lane = [&](int t) {
        for (int r = t; r < n_rep; r += T) copy(r);
        for (int r = 0; r < n_rep; r++) {
            if (cut[r + 1] == cut[r]) continue;
            const kmcpg_results &x = res[r];
            const uint64_t a = x.n_matches * (uint64_t)t / (uint64_t)T, b = x.n_matches * (uint64_t)(t + 1) / (uint64_t)T;
            kmcpg_match *dst = out->matches + mbase[r];
            const uint32_t qa = cut[r];
            for (uint64_t i = a; i < b; i++) { dst[i] = x.matches[i]; dst[i].query += qa; }
        }
    };
    pool().parallel(T, lane);
    for (int r = 0; r < n_rep; r++) {
        if (!res[r]._priv) continue;
        out->ms_gpu_total = std::max(out->ms_gpu_total, res[r].ms_gpu_total); out->ms_post = std::max(out->ms_post, res[r].ms_post);
        out->probe_row_bytes += res[r].probe_row_bytes; out->kernel_launches += res[r].kernel_launches;
        kmcpg_free_results(&res[r]);
    }
    out->ms_total = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - T0).count();
    return KMCPG_OK;
}

}  // namespace

extern "C" int kmcpg_internal_holds_whole_db(const kmcpg_ctx *ctx);

extern "C" {

int kmcpg_engine_search_replicas(kmcpg_ctx *const *ctxs, int n_ctx, const kmcpg_engine_opts *o, const uint8_t *seq, const uint64_t *off, uint32_t n_seqs,
                                 kmcpg_results *out) {
    if (!ctxs || n_ctx < 1 || n_ctx > 64 || !o || !out || (n_seqs && (!seq || !off))) return KMCPG_EINVAL;
    kmcpg_db_info_t info;
    for (int i = 0; i < n_ctx; i++) {
        if (!ctxs[i]) return KMCPG_EINVAL;
        for (int j = 0; j < i; j++) if (ctxs[j] == ctxs[i]) return KMCPG_EINVAL;
        kmcpg_db_info_t oi;
        int rc = kmcpg_db_info(ctxs[i], &oi);
        if (rc) return rc;
        if (!kmcpg_internal_holds_whole_db(ctxs[i])) return KMCPG_EINVAL;            // a shard cannot answer alone
        if (i == 0) info = oi;
        else if (oi.n_targets != info.n_targets || oi.n_blocks != info.n_blocks || oi.disk_bytes != info.disk_bytes || oi.fpr != info.fpr ||
                 oi.n_ks != info.n_ks || memcmp(oi.ks, info.ks, sizeof(info.ks)) || oi.num_hashes != info.num_hashes)
            return KMCPG_EINVAL;
    }
    memset(out, 0, sizeof(*out));
    FprCache *cache = thread_fpr_cache(info.fpr);       // one memo for all replicas
    ReplicaSearch dev = [&](int r, const uint64_t *off_r, uint32_t n_r, kmcpg_results *res) {
        return engine_search_impl(&ctxs[r], 1, o, seq, off_r, n_r, res, cache);
    };
    return replicas_impl(n_ctx, dev, o->paired != 0, off, n_seqs, out);
}

// test hook (host only): replicas_impl with stand-in replicas whose answer is a seeded function of the GLOBAL query index, so the
// concatenation over n_rep ranges must equal the answer of one replica over the whole batch.  fail_rep >= 0: that replica fails.
int kmcpg_internal_replicas_selftest(int n_rep, uint32_t nq, int paired, int fail_rep, uint64_t seed) {
    if (n_rep < 1 || n_rep > 64) return KMCPG_EINVAL;
    auto mix = [](uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return x; };
    const uint32_t step = paired ? 2 : 1, n_seqs = nq * step;
    std::vector<uint64_t> off((size_t)n_seqs + 1);
    off[0] = 1000;                                        // a batch that does not start at offset 0
    for (uint32_t i = 0; i < n_seqs; i++) {
        const uint64_t h = mix(seed * 1315423911ull + i);
        off[i + 1] = off[i] + (h % 7 == 0 ? 0 : (h % 5 == 0 ? 20000 : 100 + h % 100));
    }
    ReplicaSearch fake = [&](int r, const uint64_t *off_r, uint32_t n_r, kmcpg_results *res) {
        if (r == fail_rep) return (int)KMCPG_ECUDA;
        const uint32_t base = (uint32_t)((off_r - off.data()) / step), n = n_r / step;
        std::vector<uint32_t> cnt(n);
        uint64_t tot = 0;
        for (uint32_t q = 0; q < n; q++) { cnt[q] = (uint32_t)(mix(seed + 7 * (uint64_t)(base + q)) % 4); tot += cnt[q]; }
        *res = alloc_results(n, tot);
        uint64_t w = 0;
        for (uint32_t q = 0; q < n; q++) {
            const uint32_t g = base + q;
            res->query_len[q] = (int32_t)(off_r[(size_t)q * step + step] - off_r[(size_t)q * step]);
            res->n_kmers[q] = (int32_t)(mix(g) % 130); res->k_used[q] = 21;
            for (uint32_t j = 0; j < cnt[q]; j++) {
                kmcpg_match &m = res->matches[w++];
                m.query = q; m.target = (uint32_t)(mix(g * 31ull + j) % 1000); m.count = j + 1; m._pad = 0;
                m.fpr = 1e-9 * g; m.qcov = 0.5 + j; m.tcov = 0.25 * g; m.jacc = (double)j / (g + 1.0);
            }
            res->match_off[q + 1] = w;
        }
        res->ms_gpu_total = 1.0f + r; res->ms_post = 0.5f; res->probe_row_bytes = 10; res->kernel_launches = 2;
        std::this_thread::sleep_for(std::chrono::microseconds(mix(seed + r) % 500));
        return (int)KMCPG_OK;
    };
    kmcpg_results one, many;
    memset(&one, 0, sizeof(one)); memset(&many, 0, sizeof(many));
    int rc = replicas_impl(n_rep, fake, paired != 0, off.data(), n_seqs, &many);
    if (rc) return rc;
    const int fr = fail_rep; fail_rep = -1;
    rc = replicas_impl(1, fake, paired != 0, off.data(), n_seqs, &one);
    fail_rep = fr;
    if (rc) { kmcpg_free_results(&many); return rc; }
    bool same = one.n_queries == many.n_queries && one.n_matches == many.n_matches && one.n_queries == nq;
    if (same && nq) same = !memcmp(one.query_len, many.query_len, (size_t)nq * 4) && !memcmp(one.n_kmers, many.n_kmers, (size_t)nq * 4) &&
                           !memcmp(one.k_used, many.k_used, (size_t)nq * 4);
    if (same) same = !memcmp(one.match_off, many.match_off, ((size_t)nq + 1) * 8);
    if (same && one.n_matches) same = !memcmp(one.matches, many.matches, one.n_matches * sizeof(kmcpg_match));
    kmcpg_free_results(&one); kmcpg_free_results(&many);
    return same ? KMCPG_OK : KMCPG_EINVAL;
}

// The engine's result handling over hit lists that were produced elsewhere — by the other processes of a one-process-per-GPU run
// (each rank probes its blocks, the gathering rank holds the concatenated lists) or by a test.  Host only.  `hits` is sorted by
// (query, target); the lists are handed to the same code that kmcpg_engine_search runs behind the device call, in parts of
// part_queries queries.  Single-round configurations only (one k, no --try-se): retries need the device.
int kmcpg_engine_postfilter(const kmcpg_engine_opts *o, uint32_t n_queries, const int32_t *n_kmers, const int32_t *query_len, const kmcpg_hit *hits,
                            uint64_t n_hits, const double *target_sizes, int64_t n_targets, double fpr, int k, uint32_t part_queries, kmcpg_results *out) {
    if (!o || !out || o->try_se || (n_queries && (!n_kmers || !query_len)) || (n_hits && !hits) || !target_sizes) return KMCPG_EINVAL;
    if (part_queries == 0) part_queries = 262144;
    StandIn si;
    memset(&si.info, 0, sizeof(si.info));
    si.info.n_ks = 1; si.info.ks[0] = k; si.info.fpr = fpr; si.info.n_targets = n_targets; si.info.num_hashes = 1; si.info.n_blocks = 1;
    si.tsize = target_sizes;
    const uint32_t step = o->paired ? 2 : 1;
    si.device_call = [&](const kmcpg_search_params *, uint32_t n_seqs, kmcpg_part_cb cb, void *user, kmcpg_hits *summary) -> int {
        if (n_seqs / step != n_queries) return KMCPG_EINVAL;
        uint64_t h0 = 0;
        for (uint32_t q0 = 0; q0 < n_queries; q0 += part_queries) {
            const uint32_t nq = std::min(part_queries, n_queries - q0);
            const kmcpg_hit *e = std::lower_bound(hits + h0, hits + n_hits, (uint64_t)q0 + nq, [](const kmcpg_hit &h, uint64_t qq) { return (uint64_t)h.query < qq; });
            const uint64_t h1 = (uint64_t)(e - hits);
            kmcpg_part pt;
            pt.first_query = q0; pt.n_queries = nq;
            pt.n_kmers = n_kmers + q0; pt.query_len = query_len + q0;
            pt.hits = hits + h0; pt.n_hits = h1 - h0;
            cb(user, &pt);
            h0 = h1;
        }
        summary->n_queries = n_queries;
        return h0 == n_hits ? KMCPG_OK : KMCPG_EINVAL;
    };
    std::vector<uint64_t> off((size_t)n_queries * step + 1, 0);          // the sequences themselves never reach the stand-in
    const uint8_t dummy = 0;
    kmcpg_ctx *none = nullptr;
    tl_standin = &si;
    const int rc = engine_search_impl(&none, 1, o, &dummy, off.data(), n_queries * step, out);
    tl_standin = nullptr;
    return rc;
}

// test hook (tests/test_engine_host.py): the former name of kmcpg_engine_postfilter
int kmcpg_internal_engine_standin(const kmcpg_engine_opts *o, uint32_t n_queries, const int32_t *n_kmers, const int32_t *query_len, const kmcpg_hit *hits,
                                  uint64_t n_hits, const double *target_sizes, int64_t n_targets, double fpr, int k, uint32_t part_queries, kmcpg_results *out) {
    if (part_queries == 0) return KMCPG_EINVAL;
    return kmcpg_engine_postfilter(o, n_queries, n_kmers, query_len, hits, n_hits, target_sizes, n_targets, fpr, k, part_queries, out);
}

// Union of per-shard hit lists (each sorted by (query, target), disjoint by target) in (query, target) order: the host side of the
// block fan-out + gather of U:939-964 when the shards live in other processes.  out holds Σ n[i] records.
int kmcpg_merge_hits(const kmcpg_hit *const *lists, const uint64_t *n, int n_lists, uint32_t first_query, uint32_t n_queries, int threads, kmcpg_hit *out) {
    if (n_lists < 0 || (n_lists && (!lists || !n)) || !out) return KMCPG_EINVAL;
    for (int s = 0; s < n_lists; s++) {
        if (n[s] && !lists[s]) return KMCPG_EINVAL;
        if (n[s] && (lists[s][0].query < first_query || (uint64_t)lists[s][n[s] - 1].query >= (uint64_t)first_query + n_queries)) return KMCPG_EINVAL;
    }
    if (n_queries == 0 || n_lists == 0) return KMCPG_OK;
    merge_shard_hits_mt(lists, n, n_lists, out, first_query, n_queries, threads > 0 ? threads : (int)std::thread::hardware_concurrency());
    return KMCPG_OK;
}

// 64-bit digest of a hit list that depends on every field AND on the order: Σ_i mix(i, query, target, count) mod 2^64.  Two runs that
// return the same hit set in the same (query, target) order — e.g. the same search on 1, 2, 4 and 8 GPUs — have the same digest.
uint64_t kmcpg_hits_digest(const kmcpg_hit *hits, uint64_t n, uint64_t first_index) {
    auto mix = [](uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return x; };
    const int T = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)pool().size() + 1, n / 65536 + 1));
    std::vector<uint64_t> part((size_t)T, 0);
    std::function<void(int)> work = [&](int t) {
        const uint64_t a = n * (uint64_t)t / (uint64_t)T, b = n * (uint64_t)(t + 1) / (uint64_t)T;
        uint64_t acc = 0;
        for (uint64_t i = a; i < b; i++) {
            const kmcpg_hit &h = hits[i];
            acc += mix((first_index + i) * 0x9E3779B97F4A7C15ull ^ ((uint64_t)h.query << 32 | h.target)) ^ mix(((uint64_t)h.count << 32 | h.target) + 0x632BE59BD9B4E019ull * (first_index + i + 1));
        }
        part[t] = acc;
    };
    pool().parallel(T, work);
    uint64_t d = 0;
    for (uint64_t v : part) d += v;
    return d;
}

// test hook (kmcp-gpu search --dry-run, host only): the result of a batch in which nothing matched, so that the command's
// reader → engine → writer plumbing can run without a device
int kmcpg_internal_unmatched_results(const uint64_t *off, uint32_t n_seqs, int paired, int k, kmcpg_results *out) {
    if (!out || (n_seqs && !off) || (paired && (n_seqs & 1))) return KMCPG_EINVAL;
    const uint32_t step = paired ? 2 : 1, nq = n_seqs / step;
    *out = alloc_results(nq, 0);
    for (uint32_t q = 0; q < nq; q++) {
        out->query_len[q] = (int32_t)(off[(size_t)q * step + step] - off[(size_t)q * step]);
        out->n_kmers[q] = 0; out->k_used[q] = k;
        out->match_off[q + 1] = 0;
    }
    return KMCPG_OK;
}

// test hook (tests/test_abi.py, no GPU needed): the k-way merge the sharded engine applies to the per-shard hit lists
void kmcpg_internal_merge_hits(const kmcpg_hit *const *lists, const uint64_t *n, int k, kmcpg_hit *out, uint32_t first_query, uint32_t n_queries, int threads) {
    if (n_queries == 0) {                 // range not given: take it from the lists
        uint32_t lo = ~0u, hi = 0;
        for (int s = 0; s < k; s++) if (n[s]) { lo = std::min(lo, lists[s][0].query); hi = std::max(hi, lists[s][n[s] - 1].query); }
        if (lo == ~0u) return;
        first_query = lo; n_queries = hi - lo + 1;
    }
    merge_shard_hits_mt(lists, n, k, out, first_query, n_queries, std::max(1, threads));
}

// test hook (host only): the threaded part-by-part merger of the sharded engine driven by stand-in shards that deliver
// seeded hit lists with random delays; fail_shard >= 0 makes that shard return KMCPG_ECUDA before part fail_part.
// Returns 0 when every merged part equals the expected union (or, with a failure injected, the propagated error code).
int kmcpg_internal_sharded_selftest(int n_shards, int n_parts, int fail_shard, int fail_part, uint64_t seed) {
    if (n_shards < 1 || n_shards > 64 || n_parts < 0) return KMCPG_EINVAL;
    const uint32_t QP = 257, NT = 97;                       // queries per part, targets
    auto rnd = [](uint64_t &x) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
    uint64_t st = seed * 2654435761ull + 88172645463325252ull;
    std::vector<std::vector<kmcpg_hit>> expect((size_t)n_parts);
    std::vector<std::vector<std::vector<kmcpg_hit>>> per((size_t)n_shards, std::vector<std::vector<kmcpg_hit>>((size_t)n_parts));
    std::vector<std::vector<int32_t>> nk((size_t)n_parts), ql((size_t)n_parts);
    for (int pi = 0; pi < n_parts; pi++) {
        for (uint32_t q = 0; q < QP; q++) {
            nk[pi].push_back((int32_t)(rnd(st) % 130)); ql[pi].push_back(150);
            for (uint32_t t = 0; t < NT; t++)
                if (rnd(st) % 5 == 0) {
                    kmcpg_hit h{(uint32_t)pi * QP + q, t, (uint32_t)(rnd(st) % 130 + 1)};
                    expect[pi].push_back(h);
                    per[(t * 7 + 3) % (uint32_t)n_shards][pi].push_back(h);
                }
        }
    }
    ShardSearch fake = [&](int s, kmcpg_part_cb cb, void *user, kmcpg_hits *summary) {
        uint64_t ls = seed + 977 * (uint64_t)(s + 1);
        for (int pi = 0; pi < n_parts; pi++) {
            if (s == fail_shard && pi == fail_part) return (int)KMCPG_ECUDA;
            std::this_thread::sleep_for(std::chrono::microseconds(rnd(ls) % 300));
            std::vector<kmcpg_hit> tmp = per[s][pi];         // the executor's arrays are only valid during the callback
            kmcpg_part pt;
            pt.first_query = (uint32_t)pi * QP; pt.n_queries = QP;
            pt.n_kmers = nk[pi].data(); pt.query_len = ql[pi].data();
            pt.hits = tmp.data(); pt.n_hits = tmp.size();
            cb(user, &pt);
            std::fill(tmp.begin(), tmp.end(), kmcpg_hit{0, 0, 0});
        }
        if (s == fail_shard && fail_part >= n_parts) return (int)KMCPG_ECUDA;
        summary->ms_total = 1.0f; summary->kernel_launches = 3; summary->probe_row_bytes = 10;
        return (int)KMCPG_OK;
    };
    int seen = 0;
    bool same = true;
    std::function<void(const kmcpg_part &)> absorb = [&](const kmcpg_part &pt) {
        if (seen >= n_parts) { same = false; return; }
        const auto &e = expect[seen];
        if (pt.first_query != (uint32_t)seen * QP || pt.n_queries != QP || pt.n_hits != e.size()) same = false;
        else if (!e.empty() && memcmp(pt.hits, e.data(), e.size() * sizeof(kmcpg_hit))) same = false;
        else if (memcmp(pt.n_kmers, nk[seen].data(), QP * 4) || memcmp(pt.query_len, ql[seen].data(), QP * 4)) same = false;
        seen++;
    };
    kmcpg_results out;
    memset(&out, 0, sizeof(out));
    int rc = sharded_round(n_shards, fake, absorb, &out, 1 + (int)(seed % 4));
    if (rc) return rc;
    if (!same || seen != n_parts) return KMCPG_EINVAL;
    if (out.kernel_launches != 3u * (uint32_t)n_shards || out.probe_row_bytes != 10ull * (uint64_t)n_shards) return KMCPG_EINVAL;
    return KMCPG_OK;
}

void kmcpg_free_results(kmcpg_results *r) {
    if (!r) return;
    delete (ResPriv *)r->_priv;
    memset(r, 0, sizeof(*r));
}

}  // extern "C"

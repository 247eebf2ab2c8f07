// engine.cpp — host side of the search engine above the device path: the C++ mirror of what
// UnikIndexDBSearchEngine / UnikIndexDB.handleQuery do with the per-block match counts
// (reference kmcp/cmd/util-db-search.go: multi-k loop U:763-1025, --try-se U:808-842 + 995-1011,
//  Match fields and filters U:7466-7491, sort U:273-282 with Less functions U:105-145, top-N scores U:285-311).
// The counts themselves always come from the GPU (kmcpg_search_batch); nothing here probes an index.
//
// A "round" = one device search of the still-undecided queries with one k and one mate selection; its hit
// list (sorted by query) is filtered and sorted per query by a few host threads into one flat array.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

#include "common.h"

extern "C" const double *kmcpg_internal_target_sizes(const kmcpg_ctx *ctx);

namespace {

using namespace kmcpg;

// QueryFPRWithCacheWithConstantFPR (F:140-193) without its slot-sharing quirk: plain memo of the pure function
struct FprCache {
    static constexpr int N = 1024;
    double p = -1;
    std::vector<double> tab;
    void reset(double fpr) {
        if (fpr == p && !tab.empty()) return;
        p = fpr;
        tab.assign((size_t)(N + 1) * (N + 1), -1.0);
    }
    double get(int n, int c) {
        if (n > N || c > n) return query_fpr(n, c, p);
        double &slot = tab[(size_t)n * (N + 1) + c];
        double v = slot;
        if (v < 0) { v = query_fpr(n, c, p); slot = v; }   // idempotent value: a racing duplicate store is harmless
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

struct ResPriv {
    std::vector<int32_t> query_len, n_kmers, k_used;
    std::vector<uint64_t> match_off;
    std::vector<kmcpg_match> matches;
};

// matches of one round, flat, in local-query order
struct Round {
    std::vector<uint32_t> queries;        // local index → global query
    std::vector<uint64_t> off;            // local index → range in `matches`
    std::vector<kmcpg_match> matches;
};

// filters + sorts the hits of local queries [lo, hi) (U:7466-7491, U:273-311)
void post_filter(const kmcpg_engine_opts *o, const kmcpg_hits &hits, const std::vector<uint64_t> &hoff, const std::vector<uint32_t> &cur,
                 const double *tsize, FprCache *cache, uint32_t lo, uint32_t hi, std::vector<kmcpg_match> &out, std::vector<uint32_t> &count) {
    const Less less{o->sort_by};
    for (uint32_t l = lo; l < hi; l++) {
        const int n = hits.n_kmers[l];
        count[l] = 0;
        if (n == 0 || hoff[l] == hoff[l + 1]) continue;
        const uint32_t q = cur[l];
        const double nh = (double)n;
        const size_t start = out.size();
        for (uint64_t i = hoff[l]; i < hoff[l + 1]; i++) {
            const kmcpg_hit &h = hits.hits[i];
            const double c = (double)h.count, sz = tsize[h.target];
            const double tcov = c / sz;
            if (!(tcov >= o->min_target_cov)) continue;              // U:7473-7474
            const double fpr = cache->get(n, (int)h.count);
            if (!(fpr <= o->max_fpr)) continue;                      // U:7477-7478
            kmcpg_match m;
            m.query = q; m.target = h.target; m.count = h.count; m._pad = 0;
            m.fpr = fpr; m.qcov = c / nh; m.tcov = tcov; m.jacc = c / (nh + sz - c);   // U:7470-7488
            out.push_back(m);
        }
        size_t cnt = out.size() - start;
        if (cnt > 1 && !o->do_not_sort) std::sort(out.begin() + start, out.end(), less);   // U:273-282
        if (cnt > 0 && o->top_n_scores > 0 && !o->do_not_sort) {     // U:285-311 (kept verbatim, including [:i+1])
            int nsc = 0; double pscore = 1024; size_t i = 0; bool broke = false;
            for (i = 0; i < cnt; i++) {
                const kmcpg_match &m = out[start + i];
                double score = o->sort_by == 0 ? m.qcov : (o->sort_by == 1 ? m.tcov : m.jacc);
                if (score < pscore) { nsc++; if (nsc > o->top_n_scores) { broke = true; break; } pscore = score; }
            }
            if (broke) { out.resize(start + i + 1); cnt = i + 1; }
        }
        count[l] = (uint32_t)cnt;
    }
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
    kmcpg_db_info_t info;
    int rc = kmcpg_db_info(ctx, &info);
    if (rc) return rc;
    memset(out, 0, sizeof(*out));
    const uint32_t step = o->paired ? 2 : 1;
    const uint32_t nq = n_seqs / step;
    static thread_local FprCache tl_cache;
    tl_cache.reset(info.fpr);
    FprCache *cache = &tl_cache;          // worker threads must share THIS instance, not their own thread_local
    const double *tsize = kmcpg_internal_target_sizes(ctx);

    ResPriv *priv = new ResPriv();
    priv->query_len.assign(nq, 0); priv->n_kmers.assign(nq, 0); priv->k_used.assign(nq, info.ks[0]);
    std::vector<uint32_t> pending(nq);
    for (uint32_t q = 0; q < nq; q++) pending[q] = q;
    const int tries_max = (o->try_se && o->paired) ? 3 : 1;
    int threads = o->threads > 0 ? o->threads : (int)std::thread::hardware_concurrency();
    threads = std::max(1, std::min(threads, 32));

    std::vector<Round> rounds;
    // where the final matches of query q live: (round, local index); round -1 = unmatched
    std::vector<int32_t> q_round(nq, -1);
    std::vector<uint32_t> q_local(nq, 0);
    std::vector<uint8_t> sub_seq;
    std::vector<uint64_t> sub_off;
    for (int ik = 0; ik < info.n_ks && !pending.empty(); ik++) {
        const int k = info.ks[ik];
        std::vector<uint32_t> next_k;                 // queries that found nothing with this k
        std::vector<uint32_t> cur = pending;
        for (int tries = 0; tries < tries_max && !cur.empty(); tries++) {
            // pack the pending subset (the first pass uses the caller's buffers untouched)
            const uint8_t *bs = seq; const uint64_t *bo = off; uint32_t bn = n_seqs;
            if (cur.size() != nq) {
                sub_off.assign(1, 0); sub_seq.clear();
                for (uint32_t q : cur)
                    for (uint32_t m = 0; m < step; m++) {
                        uint64_t a = off[q * step + m], b = off[q * step + m + 1];
                        sub_seq.insert(sub_seq.end(), seq + a, seq + b);
                        sub_off.push_back(sub_seq.size());
                    }
                if (sub_seq.empty()) sub_seq.push_back(0);
                bs = sub_seq.data(); bo = sub_off.data(); bn = (uint32_t)cur.size() * step;
            }
            kmcpg_search_params p;
            kmcpg_default_params(&p);
            p.min_query_len = o->min_query_len; p.min_matched = o->min_matched; p.dedup_threshold = o->dedup_threshold;
            p.paired = o->paired; p.min_query_cov = o->min_query_cov; p.k = k; p.mate_select = tries;
            kmcpg_hits hits;
            rc = kmcpg_search_batch(ctx, &p, bs, bo, bn, &hits);
            if (rc) { delete priv; return rc; }
            out->ms_gpu_total += hits.ms_total; out->probe_row_bytes += hits.probe_row_bytes; out->kernel_launches += hits.kernel_launches;

            const uint32_t ln = (uint32_t)cur.size();
            // hit ranges per local query (hits are sorted by query)
            std::vector<uint64_t> hoff((size_t)ln + 1, 0);
            for (uint64_t i = 0; i < hits.n_hits; i++) hoff[hits.hits[i].query + 1]++;
            for (uint32_t i = 0; i < ln; i++) hoff[i + 1] += hoff[i];

            rounds.emplace_back();
            Round &R = rounds.back();
            const int ridx = (int)rounds.size() - 1;
            std::vector<uint32_t> count(ln, 0);
            int T = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)threads, (hits.n_hits + ln / 8) / 16384 + 1));
            std::vector<std::vector<kmcpg_match>> parts(T);
            std::vector<uint32_t> bounds(T + 1, ln);
            bounds[0] = 0;
            for (int t = 1; t < T; t++) {            // split by hits so the threads get equal work
                uint64_t want = hits.n_hits * (uint64_t)t / T;
                bounds[t] = (uint32_t)(std::lower_bound(hoff.begin(), hoff.end(), want) - hoff.begin());
                if (bounds[t] > ln) bounds[t] = ln;
                if (bounds[t] < bounds[t - 1]) bounds[t] = bounds[t - 1];
            }
            if (T == 1) {
                post_filter(o, hits, hoff, cur, tsize, cache, 0, ln, parts[0], count);
            } else {
                std::vector<std::thread> th;
                for (int t = 0; t < T; t++)
                    th.emplace_back([&, t] {
                        parts[t].reserve((size_t)(hoff[bounds[t + 1]] - hoff[bounds[t]]));
                        post_filter(o, hits, hoff, cur, tsize, cache, bounds[t], bounds[t + 1], parts[t], count);
                    });
                for (auto &t : th) t.join();
            }
            size_t total = 0;
            for (auto &v : parts) total += v.size();
            if (T == 1) R.matches.swap(parts[0]);
            else {
                R.matches.resize(total);
                size_t w = 0;
                for (auto &v : parts) { if (!v.empty()) memcpy(R.matches.data() + w, v.data(), v.size() * sizeof(kmcpg_match)); w += v.size(); }
            }
            R.off.resize((size_t)ln + 1);
            R.off[0] = 0;
            std::vector<uint32_t> retry;
            for (uint32_t l = 0; l < ln; l++) {
                const uint32_t q = cur[l];
                const int n = hits.n_kmers[l];
                R.off[l + 1] = R.off[l] + count[l];
                priv->query_len[q] = hits.query_len[l];
                priv->k_used[q] = k;
                if (n == 0) { if (tries == 0) priv->n_kmers[q] = 0; continue; }     // U:778-786, U:854-869: final, unmatched
                priv->n_kmers[q] = n;
                if (count[l]) { q_round[q] = ridx; q_local[q] = l; }
                else retry.push_back(q);
            }
            R.queries.swap(cur);
            kmcpg_free_hits(&hits);
            if (tries + 1 < tries_max) cur.swap(retry);          // --try-se: read1 only, then read2 only
            else { next_k.insert(next_k.end(), retry.begin(), retry.end()); cur.clear(); }
        }
        std::sort(next_k.begin(), next_k.end());
        pending.swap(next_k);                                    // U:1018-1023: try the next smaller k
    }

    priv->match_off.assign((size_t)nq + 1, 0);
    if (rounds.size() == 1 && rounds[0].queries.size() == nq) {
        // the common case: one round over all queries — its flat array already is the answer
        priv->matches.swap(rounds[0].matches);
        for (uint32_t q = 0; q < nq; q++) priv->match_off[q + 1] = rounds[0].off[q + 1];
    } else {
        for (uint32_t q = 0; q < nq; q++) {
            uint64_t c = 0;
            if (q_round[q] >= 0) { const Round &R = rounds[q_round[q]]; c = R.off[q_local[q] + 1] - R.off[q_local[q]]; }
            priv->match_off[q + 1] = priv->match_off[q] + c;
        }
        priv->matches.resize(priv->match_off[nq]);
        for (uint32_t q = 0; q < nq; q++)
            if (q_round[q] >= 0) {
                const Round &R = rounds[q_round[q]];
                memcpy(priv->matches.data() + priv->match_off[q], R.matches.data() + R.off[q_local[q]],
                       (priv->match_off[q + 1] - priv->match_off[q]) * sizeof(kmcpg_match));
            }
    }
    out->n_queries = nq; out->n_matches = priv->matches.size();
    out->query_len = priv->query_len.data(); out->n_kmers = priv->n_kmers.data(); out->k_used = priv->k_used.data();
    out->match_off = priv->match_off.data(); out->matches = priv->matches.data();
    out->_priv = priv;
    return KMCPG_OK;
}

void kmcpg_free_results(kmcpg_results *r) {
    if (!r) return;
    delete (ResPriv *)r->_priv;
    memset(r, 0, sizeof(*r));
}

}  // extern "C"

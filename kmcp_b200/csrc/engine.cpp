// engine.cpp — host side of the search engine above the device path: the C++ mirror of what
// UnikIndexDBSearchEngine / UnikIndexDB.handleQuery do with the per-block match counts
// (reference kmcp/cmd/util-db-search.go: multi-k loop U:763-1025, --try-se U:808-842 + 995-1011,
//  Match fields and filters U:7466-7491, sort U:273-282 with Less functions U:105-145, top-N scores U:285-311).
// The counts themselves always come from the GPU (kmcpg_search_batch); nothing here probes an index.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

#include "common.h"

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

    // per-target Sizes (k-mers of the target) once
    std::vector<double> tsize((size_t)info.n_targets);
    for (int64_t t = 0; t < info.n_targets; t++) {
        kmcpg_target_t tt;
        kmcpg_target(ctx, t, &tt);
        tsize[(size_t)t] = (double)tt.n_kmers;
    }

    ResPriv *priv = new ResPriv();
    priv->query_len.assign(nq, 0); priv->n_kmers.assign(nq, 0); priv->k_used.assign(nq, info.ks[0]);
    std::vector<std::vector<kmcpg_match>> per(nq);
    std::vector<uint32_t> pending(nq);
    for (uint32_t q = 0; q < nq; q++) pending[q] = q;
    const int tries_max = (o->try_se && o->paired) ? 3 : 1;
    int threads = o->threads > 0 ? o->threads : (int)std::thread::hardware_concurrency();
    if (threads < 1) threads = 1;

    std::vector<uint8_t> sub_seq;
    std::vector<uint64_t> sub_off;
    for (int ik = 0; ik < info.n_ks && !pending.empty(); ik++) {
        const int k = info.ks[ik];
        std::vector<uint32_t> next_k;                 // queries that found nothing with this k
        std::vector<uint32_t> cur = pending;
        for (int tries = 0; tries < tries_max && !cur.empty(); tries++) {
            // pack the pending subset (the first pass uses the caller's buffers untouched)
            const uint8_t *bs = seq; const uint64_t *bo = off; uint32_t bn = n_seqs;
            const bool all = cur.size() == nq;
            if (!all) {
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

            // hit ranges per local query (hits are sorted by query)
            const uint32_t ln = (uint32_t)cur.size();
            std::vector<uint64_t> hoff(ln + 1, 0);
            for (uint64_t i = 0; i < hits.n_hits; i++) hoff[hits.hits[i].query + 1]++;
            for (uint32_t i = 0; i < ln; i++) hoff[i + 1] += hoff[i];

            std::vector<uint8_t> found(ln, 0), gave_up(ln, 0);
            auto work = [&](uint32_t lo, uint32_t hi) {
                for (uint32_t l = lo; l < hi; l++) {
                    const uint32_t q = cur[l];
                    const int n = hits.n_kmers[l];
                    priv->query_len[q] = hits.query_len[l];
                    priv->k_used[q] = k;
                    if (n == 0) { gave_up[l] = 1; if (tries == 0) priv->n_kmers[q] = 0; continue; }   // U:778-786, U:854-869: final
                    priv->n_kmers[q] = n;
                    const double nh = (double)n;
                    std::vector<kmcpg_match> &ms = per[q];
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
                        ms.push_back(m);
                    }
                    if (ms.empty()) continue;
                    found[l] = 1;
                    if (ms.size() > 1 && !o->do_not_sort) std::sort(ms.begin(), ms.end(), Less{o->sort_by});   // U:273-282
                    if (o->top_n_scores > 0 && !o->do_not_sort) {                // U:285-311 (kept verbatim, including [:i+1])
                        int nsc = 0; double pscore = 1024; size_t i = 0; bool broke = false;
                        for (i = 0; i < ms.size(); i++) {
                            double score = o->sort_by == 0 ? ms[i].qcov : (o->sort_by == 1 ? ms[i].tcov : ms[i].jacc);
                            if (score < pscore) { nsc++; if (nsc > o->top_n_scores) { broke = true; break; } pscore = score; }
                        }
                        if (broke) ms.resize(i + 1);
                    }
                }
            };
            if (threads > 1 && ln > 4096) {
                std::vector<std::thread> th;
                uint32_t per_t = (ln + threads - 1) / threads;
                for (int t = 0; t < threads; t++) {
                    uint32_t lo = std::min<uint32_t>(ln, t * per_t), hi = std::min<uint32_t>(ln, lo + per_t);
                    if (lo < hi) th.emplace_back(work, lo, hi);
                }
                for (auto &t : th) t.join();
            } else {
                work(0, ln);
            }
            kmcpg_free_hits(&hits);
            std::vector<uint32_t> retry;
            for (uint32_t l = 0; l < ln; l++)
                if (!found[l] && !gave_up[l]) retry.push_back(cur[l]);
            if (tries + 1 < tries_max) cur.swap(retry);          // --try-se: read1 only, then read2 only
            else { next_k.insert(next_k.end(), retry.begin(), retry.end()); cur.clear(); }
        }
        std::sort(next_k.begin(), next_k.end());
        pending.swap(next_k);                                    // U:1018-1023: try the next smaller k
    }

    priv->match_off.assign(nq + 1, 0);
    for (uint32_t q = 0; q < nq; q++) priv->match_off[q + 1] = priv->match_off[q] + per[q].size();
    priv->matches.resize(priv->match_off[nq]);
    for (uint32_t q = 0; q < nq; q++)
        if (!per[q].empty()) memcpy(priv->matches.data() + priv->match_off[q], per[q].data(), per[q].size() * sizeof(kmcpg_match));
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

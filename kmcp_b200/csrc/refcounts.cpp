// `kmcp profile` stage 1/4 counters accumulated from engine results.
// Reference: kmcp/cmd/profile.go:761-990 (the per-query state machine and the flush into Target.Match /
// UniqMatch / UniqMatchHic) and util-profile.go:94-182 (parseMatchResult: rows with qCov < -t or FPR > -f are
// dropped before the loop; the values it sees are the %.4f / %.4e texts of the search TSV, S:536-539).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include <algorithm>

#include "common.h"

struct kmcpg_refcounts {
    kmcpg_ctx *ctx = nullptr;
    kmcpg_refcount_params p;
    struct Ref { std::string name; uint64_t gsize = 0; uint32_t n_chunks = 0; bool seen = false; std::vector<double> match, uniq, hic; };
    std::vector<Ref> refs;
    std::vector<uint32_t> ref_of_target, chunk_of_target, chunks_of_target;
    std::vector<uint64_t> gsize_of_target;
    uint64_t n_reads = 0;
    std::vector<kmcpg_refcount_row> rows;
};

namespace {
// the value `profile` parses back from the search TSV
inline double as_text(double v, const char *fmt) {
    char b[64];
    snprintf(b, sizeof(b), fmt, v);
    return strtod(b, nullptr);
}
}  // namespace

extern "C" {

void kmcpg_default_refcount_params(kmcpg_refcount_params *p) {
    if (!p) return;
    p->min_query_cov = 0.55; p->max_fpr = 0.01; p->top_n_scores = 0; p->keep_perfect = 0; p->keep_main = 0;
    p->max_qcov_gap = 0.4; p->hic_min_qcov = 0.75;
}

int kmcpg_refcounts_create(kmcpg_ctx *ctx, const char *db_dir, const kmcpg_refcount_params *p, kmcpg_refcounts **out) {
    if ((!ctx && !db_dir) || !out) return KMCPG_EINVAL;
    auto *r = new kmcpg_refcounts();
    r->ctx = ctx;
    if (p) r->p = *p; else kmcpg_default_refcount_params(&r->p);
    std::unordered_map<std::string, uint32_t> ids;
    auto add_target = [&](const char *name, uint32_t index, uint64_t gsize) {
        auto it = ids.find(name);
        if (it == ids.end()) {
            it = ids.emplace(name, (uint32_t)r->refs.size()).first;
            r->refs.emplace_back();
            r->refs.back().name = name;
        }
        r->ref_of_target.push_back(it->second);
        r->chunk_of_target.push_back(index & 0xFFFFu);          // S:532-533
        r->chunks_of_target.push_back(index >> 16);
        r->gsize_of_target.push_back(gsize);
    };
    if (ctx) {                      // the database open in ctx
        kmcpg_db_info_t info;
        if (int rc = kmcpg_db_info(ctx, &info)) { delete r; return rc; }
        for (int64_t t = 0; t < info.n_targets; t++) {
            kmcpg_target_t tg;
            if (int rc = kmcpg_target(ctx, t, &tg)) { delete r; return rc; }
            add_target(tg.name, tg.index, tg.genome_size);
        }
    } else {                        // block headers only (host; for results produced elsewhere)
        kmcpg::DbMeta m;
        std::string err;
        if (int rc = kmcpg::read_db_meta(db_dir, m, err)) { delete r; return rc; }
        for (auto &bm : m.blocks)
            for (int c = 0; c < bm.n_names; c++) add_target(bm.names[c].c_str(), bm.indices[c], bm.gsizes[c]);
    }
    *out = r;
    return KMCPG_OK;
}

int kmcpg_refcounts_add(kmcpg_refcounts *rc, const kmcpg_results *r) {
    if (!rc || !r) return KMCPG_EINVAL;
    const kmcpg_refcount_params &p = rc->p;
    const uint64_t nm = r->n_matches;
    // 1. the text round trip of qCov / FPR and parseMatchResult's two filters, on several threads
    std::vector<double> q4(nm);
    std::vector<uint8_t> pass(nm);
    {
        unsigned hw = std::thread::hardware_concurrency();
        const unsigned nt = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>({(uint64_t)(hw ? hw : 4), (uint64_t)32, nm / 4096 + 1}));
        auto work = [&](unsigned t) {
            const uint64_t lo = nm * t / nt, hi = nm * (t + 1) / nt;
            for (uint64_t i = lo; i < hi; i++) {
                const kmcpg_match &m = r->matches[i];
                const double q = as_text(m.qcov, "%.4f");
                q4[i] = q;
                pass[i] = !(q < p.min_query_cov) && !(as_text(m.fpr, "%.4e") > p.max_fpr);
            }
        };
        std::vector<std::thread> th;
        for (unsigned t = 1; t < nt; t++) th.emplace_back(work, t);
        work(0);
        for (auto &t : th) t.join();
    }
    // 2. the per-query state machine and the flush, in input order (Match sums are order-dependent doubles)
    struct Kept { uint32_t ref, target; double q; };
    std::vector<Kept> kept;
    std::vector<uint32_t> qrefs, qcount;
    for (uint32_t q = 0; q < r->n_queries; q++) {
        const uint64_t a = r->match_off[q], b = r->match_off[q + 1];
        if (a == b) continue;
        kept.clear();
        double p_score = 1024;
        int n_score = 0;
        bool process = true, first_row = true;
        for (uint64_t i = a; i < b; i++) {
            if (!pass[i]) continue;
            const double qc = q4[i];
            if (first_row) first_row = false;                                    // "new query": state was just reset
            else if (p.keep_perfect) {
                if (!process) continue;
                if (p_score == 1 && qc < 1) { process = false; continue; }
            } else if (p.keep_main && p_score <= 1) {
                if (!process) continue;
                if (p_score - qc > p.max_qcov_gap) { process = false; continue; }
            }
            if (p.top_n_scores > 0) {
                if (!process) continue;
                if (qc < p_score) {
                    if (++n_score > p.top_n_scores) { process = false; continue; }
                }
            }
            const uint32_t t = r->matches[i].target;
            if (t >= rc->ref_of_target.size()) return KMCPG_EINVAL;
            kept.push_back({rc->ref_of_target[t], t, qc});
            p_score = qc;
        }
        if (kept.empty()) continue;
        rc->n_reads++;
        qrefs.clear(); qcount.clear();
        for (auto &k : kept) {
            size_t j = 0;
            while (j < qrefs.size() && qrefs[j] != k.ref) j++;
            if (j == qrefs.size()) { qrefs.push_back(k.ref); qcount.push_back(0); }
            qcount[j]++;
        }
        for (size_t j = 0; j < qrefs.size(); j++) {
            kmcpg_refcounts::Ref &ref = rc->refs[qrefs[j]];
            const double share = 1.0 / (double)qcount[j];
            bool first = true;
            for (auto &k : kept) {
                if (k.ref != qrefs[j]) continue;
                if (!ref.seen) {                                                 // Target created from the first row seen
                    ref.seen = true;
                    ref.gsize = rc->gsize_of_target[k.target];
                    ref.n_chunks = rc->chunks_of_target[k.target];
                    ref.match.assign(ref.n_chunks, 0.0); ref.uniq.assign(ref.n_chunks, 0.0); ref.hic.assign(ref.n_chunks, 0.0);
                }
                const uint32_t c = rc->chunk_of_target[k.target];
                if (c >= ref.n_chunks) return KMCPG_EFORMAT;                      // the reference would index out of range here
                if (first) {
                    if (qrefs.size() == 1) {
                        ref.uniq[c] += 1;
                        if (k.q >= p.hic_min_qcov) ref.hic[c] += 1;
                    }
                    first = false;
                }
                ref.match[c] += share;
            }
        }
    }
    return KMCPG_OK;
}

int kmcpg_refcounts_get(kmcpg_refcounts *rc, kmcpg_refcount_table *out) {
    if (!rc || !out) return KMCPG_EINVAL;
    rc->rows.clear();
    for (auto &ref : rc->refs) {
        if (!ref.seen) continue;
        kmcpg_refcount_row row;
        row.name = ref.name.c_str(); row.genome_size = ref.gsize; row.n_chunks = ref.n_chunks; row._pad = 0;
        row.match = ref.match.data(); row.uniq_match = ref.uniq.data(); row.uniq_match_hic = ref.hic.data();
        rc->rows.push_back(row);
    }
    out->n_reads = rc->n_reads; out->n_refs = (uint32_t)rc->rows.size(); out->_pad = 0; out->rows = rc->rows.data();
    return KMCPG_OK;
}

void kmcpg_refcounts_free(kmcpg_refcounts *rc) { delete rc; }

}  // extern "C"

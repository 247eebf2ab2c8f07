// reader.cpp — the C ABI of the reader stage (kmcpg_reader_*, include/kmcp_gpu.h) over fastx_reader.h.
// Reference: the reader loop of kmcp/cmd/search.go S:793-1000.  Host only: no CUDA call in here.
#include "fastx_reader.h"

#include "../../include/kmcp_gpu.h"

namespace {
// used batches go round: their arrays (tens of MB) are touched once instead of once per batch
struct BatchPool {
    std::mutex mu;
    std::vector<fastx::Batch *> free;
    ~BatchPool() { for (auto *b : free) delete b; }
    fastx::Batch *get() {
        std::lock_guard<std::mutex> lk(mu);
        if (free.empty()) return new fastx::Batch();
        fastx::Batch *b = free.back();
        free.pop_back();
        return b;
    }
    void put(fastx::Batch *b) {
        b->clear();
        std::lock_guard<std::mutex> lk(mu);
        if (free.size() < 6) free.push_back(b); else delete b;
    }
};
struct BatchRef { fastx::Batch *b; std::shared_ptr<BatchPool> home; };     // what a kmcpg_read_batch points to
}  // namespace

#include "common.h"

namespace kmcpg {
struct FastxFile { fastx::Reader r; std::vector<char> seq; std::string id; };
FastxFile *fastx_open(const std::string &path) {
    FastxFile *f = new FastxFile();
    f->r.keep_header = true;
    fastx::Tuning t;
    t.inflate_threads = 1;              // the index builder loads many files side by side: one decoder thread per file
    if (!f->r.open(path, false, t)) { delete f; return nullptr; }
    return f;
}
int fastx_next(FastxFile *f, std::string &header, std::string &seq, std::string &err) {
    try {
        f->seq.clear();
        if (!f->r.next(f->id, f->seq)) return 0;
        header = f->r.header;
        seq.assign(f->seq.data(), f->seq.size());
        return 1;
    } catch (const std::exception &e) { err = e.what(); return -1; }
}
void fastx_close(FastxFile *f) { delete f; }
}  // namespace kmcpg

struct kmcpg_reader {
    fastx::ReaderConfig cfg;
    std::shared_ptr<BatchPool> pool = std::make_shared<BatchPool>();
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::deque<fastx::Batch *> q;       // finished batches, at most two ahead of the caller
    bool done = false, stop = false, failed = false;
    std::string err;
    uint64_t next_base = 0;
};

namespace {
struct Stopped {};                       // the caller closed the reader while batches were still being built
}

extern "C" {

void kmcpg_default_reader_opts(kmcpg_reader_opts *o) {
    if (!o) return;
    memset(o, 0, sizeof(*o));
    o->k = 21;
}

int kmcpg_reader_open(const kmcpg_reader_opts *o, kmcpg_reader **out) {
    if (!o || !out) return KMCPG_EINVAL;
    *out = nullptr;
    const bool paired = o->read1 && *o->read1 && o->read2 && *o->read2;
    if (!paired && ((o->read1 && *o->read1) || (o->read2 && *o->read2))) return KMCPG_EINVAL;      // S:383-389: both or neither
    if (!paired && (o->n_files < 1 || !o->files)) return KMCPG_EINVAL;
    if (o->k < 1) return KMCPG_EINVAL;
    kmcpg_reader *r = new kmcpg_reader();
    fastx::ReaderConfig &c = r->cfg;
    c.paired = paired;
    if (paired) { c.read1 = o->read1; c.read2 = o->read2; }
    else for (int i = 0; i < o->n_files; i++) { if (!o->files[i]) { delete r; return KMCPG_EINVAL; } c.files.push_back(o->files[i]); }
    c.whole_file = o->whole_file != 0; c.use_filename = o->use_filename != 0;
    if (o->query_id) c.query_id = o->query_id;
    c.kmax = o->k;
    if (o->batch_reads) c.batch_reads = o->batch_reads;
    if (o->batch_bytes) c.batch_bytes = (size_t)o->batch_bytes;
    c.tune.inflate_threads = o->inflate_threads; c.tune.parse_threads = o->parse_threads;
    if (o->inflate_chunk) c.tune.inflate_chunk = (size_t)o->inflate_chunk;
    if (o->inflate_cap) c.tune.inflate_cap = (size_t)o->inflate_cap;
    if (o->parse_piece) c.tune.parse_piece = (size_t)o->parse_piece;
    { std::shared_ptr<BatchPool> pool = r->pool; c.new_batch = [pool] { return pool->get(); }; }
    if (o->log) { auto fn = o->log; void *user = o->log_user; c.log = [fn, user](const char *level, const char *msg) { fn(user, level, msg); }; }
    r->th = std::thread([r] {
        try {
            fastx::read_batches(r->cfg, [r](fastx::Batch *b) {
                std::unique_lock<std::mutex> lk(r->mu);
                r->cv.wait(lk, [&] { return r->stop || r->q.size() < 2; });
                if (r->stop) { delete b; throw Stopped(); }
                b->base = r->next_base;
                r->next_base += b->n_ids();
                r->q.push_back(b);
                r->cv.notify_all();
            });
        } catch (const Stopped &) {
        } catch (const std::exception &e) {
            std::lock_guard<std::mutex> lk(r->mu);
            r->err = e.what(); r->failed = true;
        }
        std::lock_guard<std::mutex> lk(r->mu);
        r->done = true;
        r->cv.notify_all();
    });
    *out = r;
    return KMCPG_OK;
}

int kmcpg_reader_next(kmcpg_reader *r, kmcpg_read_batch *out) {
    if (!r || !out) return KMCPG_EINVAL;
    memset(out, 0, sizeof(*out));
    fastx::Batch *b = nullptr;
    {
        std::unique_lock<std::mutex> lk(r->mu);
        r->cv.wait(lk, [&] { return !r->q.empty() || r->done; });
        if (r->q.empty()) return r->failed ? KMCPG_EIO : 0;      // the batches in front of an error are handed out first
        b = r->q.front(); r->q.pop_front();
        r->cv.notify_all();
    }
    out->n_queries = (uint32_t)b->n_ids();
    out->n_seqs = (uint32_t)(b->off.size() - 1);
    out->seq = b->seq.data(); out->off = b->off.data();
    out->ids = b->id_buf.data(); out->id_off = b->id_off.data();
    out->first_query = b->base;
    out->_priv = new BatchRef{b, r->pool};
    return 1;
}

void kmcpg_reader_free_batch(kmcpg_read_batch *b) {
    if (!b || !b->_priv) return;
    BatchRef *ref = (BatchRef *)b->_priv;
    ref->home->put(ref->b);                  // the pool outlives the reader for batches released after kmcpg_reader_close
    delete ref;
    memset(b, 0, sizeof(*b));
}

// test hook: pieces parsed by the workers of the parallel parser / files handed back to the general reader, process-wide
void kmcpg_internal_reader_stats(uint64_t *pieces, uint64_t *fallbacks) {
    if (pieces) *pieces = fastx::g_stat_pieces.load();
    if (fallbacks) *fallbacks = fastx::g_stat_fallbacks.load();
}

const char *kmcpg_reader_error(const kmcpg_reader *r) { return r ? r->err.c_str() : ""; }

int kmcpg_reader_close(kmcpg_reader *r) {
    if (!r) return KMCPG_EINVAL;
    { std::lock_guard<std::mutex> lk(r->mu); r->stop = true; }
    r->cv.notify_all();
    if (r->th.joinable()) r->th.join();
    for (fastx::Batch *b : r->q) delete b;
    delete r;
    return KMCPG_OK;
}

}  // extern "C"

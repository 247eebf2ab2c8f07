// Fuzz driver of kmcp_b200/csrc/fastgz.h (built with ASan/UBSan by tests/test_fastgz.py).
// argv: <two-member .gz> <variants> <plain bytes of member 1>.  Decodes the intact file and N damaged variants in-process, with
// random input and output piece sizes.  A variant may be refused; if it is accepted its output must be what gzread would
// give for it: both members, one whole member (the other one cut off, deleted, or turned into ignored trailing garbage),
// nothing, or — when the first magic bytes were hit — the file itself (not gzip: passed through).
#include "pargz.h"
#include <stdio.h>
#include <stdlib.h>
#include <zlib.h>
static int decode(const std::vector<uint8_t>& z, std::vector<uint8_t>* out, size_t rd, size_t ck) {
    size_t pos = 0;
    fastgz::Inflater inf([&](void* p, size_t n) { n = std::min(std::min(n, rd), z.size() - pos); memcpy(p, z.data() + pos, n); pos += n; return (ssize_t)n; });
    std::vector<uint8_t> buf(ck);
    for (;;) { ssize_t r = inf.read(buf.data(), buf.size()); if (r < 0) return 1; if (!r) return 0; if (out) out->insert(out->end(), buf.begin(), buf.begin() + r); }
}
// the same bytes through the chunk-parallel decoder (needs a file: pread)
static int decode_par(const std::vector<uint8_t>& z, std::vector<uint8_t>* out, int threads, size_t par_chunk, size_t ck) {
    FILE* tf = tmpfile();
    if (!tf) return -2;
    if (!z.empty() && fwrite(z.data(), 1, z.size(), tf) != z.size()) { fclose(tf); return -2; }
    fflush(tf);
    const int fd = fileno(tf);
    int rc = 0;
    if (!fastgz::ParallelInflater::usable(fd)) rc = -1;
    else {
        fastgz::ParallelInflater inf(fd, threads, par_chunk);
        std::vector<uint8_t> buf(ck);
        for (;;) { ssize_t r = inf.read(buf.data(), buf.size()); if (r < 0) { rc = 1; break; } if (!r) break; out->insert(out->end(), buf.begin(), buf.begin() + r); }
    }
    fclose(tf);
    return rc;
}
int main(int argc, char** argv) {
    FILE* f = fopen(argv[1], "rb"); std::vector<uint8_t> z(64 << 20); z.resize(fread(z.data(), 1, z.size(), f)); fclose(f);
    int n = argc > 2 ? atoi(argv[2]) : 200;
    const size_t split_arg = argc > 3 ? (size_t)atol(argv[3]) : 0;
    std::vector<uint8_t> good; if (decode(z, &good, 1 << 20, 1 << 20)) { printf("intact file failed\n"); return 1; }
    const size_t split = std::min(split_arg, good.size());
    for (int th = 1; th <= 4; th++) {           // the intact file through the chunk-parallel decoder, several chunk sizes
        std::vector<uint8_t> po;
        const int prc = decode_par(z, &po, th, th == 4 ? (2u << 20) : 65536u * th, 100000);
        if (prc != 0 || po != good) { printf("parallel decoder failed on the intact file (threads %d, rc %d)\n", th, prc); return 1; }
    }
    uint64_t s = 88172645463325252ull; auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
    int ok = 0, err = 0, same = 0;
    for (int i = 0; i < n; i++) {
        std::vector<uint8_t> d = z;
        int kind = rnd() % 4;
        if (kind == 0) d.resize(rnd() % d.size());
        else if (kind == 1) { for (int k = 0, m = 1 + rnd() % 3; k < m; k++) d[rnd() % d.size()] ^= (uint8_t)(1u << (rnd() % 8)); }
        else if (kind == 2) { size_t a = rnd() % d.size(); for (size_t k = a; k < d.size() && k < a + 1 + rnd() % 64; k++) d[k] = (uint8_t)rnd(); }
        else { size_t a = rnd() % d.size(), b = rnd() % d.size(); if (a > b) std::swap(a, b); d.erase(d.begin() + a, d.begin() + b); }
        std::vector<uint8_t> o;
        int rc = decode(d, &o, 1 + rnd() % 70000, 1 + rnd() % 300000);
        // the chunk-parallel decoder must agree with the sequential one on every variant it can open
        std::vector<uint8_t> po;
        const int prc = decode_par(d, &po, 1 + (int)(rnd() % 3), 65536, 1 + rnd() % 300000);
        if (prc >= 0) {
            if ((prc != 0) != (rc != 0)) { printf("variant %d (kind %d): sequential rc %d, parallel rc %d\n", i, kind, rc, prc); return 1; }
            if (!rc && po != o) { printf("variant %d (kind %d): parallel decoder gave other bytes\n", i, kind); return 1; }
            if (rc && !(po.size() <= good.size() + 70000)) { printf("variant %d: runaway output\n", i); return 1; }
        }
        if (rc) { err++; continue; }
        ok++;
        const std::vector<uint8_t> A(good.begin(), good.begin() + split), B(good.begin() + split, good.end());
        if (o == good) same++;
        else if (!(o == A || o == B || o.empty() || o == d)) { printf("variant %d (kind %d) accepted with wrong bytes (%zu)\n", i, kind, o.size()); return 1; }
    }
    printf("variants %d: rejected %d, accepted %d (identical output %d)\n", n, err, ok, same);
    return 0;
}

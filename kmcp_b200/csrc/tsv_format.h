// tsv_format.h — the number formatting of the result table (search.go S:517-575: `%d` integers, `%.4f` qCov / tCov / jacc, `%.4e` FPR), written so
// that a row costs a few hundred nanoseconds instead of two snprintf calls, byte for byte what printf prints.
//
//  * %.4f: q = round(v·10^4) is found in double arithmetic and then CHECKED exactly: fma(v, 1e4, -(q ± 0.5)) has the exact sign of
//    v·10^4 - (q ± 0.5) (fma rounds once, and rounding keeps the sign), so the candidate is moved by one where the rounded product landed on
//    the wrong side of a half, and an exact tie goes to the even digit — what glibc and Go do (round half to even on the exact binary value).
//  * %.4e: the FPR of a row is a function of (k-mers of the query, matched k-mers) — a few hundred distinct values per batch — so the strings
//    printf makes are kept in a small direct-mapped table keyed by the double's bits.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

namespace tsvfmt {

inline char *put_uint(char *p, uint64_t v) {
    char tmp[24];
    int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) *p++ = tmp[--n];
    return p;
}
inline char *put_int(char *p, int64_t v) {
    if (v < 0) { *p++ = '-'; return put_uint(p, (uint64_t)(-(v + 1)) + 1); }
    return put_uint(p, (uint64_t)v);
}

// printf("%.4f", v), exactly; falls back to snprintf outside [0, 1e9) (and for NaN)
inline char *put_f4(char *p, double v) {
    if (!(v >= 0.0 && v < 1e9)) return p + snprintf(p, 64, "%.4f", v);
    double q = std::nearbyint(v * 1e4);
    for (int guard = 0; guard < 3; guard++) {
        const double lo = std::fma(v, 1e4, -(q - 0.5)), hi = std::fma(v, 1e4, -(q + 0.5));     // exact signs of v·10^4 - (q ∓ 0.5)
        const bool q_even = std::fmod(q, 2.0) == 0.0;
        if (lo < 0 || (lo == 0 && !q_even)) { q -= 1; continue; }       // below the lower half (or an exact tie that belongs to the even q - 1)
        if (hi > 0 || (hi == 0 && !q_even)) { q += 1; continue; }
        break;
    }
    const uint64_t n = (uint64_t)q;
    p = put_uint(p, n / 10000);
    const uint32_t f = (uint32_t)(n % 10000);
    *p++ = '.';
    *p++ = (char)('0' + f / 1000); *p++ = (char)('0' + f / 100 % 10); *p++ = (char)('0' + f / 10 % 10); *p++ = (char)('0' + f % 10);
    return p;
}

// printf("%.4e", v) through a direct-mapped memo (one per formatting thread)
struct E4Cache {
    static constexpr int N = 4096;
    struct Slot { uint64_t bits; uint8_t len; char s[23]; };
    Slot slots[N];
    E4Cache() { for (auto &s : slots) { s.bits = 0x7FF8DEADBEEF0001ull; s.len = 0; } }      // a NaN payload no FPR has
    char *put(char *p, double v) {
        uint64_t b;
        memcpy(&b, &v, 8);
        Slot &s = slots[(b * 0x9E3779B97F4A7C15ull) >> 52];
        if (s.bits != b) {
            const int n = snprintf(s.s, sizeof(s.s), "%.4e", v);
            s.len = (uint8_t)(n < (int)sizeof(s.s) ? n : (int)sizeof(s.s) - 1);
            s.bits = b;
        }
        memcpy(p, s.s, s.len);
        return p + s.len;
    }
};

// compares put_f4 / E4Cache with snprintf on n seeded values: uniform in [0,1], ratios c/n of small integers (what qCov / tCov / jacc are),
// values at and next to every rounding tie k + 0.5 of the fourth decimal, tiny and large values.  Returns the number of mismatches.
inline uint64_t selftest(uint64_t n, uint64_t seed) {
    uint64_t bad = 0, x = seed * 0x9E3779B97F4A7C15ull + 1;
    auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
    E4Cache cache;
    char a[96], b[96];
    auto check = [&](double v) {
        char *e = put_f4(a, v); *e = 0;
        snprintf(b, sizeof(b), "%.4f", v);
        if (strcmp(a, b)) { if (bad < 5) fprintf(stderr, "%%.4f mismatch: %.17g -> %s, printf %s\n", v, a, b); bad++; }
        e = cache.put(a, v); *e = 0;
        snprintf(b, sizeof(b), "%.4e", v);
        if (strcmp(a, b)) { if (bad < 5) fprintf(stderr, "%%.4e mismatch: %.17g -> %s, printf %s\n", v, a, b); bad++; }
    };
    for (uint64_t i = 0; i < n; i++) {
        const uint64_t r = rnd();
        switch (i % 6) {
            case 0: check((double)(r >> 11) / 9007199254740992.0); break;                                   // uniform [0,1)
            case 1: { const uint64_t d = 1 + r % 5000, c = (r >> 20) % (d + 1); check((double)c / (double)d); break; }   // c / n
            case 2: { const uint64_t d = 1 + r % 4000000, c = (r >> 24) % 1000; check((double)c / (double)d); break; }   // c / target size
            case 3: {                                                                                      // around the ties of the 4th decimal
                double v = ((double)(r % 20000) + 0.5) / 1e4;
                const int steps = (int)((r >> 40) % 5) - 2;
                for (int s = 0; s < (steps < 0 ? -steps : steps); s++) v = std::nextafter(v, steps < 0 ? 0.0 : 10.0);
                check(v);
                break;
            }
            case 4: check(std::ldexp((double)(r >> 11), -(int)(53 + r % 40))); break;                       // tiny
            default: check((double)(r % 100000000) / 997.0); break;                                         // up to 1e5
        }
    }
    const double fixed[] = {0.0, 1.0, 0.5, 0.00005, 0.00015, 0.00025, 0.99995, 0.999949999999, 1e-300, 5e-324, 0.12345, 0.12355, 2.5e-5, 7.5e-5, 123456.78905, 999999999.0};
    for (double v : fixed) check(v);
    return bad;
}

}  // namespace tsvfmt

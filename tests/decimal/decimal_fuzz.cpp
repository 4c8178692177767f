// Host-side check of csrc/kdejs_decimal.h (the conversion the DEVICE table parser uses; the same code compiles for
// both): every spelling it accepts must give the bits strtod gives.  Random (digits, exponent) pairs over the whole
// accepted range, doubles printed with 15-19 digits, integers around 2^53 and their halves (exact ties).
// usage: decimal_fuzz <count> <seed>   -> prints "checked N handled M mismatches K"
#include "../../poppunk-mod_b200/csrc/kdejs_decimal.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>

int main(int argc, char **argv) {
    const long count = argc > 1 ? atol(argv[1]) : 1000000;
    std::mt19937_64 r(argc > 2 ? (unsigned long long)atoll(argv[2]) : 1ull);
    long bad = 0, n = 0, handled = 0;
    auto check = [&](const char *s) {
        double v = 0.0;
        const bool ok = kdejs_decimal::parse_field(s, s + strlen(s), v);
        char *end = nullptr;
        const double ref = strtod(s, &end);
        ++n;
        if (!ok) return;
        ++handled;
        if (memcmp(&v, &ref, sizeof v) != 0 || *end) {
            if (bad < 10) printf("BAD %s got %a want %a\n", s, v, ref);
            ++bad;
        }
    };
    char buf[128];
    std::uniform_real_distribution<double> u(0.0, 1.0);
    for (long i = 0; i < count; ++i) {
        unsigned long long w = r();
        const int nd = 1 + (int)(r() % 19);
        unsigned long long lim = 1;
        for (int k = 0; k < nd; ++k) lim *= 10ull;
        w %= lim;
        snprintf(buf, sizeof buf, "%llue%d", w, (int)(r() % 55) - 27);
        check(buf);
        const double x = u(r) * pow(10.0, (double)((int)(r() % 30) - 15));
        static const char *const fmt[5] = {"%.17g", "%.16g", "%.15g", "%.18e", "%.19g"};
        snprintf(buf, sizeof buf, fmt[i % 5], x);
        check(buf);
        const unsigned long long base = (1ull << 53) + (r() % (1ull << 60));
        snprintf(buf, sizeof buf, (i & 1) ? "%llu" : "%llu.5", base);
        if (strlen(buf) <= 20) check(buf);
    }
    const char *fixed[] = {"0", "-0.0", "1.", ".5", "+.5", "5e3", "1E+2", "00012.50", "9007199254740993", "1e27", "1e-27",
                           "0.1", "0.30000000000000004", "1e0005"};
    for (const char *s : fixed) check(s);
    const char *rejected[] = {"", "+-1", "1e", "e5", ".", "-", "1e28", "1e-28", "12345678901234567890", "nan", "inf", "NA", "1 2", "0x10"};
    for (const char *s : rejected) {
        double v;
        if (kdejs_decimal::parse_field(s, s + strlen(s), v)) { printf("ACCEPTED %s\n", s); ++bad; }
    }
    printf("checked %ld handled %ld mismatches %ld\n", n, handled, bad);
    return bad ? 1 : 0;
}

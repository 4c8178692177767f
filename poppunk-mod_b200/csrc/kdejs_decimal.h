// kdejs_decimal.h -- decimal text -> double, correctly rounded, in integer arithmetic only (host and device).
//
// The device-side table parser (k_tsv_parse) must produce the very bits float() / np.loadtxt / pandas produce
// (poppunk_mod.py:360, scripts/pairwise_2D_distance.py:38-45): the discrepancy is defined on those doubles.  The
// spellings these tables hold are [-]digits[.digits][e[+-]digits] with at most 19 significant digits and small
// exponents, i.e. value = w * 10^q with w < 2^64 and |q| <= 27.  Then 10^|q| = 5^|q| * 2^|q| with 5^|q| < 2^63, and
//   q >= 0:  w * 5^q is an exact 128-bit integer;
//   q <  0:  floor(w' * 2^S / D') is a 64-bit quotient with its top bit set (w', D' = w, 5^|q| shifted to the top),
//            and its remainder says whether anything was cut off;
// either way the 53 leading bits, the next bit and "anything below" give round-to-nearest-even exactly, and the power
// of two that is left goes into the exponent field (no rounding: these values are far from the subnormal range).
// Numbers of up to 15 digits take Clinger's shortcut: one IEEE operation on exact operands.
// No table of rounded powers, no ambiguity, no fallback INSIDE the range; everything outside it (more digits, big
// exponents, inf / nan / missing-value spellings, junk) is reported as "not handled" and goes to the host reader.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define KDEJS_HD __host__ __device__ __forceinline__
#else
#define KDEJS_HD inline
#endif

namespace kdejs_decimal {

// 5^0 .. 5^27 and 10^0 .. 10^22 (exact).  On the device they live in constant memory: as local arrays they were
// rebuilt on the stack by every call (ncu: 15 M local stores and 770 MB of DRAM writes per wave of 16 tables).
#define KDEJS_P5_TABLE {1ull, 5ull, 25ull, 125ull, 625ull, 3125ull, 15625ull, 78125ull, 390625ull, 1953125ull, \
                        9765625ull, 48828125ull, 244140625ull, 1220703125ull, 6103515625ull, 30517578125ull, \
                        152587890625ull, 762939453125ull, 3814697265625ull, 19073486328125ull, \
                        95367431640625ull, 476837158203125ull, 2384185791015625ull, 11920928955078125ull, \
                        59604644775390625ull, 298023223876953125ull, 1490116119384765625ull, \
                        7450580596923828125ull}
#define KDEJS_P10_TABLE {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, \
                         1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22}
static const uint64_t kP5Host[28] = KDEJS_P5_TABLE;
static const double kP10Host[23] = KDEJS_P10_TABLE;
#if defined(__CUDACC__)
static __constant__ uint64_t kP5Dev[28] = KDEJS_P5_TABLE;
static __constant__ double kP10Dev[23] = KDEJS_P10_TABLE;
#endif
KDEJS_HD uint64_t pow5(int k) {
#if defined(__CUDA_ARCH__)
    return kP5Dev[k];
#else
    return kP5Host[k];
#endif
}
KDEJS_HD double pow10d(int k) {
#if defined(__CUDA_ARCH__)
    return kP10Dev[k];
#else
    return kP10Host[k];
#endif
}

KDEJS_HD int clz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __clzll((long long)x);
#else
    return __builtin_clzll(x);
#endif
}

KDEJS_HD void mul64(uint64_t a, uint64_t b, uint64_t &hi, uint64_t &lo) {
#if defined(__CUDA_ARCH__)
    lo = a * b;
    hi = __umul64hi(a, b);
#else
    const unsigned __int128 p = (unsigned __int128)a * b;
    lo = (uint64_t)p;
    hi = (uint64_t)(p >> 64);
#endif
}

// (hi:lo) / d with hi < d and d's top bit set: 64-bit quotient and remainder (Knuth D on 32-bit digits; Hacker's
// Delight "divlu").  The same code on host and device, so both produce the same bits by construction.
KDEJS_HD uint64_t div128by64(uint64_t hi, uint64_t lo, uint64_t d, uint64_t &rem) {
    const uint64_t b = 1ull << 32;
    const uint64_t d1 = d >> 32, d0 = d & 0xffffffffull;
    const uint64_t l1 = lo >> 32, l0 = lo & 0xffffffffull;
    uint64_t q1 = hi / d1, r = hi - q1 * d1;
    while (q1 >= b || q1 * d0 > ((r << 32) | l1)) {
        --q1;
        r += d1;
        if (r >= b) break;
    }
    const uint64_t mid = ((hi << 32) | l1) - q1 * d;          // modulo 2^64, as in the classic formulation
    uint64_t q0 = mid / d1;
    r = mid - q0 * d1;
    while (q0 >= b || q0 * d0 > ((r << 32) | l0)) {
        --q0;
        r += d1;
        if (r >= b) break;
    }
    rem = ((mid << 32) | l0) - q0 * d;
    return (q1 << 32) | q0;
}

KDEJS_HD double make_double(uint64_t mant53, int exp2) {       // mant53 * 2^exp2, mant53 in [2^52, 2^53]
    // exact: the result is a normal double (callers keep |value| within ~[1e-28, 1e47])
    if (mant53 == (1ull << 53)) { mant53 >>= 1; ++exp2; }
    const uint64_t bits = ((uint64_t)(exp2 + 52 + 1023) << 52) | (mant53 & ((1ull << 52) - 1));
    double v;
#if defined(__CUDA_ARCH__)
    v = __longlong_as_double((long long)bits);
#else
    __builtin_memcpy(&v, &bits, sizeof v);
#endif
    return v;
}

// w * 10^q, correctly rounded (nearest, ties to even).  Requires w != 0 and -27 <= q <= 27.
KDEJS_HD double scale_exact(uint64_t w, int q) {
    // Short numbers (the core-distance column: three or four digits): w < 2^53 and 10^|q| <= 10^22 are exact doubles,
    // so ONE IEEE multiplication or division is the correctly rounded result (Clinger) -- a tenth of the integer path
    if (w < (1ull << 53) && q >= -22 && q <= 22) {
        const double d = (double)(long long)w;
#if defined(__CUDA_ARCH__)
        return q < 0 ? __ddiv_rn(d, pow10d(-q)) : __dmul_rn(d, pow10d(q));
#else
        return q < 0 ? d / pow10d(-q) : d * pow10d(q);
#endif
    }
    if (q >= 0) {
        uint64_t hi, lo;
        mul64(w, pow5(q), hi, lo);                              // exact product, value = (hi:lo) * 2^q
        // the 64 leading bits of the product and whether anything lies below them
        uint64_t top;
        bool below;
        int shift;                                           // top = product >> shift (shift may be negative: << )
        if (hi) {
            const int lz = clz64(hi);
            top = lz ? (hi << lz) | (lo >> (64 - lz)) : hi;
            below = lz ? (lo << lz) != 0 : lo != 0;
            shift = 64 - lz;
        } else {
            const int lz = clz64(lo);
            top = lo << lz;
            below = false;
            shift = -lz;
        }
        uint64_t m = top >> 11;
        const bool round = (top >> 10) & 1, sticky = below || (top & 0x3ff) != 0;
        if (round && (sticky || (m & 1))) ++m;
        return make_double(m, shift + 11 + q);
    }
    const int k = -q;
    const uint64_t d5 = pow5(k);
    const int cw = clz64(w), cd = clz64(d5);
    const uint64_t wn = w << cw, dn = d5 << cd;
    uint64_t nh, nl;
    int S;
    if (wn >= dn) { nh = wn >> 1; nl = (wn & 1) << 63; S = 63; }
    else { nh = wn; nl = 0; S = 64; }
    uint64_t rem;
    const uint64_t Q = div128by64(nh, nl, dn, rem);         // top bit set
    uint64_t m = Q >> 11;
    const bool round = (Q >> 10) & 1, sticky = (Q & 0x3ff) != 0 || rem != 0;
    if (round && (sticky || (m & 1))) ++m;
    return make_double(m, 11 - S + cd - cw - k);
}

// One field [b, e) (already trimmed of blanks): true and the value when it is a plain decimal inside the range
// above; false for everything else.
KDEJS_HD bool parse_field(const char *b, const char *e, double &v) {
    const char *p = b;
    bool neg = false;
    if (p < e && (*p == '+' || *p == '-')) { neg = (*p == '-'); ++p; }
    uint64_t w = 0;
    int sig = 0, frac = 0, ndig = 0;
    for (; p < e && (unsigned)(*p - '0') < 10u; ++p, ++ndig) {
        if (sig == 0 && *p == '0') continue;
        if (++sig > 19) return false;
        w = w * 10u + (unsigned)(*p - '0');
    }
    if (p < e && *p == '.') {
        ++p;
        for (; p < e && (unsigned)(*p - '0') < 10u; ++p, ++ndig) {
            ++frac;
            if (sig == 0 && *p == '0') continue;
            if (++sig > 19) return false;
            w = w * 10u + (unsigned)(*p - '0');
        }
    }
    if (ndig == 0) return false;
    int ex = 0;
    if (p < e && (*p == 'e' || *p == 'E')) {
        ++p;
        bool eneg = false;
        if (p < e && (*p == '+' || *p == '-')) { eneg = (*p == '-'); ++p; }
        int nd = 0;
        for (; p < e && (unsigned)(*p - '0') < 10u; ++p, ++nd) {
            if (nd >= 4) return false;
            ex = ex * 10 + (*p - '0');
        }
        if (nd == 0) return false;
        if (eneg) ex = -ex;
    }
    if (p != e) return false;
    if (w == 0) { v = neg ? -0.0 : 0.0; return true; }
    const int q = ex - frac;
    if (q < -27 || q > 27) return false;
    const double r = scale_exact(w, q);
    v = neg ? -r : r;
    return true;
}

}  // namespace kdejs_decimal

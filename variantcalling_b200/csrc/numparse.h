// numparse.h -- exact decimal -> binary64 (then float32) conversion, host + device.
//
// htslib stores INFO/FORMAT Float values and QUAL as float32 obtained from strtod(), i.e.
// float(double(text)) with both roundings to nearest-even.  The reference's features are
// built from those values, so K1 must reproduce strtod bit for bit.  Two exact paths:
//   * Clinger's fast path: mantissa <= 2^53 and |exp10| <= 22 -> one IEEE multiply / divide;
//   * Eisel-Lemire with the 128-bit power-of-five table (no fallback needed for mantissas
//     that fit 64 bits, cf. Mushtak & Lemire, "Fast number parsing without fallback").
// More than 19 significant digits: both w and w+1 are converted; if they disagree the
// literal is reported as NUM_BAD (the caller raises) rather than rounded approximately.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define UGVC_HD __host__ __device__ __forceinline__
#define UGVC_HDM __host__ __device__ __forceinline__
#define UGVC_TABLE static __device__ const
#else
#define UGVC_HD static inline
#define UGVC_HDM inline
#define UGVC_TABLE static const
#endif

UGVC_TABLE uint64_t ugvc_pow5_128[651 * 2] = {
#include "pow5_table.inc"
};
#define UGVC_TENS {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11, \
                   1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22}
UGVC_TABLE double ugvc_tens[23] = UGVC_TENS;
#if defined(__CUDACC__)
static const uint64_t ugvc_pow5_128_host[651 * 2] = {
#include "pow5_table.inc"
};
static const double ugvc_tens_host[23] = UGVC_TENS;
#endif

enum { NUM_OK = 0, NUM_MISSING = 1, NUM_BAD = 2 };

UGVC_HD void ugvc_mul64(uint64_t a, uint64_t b, uint64_t& hi, uint64_t& lo) {
#if defined(__CUDA_ARCH__)
    lo = a * b;
    hi = __umul64hi(a, b);
#else
    const unsigned __int128 p = (unsigned __int128)a * b;
    lo = (uint64_t)p;
    hi = (uint64_t)(p >> 64);
#endif
}

UGVC_HD int ugvc_clz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __clzll((long long)x);
#else
    return __builtin_clzll(x);
#endif
}

UGVC_HD double ugvc_bits_to_double(uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)b);
#else
    double d;
    __builtin_memcpy(&d, &b, 8);
    return d;
#endif
}

UGVC_HD uint64_t ugvc_pow5(int idx) {
#if defined(__CUDA_ARCH__)
    return ugvc_pow5_128[idx];
#elif defined(__CUDACC__)
    return ugvc_pow5_128_host[idx];
#else
    return ugvc_pow5_128[idx];
#endif
}

// Correctly rounded w * 10^q (w != 0) as IEEE binary64 bits (positive).
UGVC_HD uint64_t ugvc_eisel_lemire(uint64_t w, int q) {
    if (q < -342) return 0;
    if (q > 308) return 0x7FF0000000000000ull;
    int lz = ugvc_clz64(w);
    w <<= lz;
    const int idx = 2 * (q + 342);
    uint64_t hi, lo;
    ugvc_mul64(w, ugvc_pow5(idx), hi, lo);
    if ((hi & 0x1FFull) == 0x1FFull) {
        uint64_t hi2, lo2;
        ugvc_mul64(w, ugvc_pow5(idx + 1), hi2, lo2);
        lo += hi2;
        if (hi2 > lo) ++hi;
    }
    const int upperbit = (int)(hi >> 63);
    const int shift = upperbit + 64 - 52 - 3;
    uint64_t mantissa = hi >> shift;
    // floor(log2(10^q)) + 63, valid for |q| <= 400
    int power2 = (int)(((152170ll + 65536ll) * (long long)q) >> 16) + 63 + upperbit - lz + 1023;
    if (power2 <= 0) {  // subnormal
        if (-power2 + 1 >= 64) return 0;
        mantissa >>= -power2 + 1;
        mantissa += (mantissa & 1);
        mantissa >>= 1;
        power2 = (mantissa < (1ull << 52)) ? 0 : 1;
        return ((uint64_t)power2 << 52) | (mantissa & ~(1ull << 52));
    }
    if (lo <= 1 && q >= -4 && q <= 23 && ((mantissa & 3) == 1)) {
        if ((mantissa << shift) == hi) mantissa &= ~1ull;  // exact tie: round to even
    }
    mantissa += (mantissa & 1);
    mantissa >>= 1;
    if (mantissa >= (2ull << 52)) {
        mantissa = 1ull << 52;
        ++power2;
    }
    mantissa &= ~(1ull << 52);
    if (power2 >= 0x7FF) return 0x7FF0000000000000ull;
    return ((uint64_t)power2 << 52) | mantissa;
}

UGVC_HD double ugvc_ten(int k) {
#if defined(__CUDA_ARCH__)
    return ugvc_tens[k];
#elif defined(__CUDACC__)
    return ugvc_tens_host[k];
#else
    return ugvc_tens[k];
#endif
}

UGVC_HD bool ugvc_is_digit(unsigned c) { return (c - '0') <= 9u; }
UGVC_HD unsigned ugvc_lower(unsigned c) { return c | 0x20u; }

// Parse one numeric token from a byte source.  Src: unsigned peek() const; void adv().
// On return the source stands on the first byte after the token.
template <class Src>
UGVC_HD int ugvc_parse_num(Src& s, double& out) {
    unsigned c = s.peek();
    bool neg = false;
    if (c == '-' || c == '+') {
        neg = (c == '-');
        s.adv();
        c = s.peek();
    }
    uint64_t m = 0;
    int exp10 = 0;
    bool any = false, truncated = false;
    while (ugvc_is_digit(c)) {
        any = true;
        const unsigned d = c - '0';
        if (m < 1000000000000000000ull) m = m * 10 + d;  // up to 19 significant digits
        else {
            ++exp10;
            truncated |= (d != 0);
        }
        s.adv();
        c = s.peek();
    }
    if (c == '.') {
        s.adv();
        c = s.peek();
        while (ugvc_is_digit(c)) {
            any = true;
            const unsigned d = c - '0';
            if (m < 1000000000000000000ull) {
                m = m * 10 + d;
                --exp10;
            } else
                truncated |= (d != 0);
            s.adv();
            c = s.peek();
        }
        if (!any) {  // a lone "."
            out = 0.0;
            return neg ? NUM_BAD : NUM_MISSING;
        }
    }
    if (!any) {
        // nan / inf / infinity, any case, as strtod accepts them
        const unsigned a = ugvc_lower(c);
        if (a == 'n') {
            s.adv();
            if (ugvc_lower(s.peek()) != 'a') { out = 0.0; return NUM_BAD; }
            s.adv();
            if (ugvc_lower(s.peek()) != 'n') { out = 0.0; return NUM_BAD; }
            s.adv();
            out = ugvc_bits_to_double(0x7FF8000000000000ull);
            return NUM_OK;
        }
        if (a == 'i') {
            s.adv();
            if (ugvc_lower(s.peek()) != 'n') { out = 0.0; return NUM_BAD; }
            s.adv();
            if (ugvc_lower(s.peek()) != 'f') { out = 0.0; return NUM_BAD; }
            s.adv();
            if (ugvc_lower(s.peek()) == 'i') {  // "inity"
                const char* rest = "inity";
                for (int i = 0; i < 5; ++i) {
                    if (ugvc_lower(s.peek()) != (unsigned)rest[i]) { out = 0.0; return NUM_BAD; }
                    s.adv();
                }
            }
            out = ugvc_bits_to_double(neg ? 0xFFF0000000000000ull : 0x7FF0000000000000ull);
            return NUM_OK;
        }
        out = 0.0;
        return NUM_BAD;
    }
    if (ugvc_lower(c) == 'e') {
        s.adv();
        unsigned e = s.peek();
        bool eneg = false;
        if (e == '-' || e == '+') {
            eneg = (e == '-');
            s.adv();
            e = s.peek();
        }
        if (!ugvc_is_digit(e)) {  // "1e", "1e+": not a number for htslib's field parser
            out = 0.0;
            return NUM_BAD;
        }
        int ev = 0;
        while (ugvc_is_digit(e)) {
            if (ev < 100000) ev = ev * 10 + (int)(e - '0');
            s.adv();
            e = s.peek();
        }
        exp10 += eneg ? -ev : ev;
    }
    uint64_t bits;
    if (m == 0) {
        bits = 0;
    } else if (!truncated && m <= (1ull << 53) && exp10 >= -22 && exp10 <= 22) {
        // Clinger: exactly representable operands, one correctly rounded IEEE operation
        const double p = ugvc_ten(exp10 < 0 ? -exp10 : exp10);
        const double v = exp10 < 0 ? (double)m / p : (double)m * p;
        out = neg ? -v : v;
        return NUM_OK;
    } else {
        if (exp10 < -400) exp10 = -400;
        if (exp10 > 400) exp10 = 400;
        bits = ugvc_eisel_lemire(m, exp10);
        if (truncated && bits != ugvc_eisel_lemire(m + 1, exp10)) {
            out = 0.0;
            return NUM_BAD;  // > 19 digits and the tail decides the rounding
        }
    }
    out = ugvc_bits_to_double(bits | (neg ? 0x8000000000000000ull : 0));
    return NUM_OK;
}

struct UgvcPtrSrc {
    const uint8_t* p;
    UGVC_HDM unsigned peek() const { return *p; }
    UGVC_HDM void adv() { ++p; }
};

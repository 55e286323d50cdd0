// Goldilocks field  p = 2^64 - 2^32 + 1  (the reference's gl64_t, ff/gl64_t.cuh:39-298).
// Memory format is the reference's: one canonical uint64_t (< p), not Montgomery.
//
// Internal representation here: any uint64_t ("loose", value mod p).  mul() and the
// final canon() return canonical values; add()/sub() require their SECOND operand to be
// canonical and accept a loose first operand -- exactly the shape of a radix-2 butterfly
// (u loose, t = v*w canonical), which saves the conditional subtraction on every
// add/sub.  2^64 = 2^32 - 1 =: EPS (mod p), 2^96 = -1 (mod p).
#pragma once
#include "../util/hd.cuh"

struct gl64 {
    typedef uint64_t T;          // storage type in HBM
    static constexpr uint64_t P = 0xffffffff00000001ULL;
    static constexpr uint64_t EPS = 0xffffffffULL;
    static constexpr int MAX_LG = 32;
    static constexpr int LG_BYTES = 3;

    static HD T canon(T a) { return a >= P ? a - P : a; }
    static HD T load(T a) { return a; }     // memory -> register (accepts non-canonical input)
    static HD T one() { return 1; }

    // a loose, b canonical -> loose
    static HD T add(T a, T b)
    {
        T s = a + b;
        return s < a ? s + EPS : s;           // wrapped: +2^64 == +EPS; cannot wrap twice as b < p
    }
    // a loose, b canonical -> loose
    static HD T sub(T a, T b)
    {
        T d = a - b;
        return a < b ? d - EPS : d;           // borrowed: -2^64 == -EPS; cannot borrow twice as b < p
    }
    static HD T reduce128(T lo, T hi)
    {
        // lo + hi_lo*2^64 + hi_hi*2^96 == lo - hi_hi + hi_lo*EPS
        T hh = hi >> 32, hl = hi & EPS;
        T t = lo - hh;
        if (lo < hh) t -= EPS;
        T m = hl * EPS;                       // < 2^64
        T r = t + m;
        if (r < t) r += EPS;
        return canon(r);
    }
    // loose x loose -> canonical
    static HD T mul(T a, T b)
    {
#if defined(__CUDA_ARCH__)
        return reduce128(a * b, __umul64hi(a, b));
#else
        unsigned __int128 x = (unsigned __int128)a * b;
        return reduce128((T)x, (T)(x >> 64));
#endif
    }
    static HD T pow(T b, uint64_t e)
    {
        T r = 1;
        b = canon(b);
        for (; e; e >>= 1, b = mul(b, b))
            if (e & 1) r = mul(r, b);
        return r;
    }
    // parameters: ntt/parameters/goldilocks.h:84-160 (default, non-PLONKY2 branch):
    // group_gen = 7, w_(2^32) = 7^((p-1)/2^32); values re-derived, pinned by tests/test_params_pin.py
    static HD T group_gen() { return 7; }
    static HD T root_of_unity_max() { return 0x185629dcda58878cULL; }   // order 2^32
    static HD T inv(T a) { return pow(a, P - 2); }
};

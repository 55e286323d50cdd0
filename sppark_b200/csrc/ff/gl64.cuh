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
    static constexpr uint32_t NTT_MAX_LG_R = 12;   // largest sub-NTT per tile
    static constexpr uint32_t NTT_MAX_THREADS = 1024;
    static constexpr uint32_t LG_EPT = 4;          // NTT: elements per thread per register step
    static constexpr int LG_BYTES = 3;

    static HD T canon(T a) { return a >= P ? a - P : a; }
    static HD T load(T a) { return a; }     // memory -> register (accepts non-canonical input)
    static HD T one() { return 1; }

    // a loose, b canonical -> loose
    static HD T add(T a, T b)
    {
#if defined(__CUDA_ARCH__)
        uint32_t lo, hi, c;
        asm("{ .reg .u32 a0, a1, b0, b1;\n\t"
            "mov.b64 {a0, a1}, %3; mov.b64 {b0, b1}, %4;\n\t"
            "add.cc.u32 %0, a0, b0; addc.cc.u32 %1, a1, b1; addc.u32 %2, 0, 0;\n\t"
            "neg.s32 %2, %2;\n\t"                        // 0 or 0xffffffff == EPS
            "add.cc.u32 %0, %0, %2; addc.u32 %1, %1, 0; }"
            : "=r"(lo), "=r"(hi), "=r"(c) : "l"(a), "l"(b));
        return ((T)hi << 32) | lo;
#else
        T s = a + b;
        return s < a ? s + EPS : s;           // wrapped: +2^64 == +EPS; cannot wrap twice as b < p
#endif
    }
    // a loose, b canonical -> loose
    static HD T sub(T a, T b)
    {
#if defined(__CUDA_ARCH__)
        uint32_t lo, hi, c;
        asm("{ .reg .u32 a0, a1, b0, b1;\n\t"
            "mov.b64 {a0, a1}, %3; mov.b64 {b0, b1}, %4;\n\t"
            "sub.cc.u32 %0, a0, b0; subc.cc.u32 %1, a1, b1; subc.u32 %2, 0, 0;\n\t"   // 0 or EPS
            "sub.cc.u32 %0, %0, %2; subc.u32 %1, %1, 0; }"
            : "=r"(lo), "=r"(hi), "=r"(c) : "l"(a), "l"(b));
        return ((T)hi << 32) | lo;
#else
        T d = a - b;
        return a < b ? d - EPS : d;           // borrowed: -2^64 == -EPS; cannot borrow twice as b < p
#endif
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
        // 4 wide products (each mad.lo.cc/madc.hi.cc pair is one IMAD.WIDE.U32) -> r3:r2:r1:r0,
        // then r0 + r1*2^32 + r2*EPS - r3, one conditional +-EPS per wrap, final canonical fix
        uint32_t r0, r1;
        asm("{ .reg .u32 a0, a1, b0, b1, r2, r3, t;\n\t"
            ".reg .pred q;\n\t"
            "mov.b64 {a0, a1}, %2; mov.b64 {b0, b1}, %3;\n\t"
            "mul.lo.u32 %0, a0, b0; mul.hi.u32 %1, a0, b0;\n\t"
            "mul.lo.u32 r2, a1, b1; mul.hi.u32 r3, a1, b1;\n\t"
            "mad.lo.cc.u32 %1, a0, b1, %1; madc.hi.cc.u32 r2, a0, b1, r2; addc.u32 r3, r3, 0;\n\t"
            "mad.lo.cc.u32 %1, a1, b0, %1; madc.hi.cc.u32 r2, a1, b0, r2; addc.u32 r3, r3, 0;\n\t"
            "sub.cc.u32 %0, %0, r3; subc.cc.u32 %1, %1, 0; subc.u32 t, 0, 0;\n\t"      // - r3*2^96
            "sub.cc.u32 %0, %0, t; subc.u32 %1, %1, 0;\n\t"
            "mad.lo.cc.u32 %0, r2, 0xffffffff, %0; madc.hi.cc.u32 %1, r2, 0xffffffff, %1; addc.u32 t, 0, 0;\n\t"
            "neg.s32 t, t;\n\t"
            "add.cc.u32 %0, %0, t; addc.u32 %1, %1, 0;\n\t"
            "setp.eq.u32 q, %1, 0xffffffff;\n\t"
            "@q setp.ne.u32 q, %0, 0;\n\t"
            "@q sub.u32 %0, %0, 1;\n\t"
            "@q mov.u32 %1, 0; }"
            : "=r"(r0), "=r"(r1) : "l"(a), "l"(b));
        return ((T)r1 << 32) | r0;
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

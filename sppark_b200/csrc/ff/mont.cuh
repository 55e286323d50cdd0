// N x 32-bit Montgomery prime field, R = 2^(32N).
//
// Same memory format as the reference's mont_t (ff/mont_t.cuh:33-44: `uint32_t even[n]`,
// little-endian limbs, Montgomery form), so fp/fr/affine/xyzz buffers are ABI-compatible.
// The arithmetic is this library's own:
//   * multiplication is one fused CIOS ladder over two accumulator files E (pairs at even
//     limb positions) and O (pairs at odd positions).  Every 32x32 product is a single
//     mad.lo.cc/madc.hi.cc pair that ptxas fuses into IMAD.WIDE.U32(.X), all pairs of one
//     row form one carry chain, and the cross-file carry at the vanishing limb is tracked as
//     a 2-bit integer instead of being rippled.  Modulus limbs are compile-time constants and
//     become immediates in SASS.
//   * the portable branch (no __CUDA_ARCH__) is the same ladder on uint64_t, used by the CPU
//     single-stepper in tests/emu and for host-side constants only.
#pragma once
#include "../util/hd.cuh"

namespace ff {

#if defined(__CUDA_ARCH__)
namespace ptx {
DEV void mad_lo_cc(uint32_t& d, uint32_t a, uint32_t b, uint32_t c)
{   asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));   }
DEV void madc_lo_cc(uint32_t& d, uint32_t a, uint32_t b, uint32_t c)
{   asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));   }
DEV void madc_hi_cc(uint32_t& d, uint32_t a, uint32_t b, uint32_t c)
{   asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));   }
DEV void addc(uint32_t& d, uint32_t a, uint32_t b)
{   asm volatile("addc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));   }
DEV void add_cc(uint32_t& d, uint32_t a, uint32_t b)
{   asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));   }
DEV void addc_cc(uint32_t& d, uint32_t a, uint32_t b)
{   asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));   }
DEV void sub_cc(uint32_t& d, uint32_t a, uint32_t b)
{   asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));   }
DEV void subc_cc(uint32_t& d, uint32_t a, uint32_t b)
{   asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));   }
DEV void subc(uint32_t& d, uint32_t a, uint32_t b)
{   asm volatile("subc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));   }
}  // namespace ptx
#endif

// C supplies: static constexpr int N; constexpr accessors P(i), ONE(i), RR(i); M0.
template<class C>
struct mont_t {
    static constexpr int N = C::N;
    uint32_t l[N];

    static HD mont_t zero()
    {
        mont_t r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = 0;
        return r;
    }
    static HD mont_t one()
    {
        mont_t r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = C::ONE(i);
        return r;
    }
    static HD mont_t rr()
    {
        mont_t r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = C::RR(i);
        return r;
    }
    HD bool is_zero() const
    {
        uint32_t acc = 0;
#pragma unroll
        for (int i = 0; i < N; i++) acc |= l[i];
        return acc == 0;
    }
    friend HD bool operator==(const mont_t& a, const mont_t& b)
    {
        uint32_t acc = 0;
#pragma unroll
        for (int i = 0; i < N; i++) acc |= a.l[i] ^ b.l[i];
        return acc == 0;
    }

    // r = a - p if a >= p else a        (a < 2p)
    static HD mont_t final_sub(const mont_t& a, uint32_t top_carry = 0)
    {
        mont_t r;
#if defined(__CUDA_ARCH__)
        uint32_t t[N], borrow;
        ptx::sub_cc(t[0], a.l[0], C::P(0));
#pragma unroll
        for (int i = 1; i < N; i++) ptx::subc_cc(t[i], a.l[i], C::P(i));
        ptx::subc(borrow, top_carry, 0);          // 0 -> a >= p (take t), 0xffffffff -> keep a
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = borrow ? a.l[i] : t[i];
#else
        uint32_t t[N];
        int64_t br = 0;
        for (int i = 0; i < N; i++) {
            int64_t d = (int64_t)a.l[i] - C::P(i) + br;
            t[i] = (uint32_t)d;
            br = d >> 32;
        }
        br += top_carry;
        for (int i = 0; i < N; i++) r.l[i] = br < 0 ? a.l[i] : t[i];
#endif
        return r;
    }

    friend HD mont_t operator+(const mont_t& a, const mont_t& b)
    {
        mont_t s;
        uint32_t carry;
#if defined(__CUDA_ARCH__)
        ptx::add_cc(s.l[0], a.l[0], b.l[0]);
#pragma unroll
        for (int i = 1; i < N; i++) ptx::addc_cc(s.l[i], a.l[i], b.l[i]);
        ptx::addc(carry, 0, 0);
#else
        uint64_t c = 0;
        for (int i = 0; i < N; i++) {
            c += (uint64_t)a.l[i] + b.l[i];
            s.l[i] = (uint32_t)c;
            c >>= 32;
        }
        carry = (uint32_t)c;
#endif
        return final_sub(s, carry);
    }

    friend HD mont_t operator-(const mont_t& a, const mont_t& b)
    {
        mont_t d;
#if defined(__CUDA_ARCH__)
        uint32_t borrow;
        ptx::sub_cc(d.l[0], a.l[0], b.l[0]);
#pragma unroll
        for (int i = 1; i < N; i++) ptx::subc_cc(d.l[i], a.l[i], b.l[i]);
        ptx::subc(borrow, 0, 0);                  // 0 or 0xffffffff
        ptx::add_cc(d.l[0], d.l[0], C::P(0) & borrow);
#pragma unroll
        for (int i = 1; i < N - 1; i++) ptx::addc_cc(d.l[i], d.l[i], C::P(i) & borrow);
        ptx::addc(d.l[N - 1], d.l[N - 1], C::P(N - 1) & borrow);
#else
        int64_t br = 0;
        for (int i = 0; i < N; i++) {
            int64_t x = (int64_t)a.l[i] - b.l[i] + br;
            d.l[i] = (uint32_t)x;
            br = x >> 32;
        }
        if (br) {
            uint64_t c = 0;
            for (int i = 0; i < N; i++) {
                c += (uint64_t)d.l[i] + C::P(i);
                d.l[i] = (uint32_t)c;
                c >>= 32;
            }
        }
#endif
        return d;
    }

    HD mont_t neg() const { return is_zero() ? *this : zero() - *this; }
    HD mont_t cneg(bool flag) const { return flag ? neg() : *this; }
    HD mont_t dbl() const { return *this + *this; }

    // ---- Montgomery product a*b*R^-1 mod p ------------------------------------------
#if defined(__CUDA_ARCH__)
    // one row: acc(coords i..i+N+1) += x * v[0..N-1] placed at coord i.  Products with even
    // j go to the file whose pairs start at coord parity(i), odd j to the other one.
    template<int I, class V>
    static DEV void row(uint32_t (&E)[2 * N + 2], uint32_t (&O)[2 * N + 2], uint32_t x, const V& v)
    {
        constexpr bool even_row = (I & 1) == 0;
        uint32_t (&A1)[2 * N + 2] = even_row ? E : O;      // receives even j, pairs at (I+j, I+j+1)
        uint32_t (&A2)[2 * N + 2] = even_row ? O : E;      // receives odd j
        ptx::mad_lo_cc(A1[I], x, v[0], A1[I]);
        ptx::madc_hi_cc(A1[I + 1], x, v[0], A1[I + 1]);
#pragma unroll
        for (int j = 2; j < N; j += 2) {
            ptx::madc_lo_cc(A1[I + j], x, v[j], A1[I + j]);
            ptx::madc_hi_cc(A1[I + j + 1], x, v[j], A1[I + j + 1]);
        }
        ptx::addc(A1[I + N], A1[I + N], 0);
        ptx::mad_lo_cc(A2[I + 1], x, v[1], A2[I + 1]);
        ptx::madc_hi_cc(A2[I + 2], x, v[1], A2[I + 2]);
#pragma unroll
        for (int j = 3; j < N; j += 2) {
            ptx::madc_lo_cc(A2[I + j], x, v[j], A2[I + j]);
            ptx::madc_hi_cc(A2[I + j + 1], x, v[j], A2[I + j + 1]);
        }
        ptx::addc(A2[I + N + 1], A2[I + N + 1], 0);
    }

    struct modulus_view {
        DEV uint32_t operator[](int j) const { return C::P(j); }
    };

    template<int I>
    static DEV void rows(uint32_t (&E)[2 * N + 2], uint32_t (&O)[2 * N + 2], uint32_t& c,
                         const mont_t& a, const mont_t& b)
    {
        if constexpr (I < N) {
            row<I>(E, O, a.l[I], b.l);
            uint32_t m = (E[I] + O[I] + c) * C::M0;
            row<I>(E, O, m, modulus_view());
            // limb I of the running total is now 0 mod 2^32; its carry moves up as an integer
            uint64_t s = (uint64_t)E[I] + O[I] + c;
            c = (uint32_t)(s >> 32);
            rows<I + 1>(E, O, c, a, b);
        }
    }
#endif

    friend HD mont_t operator*(const mont_t& a, const mont_t& b) { return mul_inline(a, b); }

    // mul_shared: ONE copy of the ladder per kernel, called (not inlined) from the hot loop.  The
    // unrolled product is ~450 instructions and a mixed add has ten of them; inlining all ten
    // overflows the instruction cache (ncu: no_instruction was the top stall, profiles/).
    // Only used from code that is itself inlined into the kernel, so the call depth is one.
#if defined(__CUDA_ARCH__)
    static __device__ __noinline__ mont_t mul_shared(mont_t a, mont_t b) { return mul_inline(a, b); }
#else
    static inline mont_t mul_shared(const mont_t& a, const mont_t& b) { return mul_inline(a, b); }
#endif

    static HD mont_t mul_inline(const mont_t& a, const mont_t& b)
    {
        mont_t r;
#if defined(__CUDA_ARCH__)
        uint32_t E[2 * N + 2], O[2 * N + 2], c = 0;
#pragma unroll
        for (int i = 0; i < 2 * N + 2; i++) E[i] = O[i] = 0;
        rows<0>(E, O, c, a, b);
        ptx::add_cc(r.l[0], E[N], c);
#pragma unroll
        for (int i = 1; i < N; i++) ptx::addc_cc(r.l[i], E[N + i], 0);
        ptx::add_cc(r.l[0], r.l[0], O[N]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) ptx::addc_cc(r.l[i], r.l[i], O[N + i]);
        ptx::addc(r.l[N - 1], r.l[N - 1], O[2 * N - 1]);
#else
        uint32_t t[N + 2];
        for (int i = 0; i < N + 2; i++) t[i] = 0;
        for (int i = 0; i < N; i++) {
            uint64_t c = 0;
            for (int j = 0; j < N; j++) {
                c += (uint64_t)a.l[j] * b.l[i] + t[j];
                t[j] = (uint32_t)c;
                c >>= 32;
            }
            c += t[N];
            t[N] = (uint32_t)c;
            t[N + 1] = (uint32_t)(c >> 32);
            uint32_t m = t[0] * C::M0;
            c = ((uint64_t)m * C::P(0) + t[0]) >> 32;
            for (int j = 1; j < N; j++) {
                c += (uint64_t)m * C::P(j) + t[j];
                t[j - 1] = (uint32_t)c;
                c >>= 32;
            }
            c += t[N];
            t[N - 1] = (uint32_t)c;
            t[N] = t[N + 1] + (uint32_t)(c >> 32);
        }
        for (int i = 0; i < N; i++) r.l[i] = t[i];
        return final_sub(r, t[N]);
#endif
        return final_sub(r);
    }

    HD mont_t sqr() const { return *this * *this; }

    HD mont_t to_mont() const { return *this * rr(); }
    HD mont_t from_mont() const
    {
        mont_t o = zero();
        o.l[0] = 1;
        return *this * o;
    }
    HD mont_t inv() const               // Fermat; off the hot path (final normalisations only)
    {
        mont_t acc = one(), base = *this;
        uint32_t e[N];
        uint32_t borrow = 2;                          // e = p - 2, with borrow (P(0) may be 1)
        for (int i = 0; i < N; i++) {
            uint32_t pi = C::P(i);
            e[i] = pi - borrow;
            borrow = pi < borrow ? 1 : 0;
        }
        for (int i = N * 32 - 1; i >= 0; i--) {
            acc = acc.sqr();
            if ((e[i / 32] >> (i % 32)) & 1) acc = acc * base;
        }
        return acc;
    }
};

}  // namespace ff

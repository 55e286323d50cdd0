// Quadratic extension Fp2 = Fp[u]/(u^2 + BETA) for the G2 MSM (the reference's fp2_t:
// ff/bls12-381-fp2.hpp:25-153 and ff/alt_bn128-fp2.hpp with BETA = 1, ff/bls12-377-fp2.hpp with
// BETA = 5).  Memory format = two consecutive Fp elements (c0, c1), i.e.
// blst_fp2 / arkworks Fq2, so G2 affine / Jacobian buffers are ABI-compatible.  The reference
// spreads an Fp2 element over a pair of lanes (`degree = 2`); here one lane holds both halves and
// the three base-field products of a multiplication go through the shared Montgomery ladder.
// Exposes the same interface as ff::mont_t so that ec::xyzz_t and the MSM kernels are reused.
#pragma once
#include "mont.cuh"

namespace ff {

template<class Fp, int BETA = 1>
struct fp2_t {
    static_assert(BETA == 1 || BETA == 5, "u^2 = -1 (BLS12-381, BN254) or u^2 = -5 (BLS12-377)");
    // BETA * a and (BETA - 1) * a by additions
    static HD Fp times_beta(const Fp& a) { return BETA == 1 ? a : a.dbl().dbl() + a; }
    static HD Fp times_beta_minus_1(const Fp& a) { return a.dbl().dbl(); }
    static constexpr int N = 2 * Fp::N;
    uint32_t l[N];

    HD Fp c0() const
    {
        Fp r;
#pragma unroll
        for (int i = 0; i < Fp::N; i++) r.l[i] = l[i];
        return r;
    }
    HD Fp c1() const
    {
        Fp r;
#pragma unroll
        for (int i = 0; i < Fp::N; i++) r.l[i] = l[Fp::N + i];
        return r;
    }
    static HD fp2_t make(const Fp& a, const Fp& b)
    {
        fp2_t r;
#pragma unroll
        for (int i = 0; i < Fp::N; i++) { r.l[i] = a.l[i]; r.l[Fp::N + i] = b.l[i]; }
        return r;
    }
    static HD fp2_t zero() { return make(Fp::zero(), Fp::zero()); }
    static HD fp2_t one() { return make(Fp::one(), Fp::zero()); }
    HD bool is_zero() const
    {
        uint32_t acc = 0;
#pragma unroll
        for (int i = 0; i < N; i++) acc |= l[i];
        return acc == 0;
    }
    friend HD bool operator==(const fp2_t& a, const fp2_t& b)
    {
        uint32_t acc = 0;
#pragma unroll
        for (int i = 0; i < N; i++) acc |= a.l[i] ^ b.l[i];
        return acc == 0;
    }
    friend HD fp2_t operator+(const fp2_t& a, const fp2_t& b) { return make(a.c0() + b.c0(), a.c1() + b.c1()); }
    friend HD fp2_t operator-(const fp2_t& a, const fp2_t& b) { return make(a.c0() - b.c0(), a.c1() - b.c1()); }
    HD fp2_t neg() const { return make(c0().neg(), c1().neg()); }
    HD fp2_t dbl() const { return *this + *this; }

    // (a0 + a1 u)(b0 + b1 u) = (a0 b0 - BETA a1 b1) + ((a0 + a1)(b0 + b1) - a0 b0 - a1 b1) u
    template<bool SHARED>
    static HD fp2_t mul_impl(const fp2_t& a, const fp2_t& b)
    {
        const Fp a0 = a.c0(), a1 = a.c1(), b0 = b.c0(), b1 = b.c1();
        Fp v0, v1, m;
        if (SHARED) {
            v0 = Fp::mul_shared(a0, b0);
            v1 = Fp::mul_shared(a1, b1);
            m = Fp::mul_shared(a0 + a1, b0 + b1);
        } else {
            v0 = a0 * b0;
            v1 = a1 * b1;
            m = (a0 + a1) * (b0 + b1);
        }
        return make(v0 - times_beta(v1), m - v0 - v1);
    }
    friend HD fp2_t operator*(const fp2_t& a, const fp2_t& b) { return mul_impl<false>(a, b); }
    static HD fp2_t mul_shared(const fp2_t& a, const fp2_t& b) { return mul_impl<true>(a, b); }
    // (a0 + a1 u)^2 = (a0 + a1)(a0 - BETA a1) + (BETA - 1) a0 a1 + 2 a0 a1 u
    template<bool SHARED>
    static HD fp2_t sqr_impl(const fp2_t& a)
    {
        const Fp a0 = a.c0(), a1 = a.c1();
        const Fp t = SHARED ? Fp::mul_shared(a0 + a1, a0 - times_beta(a1)) : (a0 + a1) * (a0 - times_beta(a1));
        const Fp v = SHARED ? Fp::mul_shared(a0, a1) : a0 * a1;
        return make(BETA == 1 ? t : t + times_beta_minus_1(v), v.dbl());
    }
    HD fp2_t sqr() const { return sqr_impl<false>(*this); }
    static HD fp2_t sqr_shared(const fp2_t& a) { return sqr_impl<true>(a); }
    static HD fp2_t msub_shared(const fp2_t& a, const fp2_t& b, const fp2_t& c, const fp2_t& d)
    {   return mul_shared(a, b) - mul_shared(c, d);   }

    // 1/(a0 + a1 u) = (a0 - a1 u) / (a0^2 + BETA a1^2)
    HD fp2_t inv() const
    {
        const Fp a0 = c0(), a1 = c1();
        const Fp d = (a0 * a0 + times_beta(a1 * a1)).inv();
        return make(a0 * d, (a1 * d).neg());
    }
};

}  // namespace ff

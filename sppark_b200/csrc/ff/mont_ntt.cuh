// NTT view of a 256-bit Montgomery field: the interface ntt_core.cuh expects from a field
// (storage type T, load/canon/add/sub/mul/one/pow/inv, root of unity, coset generator) on top
// of ff::mont_t.  Memory format = the reference's fr_t for these fields: 8 x uint32 Montgomery
// limbs (ff/bls12-381.hpp:32-52, ff/pasta.hpp), so `compute_ntt` buffers are ABI-compatible.
// Replaces the "wide" kernels of the reference (ntt/kernels/{ct,gs}_mixed_radix_wide.cu).
#pragma once
#include "fields.cuh"

namespace ff {

template<class P>
struct mont_ntt {
    typedef mont_t<P> M;
    struct alignas(16) T { uint32_t l[P::N]; };
    static constexpr int MAX_LG = P::TWO_ADICITY;
    static constexpr uint32_t LG_EPT = 2;            // 4 elements (32 words) per thread per step
    static constexpr uint32_t NTT_MAX_LG_R = 11;     // 2^11 x 32 B = 64 KiB tile + 64 KiB of twiddles
    static constexpr uint32_t NTT_MAX_THREADS = 512;
    static constexpr int LG_BYTES = 5;

    static HD M in(const T& a)
    {
        M m;
#pragma unroll
        for (int i = 0; i < P::N; i++) m.l[i] = a.l[i];
        return m;
    }
    static HD T out(const M& m)
    {
        T a;
#pragma unroll
        for (int i = 0; i < P::N; i++) a.l[i] = m.l[i];
        return a;
    }
    static HD T canon(const T& a) { return a; }
    static HD T load(const T& a) { return a; }        // inputs are Montgomery residues < p, as in the reference
    static HD T one() { return out(M::one()); }
    static HD T add(const T& a, const T& b) { return out(in(a) + in(b)); }
    static HD T sub(const T& a, const T& b) { return out(in(a) - in(b)); }
    static HD T mul(const T& a, const T& b) { return out(in(a) * in(b)); }
    static HD T pow(const T& b, uint64_t e)
    {
        M r = M::one(), x = in(b);
        for (; e; e >>= 1, x = x * x)
            if (e & 1) r = r * x;
        return out(r);
    }
    static HD T inv(const T& a) { return out(in(a).inv()); }
    static HD T group_gen()
    {
        T g;
#pragma unroll
        for (int i = 0; i < P::N; i++) g.l[i] = P::GEN(i);
        return g;
    }
    static HD T root_of_unity_max()                  // order 2^TWO_ADICITY
    {
        T w;
#pragma unroll
        for (int i = 0; i < P::N; i++) w.l[i] = P::ROOT(i);
        return w;
    }
};

typedef mont_ntt<bls12_381_fr_params> bls12_381_fr_ntt;   // FEATURE_BLS12_381
typedef mont_ntt<vesta_fp_params> pallas_fr_ntt;          // FEATURE_PALLAS: fr = Vesta's base field
typedef mont_ntt<pallas_fp_params> vesta_fr_ntt;          // FEATURE_VESTA:  fr = Pallas' base field
typedef mont_ntt<bn254_fr_params> bn254_fr_ntt;           // FEATURE_BN254     (2-adicity 28)
typedef mont_ntt<bls12_377_fr_params> bls12_377_fr_ntt;   // FEATURE_BLS12_377 (2-adicity 47)

}  // namespace ff

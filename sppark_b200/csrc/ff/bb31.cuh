// BabyBear field  p = 15*2^27 + 1 = 0x78000001 in Montgomery form with R = 2^32
// (the reference's bb31_t = mont32_t<31,0x78000001,0x77ffffff,0x45dddde3,0x0ffffffe>,
// ff/baby_bear.hpp:19, ff/mont32_t.cuh:196-211).  Memory words are Montgomery residues,
// exactly as the reference stores them; all values here are canonical (< p).
#pragma once
#include "../util/hd.cuh"

struct bb31 {
    typedef uint32_t T;
    static constexpr uint32_t P = 0x78000001u;
    static constexpr uint32_t M0 = 0x77ffffffu;      // -p^-1 mod 2^32
    static constexpr uint32_t RR = 0x45dddde3u;      // 2^64 mod p
    static constexpr uint32_t ONE = 0x0ffffffeu;     // 2^32 mod p
    static constexpr int MAX_LG = 27;
    static constexpr uint32_t NTT_MAX_LG_R = 12;
    static constexpr uint32_t NTT_MAX_THREADS = 1024;
    static constexpr uint32_t LG_EPT = 4;          // NTT: elements per thread per register step
    static constexpr int LG_BYTES = 2;

    static HD T canon(T a) { return a; }
    static HD T tight(T a) { return a; }                   // every value is canonical
    static HD T load(T a) { return a >= P ? a - P : a; }   // reference test inputs are already < p
    static HD T one() { return ONE; }
    static HD T add(T a, T b)
    {
        T s = a + b;
        return s >= P ? s - P : s;
    }
    static HD T sub(T a, T b) { return a >= b ? a - b : a + (P - b); }
    static HD T mul(T a, T b)
    {
        uint64_t x = (uint64_t)a * b;
        uint32_t m = (uint32_t)x * M0;
        uint64_t y = x + (uint64_t)m * P;    // < 2^64: x < p^2 < 2^62, m*p < 2^63
        uint32_t r = (uint32_t)(y >> 32);
        return r >= P ? r - P : r;
    }
    static HD T to_mont(uint32_t a) { return mul(a % P, RR); }
    static HD T pow(T b, uint64_t e)
    {
        T r = ONE;
        for (; e; e >>= 1, b = mul(b, b))
            if (e & 1) r = mul(r, b);
        return r;
    }
    // ntt/parameters/baby_bear.h:76-143 (default, non-CANONICAL branch): group_gen = 3,
    // w_(2^27) = 137 (true values); re-derived, pinned by tests/test_params_pin.py
    static HD T group_gen() { return to_mont(3); }
    static HD T root_of_unity_max() { return to_mont(137); }            // order 2^27
    static HD T inv(T a) { return pow(a, P - 2); }
};

/*
 * oracle/ff.c -- TEST INFRASTRUCTURE ONLY (CPU oracle).  See ff.h.
 * Word-serial (CIOS) Montgomery arithmetic; stands in for blst/semolina host
 * fields that the reference does not vendor.
 */
#include "ff.h"
#include <string.h>

typedef unsigned __int128 u128;

static int ge_p(const ff_ctx *c, const uint64_t *a)
{
    for (int i = c->n; i--;) {
        if (a[i] != c->p[i])
            return a[i] > c->p[i];
    }
    return 1;
}

static void sub_p(const ff_ctx *c, uint64_t *a)
{
    uint64_t borrow = 0;
    for (int i = 0; i < c->n; i++) {
        u128 t = (u128)a[i] - c->p[i] - borrow;
        a[i] = (uint64_t)t;
        borrow = (uint64_t)(t >> 64) & 1;
    }
}

void ff_set_zero(ff_t *r) { memset(r, 0, sizeof(*r)); }

void ff_set_one(const ff_ctx *c, ff_t *r)
{
    ff_set_zero(r);
    memcpy(r->l, c->one, sizeof(uint64_t) * c->n);
}

int ff_is_zero(const ff_ctx *c, const ff_t *a)
{
    uint64_t acc = 0;
    for (int i = 0; i < c->n; i++)
        acc |= a->l[i];
    return acc == 0;
}

int ff_eq(const ff_ctx *c, const ff_t *a, const ff_t *b)
{
    return memcmp(a->l, b->l, sizeof(uint64_t) * c->n) == 0;
}

void ff_add(const ff_ctx *c, ff_t *r, const ff_t *a, const ff_t *b)
{
    uint64_t t[FF_MAX_LIMBS] = {0}, carry = 0;
    for (int i = 0; i < c->n; i++) {
        u128 s = (u128)a->l[i] + b->l[i] + carry;
        t[i] = (uint64_t)s;
        carry = (uint64_t)(s >> 64);
    }
    if (carry || ge_p(c, t))
        sub_p(c, t);
    memcpy(r->l, t, sizeof(t));
}

void ff_sub(const ff_ctx *c, ff_t *r, const ff_t *a, const ff_t *b)
{
    uint64_t t[FF_MAX_LIMBS] = {0}, borrow = 0;
    for (int i = 0; i < c->n; i++) {
        u128 s = (u128)a->l[i] - b->l[i] - borrow;
        t[i] = (uint64_t)s;
        borrow = (uint64_t)(s >> 64) & 1;
    }
    if (borrow) {
        uint64_t carry = 0;
        for (int i = 0; i < c->n; i++) {
            u128 s = (u128)t[i] + c->p[i] + carry;
            t[i] = (uint64_t)s;
            carry = (uint64_t)(s >> 64);
        }
    }
    memcpy(r->l, t, sizeof(t));
}

void ff_neg(const ff_ctx *c, ff_t *r, const ff_t *a)
{
    ff_t z;
    ff_set_zero(&z);
    if (ff_is_zero(c, a))
        *r = z;
    else
        ff_sub(c, r, &z, a);
}

void ff_mul(const ff_ctx *c, ff_t *r, const ff_t *a, const ff_t *b)
{
    const int n = c->n;
    uint64_t t[FF_MAX_LIMBS + 2] = {0};

    for (int i = 0; i < n; i++) {
        uint64_t carry = 0;
        for (int j = 0; j < n; j++) {
            u128 x = (u128)a->l[j] * b->l[i] + t[j] + carry;
            t[j] = (uint64_t)x;
            carry = (uint64_t)(x >> 64);
        }
        u128 x = (u128)t[n] + carry;
        t[n] = (uint64_t)x;
        t[n + 1] = (uint64_t)(x >> 64);

        uint64_t m = t[0] * c->m0;
        x = (u128)m * c->p[0] + t[0];
        carry = (uint64_t)(x >> 64);
        for (int j = 1; j < n; j++) {
            x = (u128)m * c->p[j] + t[j] + carry;
            t[j - 1] = (uint64_t)x;
            carry = (uint64_t)(x >> 64);
        }
        x = (u128)t[n] + carry;
        t[n - 1] = (uint64_t)x;
        t[n] = t[n + 1] + (uint64_t)(x >> 64);
    }
    if (t[n] || ge_p(c, t))
        sub_p(c, t);
    ff_set_zero(r);
    memcpy(r->l, t, sizeof(uint64_t) * n);
}

void ff_sqr(const ff_ctx *c, ff_t *r, const ff_t *a) { ff_mul(c, r, a, a); }

void ff_to_mont(const ff_ctx *c, ff_t *r, const ff_t *a)
{
    ff_t rr;
    ff_set_zero(&rr);
    memcpy(rr.l, c->rr, sizeof(uint64_t) * c->n);
    ff_mul(c, r, a, &rr);
}

void ff_from_mont(const ff_ctx *c, ff_t *r, const ff_t *a)
{
    ff_t one;
    ff_set_zero(&one);
    one.l[0] = 1;
    ff_mul(c, r, a, &one);
}

void ff_inv(const ff_ctx *c, ff_t *r, const ff_t *a)
{
    /* a^(p-2), square-and-multiply from the top bit */
    uint64_t e[FF_MAX_LIMBS];
    memcpy(e, c->p, sizeof(e));
    e[0] -= 2; /* p is odd and > 2: no borrow */
    ff_t acc, base = *a;
    ff_set_one(c, &acc);
    for (int i = c->n * 64; i--;) {
        ff_sqr(c, &acc, &acc);
        if ((e[i / 64] >> (i % 64)) & 1)
            ff_mul(c, &acc, &acc, &base);
    }
    *r = acc;
}

void ff_ctx_init(ff_ctx *c, const uint64_t *p, int n, int nbits)
{
    memset(c, 0, sizeof(*c));
    c->n = n;
    c->nbits = nbits;
    memcpy(c->p, p, sizeof(uint64_t) * n);

    /* m0 = -p^-1 mod 2^64 by Newton iteration (p odd) */
    uint64_t inv = 1;
    for (int i = 0; i < 6; i++)
        inv *= 2 - p[0] * inv;
    c->m0 = 0 - inv;

    /* ONE = 2^(64n) mod p, RR = 2^(128n) mod p by repeated modular doubling */
    uint64_t t[FF_MAX_LIMBS] = {0};
    t[0] = 1;
    for (int i = 0; i < 128 * n; i++) {
        uint64_t carry = 0;
        for (int j = 0; j < n; j++) {
            uint64_t hi = t[j] >> 63;
            t[j] = (t[j] << 1) | carry;
            carry = hi;
        }
        if (carry || ge_p(c, t))
            sub_p(c, t);
        if (i == 64 * n - 1)
            memcpy(c->one, t, sizeof(uint64_t) * n);
    }
    memcpy(c->rr, t, sizeof(uint64_t) * n);
}

#define DEFINE_FIELD(name, n, nbits, ...)                          \
    const ff_ctx *name(void)                                       \
    {                                                              \
        static ff_ctx ctx;                                         \
        static int ready;                                          \
        if (!ready) {                                              \
            static const uint64_t p[] = {__VA_ARGS__};             \
            ff_ctx_init(&ctx, p, n, nbits);                        \
            __atomic_store_n(&ready, 1, __ATOMIC_RELEASE);         \
        }                                                          \
        return &ctx;                                               \
    }

/* moduli: ff/bls12-381.hpp:100-104,120-123; ff/pasta.hpp:14-19,33-38 */
DEFINE_FIELD(ff_bls12_381_fp, 6, 381,
             0xb9feffffffffaaabULL, 0x1eabfffeb153ffffULL, 0x6730d2a0f6b0f624ULL,
             0x64774b84f38512bfULL, 0x4b1ba7b6434bacd7ULL, 0x1a0111ea397fe69aULL)
DEFINE_FIELD(ff_bls12_381_fr, 4, 255,
             0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL,
             0x73eda753299d7d48ULL)
DEFINE_FIELD(ff_pallas_fp, 4, 255,
             0x992d30ed00000001ULL, 0x224698fc094cf91bULL, 0x0000000000000000ULL,
             0x4000000000000000ULL)
DEFINE_FIELD(ff_vesta_fp, 4, 255,
             0x8c46eb2100000001ULL, 0x224698fc0994a8ddULL, 0x0000000000000000ULL,
             0x4000000000000000ULL)
/* ff/alt_bn128.hpp:13-16,31-34; ff/bls12-377.hpp:13-17,35-38 */
DEFINE_FIELD(ff_bn254_fp, 4, 254,
             0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL,
             0x30644e72e131a029ULL)
DEFINE_FIELD(ff_bn254_fr, 4, 254,
             0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL,
             0x30644e72e131a029ULL)
DEFINE_FIELD(ff_bls12_377_fp, 6, 377,
             0x8508c00000000001ULL, 0x170b5d4430000000ULL, 0x1ef3622fba094800ULL,
             0x1a22d9f300f5138fULL, 0xc63b05c06ca1493bULL, 0x01ae3a4617c510eaULL)
DEFINE_FIELD(ff_bls12_377_fr, 4, 253,
             0x0a11800000000001ULL, 0x59aa76fed0000001ULL, 0x60b44d1e5c37b001ULL,
             0x12ab655e9a2ca556ULL)

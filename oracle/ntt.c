/*
 * oracle/ntt.c -- TEST INFRASTRUCTURE ONLY (CPU oracle).  See ntt.h for the
 * definition being restated and the reference lines it follows.
 */
#include "ntt.h"
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;

/* ------------------------------------------------------------ fields -- */

#define GL_P 0xffffffff00000001ULL

static inline uint64_t gl_add(uint64_t a, uint64_t b)
{
    uint64_t s = a + b;
    if (s < a || s >= GL_P)
        s -= GL_P;
    return s;
}
static inline uint64_t gl_sub(uint64_t a, uint64_t b) { return a >= b ? a - b : a + (GL_P - b); }
static inline uint64_t gl_mul(uint64_t a, uint64_t b)
{
    /* 2^64 = 2^32 - 1, 2^96 = -1 (mod p) */
    u128 x = (u128)a * b;
    uint64_t lo = (uint64_t)x, hi = (uint64_t)(x >> 64);
    uint64_t hh = hi >> 32, hl = hi & 0xffffffffULL;
    uint64_t t = gl_sub(lo >= GL_P ? lo - GL_P : lo, hh);
    return gl_add(t, hl * 0xffffffffULL);
}
uint64_t oracle_gl64_mul(uint64_t a, uint64_t b) { return gl_mul(a % GL_P, b % GL_P); }

#define BB_P 0x78000001U
static inline uint32_t bb_add(uint32_t a, uint32_t b)
{
    uint32_t s = a + b;
    return s >= BB_P ? s - BB_P : s;
}
static inline uint32_t bb_sub(uint32_t a, uint32_t b) { return a >= b ? a - b : a + (BB_P - b); }
static inline uint32_t bb_mul(uint32_t a, uint32_t b) { return (uint32_t)((uint64_t)a * b % BB_P); }

static uint64_t gl_pow(uint64_t b, uint64_t e)
{
    uint64_t r = 1;
    for (; e; e >>= 1, b = gl_mul(b, b))
        if (e & 1)
            r = gl_mul(r, b);
    return r;
}
static uint32_t bb_pow(uint32_t b, uint64_t e)
{
    uint32_t r = 1;
    for (; e; e >>= 1, b = bb_mul(b, b))
        if (e & 1)
            r = bb_mul(r, b);
    return r;
}

uint64_t oracle_gl64_root(unsigned lg_n, int inverse)
{
    uint64_t w = gl_pow(7, (GL_P - 1) >> 32);          /* primitive 2^32-th root */
    for (unsigned i = 32; i > lg_n; i--)
        w = gl_mul(w, w);
    return inverse ? gl_pow(w, GL_P - 2) : w;
}
uint32_t oracle_bb31_root(unsigned lg_n, int inverse)
{
    uint32_t w = 137;                                  /* primitive 2^27-th root */
    for (unsigned i = 27; i > lg_n; i--)
        w = bb_mul(w, w);
    return inverse ? bb_pow(w, BB_P - 2) : w;
}

static inline size_t bitrev(size_t i, unsigned lg)
{
    size_t r = 0;
    for (unsigned b = 0; b < lg; b++)
        r |= ((i >> b) & 1) << (lg - 1 - b);
    return r;
}

/* ------------------------------------------------- generic by macro -- */

#define NTT_IMPL(NAME, T, ADD, SUB, MUL, POW, ROOT, GEN, PM2, MAXLG, REDUCE)                   \
    typedef struct {                                                                           \
        T *a;                                                                                  \
        const T *tw;                                                                           \
        size_t n, half, lo, hi;                                                                \
    } NAME##_job;                                                                              \
    static void *NAME##_stage(void *arg)                                                       \
    {                                                                                          \
        NAME##_job *j = arg;                                                                   \
        size_t half = j->half, step = j->n / (2 * half);                                       \
        for (size_t b = j->lo; b < j->hi; b++) {                                               \
            size_t blk = b / half, k = b % half, i0 = blk * 2 * half + k;                      \
            T u = j->a[i0], v = MUL(j->a[i0 + half], j->tw[k * step]);                         \
            j->a[i0] = ADD(u, v);                                                              \
            j->a[i0 + half] = SUB(u, v);                                                       \
        }                                                                                      \
        return NULL;                                                                           \
    }                                                                                          \
    /* natural in, natural out, X[k] = sum x[j] w^(jk) */                                      \
    static void NAME##_fast(T *a, unsigned lg, T w, int nthreads)                              \
    {                                                                                          \
        size_t n = (size_t)1 << lg;                                                            \
        for (size_t i = 0; i < n; i++) {                                                       \
            size_t r = bitrev(i, lg);                                                          \
            if (r > i) {                                                                       \
                T t = a[i];                                                                    \
                a[i] = a[r];                                                                   \
                a[r] = t;                                                                      \
            }                                                                                  \
        }                                                                                      \
        T *tw = malloc(sizeof(T) * (n / 2 ? n / 2 : 1));                                       \
        tw[0] = 1;                                                                             \
        for (size_t i = 1; i < n / 2; i++)                                                     \
            tw[i] = MUL(tw[i - 1], w);                                                         \
        if (nthreads < 1)                                                                      \
            nthreads = 1;                                                                      \
        if (n < 4096)                                                                          \
            nthreads = 1;                                                                      \
        pthread_t th[64];                                                                      \
        NAME##_job jobs[64];                                                                   \
        if (nthreads > 64)                                                                     \
            nthreads = 64;                                                                     \
        for (size_t half = 1; half < n; half <<= 1) {                                          \
            size_t total = n / 2, per = (total + nthreads - 1) / nthreads;                     \
            for (int t = 0; t < nthreads; t++) {                                               \
                size_t lo = t * per, hi = lo + per > total ? total : lo + per;                 \
                if (lo > total)                                                                \
                    lo = total;                                                                \
                jobs[t] = (NAME##_job){a, tw, n, half, lo, hi};                                \
                if (nthreads == 1)                                                             \
                    NAME##_stage(&jobs[t]);                                                    \
                else                                                                           \
                    pthread_create(&th[t], NULL, NAME##_stage, &jobs[t]);                      \
            }                                                                                  \
            if (nthreads > 1)                                                                  \
                for (int t = 0; t < nthreads; t++)                                             \
                    pthread_join(th[t], NULL);                                                 \
        }                                                                                      \
        free(tw);                                                                              \
    }                                                                                          \
    static void NAME##_dft(T *a, unsigned lg, T w)                                             \
    {                                                                                          \
        size_t n = (size_t)1 << lg;                                                            \
        T *out = malloc(sizeof(T) * n), *pw = malloc(sizeof(T) * n);                           \
        pw[0] = 1;                                                                             \
        for (size_t i = 1; i < n; i++)                                                         \
            pw[i] = MUL(pw[i - 1], w);                                                         \
        for (size_t k = 0; k < n; k++) {                                                       \
            T acc = 0;                                                                         \
            for (size_t j = 0; j < n; j++)                                                     \
                acc = ADD(acc, MUL(a[j], pw[(j * k) & (n - 1)]));                              \
            out[k] = acc;                                                                      \
        }                                                                                      \
        memcpy(a, out, sizeof(T) * n);                                                         \
        free(out);                                                                             \
        free(pw);                                                                              \
    }                                                                                          \
    int NAME(T *a, unsigned lg, int order, int direction, int type, int algo, int nthreads)    \
    {                                                                                          \
        if (lg == 0)                                                                           \
            return 0;                                                                          \
        if (lg > MAXLG)                                                                        \
            return -1;                                                                         \
        size_t n = (size_t)1 << lg;                                                            \
        for (size_t i = 0; i < n; i++)                                                         \
            a[i] = REDUCE(a[i]);                                                               \
        /* reference semantics (ntt/ntt.cuh:174-212): RR = GS on NATURAL input + bit_rev,    */ \
        /* i.e. the same transform as NN (its own tests assert NN == RR, tests/ntt.rs:28-30); */ \
        /* only its coset exponents are bit-reversed (LDE_powers is handed bitrev = true).    */ \
        /* ORACLE_BB is the strict bit-reversed-in / bit-reversed-out transform.              */ \
        int in_rev = order == ORACLE_RN || order == ORACLE_BB;                                 \
        int out_rev = order == ORACLE_NR || order == ORACLE_BB;                                \
        int quirk = order == ORACLE_RR;                                                        \
        if (in_rev)                                                                            \
            for (size_t i = 0; i < n; i++) {                                                   \
                size_t r = bitrev(i, lg);                                                      \
                if (r > i) {                                                                   \
                    T t = a[i];                                                                \
                    a[i] = a[r];                                                               \
                    a[r] = t;                                                                  \
                }                                                                              \
            }                                                                                  \
        if (!direction && type) {                                                              \
            T g = GEN, pw = 1;                                                                 \
            for (size_t i = 0; i < n; i++, pw = MUL(pw, g)) {                                  \
                size_t k = quirk ? bitrev(i, lg) : i;                                          \
                a[k] = MUL(a[k], pw);                                                          \
            }                                                                                  \
        }                                                                                      \
        T w = ROOT(lg, direction);                                                             \
        if (algo == 1)                                                                         \
            NAME##_dft(a, lg, w);                                                              \
        else                                                                                   \
            NAME##_fast(a, lg, w, nthreads);                                                   \
        if (direction) {                                                                       \
            T ninv = POW(POW(2, PM2), lg);             /* 2^-lg */                              \
            for (size_t i = 0; i < n; i++)                                                     \
                a[i] = MUL(a[i], ninv);                                                        \
            if (type) {                                                                        \
                T gi = POW(GEN, PM2), pw = 1;                                                  \
                for (size_t i = 0; i < n; i++, pw = MUL(pw, gi)) {                             \
                    size_t k = quirk ? bitrev(i, lg) : i;                                      \
                    a[k] = MUL(a[k], pw);                                                      \
                }                                                                              \
            }                                                                                  \
        }                                                                                      \
        if (out_rev)                                                                           \
            for (size_t i = 0; i < n; i++) {                                                   \
                size_t r = bitrev(i, lg);                                                      \
                if (r > i) {                                                                   \
                    T t = a[i];                                                                \
                    a[i] = a[r];                                                               \
                    a[r] = t;                                                                  \
                }                                                                              \
            }                                                                                  \
        return 0;                                                                              \
    }

#define GL_REDUCE(x) ((x) >= GL_P ? (x) - GL_P : (x))
#define BB_REDUCE(x) ((x) % BB_P)

NTT_IMPL(oracle_ntt_gl64, uint64_t, gl_add, gl_sub, gl_mul, gl_pow, oracle_gl64_root, 7,
         GL_P - 2, 32, GL_REDUCE)
NTT_IMPL(oracle_ntt_bb31, uint32_t, bb_add, bb_sub, bb_mul, bb_pow, oracle_bb31_root, 3,
         BB_P - 2, 27, BB_REDUCE)

/* ---- 256-bit Montgomery fields (BLS12-381 fr, Pallas/Vesta base fields) -------------------
 * Same definition as above over ff.h arithmetic; memory words are Montgomery residues as in the
 * reference's fr_t.  group_gen = 7 (BLS12-381 fr) / 5 (Pasta), S = 32, w_(2^32) = gen^((p-1)/2^32)
 * -- ntt/parameters/bls12_381.h:11-16, pallas.h:11-16, vesta.h:11-16 (tables checked by
 * tests/test_params_pin.py). */
#include "ff.h"

/* multiplicative generator and 2-adicity: ntt/parameters/{bls12_381,pallas,vesta,alt_bn128,
 * bls12_377}.h (group_gen, S) */
static const ff_ctx *ntt_field(int id, uint64_t *gen, unsigned *two_adicity)
{
    *two_adicity = 32;
    switch (id) {
    case 1: *gen = 7; return ff_bls12_381_fr();
    case 2: *gen = 5; return ff_pallas_fp();
    case 5: *gen = 5; *two_adicity = 28; return ff_bn254_fr();
    case 7: *gen = 22; *two_adicity = 47; return ff_bls12_377_fr();
    default: *gen = 5; return ff_vesta_fp();
    }
}

static void ffx_pow(const ff_ctx *c, ff_t *r, const ff_t *b, const uint64_t *e, int nlimbs)
{
    ff_t acc, base = *b;
    ff_set_one(c, &acc);
    for (int i = nlimbs * 64; i--;) {
        ff_sqr(c, &acc, &acc);
        if ((e[i / 64] >> (i % 64)) & 1)
            ff_mul(c, &acc, &acc, &base);
    }
    *r = acc;
}

static void ffx_small(const ff_ctx *c, ff_t *r, uint64_t v)
{
    ff_t t;
    ff_set_zero(&t);
    t.l[0] = v;
    ff_to_mont(c, r, &t);
}

static void ffx_root(const ff_ctx *c, uint64_t gen, unsigned S, unsigned lg, int inverse, ff_t *w)
{
    uint64_t e[FF_MAX_LIMBS] = {0};              /* (p - 1) >> S, 0 < S < 64 */
    for (int i = 0; i < c->n; i++) {
        uint64_t lo = c->p[i] >> S, hi = i + 1 < c->n ? c->p[i + 1] << (64 - S) : 0;
        e[i] = lo | hi;
    }
    ff_t g;
    ffx_small(c, &g, gen);
    ffx_pow(c, w, &g, e, c->n);
    for (unsigned i = S; i > lg; i--)
        ff_sqr(c, w, w);
    if (inverse)
        ff_inv(c, w, w);
}

int oracle_ntt_ff(int field_id, uint64_t *data, unsigned lg, int order, int direction, int type, int algo)
{
    if (lg == 0)
        return 0;
    if (lg > 26)
        return -1;
    uint64_t gen;
    unsigned S;
    const ff_ctx *c = ntt_field(field_id, &gen, &S);
    if (lg > S)
        return -1;
    const int nl = c->n;
    size_t n = (size_t)1 << lg;
    ff_t *a = malloc(n * sizeof(ff_t)), *tmp = malloc(n * sizeof(ff_t));
    for (size_t i = 0; i < n; i++) {
        ff_set_zero(&a[i]);
        memcpy(a[i].l, data + i * nl, 8 * nl);
    }
    int in_rev = order == ORACLE_RN || order == ORACLE_BB;
    int out_rev = order == ORACLE_NR || order == ORACLE_BB;
    int quirk = order == ORACLE_RR;
    if (in_rev) {
        for (size_t i = 0; i < n; i++) tmp[bitrev(i, lg)] = a[i];
        memcpy(a, tmp, n * sizeof(ff_t));
    }
    ff_t g, pw;
    if (!direction && type) {
        ffx_small(c, &g, gen);
        ff_set_one(c, &pw);
        for (size_t i = 0; i < n; i++) {
            size_t k = quirk ? bitrev(i, lg) : i;
            ff_mul(c, &a[k], &a[k], &pw);
            ff_mul(c, &pw, &pw, &g);
        }
    }
    ff_t w;
    ffx_root(c, gen, S, lg, direction, &w);
    ff_t *pwr = malloc(n * sizeof(ff_t));
    ff_set_one(c, &pwr[0]);
    for (size_t i = 1; i < n; i++) ff_mul(c, &pwr[i], &pwr[i - 1], &w);
    if (algo == 1) {                              /* definition */
        for (size_t k = 0; k < n; k++) {
            ff_t acc, t;
            ff_set_zero(&acc);
            for (size_t j = 0; j < n; j++) {
                ff_mul(c, &t, &a[j], &pwr[(j * k) & (n - 1)]);
                ff_add(c, &acc, &acc, &t);
            }
            tmp[k] = acc;
        }
        memcpy(a, tmp, n * sizeof(ff_t));
    } else {                                      /* radix-2 DIT */
        for (size_t i = 0; i < n; i++) tmp[bitrev(i, lg)] = a[i];
        memcpy(a, tmp, n * sizeof(ff_t));
        for (size_t half = 1; half < n; half <<= 1) {
            size_t step = n / (2 * half);
            for (size_t blk = 0; blk < n; blk += 2 * half)
                for (size_t k = 0; k < half; k++) {
                    ff_t u = a[blk + k], v;
                    ff_mul(c, &v, &a[blk + k + half], &pwr[k * step]);
                    ff_add(c, &a[blk + k], &u, &v);
                    ff_sub(c, &a[blk + k + half], &u, &v);
                }
        }
    }
    if (direction) {
        ff_t two, half_, ninv;
        ffx_small(c, &two, 2);
        ff_inv(c, &half_, &two);
        ff_set_one(c, &ninv);
        for (unsigned i = 0; i < lg; i++) ff_mul(c, &ninv, &ninv, &half_);
        for (size_t i = 0; i < n; i++) ff_mul(c, &a[i], &a[i], &ninv);
        if (type) {
            ffx_small(c, &g, gen);
            ff_inv(c, &g, &g);
            ff_set_one(c, &pw);
            for (size_t i = 0; i < n; i++) {
                size_t k = quirk ? bitrev(i, lg) : i;
                ff_mul(c, &a[k], &a[k], &pw);
                ff_mul(c, &pw, &pw, &g);
            }
        }
    }
    if (out_rev) {
        for (size_t i = 0; i < n; i++) tmp[bitrev(i, lg)] = a[i];
        memcpy(a, tmp, n * sizeof(ff_t));
    }
    for (size_t i = 0; i < n; i++)
        memcpy(data + i * nl, a[i].l, 8 * nl);
    free(pwr);
    free(tmp);
    free(a);
    return 0;
}

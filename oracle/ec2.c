/* oracle/ec2.c -- TEST INFRASTRUCTURE ONLY.  See ec2.h. */
#include "ec2.h"
#include <stdlib.h>
#include <string.h>

typedef struct { ff_t c0, c1; } f2_t;
typedef struct { f2_t X, Y; } g2_aff;
typedef struct { f2_t X, Y, Z; } g2_jac;

#define FP ff_bls12_381_fp()

static void f2_add(f2_t *r, const f2_t *a, const f2_t *b)
{   ff_add(FP, &r->c0, &a->c0, &b->c0); ff_add(FP, &r->c1, &a->c1, &b->c1);   }
static void f2_sub(f2_t *r, const f2_t *a, const f2_t *b)
{   ff_sub(FP, &r->c0, &a->c0, &b->c0); ff_sub(FP, &r->c1, &a->c1, &b->c1);   }
/* schoolbook (a0 b0 - a1 b1) + (a0 b1 + a1 b0) u */
static void f2_mul(f2_t *r, const f2_t *a, const f2_t *b)
{
    ff_t t0, t1, t2, t3;
    ff_mul(FP, &t0, &a->c0, &b->c0);
    ff_mul(FP, &t1, &a->c1, &b->c1);
    ff_mul(FP, &t2, &a->c0, &b->c1);
    ff_mul(FP, &t3, &a->c1, &b->c0);
    ff_sub(FP, &r->c0, &t0, &t1);
    ff_add(FP, &r->c1, &t2, &t3);
}
static void f2_sqr(f2_t *r, const f2_t *a) { f2_t t = *a; f2_mul(r, &t, &t); }
static void f2_dbl(f2_t *r, const f2_t *a) { f2_t t = *a; f2_add(r, &t, &t); }
/* ff/bls12-381-fp2.hpp:366-377: 1/(a0 + a1 u) = (a0 - a1 u)/(a0^2 + a1^2) */
static void f2_inv(f2_t *r, const f2_t *a)
{
    ff_t t0, t1;
    ff_sqr(FP, &t0, &a->c0);
    ff_sqr(FP, &t1, &a->c1);
    ff_add(FP, &t0, &t0, &t1);
    ff_inv(FP, &t1, &t0);
    ff_mul(FP, &r->c0, &a->c0, &t1);
    ff_mul(FP, &t0, &a->c1, &t1);
    ff_neg(FP, &r->c1, &t0);
}
static int f2_is_zero(const f2_t *a) { return ff_is_zero(FP, &a->c0) && ff_is_zero(FP, &a->c1); }
static int f2_eq(const f2_t *a, const f2_t *b) { return ff_eq(FP, &a->c0, &b->c0) && ff_eq(FP, &a->c1, &b->c1); }
static void f2_zero(f2_t *r) { ff_set_zero(&r->c0); ff_set_zero(&r->c1); }
static void f2_one(f2_t *r) { ff_set_one(FP, &r->c0); ff_set_zero(&r->c1); }
static void f2_load(f2_t *r, const uint64_t *p)
{   f2_zero(r); memcpy(r->c0.l, p, 48); memcpy(r->c1.l, p + 6, 48);   }
static void f2_store(uint64_t *p, const f2_t *a) { memcpy(p, a->c0.l, 48); memcpy(p + 6, a->c1.l, 48); }

static void f2_from_words(f2_t *r, const uint64_t *c0, const uint64_t *c1)
{
    ff_t t;
    ff_set_zero(&t); memcpy(t.l, c0, 48); ff_to_mont(FP, &r->c0, &t);
    ff_set_zero(&t); memcpy(t.l, c1, 48); ff_to_mont(FP, &r->c1, &t);
}

/* generator of G2 (IETF pairing-friendly-curves draft, sec. 4.2.1), little-endian limbs */
static const uint64_t G2X0[6] = {0xd48056c8c121bdb8ULL, 0x0bac0326a805bbefULL, 0xb4510b647ae3d177ULL,
                                 0xc6e47ad4fa403b02ULL, 0x260805272dc51051ULL, 0x024aa2b2f08f0a91ULL};
static const uint64_t G2X1[6] = {0xe5ac7d055d042b7eULL, 0x334cf11213945d57ULL, 0xb5da61bbdc7f5049ULL,
                                 0x596bd0d09920b61aULL, 0x7dacd3a088274f65ULL, 0x13e02b6052719f60ULL};
static const uint64_t G2Y0[6] = {0xe193548608b82801ULL, 0x923ac9cc3baca289ULL, 0x6d429a695160d12cULL,
                                 0xadfd9baa8cbdd3a7ULL, 0x8cc9cdc6da2e351aULL, 0x0ce5d527727d6e11ULL};
static const uint64_t G2Y1[6] = {0xaaa9075ff05f79beULL, 0x3f370d275cec1da1ULL, 0x267492ab572e99abULL,
                                 0xcb3e287e85a763afULL, 0x32acd2b02bc28b99ULL, 0x0606c4a02ea734ccULL};

static void g2_generator(g2_aff *g)
{
    f2_from_words(&g->X, G2X0, G2X1);
    f2_from_words(&g->Y, G2Y0, G2Y1);
}

static int aff_is_inf(const g2_aff *p) { return f2_is_zero(&p->X) && f2_is_zero(&p->Y); }
static void jac_inf(g2_jac *p) { memset(p, 0, sizeof(*p)); }
static int jac_is_inf(const g2_jac *p) { return f2_is_zero(&p->Z); }
static void jac_from_affine(g2_jac *r, const g2_aff *a)
{
    if (aff_is_inf(a)) { jac_inf(r); return; }
    r->X = a->X; r->Y = a->Y; f2_one(&r->Z);
}

/* dbl-2009-l (a = 0), the formula of ec/jacobian_t.hpp:355-392 */
static void jac_dbl(g2_jac *p)
{
    if (jac_is_inf(p)) return;
    f2_t A, B, C, D, E, F, t;
    f2_sqr(&A, &p->X);
    f2_sqr(&B, &p->Y);
    f2_sqr(&C, &B);
    f2_add(&t, &p->X, &B); f2_sqr(&t, &t); f2_sub(&t, &t, &A); f2_sub(&t, &t, &C); f2_dbl(&D, &t);
    f2_dbl(&E, &A); f2_add(&E, &E, &A);
    f2_sqr(&F, &E);
    f2_mul(&t, &p->Y, &p->Z); f2_dbl(&p->Z, &t);
    f2_dbl(&t, &D); f2_sub(&p->X, &F, &t);
    f2_sub(&t, &D, &p->X); f2_mul(&t, &E, &t);
    f2_dbl(&C, &C); f2_dbl(&C, &C); f2_dbl(&C, &C);
    f2_sub(&p->Y, &t, &C);
}

/* add-2007-bl with the doubling / cancellation cases of ec/jacobian_t.hpp:397-482 */
static void jac_add(g2_jac *p1, const g2_jac *p2)
{
    if (jac_is_inf(p2)) return;
    if (jac_is_inf(p1)) { *p1 = *p2; return; }
    f2_t Z1Z1, Z2Z2, U1, U2, S1, S2, H, I, J, r, V, t;
    f2_sqr(&Z1Z1, &p1->Z);
    f2_sqr(&Z2Z2, &p2->Z);
    f2_mul(&U1, &p1->X, &Z2Z2);
    f2_mul(&U2, &p2->X, &Z1Z1);
    f2_mul(&S1, &p1->Y, &p2->Z); f2_mul(&S1, &S1, &Z2Z2);
    f2_mul(&S2, &p2->Y, &p1->Z); f2_mul(&S2, &S2, &Z1Z1);
    f2_sub(&H, &U2, &U1);
    f2_sub(&r, &S2, &S1);
    if (f2_is_zero(&H)) {
        if (f2_is_zero(&r)) jac_dbl(p1);
        else jac_inf(p1);
        return;
    }
    f2_dbl(&r, &r);
    f2_dbl(&I, &H); f2_sqr(&I, &I);
    f2_mul(&J, &H, &I);
    f2_mul(&V, &U1, &I);
    f2_add(&t, &p1->Z, &p2->Z); f2_sqr(&t, &t); f2_sub(&t, &t, &Z1Z1); f2_sub(&t, &t, &Z2Z2);
    f2_mul(&p1->Z, &t, &H);
    f2_sqr(&t, &r); f2_sub(&t, &t, &J); f2_sub(&t, &t, &V); f2_sub(&p1->X, &t, &V);
    f2_sub(&t, &V, &p1->X); f2_mul(&t, &r, &t);
    f2_mul(&S1, &S1, &J); f2_dbl(&S1, &S1);
    f2_sub(&p1->Y, &t, &S1);
}

static void jac_to_affine(g2_aff *r, const g2_jac *p)
{
    if (jac_is_inf(p)) { memset(r, 0, sizeof(*r)); return; }
    f2_t zi, zi2;
    f2_inv(&zi, &p->Z);
    f2_sqr(&zi2, &zi);
    f2_mul(&r->X, &p->X, &zi2);
    f2_mul(&zi2, &zi2, &zi);
    f2_mul(&r->Y, &p->Y, &zi2);
}

static void jac_mul(g2_jac *r, const g2_aff *p, const unsigned char *scalar, size_t nbits)
{
    g2_jac base;
    jac_from_affine(&base, p);
    jac_inf(r);
    for (size_t i = nbits; i--;) {
        jac_dbl(r);
        if ((scalar[i / 8] >> (i % 8)) & 1) jac_add(r, &base);
    }
}

static void unpack(g2_aff *out, const void *points, size_t stride, int has_flag, size_t n)
{
    for (size_t i = 0; i < n; i++) {
        const unsigned char *src = (const unsigned char *)points + i * stride;
        uint64_t w[24];
        memcpy(w, src, 192);
        f2_load(&out[i].X, w);
        f2_load(&out[i].Y, w + 12);
        if (has_flag && src[192]) memset(&out[i], 0, sizeof(out[i]));
    }
}

static unsigned get_window(const unsigned char *s, unsigned off, unsigned bits)
{
    unsigned v = 0;
    for (unsigned i = 0; i < bits && off + i < 256; i++)
        v |= (unsigned)((s[(off + i) / 8] >> ((off + i) % 8)) & 1) << i;
    return v;
}

static void msm_buckets(g2_jac *ret, const g2_aff *pts, size_t n, const unsigned char *scalars)
{
    unsigned c = n < 64 ? 4 : n < 1024 ? 8 : n < 32768 ? 10 : 12;
    size_t nb = (size_t)1 << c;
    g2_jac *bk = malloc(nb * sizeof(g2_jac));
    jac_inf(ret);
    for (int w = (int)((256 + c - 1) / c) - 1; w >= 0; w--) {
        for (unsigned k = 0; k < c; k++) jac_dbl(ret);
        for (size_t b = 0; b < nb; b++) jac_inf(&bk[b]);
        for (size_t i = 0; i < n; i++) {
            unsigned d = get_window(scalars + 32 * i, (unsigned)w * c, c);
            if (d) {
                g2_jac t;
                jac_from_affine(&t, &pts[i]);
                jac_add(&bk[d], &t);
            }
        }
        g2_jac run, acc;
        jac_inf(&run);
        jac_inf(&acc);
        for (size_t b = nb - 1; b >= 1; b--) {
            jac_add(&run, &bk[b]);
            jac_add(&acc, &run);
        }
        jac_add(ret, &acc);
    }
    free(bk);
}

void oracle_fp2_op(int op, uint64_t *r, const uint64_t *a, const uint64_t *b)
{
    f2_t x, y, z;
    f2_load(&x, a);
    if (b) f2_load(&y, b); else f2_zero(&y);
    switch (op) {
    case 0: f2_mul(&z, &x, &y); break;
    case 1: f2_add(&z, &x, &y); break;
    case 2: f2_sub(&z, &x, &y); break;
    case 3: f2_sqr(&z, &x); break;
    default: f2_inv(&z, &x); break;
    }
    f2_store(r, &z);
}

void oracle_g2_points(uint64_t *out, size_t n)
{
    g2_aff g, a;
    g2_jac acc, gj;
    g2_generator(&g);
    jac_from_affine(&gj, &g);
    jac_inf(&acc);
    for (size_t i = 0; i < n; i++) {
        jac_add(&acc, &gj);
        jac_to_affine(&a, &acc);
        f2_store(out + 24 * i, &a.X);
        f2_store(out + 24 * i + 12, &a.Y);
    }
}

void oracle_g2_msm(int algo, uint64_t *out_jac, const void *points, size_t stride, int has_flag,
                   size_t npoints, const unsigned char *scalars)
{
    g2_aff *pts = malloc((npoints ? npoints : 1) * sizeof(g2_aff));
    unpack(pts, points, stride, has_flag, npoints);
    g2_jac r;
    if (algo == 0) {
        jac_inf(&r);
        for (size_t i = 0; i < npoints; i++) {
            g2_jac t;
            jac_mul(&t, &pts[i], scalars + 32 * i, 256);
            jac_add(&r, &t);
        }
    } else {
        msm_buckets(&r, pts, npoints, scalars);
    }
    free(pts);
    f2_store(out_jac, &r.X);
    f2_store(out_jac + 12, &r.Y);
    f2_store(out_jac + 24, &r.Z);
}

void oracle_g2_jac_to_affine(uint64_t *out_xy, const uint64_t *jac)
{
    g2_jac p;
    g2_aff a;
    f2_load(&p.X, jac); f2_load(&p.Y, jac + 12); f2_load(&p.Z, jac + 24);
    jac_to_affine(&a, &p);
    f2_store(out_xy, &a.X);
    f2_store(out_xy + 12, &a.Y);
}

int oracle_g2_on_curve(const uint64_t *xy)
{
    g2_aff p;
    f2_load(&p.X, xy); f2_load(&p.Y, xy + 12);
    if (aff_is_inf(&p)) return 1;
    f2_t l, r, b;
    static const uint64_t four[6] = {4, 0, 0, 0, 0, 0};
    f2_from_words(&b, four, four);                 /* b' = 4 + 4u */
    f2_sqr(&l, &p.Y);
    f2_sqr(&r, &p.X); f2_mul(&r, &r, &p.X); f2_add(&r, &r, &b);
    return f2_eq(&l, &r);
}

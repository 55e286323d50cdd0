/*
 * oracle/ec.c -- TEST INFRASTRUCTURE ONLY (CPU oracle).  See ec.h.
 *
 * Formulae are the EFD ones the reference cites (shortw-xyzz madd/add/dbl-2008-s,
 * shortw-jacobian-0 dbl-2009-l / add-2007-bl), with the same special-case
 * behaviour as the reference: either operand at infinity, P+P falls through to
 * doubling, P+(-P) gives infinity.
 */
#include "ec.h"
#include <string.h>

#define F (c->fp)

static void set_small(const ec_curve *c, ff_t *r, uint64_t v)
{
    ff_t t;
    ff_set_zero(&t);
    t.l[0] = v;
    ff_to_mont(F, r, &t);
}

static void curve_init(ec_curve *c, const ff_ctx *fp, const ff_ctx *fr, uint64_t b,
                       const uint64_t *gx, const uint64_t *gy)
{
    c->fp = fp;
    c->fr = fr;
    set_small(c, &c->b, b);
    ff_t t;
    ff_set_zero(&t);
    memcpy(t.l, gx, sizeof(uint64_t) * fp->n);
    ff_to_mont(fp, &c->gx, &t);
    ff_set_zero(&t);
    memcpy(t.l, gy, sizeof(uint64_t) * fp->n);
    ff_to_mont(fp, &c->gy, &t);
}

const ec_curve *ec_bls12_381_g1(void)
{
    static ec_curve c;
    static int ready;
    if (!ready) {
        /* standard G1 generator (IETF pairing-friendly-curves draft, sec. 4.2.1) */
        static const uint64_t gx[6] = {0xfb3af00adb22c6bbULL, 0x6c55e83ff97a1aefULL,
                                       0xa14e3a3f171bac58ULL, 0xc3688c4f9774b905ULL,
                                       0x2695638c4fa9ac0fULL, 0x17f1d3a73197d794ULL};
        static const uint64_t gy[6] = {0x0caa232946c5e7e1ULL, 0xd03cc744a2888ae4ULL,
                                       0x00db18cb2c04b3edULL, 0xfcf5e095d5d00af6ULL,
                                       0xa09e30ed741d8ae4ULL, 0x08b3f481e3aaa0f1ULL};
        curve_init(&c, ff_bls12_381_fp(), ff_bls12_381_fr(), 4, gx, gy);
        __atomic_store_n(&ready, 1, __ATOMIC_RELEASE);
    }
    return &c;
}

static void pasta_init(ec_curve *c, const ff_ctx *fp, const ff_ctx *fr)
{
    /* y^2 = x^3 + 5, generator (-1, 2) */
    uint64_t gx[4], gy[4] = {2, 0, 0, 0};
    memcpy(gx, fp->p, sizeof(gx));
    gx[0] -= 1;
    curve_init(c, fp, fr, 5, gx, gy);
}

const ec_curve *ec_pallas(void)
{
    static ec_curve c;
    static int ready;
    if (!ready) {
        pasta_init(&c, ff_pallas_fp(), ff_vesta_fp());
        __atomic_store_n(&ready, 1, __ATOMIC_RELEASE);
    }
    return &c;
}

const ec_curve *ec_vesta(void)
{
    static ec_curve c;
    static int ready;
    if (!ready) {
        pasta_init(&c, ff_vesta_fp(), ff_pallas_fp());
        __atomic_store_n(&ready, 1, __ATOMIC_RELEASE);
    }
    return &c;
}

int ec_affine_is_inf(const ec_curve *c, const ec_affine *p)
{
    return ff_is_zero(F, &p->X) & ff_is_zero(F, &p->Y);
}

int ec_affine_on_curve(const ec_curve *c, const ec_affine *p)
{
    if (ec_affine_is_inf(c, p))
        return 1;
    ff_t l, r;
    ff_sqr(F, &l, &p->Y);
    ff_sqr(F, &r, &p->X);
    ff_mul(F, &r, &r, &p->X);
    ff_add(F, &r, &r, &c->b);
    return ff_eq(F, &l, &r);
}

/* ---------------------------------------------------------------- xyzz -- */

void ec_xyzz_inf(ec_xyzz *p) { memset(p, 0, sizeof(*p)); }

int ec_xyzz_is_inf(const ec_curve *c, const ec_xyzz *p)
{
    return ff_is_zero(F, &p->ZZZ) & ff_is_zero(F, &p->ZZ);
}

void ec_xyzz_from_affine(const ec_curve *c, ec_xyzz *r, const ec_affine *a)
{
    r->X = a->X;
    r->Y = a->Y;
    if (ec_affine_is_inf(c, a)) {
        ff_set_zero(&r->ZZZ);
        ff_set_zero(&r->ZZ);
    } else {
        ff_set_one(F, &r->ZZZ);
        ff_set_one(F, &r->ZZ);
    }
}

void ec_xyzz_madd(const ec_curve *c, ec_xyzz *p1, const ec_affine *p2, int subtract)
{
    if (ec_affine_is_inf(c, p2))
        return;
    if (ec_xyzz_is_inf(c, p1)) {
        ec_xyzz_from_affine(c, p1, p2);
        if (subtract)
            ff_neg(F, &p1->ZZZ, &p1->ZZZ);     /* -P == (X, Y, -ZZZ, ZZ) */
        return;
    }

    ff_t P, R, PP, PPP, Q, t;

    ff_mul(F, &R, &p2->Y, &p1->ZZZ);           /* S2 = Y2*ZZZ1 */
    if (subtract)
        ff_neg(F, &R, &R);
    ff_sub(F, &R, &R, &p1->Y);                 /* R = S2 - Y1  */
    ff_mul(F, &P, &p2->X, &p1->ZZ);            /* U2 = X2*ZZ1  */
    ff_sub(F, &P, &P, &p1->X);                 /* P = U2 - X1  */

    if (!ff_is_zero(F, &P)) {
        ff_sqr(F, &PP, &P);
        ff_mul(F, &PPP, &P, &PP);
        ff_mul(F, &p1->ZZ, &p1->ZZ, &PP);
        ff_mul(F, &p1->ZZZ, &p1->ZZZ, &PPP);
        ff_mul(F, &Q, &p1->X, &PP);
        ff_sqr(F, &t, &R);
        ff_sub(F, &t, &t, &PPP);
        ff_sub(F, &t, &t, &Q);
        ff_sub(F, &t, &t, &Q);                 /* X3 = R^2 - PPP - 2Q */
        p1->X = t;
        ff_sub(F, &Q, &Q, &t);
        ff_mul(F, &Q, &Q, &R);                 /* R*(Q - X3)          */
        ff_mul(F, &t, &p1->Y, &PPP);
        ff_sub(F, &p1->Y, &Q, &t);             /* Y3 = R*(Q-X3) - Y1*PPP */
    } else if (ff_is_zero(F, &R)) {
        /* same point: double the affine operand (mdbl-2008-s-1) */
        ff_t U, V, W, S, M;
        ff_add(F, &U, &p2->Y, &p2->Y);
        ff_sqr(F, &V, &U);
        ff_mul(F, &W, &U, &V);
        ff_mul(F, &S, &p2->X, &V);
        ff_sqr(F, &M, &p2->X);
        ff_add(F, &t, &M, &M);
        ff_add(F, &M, &t, &M);                 /* M = 3*X^2 (a = 0)   */
        ff_sqr(F, &t, &M);
        ff_sub(F, &t, &t, &S);
        ff_sub(F, &t, &t, &S);                 /* X3 = M^2 - 2S       */
        p1->X = t;
        ff_sub(F, &S, &S, &t);
        ff_mul(F, &S, &S, &M);
        ff_mul(F, &t, &W, &p2->Y);
        ff_sub(F, &p1->Y, &S, &t);             /* Y3 = M*(S-X3) - W*Y */
        p1->ZZ = V;
        p1->ZZZ = W;
        if (subtract)
            ff_neg(F, &p1->ZZZ, &p1->ZZZ);
    } else {
        ec_xyzz_inf(p1);
    }
}

void ec_xyzz_add(const ec_curve *c, ec_xyzz *p1, const ec_xyzz *p2)
{
    if (ec_xyzz_is_inf(c, p2))
        return;
    if (ec_xyzz_is_inf(c, p1)) {
        *p1 = *p2;
        return;
    }

    ff_t U, S, P, R, PP, PPP, Q, t;

    ff_mul(F, &U, &p1->X, &p2->ZZ);            /* U1 */
    ff_mul(F, &S, &p1->Y, &p2->ZZZ);           /* S1 */
    ff_mul(F, &P, &p2->X, &p1->ZZ);            /* U2 */
    ff_mul(F, &R, &p2->Y, &p1->ZZZ);           /* S2 */
    ff_sub(F, &P, &P, &U);
    ff_sub(F, &R, &R, &S);

    if (!ff_is_zero(F, &P)) {
        ff_sqr(F, &PP, &P);
        ff_mul(F, &PPP, &P, &PP);
        ff_mul(F, &p1->ZZ, &p1->ZZ, &PP);
        ff_mul(F, &p1->ZZZ, &p1->ZZZ, &PPP);
        ff_mul(F, &Q, &U, &PP);
        ff_sqr(F, &t, &R);
        ff_sub(F, &t, &t, &PPP);
        ff_sub(F, &t, &t, &Q);
        ff_sub(F, &t, &t, &Q);
        p1->X = t;
        ff_sub(F, &Q, &Q, &t);
        ff_mul(F, &Q, &Q, &R);
        ff_mul(F, &t, &S, &PPP);
        ff_sub(F, &p1->Y, &Q, &t);
        ff_mul(F, &p1->ZZ, &p1->ZZ, &p2->ZZ);
        ff_mul(F, &p1->ZZZ, &p1->ZZZ, &p2->ZZZ);
    } else if (ff_is_zero(F, &R)) {
        /* same point: double p1 (dbl-2008-s-1, a = 0) */
        ff_t V, W, M;
        ff_add(F, &U, &p1->Y, &p1->Y);
        ff_sqr(F, &V, &U);
        ff_mul(F, &W, &U, &V);
        ff_mul(F, &S, &p1->X, &V);
        ff_sqr(F, &M, &p1->X);
        ff_add(F, &t, &M, &M);
        ff_add(F, &M, &t, &M);
        ff_sqr(F, &t, &M);
        ff_sub(F, &t, &t, &S);
        ff_sub(F, &t, &t, &S);
        p1->X = t;
        ff_mul(F, &Q, &W, &p1->Y);
        ff_sub(F, &S, &S, &t);
        ff_mul(F, &S, &S, &M);
        ff_sub(F, &p1->Y, &S, &Q);
        ff_mul(F, &p1->ZZ, &p1->ZZ, &V);
        ff_mul(F, &p1->ZZZ, &p1->ZZZ, &W);
    } else {
        ec_xyzz_inf(p1);
    }
}

void ec_xyzz_to_jac(const ec_curve *c, ec_jac *r, const ec_xyzz *p)
{
    /* (X*ZZ, Y*ZZZ, ZZ): Z := ZZ gives x = X*ZZ/ZZ^2, y = Y*ZZZ/ZZ^3 */
    ff_mul(F, &r->X, &p->X, &p->ZZ);
    ff_mul(F, &r->Y, &p->Y, &p->ZZZ);
    r->Z = p->ZZ;
}

/* ------------------------------------------------------------ jacobian -- */

void ec_jac_inf(ec_jac *p) { memset(p, 0, sizeof(*p)); }

int ec_jac_is_inf(const ec_curve *c, const ec_jac *p) { return ff_is_zero(F, &p->Z); }

void ec_jac_from_affine(const ec_curve *c, ec_jac *r, const ec_affine *a)
{
    r->X = a->X;
    r->Y = a->Y;
    if (ec_affine_is_inf(c, a))
        ff_set_zero(&r->Z);
    else
        ff_set_one(F, &r->Z);
}

void ec_jac_dbl(const ec_curve *c, ec_jac *p)
{
    ff_t A, B, C, D, E, Fq, t;
    ff_sqr(F, &A, &p->X);
    ff_sqr(F, &B, &p->Y);
    ff_sqr(F, &C, &B);
    ff_add(F, &D, &p->X, &B);
    ff_sqr(F, &D, &D);
    ff_sub(F, &D, &D, &A);
    ff_sub(F, &D, &D, &C);
    ff_add(F, &D, &D, &D);                     /* D = 2*((X+B)^2 - A - C) */
    ff_add(F, &E, &A, &A);
    ff_add(F, &E, &E, &A);                     /* E = 3*A                 */
    ff_sqr(F, &Fq, &E);
    ff_mul(F, &t, &p->Z, &p->Y);
    ff_add(F, &p->Z, &t, &t);                  /* Z3 = 2*Y*Z              */
    ff_sub(F, &Fq, &Fq, &D);
    ff_sub(F, &p->X, &Fq, &D);                 /* X3 = E^2 - 2D           */
    ff_add(F, &C, &C, &C);
    ff_add(F, &C, &C, &C);
    ff_add(F, &C, &C, &C);                     /* 8*C                     */
    ff_sub(F, &t, &D, &p->X);
    ff_mul(F, &t, &t, &E);
    ff_sub(F, &p->Y, &t, &C);
}

void ec_jac_add(const ec_curve *c, ec_jac *p1, const ec_jac *p2)
{
    if (ec_jac_is_inf(c, p2))
        return;
    if (ec_jac_is_inf(c, p1)) {
        *p1 = *p2;
        return;
    }
    ff_t Z1Z1, Z2Z2, U1, U2, S1, S2, H, r, I, J, V, t;
    ff_sqr(F, &Z1Z1, &p1->Z);
    ff_sqr(F, &Z2Z2, &p2->Z);
    ff_mul(F, &U1, &p1->X, &Z2Z2);
    ff_mul(F, &U2, &p2->X, &Z1Z1);
    ff_mul(F, &S1, &Z2Z2, &p2->Z);
    ff_mul(F, &S1, &S1, &p1->Y);
    ff_mul(F, &S2, &Z1Z1, &p1->Z);
    ff_mul(F, &S2, &S2, &p2->Y);
    ff_sub(F, &H, &U2, &U1);
    ff_sub(F, &r, &S2, &S1);

    if (ff_is_zero(F, &H) && ff_is_zero(F, &r)) {
        ec_jac_dbl(c, p1);
        return;
    }
    /* add-2007-bl; H == 0 with r != 0 yields Z3 = 0 (infinity) as it must */
    ff_add(F, &I, &H, &H);
    ff_sqr(F, &I, &I);
    ff_mul(F, &J, &H, &I);
    ff_add(F, &r, &r, &r);
    ff_mul(F, &V, &U1, &I);
    ff_t X3, Y3, Z3;
    ff_sqr(F, &X3, &r);
    ff_sub(F, &X3, &X3, &J);
    ff_sub(F, &X3, &X3, &V);
    ff_sub(F, &X3, &X3, &V);
    ff_sub(F, &Y3, &V, &X3);
    ff_mul(F, &Y3, &Y3, &r);
    ff_mul(F, &t, &S1, &J);
    ff_sub(F, &Y3, &Y3, &t);
    ff_sub(F, &Y3, &Y3, &t);
    ff_add(F, &Z3, &p1->Z, &p2->Z);
    ff_sqr(F, &Z3, &Z3);
    ff_sub(F, &Z3, &Z3, &Z1Z1);
    ff_sub(F, &Z3, &Z3, &Z2Z2);
    ff_mul(F, &Z3, &Z3, &H);
    p1->X = X3;
    p1->Y = Y3;
    p1->Z = Z3;
}

void ec_jac_to_affine(const ec_curve *c, ec_affine *r, const ec_jac *p)
{
    if (ec_jac_is_inf(c, p)) {
        memset(r, 0, sizeof(*r));
        return;
    }
    ff_t zi, zi2, zi3;
    ff_inv(F, &zi, &p->Z);
    ff_sqr(F, &zi2, &zi);
    ff_mul(F, &zi3, &zi2, &zi);
    ff_mul(F, &r->X, &p->X, &zi2);
    ff_mul(F, &r->Y, &p->Y, &zi3);
}

int ec_jac_eq(const ec_curve *c, const ec_jac *a, const ec_jac *b)
{
    int ia = ec_jac_is_inf(c, a), ib = ec_jac_is_inf(c, b);
    if (ia || ib)
        return ia == ib;
    ff_t za2, zb2, za3, zb3, l, r;
    ff_sqr(F, &za2, &a->Z);
    ff_sqr(F, &zb2, &b->Z);
    ff_mul(F, &l, &a->X, &zb2);
    ff_mul(F, &r, &b->X, &za2);
    if (!ff_eq(F, &l, &r))
        return 0;
    ff_mul(F, &za3, &za2, &a->Z);
    ff_mul(F, &zb3, &zb2, &b->Z);
    ff_mul(F, &l, &a->Y, &zb3);
    ff_mul(F, &r, &b->Y, &za3);
    return ff_eq(F, &l, &r);
}

void ec_jac_mul(const ec_curve *c, ec_jac *r, const ec_affine *p,
                const unsigned char *scalar, size_t nbits)
{
    ec_jac acc, base;
    ec_jac_inf(&acc);
    ec_jac_from_affine(c, &base, p);
    for (size_t i = nbits; i--;) {
        ec_jac_dbl(c, &acc);
        if ((scalar[i / 8] >> (i % 8)) & 1)
            ec_jac_add(c, &acc, &base);
    }
    *r = acc;
}

/* BN254 (alt_bn128): y^2 = x^3 + 3, generator (1, 2) */
const ec_curve *ec_bn254_g1(void)
{
    static ec_curve c;
    static int ready;
    if (!ready) {
        static const uint64_t gx[4] = {1, 0, 0, 0}, gy[4] = {2, 0, 0, 0};
        curve_init(&c, ff_bn254_fp(), ff_bn254_fr(), 3, gx, gy);
        __atomic_store_n(&ready, 1, __ATOMIC_RELEASE);
    }
    return &c;
}

/* BLS12-377 G1: y^2 = x^3 + 1, the arkworks generator */
const ec_curve *ec_bls12_377_g1(void)
{
    static ec_curve c;
    static int ready;
    if (!ready) {
        static const uint64_t gx[6] = {0xeab9b16eb21be9efULL, 0xd5481512ffcd394eULL, 0x188282c8bd37cb5cULL,
                                       0x85951e2caa9d41bbULL, 0xc8fc6225bf87ff54ULL, 0x008848defe740a67ULL};
        static const uint64_t gy[6] = {0xfd82de55559c8ea6ULL, 0xc2fe3d3634a9591aULL, 0x6d182ad44fb82305ULL,
                                       0xbd7fb348ca3e52d9ULL, 0x1f674f5d30afeec4ULL, 0x01914a69c5102effULL};
        curve_init(&c, ff_bls12_377_fp(), ff_bls12_377_fr(), 1, gx, gy);
        __atomic_store_n(&ready, 1, __ATOMIC_RELEASE);
    }
    return &c;
}

/*
 * oracle/ec.h -- TEST INFRASTRUCTURE ONLY (CPU oracle).
 *
 * Short-Weierstrass (a = 0) point arithmetic restating the reference's
 *   ec/affine_t.hpp:19-72     Affine_t   {X,Y}, infinity = X==Y==0
 *   ec/xyzz_t.hpp:16-17,94-101 xyzz_t    {X,Y,ZZZ,ZZ}, infinity = ZZZ==ZZ==0
 *   ec/jacobian_t.hpp:16-58   jacobian_t {X,Y,Z}, infinity = Z==0
 * All coordinates are Montgomery residues (ff.h).
 */
#ifndef ORACLE_EC_H
#define ORACLE_EC_H
#include "ff.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    const ff_ctx *fp;   /* base field                     */
    const ff_ctx *fr;   /* scalar field (group order)     */
    ff_t b;             /* y^2 = x^3 + b, Montgomery form */
    ff_t gx, gy;        /* a generator, Montgomery form   */
} ec_curve;

typedef struct { ff_t X, Y; } ec_affine;
typedef struct { ff_t X, Y, ZZZ, ZZ; } ec_xyzz;   /* field order as in xyzz_t.hpp:17 */
typedef struct { ff_t X, Y, Z; } ec_jac;

const ec_curve *ec_bls12_381_g1(void);
const ec_curve *ec_pallas(void);
const ec_curve *ec_vesta(void);
const ec_curve *ec_bn254_g1(void);
const ec_curve *ec_bls12_377_g1(void);

int  ec_affine_is_inf(const ec_curve *c, const ec_affine *p);
int  ec_affine_on_curve(const ec_curve *c, const ec_affine *p);

void ec_xyzz_inf(ec_xyzz *p);
int  ec_xyzz_is_inf(const ec_curve *c, const ec_xyzz *p);
void ec_xyzz_from_affine(const ec_curve *c, ec_xyzz *r, const ec_affine *a);
/* xyzz_t::add(affine, subtract)  -- ec/xyzz_t.hpp:352-429 */
void ec_xyzz_madd(const ec_curve *c, ec_xyzz *p1, const ec_affine *p2, int subtract);
/* xyzz_t::add(xyzz)              -- ec/xyzz_t.hpp:117-200 */
void ec_xyzz_add(const ec_curve *c, ec_xyzz *p1, const ec_xyzz *p2);
/* xyzz_t -> jacobian_t           -- ec/xyzz_t.hpp:87-90   */
void ec_xyzz_to_jac(const ec_curve *c, ec_jac *r, const ec_xyzz *p);

void ec_jac_inf(ec_jac *p);
int  ec_jac_is_inf(const ec_curve *c, const ec_jac *p);
void ec_jac_from_affine(const ec_curve *c, ec_jac *r, const ec_affine *a);
/* jacobian_t::dbl  -- ec/jacobian_t.hpp:355-392 (dbl-2009-l) */
void ec_jac_dbl(const ec_curve *c, ec_jac *p);
/* jacobian_t::add(jacobian) -- ec/jacobian_t.hpp:397-482 */
void ec_jac_add(const ec_curve *c, ec_jac *p1, const ec_jac *p2);
/* jacobian_t -> Affine_t -- ec/jacobian_t.hpp:31-39; infinity maps to (0,0) */
void ec_jac_to_affine(const ec_curve *c, ec_affine *r, const ec_jac *p);
/* jacobian_t::operator== -- ec/jacobian_t.hpp:563-570 */
int  ec_jac_eq(const ec_curve *c, const ec_jac *a, const ec_jac *b);

/* double-and-add, as the reference's `mult` (msm/pippenger.hpp:192-214);
 * scalar = little-endian bytes, nbits significant */
void ec_jac_mul(const ec_curve *c, ec_jac *r, const ec_affine *p,
                const unsigned char *scalar, size_t nbits);

#ifdef __cplusplus
}
#endif
#endif

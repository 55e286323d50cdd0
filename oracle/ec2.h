/*
 * oracle/ec2.h -- TEST INFRASTRUCTURE ONLY (CPU oracle).
 *
 * BLS12-381 G2: the twist y^2 = x^3 + 4(1 + u) over Fp2 = Fp[u]/(u^2 + 1).  Restates what the
 * reference instantiates for mult_pippenger_fp2_inf (poc/msm-cuda/cuda/pippenger_inf.cu:36-47):
 *   ff/bls12-381-fp2.hpp:158-417   host fp2_t (blst vec384x: c0, c1 consecutive, Montgomery)
 *   ec/jacobian_t.hpp:355-482      dbl / add over that field
 *   msm/pippenger.hpp:192-214      double-and-add `mult`
 * Flat uint64 layouts: Fp2 = 12 limbs (c0, c1), affine = 24, Jacobian = 36.
 */
#ifndef ORACLE_EC2_H
#define ORACLE_EC2_H
#include "ff.h"

#ifdef __cplusplus
extern "C" {
#endif

/* op: 0 mul, 1 add, 2 sub, 3 sqr, 4 inv (b ignored).  Montgomery residues in and out. */
void oracle_fp2_op(int op, uint64_t *r, const uint64_t *a, const uint64_t *b);
/* out[i] = (i+1)*G2, affine, i < n */
void oracle_g2_points(uint64_t *out, size_t n);
/* algo 0: naive sum of double-and-add products; 1: bucket method (unsigned 8..12-bit windows,
 * Jacobian buckets) -- two independent routes to the same point.
 * points: `stride` bytes apart, X (96 B), Y (96 B) and, when has_flag, an infinity flag byte at
 * offset 192 (arkworks G2Affine); X == Y == 0 is infinity as well (ec/affine_t.hpp:56-60).
 * scalars: 32 little-endian bytes each. */
void oracle_g2_msm(int algo, uint64_t *out_jac, const void *points, size_t stride, int has_flag,
                   size_t npoints, const unsigned char *scalars);
void oracle_g2_jac_to_affine(uint64_t *out_xy, const uint64_t *jac);   /* infinity -> (0,0) */
int  oracle_g2_on_curve(const uint64_t *xy);

#ifdef __cplusplus
}
#endif
#endif

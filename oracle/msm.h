/*
 * oracle/msm.h -- TEST INFRASTRUCTURE ONLY (CPU oracle).
 * CPU Pippenger restating msm/pippenger.hpp (see msm.c for line citations).
 */
#ifndef ORACLE_MSM_H
#define ORACLE_MSM_H
#include "ec.h"

#ifdef __cplusplus
extern "C" {
#endif

size_t oracle_msm_window_size(size_t npoints);
void oracle_msm_pippenger_serial(const ec_curve *c, ec_jac *ret, const ec_affine *points,
                                 size_t npoints, const unsigned char *scalars);
void oracle_msm_pippenger(const ec_curve *c, ec_jac *ret, const ec_affine *points,
                          size_t npoints, const unsigned char *scalars, size_t ncpus);
void oracle_msm_naive(const ec_curve *c, ec_jac *ret, const ec_affine *points, size_t npoints,
                      const unsigned char *scalars);
void oracle_gen_points(const ec_curve *c, ec_affine *out, size_t ndistinct);

/* flat-buffer (ctypes) entry points; curve_id 0 = BLS12-381 G1, 1 = Pallas, 2 = Vesta;
 * algo 0 = threaded Pippenger, 1 = serial Pippenger, 2 = naive double-and-add */
int  oracle_msm(int curve_id, uint64_t *out_jac, const void *points, size_t stride_bytes,
                size_t npoints, const unsigned char *scalars, int ncpus, int algo);
void oracle_points(int curve_id, uint64_t *out, size_t ndistinct);
void oracle_jac_to_affine(int curve_id, uint64_t *out_xy, const uint64_t *jac);
int  oracle_affine_on_curve(int curve_id, const uint64_t *xy);
void oracle_ff_op(int field_id, int op, uint64_t *r, const uint64_t *a, const uint64_t *b);
void oracle_ff_consts(int field_id, uint64_t *p, uint64_t *m0, uint64_t *rr, uint64_t *one);

#ifdef __cplusplus
}
#endif
#endif

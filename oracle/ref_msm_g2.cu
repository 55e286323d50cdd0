/*
 * oracle/ref_msm_g2.cu -- TEST INFRASTRUCTURE ONLY.
 *
 * Harness around the reference's OWN templates (msm/pippenger.cuh, ff/bls12-381-fp2.hpp, ec/ *.hpp,
 * included from where they lie under $(REF)) instantiating the G2 MSM on PACKED affine points
 * (xyzz_t<fp2_t>::affine_t, 192 bytes, infinity = X == Y == 0).
 *
 * Why not the reference's mult_pippenger_fp2_inf itself (poc/msm-cuda/cuda/pippenger_inf.cu:41-47,
 * built as _ref/libref_msm_g2_gpu.so)?  Its Affine_inf_t<fp2_t>::mem_t (ec/affine_t.hpp:95-97) is
 * sized from sizeof(field_t), which is 96 bytes in nvcc's host pass (host fp2_t = vec384x) but 48
 * in the device pass (device fp2_t = one fp_mont per lane): the host copies points with a
 * 224-byte pitch, the kernels index them with 208.  Recorded on a B200 (tests/golden/
 * make_golden.py g2): n = 1 is right, n = 2 returns s0*P0, n >= 33 is not even on the curve.
 * The packed affine_t has no such field, so this instantiation is consistent and pins G2.
 */
#include <cuda.h>

#include <ff/bls12-381-fp2.hpp>

#include <ec/jacobian_t.hpp>
#include <ec/xyzz_t.hpp>

typedef jacobian_t<fp_t> point_t;
typedef xyzz_t<fp_t> bucket_t;
typedef bucket_t::affine_t affine_t;
typedef fr_t scalar_t;

#include <msm/pippenger.cuh>

typedef jacobian_t<fp2_t> g2_point_t;
typedef xyzz_t<fp2_t> g2_bucket_t;
typedef g2_bucket_t::affine_t g2_affine_t;

/* not guarded by __CUDA_ARCH__: the device pass must see the call to instantiate the fp2 kernels
 * (the reference's explicit instantiations, msm/pippenger.cuh:299-316, cover bucket_t only) */
extern "C" RustError::by_value ref_mult_pippenger_fp2(g2_point_t* out, const g2_affine_t points[],
                                            size_t npoints, const scalar_t scalars[])
{
    return mult_pippenger<g2_bucket_t>(out, points, npoints, scalars, false);
}

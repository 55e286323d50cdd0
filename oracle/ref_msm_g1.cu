/*
 * oracle/ref_msm_g1.cu -- TEST INFRASTRUCTURE ONLY.
 *
 * The G1 half of the reference's poc/msm-cuda/cuda/pippenger_inf.cu (:20-34) for the crate's
 * bn254 / bls12_377 features, around the reference's OWN templates included from where they lie
 * under $(REF).  The file itself cannot be built for these features with the blst stand-in
 * (oracle/shim/blst_t.hpp): its G2 half needs blst's 256-bit vector API for the host fp2_t, and
 * its G2 entry point has the host/device layout mismatch described in ref_msm_g2.cu anyway.
 */
#include <cuda.h>

#if defined(FEATURE_BN254)
# include <ff/alt_bn128.hpp>
#elif defined(FEATURE_BLS12_377)
# include <ff/bls12-377.hpp>
#else
# error "FEATURE_BN254 or FEATURE_BLS12_377"
#endif

#include <ec/jacobian_t.hpp>
#include <ec/xyzz_t.hpp>

typedef jacobian_t<fp_t> point_t;
typedef xyzz_t<fp_t> bucket_t;
typedef bucket_t::affine_inf_t affine_t;
typedef fr_t scalar_t;

#include <msm/pippenger.cuh>

#ifndef __CUDA_ARCH__
extern "C" RustError::by_value mult_pippenger_inf(point_t* out, const affine_t points[], size_t npoints,
                                                  const scalar_t scalars[], size_t ffi_affine_sz)
{
    return mult_pippenger<bucket_t>(out, points, npoints, scalars, false, ffi_affine_sz);
}
#endif

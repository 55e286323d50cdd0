/*
 * oracle/ref_msm_g1.cu -- TEST INFRASTRUCTURE ONLY.
 *
 * The G1 half of the reference's poc/msm-cuda/cuda/pippenger_inf.cu (:20-34) for the crate's
 * bn254 / bls12_377 features (and for the Pasta curves of ff/pasta.hpp), around the reference's OWN templates included from where they lie
 * under $(REF).  The file itself cannot be built for these features with the blst stand-in
 * (oracle/shim/blst_t.hpp): its G2 half needs blst's 256-bit vector API for the host fp2_t, and
 * its G2 entry point has the host/device layout mismatch described in ref_msm_g2.cu anyway.
 */
#include <cuda.h>

#if defined(FEATURE_BN254)
# include <ff/alt_bn128.hpp>
#elif defined(FEATURE_BLS12_377)
# include <ff/bls12-377.hpp>
#elif defined(FEATURE_PALLAS) || defined(FEATURE_VESTA)
/* the Pasta cycle (ff/pasta.hpp:82-103): the reference has no PoC boundary for these curves, its
 * templates are instantiated here exactly as for the other features; host field = shim/pasta_t.hpp */
# include <ff/pasta.hpp>
#else
# error "FEATURE_BN254, FEATURE_BLS12_377, FEATURE_PALLAS or FEATURE_VESTA"
#endif
#include <ec/jacobian_t.hpp>
#include <ec/xyzz_t.hpp>

/* msm/pippenger.cuh explicitly instantiates its kernels for these four global names */
using bucket_t = xyzz_t<fp_t>;
using point_t = jacobian_t<fp_t>;
using affine_t = bucket_t::affine_inf_t;      /* arkworks rows: x, y, infinity flag */
using scalar_t = fr_t;

#include <msm/pippenger.cuh>

#ifndef __CUDA_ARCH__
/* same name and signature as the reference's entry point, so the golden generator and the
 * probes load either library the same way; scalars are plain integers (mont = false) */
extern "C" RustError::by_value mult_pippenger_inf(point_t* sum, const affine_t* rows, size_t n,
                                                  const scalar_t* k, size_t row_bytes)
{
    const bool scalars_in_montgomery_form = false;
    RustError status = mult_pippenger<bucket_t>(sum, rows, n, k, scalars_in_montgomery_form, row_bytes);
    return status;
}
#endif

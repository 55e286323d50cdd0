// Curve dispatch for the extended MSM entry points (include/sppark_b200.h).
#include "../util/gpu.cuh"

RustError msm_host_bls12_381(void*, const void*, size_t, const void*, size_t, bool);
RustError msm_dev_bls12_381(void*, const void*, size_t, const void*, void*);
RustError msm_host_pallas(void*, const void*, size_t, const void*, size_t, bool);
RustError msm_dev_pallas(void*, const void*, size_t, const void*, void*);
RustError msm_host_vesta(void*, const void*, size_t, const void*, size_t, bool);
RustError msm_dev_vesta(void*, const void*, size_t, const void*, void*);

extern "C" RustError sppark_b200_msm(int curve, void* out, const void* points, size_t npoints,
                                     const void* scalars, size_t ffi_affine_sz)
{
    switch (curve) {
    case SPPARK_CURVE_BLS12_381_G1:
        return msm_host_bls12_381(out, points, npoints, scalars, ffi_affine_sz ? ffi_affine_sz : 96, ffi_affine_sz > 96);
    case SPPARK_CURVE_PALLAS:
        return msm_host_pallas(out, points, npoints, scalars, ffi_affine_sz ? ffi_affine_sz : 64, ffi_affine_sz > 64);
    case SPPARK_CURVE_VESTA:
        return msm_host_vesta(out, points, npoints, scalars, ffi_affine_sz ? ffi_affine_sz : 64, ffi_affine_sz > 64);
    default:
        return rust_err(-(int)cudaErrorInvalidValue, "sppark_b200_msm: unknown curve");
    }
}

extern "C" RustError sppark_b200_msm_dev(int curve, void* out, const void* d_points, size_t npoints,
                                         const void* d_scalars, void* stream)
{
    switch (curve) {
    case SPPARK_CURVE_BLS12_381_G1: return msm_dev_bls12_381(out, d_points, npoints, d_scalars, stream);
    case SPPARK_CURVE_PALLAS: return msm_dev_pallas(out, d_points, npoints, d_scalars, stream);
    case SPPARK_CURVE_VESTA: return msm_dev_vesta(out, d_points, npoints, d_scalars, stream);
    default: return rust_err(-(int)cudaErrorInvalidValue, "sppark_b200_msm_dev: unknown curve");
    }
}

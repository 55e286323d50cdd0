// Curve dispatch for the extended MSM entry points (include/sppark_b200.h).
#include "../util/gpu.cuh"

RustError msm_host_bls12_381(void*, const void*, size_t, const void*, size_t, bool, bool);
RustError msm_dev_bls12_381(void*, const void*, size_t, const void*, void*);
RustError msm_host_pallas(void*, const void*, size_t, const void*, size_t, bool, bool);
RustError msm_dev_pallas(void*, const void*, size_t, const void*, void*);
RustError msm_host_vesta(void*, const void*, size_t, const void*, size_t, bool, bool);
RustError msm_dev_vesta(void*, const void*, size_t, const void*, void*);

RustError msm_host_bls12_381_g2(void*, const void*, size_t, const void*, size_t, bool, bool);
RustError msm_dev_bls12_381_g2(void*, const void*, size_t, const void*, void*);
RustError gen_points_bls12_381_g2(void*, size_t, void*);
RustError combine_bls12_381_g2(void*, const void*, size_t);
RustError msm_host_bn254(void*, const void*, size_t, const void*, size_t, bool, bool);
RustError msm_dev_bn254(void*, const void*, size_t, const void*, void*);
RustError gen_points_bn254(void*, size_t, void*);
RustError combine_bn254(void*, const void*, size_t);
RustError msm_host_bls12_377(void*, const void*, size_t, const void*, size_t, bool, bool);
RustError msm_dev_bls12_377(void*, const void*, size_t, const void*, void*);
RustError gen_points_bls12_377(void*, size_t, void*);
RustError combine_bls12_377(void*, const void*, size_t);
RustError msm_preload_bls12_381(const void*, size_t, size_t, bool, void**);
RustError msm_resident_bls12_381(void*, const void*, size_t, const void*, bool);
RustError msm_preload_bls12_381_g2(const void*, size_t, size_t, bool, void**);
RustError msm_resident_bls12_381_g2(void*, const void*, size_t, const void*, bool);
RustError msm_preload_pallas(const void*, size_t, size_t, bool, void**);
RustError msm_resident_pallas(void*, const void*, size_t, const void*, bool);
RustError msm_preload_vesta(const void*, size_t, size_t, bool, void**);
RustError msm_resident_vesta(void*, const void*, size_t, const void*, bool);
RustError msm_preload_bn254(const void*, size_t, size_t, bool, void**);
RustError msm_resident_bn254(void*, const void*, size_t, const void*, bool);
RustError msm_preload_bls12_377(const void*, size_t, size_t, bool, void**);
RustError msm_resident_bls12_377(void*, const void*, size_t, const void*, bool);
RustError gen_points_bls12_381(void*, size_t, void*);
RustError gen_points_pallas(void*, size_t, void*);
RustError gen_points_vesta(void*, size_t, void*);
RustError combine_bls12_381(void*, const void*, size_t);
RustError combine_pallas(void*, const void*, size_t);
RustError combine_vesta(void*, const void*, size_t);

extern "C" RustError sppark_b200_generate_points_dev(int curve, void* d_out, size_t n, void* stream)
{
    if (n >= (1ull << 31)) return rust_err(-(int)cudaErrorInvalidValue, "generate_points: n too large");
    switch (curve) {
    case SPPARK_CURVE_BLS12_381_G1: return gen_points_bls12_381(d_out, n, stream);
    case SPPARK_CURVE_PALLAS: return gen_points_pallas(d_out, n, stream);
    case SPPARK_CURVE_VESTA: return gen_points_vesta(d_out, n, stream);
    case SPPARK_CURVE_BLS12_381_G2: return gen_points_bls12_381_g2(d_out, n, stream);
    case SPPARK_CURVE_BN254_G1: return gen_points_bn254(d_out, n, stream);
    case SPPARK_CURVE_BLS12_377_G1: return gen_points_bls12_377(d_out, n, stream);
    default: return rust_err(-(int)cudaErrorInvalidValue, "generate_points: unknown curve");
    }
}

extern "C" RustError sppark_b200_msm_combine(int curve, void* out, const void* partials, size_t count)
{
    switch (curve) {
    case SPPARK_CURVE_BLS12_381_G1: return combine_bls12_381(out, partials, count);
    case SPPARK_CURVE_PALLAS: return combine_pallas(out, partials, count);
    case SPPARK_CURVE_VESTA: return combine_vesta(out, partials, count);
    case SPPARK_CURVE_BLS12_381_G2: return combine_bls12_381_g2(out, partials, count);
    case SPPARK_CURVE_BN254_G1: return combine_bn254(out, partials, count);
    case SPPARK_CURVE_BLS12_377_G1: return combine_bls12_377(out, partials, count);
    default: return rust_err(-(int)cudaErrorInvalidValue, "msm_combine: unknown curve");
    }
}

static RustError msm_any(int curve, void* out, const void* points, size_t npoints, const void* scalars,
                         size_t ffi_affine_sz, bool mont)
{
    switch (curve) {
    case SPPARK_CURVE_BLS12_381_G1:
        return msm_host_bls12_381(out, points, npoints, scalars, ffi_affine_sz ? ffi_affine_sz : 96, ffi_affine_sz > 96, mont);
    case SPPARK_CURVE_PALLAS:
        return msm_host_pallas(out, points, npoints, scalars, ffi_affine_sz ? ffi_affine_sz : 64, ffi_affine_sz > 64, mont);
    case SPPARK_CURVE_VESTA:
        return msm_host_vesta(out, points, npoints, scalars, ffi_affine_sz ? ffi_affine_sz : 64, ffi_affine_sz > 64, mont);
    case SPPARK_CURVE_BLS12_381_G2:
        return msm_host_bls12_381_g2(out, points, npoints, scalars, ffi_affine_sz ? ffi_affine_sz : 192, ffi_affine_sz > 192, mont);
    case SPPARK_CURVE_BN254_G1:
        return msm_host_bn254(out, points, npoints, scalars, ffi_affine_sz ? ffi_affine_sz : 64, ffi_affine_sz > 64, mont);
    case SPPARK_CURVE_BLS12_377_G1:
        return msm_host_bls12_377(out, points, npoints, scalars, ffi_affine_sz ? ffi_affine_sz : 96, ffi_affine_sz > 96, mont);
    default:
        return rust_err(-(int)cudaErrorInvalidValue, "sppark_b200_msm: unknown curve");
    }
}

extern "C" RustError sppark_b200_msm(int curve, void* out, const void* points, size_t npoints,
                                     const void* scalars, size_t ffi_affine_sz)
{   return msm_any(curve, out, points, npoints, scalars, ffi_affine_sz, false);   }

extern "C" RustError sppark_b200_msm_ex(int curve, void* out, const void* points, size_t npoints,
                                        const void* scalars, size_t ffi_affine_sz, int scalars_mont)
{   return msm_any(curve, out, points, npoints, scalars, ffi_affine_sz, scalars_mont != 0);   }

extern "C" RustError sppark_b200_msm_dev(int curve, void* out, const void* d_points, size_t npoints,
                                         const void* d_scalars, void* stream)
{
    switch (curve) {
    case SPPARK_CURVE_BLS12_381_G1: return msm_dev_bls12_381(out, d_points, npoints, d_scalars, stream);
    case SPPARK_CURVE_PALLAS: return msm_dev_pallas(out, d_points, npoints, d_scalars, stream);
    case SPPARK_CURVE_VESTA: return msm_dev_vesta(out, d_points, npoints, d_scalars, stream);
    case SPPARK_CURVE_BLS12_381_G2: return msm_dev_bls12_381_g2(out, d_points, npoints, d_scalars, stream);
    case SPPARK_CURVE_BN254_G1: return msm_dev_bn254(out, d_points, npoints, d_scalars, stream);
    case SPPARK_CURVE_BLS12_377_G1: return msm_dev_bls12_377(out, d_points, npoints, d_scalars, stream);
    default: return rust_err(-(int)cudaErrorInvalidValue, "sppark_b200_msm_dev: unknown curve");
    }
}

// ---- preloaded points (the reference's msm_t{points, npoints} + invoke(out, scalars),
// msm/pippenger.cuh:377-390,582-601): the SRS stays on the device, only scalars cross PCIe ------
struct sppark_b200_msm_ctx {
    int curve, device;
    void* d_points;
    size_t npoints;
};

extern "C" RustError sppark_b200_msm_ctx_create(int curve, const void* points, size_t npoints,
                                                size_t ffi_affine_sz, sppark_b200_msm_ctx** out)
{
    if (out == nullptr) return rust_err(-(int)cudaErrorInvalidValue, "msm_ctx_create: null output");
    *out = nullptr;
    void* d = nullptr;
    RustError e;
    switch (curve) {
    case SPPARK_CURVE_BLS12_381_G1: e = msm_preload_bls12_381(points, npoints, ffi_affine_sz ? ffi_affine_sz : 96, ffi_affine_sz > 96, &d); break;
    case SPPARK_CURVE_PALLAS: e = msm_preload_pallas(points, npoints, ffi_affine_sz ? ffi_affine_sz : 64, ffi_affine_sz > 64, &d); break;
    case SPPARK_CURVE_VESTA: e = msm_preload_vesta(points, npoints, ffi_affine_sz ? ffi_affine_sz : 64, ffi_affine_sz > 64, &d); break;
    case SPPARK_CURVE_BLS12_381_G2: e = msm_preload_bls12_381_g2(points, npoints, ffi_affine_sz ? ffi_affine_sz : 192, ffi_affine_sz > 192, &d); break;
    case SPPARK_CURVE_BN254_G1: e = msm_preload_bn254(points, npoints, ffi_affine_sz ? ffi_affine_sz : 64, ffi_affine_sz > 64, &d); break;
    case SPPARK_CURVE_BLS12_377_G1: e = msm_preload_bls12_377(points, npoints, ffi_affine_sz ? ffi_affine_sz : 96, ffi_affine_sz > 96, &d); break;
    default: return rust_err(-(int)cudaErrorInvalidValue, "msm_ctx_create: unknown curve");
    }
    if (e.code != 0) return e;
    int dev = 0;
    (void)cudaGetDevice(&dev);
    *out = new sppark_b200_msm_ctx{curve, dev, d, npoints};
    return rust_ok();
}

extern "C" RustError sppark_b200_msm_ctx_invoke(sppark_b200_msm_ctx* ctx, void* out, const void* scalars,
                                                size_t npoints, int scalars_mont)
{
    if (ctx == nullptr || npoints > ctx->npoints)
        return rust_err(-(int)cudaErrorInvalidValue, "msm_ctx_invoke: more scalars than preloaded points");
    int cur = 0;
    (void)cudaGetDevice(&cur);
    if (cur != ctx->device) return rust_err(-(int)cudaErrorInvalidDevice, "msm_ctx_invoke: the points live on another device");
    const bool mont = scalars_mont != 0;
    switch (ctx->curve) {
    case SPPARK_CURVE_BLS12_381_G1: return msm_resident_bls12_381(out, ctx->d_points, npoints, scalars, mont);
    case SPPARK_CURVE_PALLAS: return msm_resident_pallas(out, ctx->d_points, npoints, scalars, mont);
    case SPPARK_CURVE_VESTA: return msm_resident_vesta(out, ctx->d_points, npoints, scalars, mont);
    case SPPARK_CURVE_BLS12_381_G2: return msm_resident_bls12_381_g2(out, ctx->d_points, npoints, scalars, mont);
    case SPPARK_CURVE_BN254_G1: return msm_resident_bn254(out, ctx->d_points, npoints, scalars, mont);
    default: return msm_resident_bls12_377(out, ctx->d_points, npoints, scalars, mont);
    }
}

extern "C" void sppark_b200_msm_ctx_free(sppark_b200_msm_ctx* ctx)
{
    if (ctx == nullptr) return;
    int cur = 0;
    (void)cudaGetDevice(&cur);
    if (cur != ctx->device) (void)cudaSetDevice(ctx->device);
    (void)cudaFree(ctx->d_points);
    if (cur != ctx->device) (void)cudaSetDevice(cur);
    delete ctx;
}

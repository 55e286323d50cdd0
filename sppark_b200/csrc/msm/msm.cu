// Curve dispatch for the extended MSM entry points (include/sppark_b200.h).
#include "../util/gpu.cuh"
#include <thread>
#include <vector>

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

// ---- one MSM sharded by point-chunk over several GPUs of this process --------
// chunk i of the points / scalars runs on device_ids[i] (one host thread per distinct device, the
// chunks of a device one after the other; the host-pointer pipeline of msm_host on each device's
// own PCIe link), the partial results -- one Jacobian point each -- are added on the first device.
// No collective library is needed inside one process: the "all-gather" of the multi-process
// route (sppark_b200/parallel.py over NCCL) is a host array here.
static size_t jacobian_bytes(int curve)
{
    switch (curve) {
    case SPPARK_CURVE_BLS12_381_G1: case SPPARK_CURVE_BLS12_377_G1: return 144;
    case SPPARK_CURVE_BLS12_381_G2: return 288;
    default: return 96;
    }
}
static size_t packed_affine_bytes(int curve)
{
    switch (curve) {
    case SPPARK_CURVE_BLS12_381_G1: case SPPARK_CURVE_BLS12_377_G1: return 96;
    case SPPARK_CURVE_BLS12_381_G2: return 192;
    default: return 64;
    }
}

extern "C" RustError sppark_b200_msm_sharded(int curve, void* out, const void* points, size_t npoints,
                                             const void* scalars, size_t ffi_affine_sz, int scalars_mont,
                                             const int* device_ids, size_t ndev)
{
    if (curve < 0 || curve > SPPARK_CURVE_BLS12_377_G1)
        return rust_err(-(int)cudaErrorInvalidValue, "msm_sharded: unknown curve");
    if (out == nullptr || ndev == 0 || ndev > 64 || device_ids == nullptr)
        return rust_err(-(int)cudaErrorInvalidValue, "msm_sharded: need 1..64 device ids");
    const size_t jb = jacobian_bytes(curve), stride = ffi_affine_sz ? ffi_affine_sz : packed_affine_bytes(curve);
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) return rust_err(-(int)cudaErrorNoDevice, "msm_sharded: no CUDA device");
    for (size_t i = 0; i < ndev; i++)
        if (device_ids[i] < 0 || device_ids[i] >= count)
            return rust_err(-(int)cudaErrorInvalidDevice, "msm_sharded: no such device");
    int home = 0;
    (void)cudaGetDevice(&home);

    std::vector<uint8_t> partials(ndev * jb, 0);
    std::vector<RustError> status(ndev, rust_ok());
    const size_t chunk = (npoints + ndev - 1) / ndev;
    auto run_device = [&](int dev) {
        if (cudaSetDevice(dev) != cudaSuccess) {
            for (size_t i = 0; i < ndev; i++)
                if (device_ids[i] == dev) status[i] = rust_err(-(int)cudaErrorInvalidDevice, "msm_sharded: cudaSetDevice failed");
            return;
        }
        for (size_t i = 0; i < ndev; i++) {
            if (device_ids[i] != dev) continue;
            const size_t first = i * chunk < npoints ? i * chunk : npoints;
            const size_t n = npoints - first < chunk ? npoints - first : chunk;
            status[i] = msm_any(curve, partials.data() + i * jb, (const uint8_t*)points + first * stride, n,
                                (const uint8_t*)scalars + first * 32, ffi_affine_sz, scalars_mont != 0);
        }
    };
    std::vector<int> distinct;
    for (size_t i = 0; i < ndev; i++) {
        bool seen = false;
        for (int d : distinct) seen |= d == device_ids[i];
        if (!seen) distinct.push_back(device_ids[i]);
    }
    std::vector<std::thread> workers;
    for (size_t k = 1; k < distinct.size(); k++) workers.emplace_back(run_device, distinct[k]);
    run_device(distinct[0]);
    for (auto& t : workers) t.join();
    (void)cudaSetDevice(distinct[0]);
    RustError result = rust_ok();
    for (size_t i = 0; i < ndev; i++) {
        if (status[i].code != 0 && result.code == 0) result = status[i];
        else if (status[i].message) free(status[i].message);
    }
    if (result.code == 0) result = sppark_b200_msm_combine(curve, out, partials.data(), ndev);
    (void)cudaSetDevice(home);
    return result;
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

// Curve dispatch for the extended MSM entry points (include/sppark_b200.h).
#include "../util/gpu.cuh"
#include <thread>
#include <vector>

// one row per curve id (SPPARK_CURVE_*): the six entry points its translation unit defines
// (msm_bls12_381.cu, msm_pasta.cu, msm_bn254_bls12_377.cu, msm_*_g2.cu) and its packed layouts
#define CURVE_DECLS(name)                                                                          \
    RustError msm_host_##name(void*, const void*, size_t, const void*, size_t, bool, bool);        \
    RustError msm_dev_##name(void*, const void*, size_t, const void*, void*);                      \
    RustError gen_points_##name(void*, size_t, void*);                                             \
    RustError combine_##name(void*, const void*, size_t);                                          \
    RustError msm_preload_##name(const void*, size_t, size_t, bool, void**);                       \
    RustError msm_resident_##name(void*, const void*, size_t, const void*, bool);
CURVE_DECLS(bls12_381) CURVE_DECLS(pallas) CURVE_DECLS(vesta) CURVE_DECLS(bls12_381_g2)
CURVE_DECLS(bn254) CURVE_DECLS(bls12_377) CURVE_DECLS(bn254_g2) CURVE_DECLS(bls12_377_g2)

struct curve_ops {
    RustError (*host)(void*, const void*, size_t, const void*, size_t, bool, bool);
    RustError (*dev)(void*, const void*, size_t, const void*, void*);
    RustError (*gen)(void*, size_t, void*);
    RustError (*combine)(void*, const void*, size_t);
    RustError (*preload)(const void*, size_t, size_t, bool, void**);
    RustError (*resident)(void*, const void*, size_t, const void*, bool);
    size_t affine_bytes, jacobian_bytes;        // packed {X, Y} and {X, Y, Z}
};
#define CURVE_ROW(name, affine, jac)                                                               \
    {msm_host_##name, msm_dev_##name, gen_points_##name, combine_##name, msm_preload_##name,       \
     msm_resident_##name, affine, jac}
static const curve_ops CURVES[] = {
    CURVE_ROW(bls12_381, 96, 144),        // SPPARK_CURVE_BLS12_381_G1
    CURVE_ROW(pallas, 64, 96),            // SPPARK_CURVE_PALLAS
    CURVE_ROW(vesta, 64, 96),             // SPPARK_CURVE_VESTA
    CURVE_ROW(bls12_381_g2, 192, 288),    // SPPARK_CURVE_BLS12_381_G2
    CURVE_ROW(bn254, 64, 96),             // SPPARK_CURVE_BN254_G1
    CURVE_ROW(bls12_377, 96, 144),        // SPPARK_CURVE_BLS12_377_G1
    CURVE_ROW(bn254_g2, 128, 192),        // SPPARK_CURVE_BN254_G2
    CURVE_ROW(bls12_377_g2, 192, 288),    // SPPARK_CURVE_BLS12_377_G2
};
static_assert(sizeof(CURVES) / sizeof(CURVES[0]) == SPPARK_CURVE_BLS12_377_G2 + 1, "one row per curve id");
static const curve_ops* curve_of(int curve)
{   return curve < 0 || curve > SPPARK_CURVE_BLS12_377_G2 ? nullptr : &CURVES[curve];   }

extern "C" RustError sppark_b200_generate_points_dev(int curve, void* d_out, size_t n, void* stream)
{
    if (n >= (1ull << 31)) return rust_err(-(int)cudaErrorInvalidValue, "generate_points: n too large");
    const curve_ops* c = curve_of(curve);
    if (c == nullptr) return rust_err(-(int)cudaErrorInvalidValue, "generate_points: unknown curve");
    return c->gen(d_out, n, stream);
}

extern "C" RustError sppark_b200_msm_combine(int curve, void* out, const void* partials, size_t count)
{
    const curve_ops* c = curve_of(curve);
    if (c == nullptr) return rust_err(-(int)cudaErrorInvalidValue, "msm_combine: unknown curve");
    return c->combine(out, partials, count);
}

// ffi_affine_sz: 0 = packed {X, Y}; larger than that = arkworks rows with an infinity flag after Y
static RustError msm_any(int curve, void* out, const void* points, size_t npoints, const void* scalars,
                         size_t ffi_affine_sz, bool mont)
{
    const curve_ops* c = curve_of(curve);
    if (c == nullptr) return rust_err(-(int)cudaErrorInvalidValue, "sppark_b200_msm: unknown curve");
    return c->host(out, points, npoints, scalars, ffi_affine_sz ? ffi_affine_sz : c->affine_bytes,
                   ffi_affine_sz > c->affine_bytes, mont);
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
    const curve_ops* c = curve_of(curve);
    if (c == nullptr) return rust_err(-(int)cudaErrorInvalidValue, "sppark_b200_msm_dev: unknown curve");
    return c->dev(out, d_points, npoints, d_scalars, stream);
}

// ---- one MSM sharded by point-chunk over several GPUs of this process --------
// chunk i of the points / scalars runs on device_ids[i] (one host thread per distinct device, the
// chunks of a device one after the other; the host-pointer pipeline of msm_host on each device's
// own PCIe link), the partial results -- one Jacobian point each -- are added on the first device.
// No collective library is needed inside one process: the "all-gather" of the multi-process
// route (sppark_b200/parallel.py over NCCL) is a host array here.
extern "C" RustError sppark_b200_msm_sharded(int curve, void* out, const void* points, size_t npoints,
                                             const void* scalars, size_t ffi_affine_sz, int scalars_mont,
                                             const int* device_ids, size_t ndev)
{
    const curve_ops* c = curve_of(curve);
    if (c == nullptr) return rust_err(-(int)cudaErrorInvalidValue, "msm_sharded: unknown curve");
    if (out == nullptr || ndev == 0 || ndev > 64 || device_ids == nullptr)
        return rust_err(-(int)cudaErrorInvalidValue, "msm_sharded: need 1..64 device ids");
    const size_t jb = c->jacobian_bytes, stride = ffi_affine_sz ? ffi_affine_sz : c->affine_bytes;
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
    const curve_ops* c = curve_of(curve);
    if (c == nullptr) return rust_err(-(int)cudaErrorInvalidValue, "msm_ctx_create: unknown curve");
    RustError e = c->preload(points, npoints, ffi_affine_sz ? ffi_affine_sz : c->affine_bytes,
                             ffi_affine_sz > c->affine_bytes, &d);
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
    return curve_of(ctx->curve)->resident(out, ctx->d_points, npoints, scalars, scalars_mont != 0);
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

// BLS12-381 G1 MSM: the drop-in entry points of poc/msm-cuda (see msm_host.cuh).
#include "msm_host.cuh"

RustError msm_host_bls12_381(void* out, const void* points, size_t npoints, const void* scalars,
                             size_t stride, bool has_flag, bool mont)
{
    return msm_host<ff::bls12_381_fp_t>(out, points, npoints, scalars, stride, has_flag,
                                        mont ? scalars_from_mont<ff::bls12_381_fr_t> : nullptr);
}
RustError msm_dev_bls12_381(void* out, const void* d_points, size_t npoints, const void* d_scalars, void* stream)
{   return msm_dev<ff::bls12_381_fp_t>(out, d_points, npoints, d_scalars, stream);   }

RustError gen_points_bls12_381(void* d_out, size_t n, void* stream)
{   return gen_points_dev<ff::bls12_381_g1_gen>(d_out, n, stream);   }
RustError combine_bls12_381(void* out, const void* partials, size_t count)
{   return combine_host<ff::bls12_381_fp_t>(out, partials, count);   }

extern "C" RustError mult_pippenger(void* out, const void* points, size_t npoints, const void* scalars)
{   return msm_host_bls12_381(out, points, npoints, scalars, 96, false, false);   }

extern "C" RustError mult_pippenger_inf(void* out, const void* points, size_t npoints,
                                        const void* scalars, size_t ffi_affine_sz)
{   return msm_host_bls12_381(out, points, npoints, scalars, ffi_affine_sz, true, false);   }

// ---- device self-test hook: element-wise field ops through the PTX arithmetic -------------
// op 0 mul, 1 add, 2 sub, 3 sqr, 4 mul_shared, 5 sqr_shared, 6 msub_shared(x,y,y,x^2).  Host arrays of n 48-byte elements.  Used by the GPU KAT tests.
template<class F>
__global__ void selftest_kernel(int op, size_t n, uint32_t* r, const uint32_t* a, const uint32_t* b)
{
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    F x, y, z;
    for (int k = 0; k < F::N; k++) { x.l[k] = a[i * F::N + k]; y.l[k] = b[i * F::N + k]; }
    switch (op) {
    case 0: z = x * y; break;
    case 1: z = x + y; break;
    case 2: z = x - y; break;
    case 3: z = x.sqr(); break;
    case 4: z = F::mul_shared(x, y); break;            // Karatsuba wide product + separate reduction
    case 5: z = F::sqr_shared(x); break;               // dedicated squaring
    default: z = F::msub_shared(x, y, y, x.sqr()); break;   // x*y - y*x^2, one reduction
    }
    for (int k = 0; k < F::N; k++) r[i * F::N + k] = z.l[k];
}

template<class F>
static RustError selftest(int op, size_t n, void* r, const void* a, const void* b)
{
    try {
        const gpu_t& gpu = select_gpu(-1);
        const stream_t& s = gpu[0];
        size_t bytes = n * F::N * 4;
        dev_ptr_t<uint32_t> da(n * F::N, s), db(n * F::N, s), dr(n * F::N, s);
        s.HtoD(da, a, bytes);
        s.HtoD(db, b, bytes);
        selftest_kernel<F><<<(unsigned)((n + 127) / 128), 128, 0, s>>>(op, n, dr, da, db);
        COUNT_LAUNCH();
        CUDA_OK(cudaGetLastError());
        s.DtoH(r, dr, bytes);
        s.sync();
    } catch (const cuda_error& e) {
        return rust_err(e.code(), e.what());
    }
    return rust_ok();
}

extern "C" RustError sppark_b200_selftest_field(int field, int op, size_t n, void* r, const void* a, const void* b)
{
    switch (field) {
    case 0: return selftest<ff::bls12_381_fp_t>(op, n, r, a, b);
    case 1: return selftest<ff::bls12_381_fr_t>(op, n, r, a, b);
    case 2: return selftest<ff::pallas_fp_t>(op, n, r, a, b);
    case 3: return selftest<ff::vesta_fp_t>(op, n, r, a, b);
    case 4: return selftest<ff::bn254_fp_t>(op, n, r, a, b);
    case 5: return selftest<ff::bn254_fr_t>(op, n, r, a, b);
    case 6: return selftest<ff::bls12_377_fp_t>(op, n, r, a, b);
    case 7: return selftest<ff::bls12_377_fr_t>(op, n, r, a, b);
    default: return rust_err(-(int)cudaErrorInvalidValue, "selftest: unknown field");
    }
}

RustError msm_preload_bls12_381(const void* points, size_t npoints, size_t stride, bool has_flag, void** d_points)
{   return msm_preload<ff::bls12_381_fp_t>(points, npoints, stride, has_flag, d_points);   }
RustError msm_resident_bls12_381(void* out, const void* d_points, size_t npoints, const void* scalars, bool mont)
{
    return msm_host<ff::bls12_381_fp_t>(out, nullptr, npoints, scalars, 0, false, mont ? scalars_from_mont<ff::bls12_381_fr_t> : nullptr,
                      (const uint32_t*)d_points);
}
